"""Developer tool: kernel-only timings of configs 2 (fast / canonical), 3, 4 (quarter), 5 (1 M) for A/B runs of library variants."""
import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sccav_cbf_b200 import scenarios as sc
from sccav_cbf_b200.rollout import ClosedLoopRollout
dev = torch.device("cuda", 0)
def t(batch, flags, dtype=torch.float64, reps=3):
    batch.params = dict(batch.params, flags=flags)
    cl = ClosedLoopRollout(batch, dtype=dtype, device=dev, pin=False)
    cl.run(); torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        cl.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); cl.launch(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return min(ms)
print("config2 flags5 %.3f flags0 %.3f" % (t(sc.config2(n_total=65536, M=8, T=1000), 5), t(sc.config2(n_total=65536, M=8, T=1000), 0)))
print("config3 flags0 %.2f flags21 %.2f" % (t(sc.config3(n_total=262144, M=16, T=600), 0), t(sc.config3(n_total=262144, M=16, T=600), 21)))
print("config4/4 flags5 %.2f flags0 %.2f" % (t(sc.config4(n_total=262144, M=8, T=1000), 5), t(sc.config4(n_total=262144, M=8, T=1000), 0)))
print("config5/2 flags5 %.2f" % t(sc.config5(n_total=16777216, T=300, lo=0, hi=1048576), 5))
