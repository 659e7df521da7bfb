"""Developer tool: time the 64-roads launch (sccav_rollout_roads_*) and print its launch geometry (SCCAV_DEBUG_LAUNCH=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sccav_cbf_b200 import ops, scenarios as sc, _native as nv
dev = torch.device("cuda", 0)
n_roads, per_road = 64, 1024
(rcx, rcy, rcyaw, rnp), rnph, rs, ro = sc.roads(n_roads, per_road, M=8, seed=4, dtype=torch.float64, device=dev)
d_rs = torch.from_numpy(rs).to(dev); d_ro = torch.from_numpy(ro).to(dev)
rsd = [nv.SLOT_ELLIPSE | nv.SLOT_STATIC] * 8
for fl in (5, 1, 0):
    prm = ops.make_params(flags=fl)
    out = {}
    for _ in range(2):
        rr = ops.rollout(prm, rsd, d_rs, d_ro, (rcx, rcy, rcyaw), 1000, course_np=rnp, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rr = ops.rollout(prm, rsd, d_rs, d_ro, (rcx, rcy, rcyaw), 1000, course_np=rnp, out=out); e1.record()
    torch.cuda.synchronize()
    print("flags", fl, "ms %.3f" % e0.elapsed_time(e1), "np max", int(rnph.max()), "evals/step %.1f" % (float(rr["n_evals"].double().sum()) / float(rr["steps"].sum())), flush=True)
