"""GPU diagnostic: config #2 prefix vs the C oracle, divergence statistics + kernel launch info."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import c_oracle as co
from sccav_cbf_b200 import ops, scenarios as sc

dev = torch.device("cuda", 0)
for N in (64, 2048, 65536, 1 << 20):
    print("launch info N=%d" % N, ops.rollout_launch_info(8, N, 2034))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
b = sc.config2(n_total=65536, M=8, T=T, lo=0, hi=n)
prm = ops.make_params(**b.params)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
g = ops.rollout(prm, b.slot_desc, t(b.state), t(b.obst), tuple(t(c) for c in b.course), T, record_stride=10)
torch.cuda.synchronize()
g = {k: v.cpu().numpy() for k, v in g.items()}
r = co.rollout(co.default_params(**b.params), b.slot_desc, b.state, b.obst, b.course, T, record_stride=10)
exact = (g["steps"] == r["steps"]) & (g["target_idx"] == r["target_idx"]) & (g["n_active"] == r["n_active"]) & (g["n_infeasible"] == r["n_infeasible"])
print("exact bookkeeping frac", exact.mean(), "evals/step mean", (g["n_evals"] / np.maximum(g["steps"], 1)).mean())
err = (np.abs(g["state"] - r["state"]) / (1 + np.abs(r["state"]))).max(axis=0)
print("state err percentiles", np.percentile(err, [50, 90, 99, 99.9, 100]))
bad = np.where(err > 1e-6)[0]
print("vehicles with err>1e-6:", len(bad), bad[:20], "of which exact:", exact[bad].sum())
tr_g, tr_r = g["traj"], r["traj"]
e_t = np.nanmax(np.abs(tr_g[:, :4] - tr_r[:, :4]) / (1 + np.abs(tr_r[:, :4])), axis=1)     # [Trec, N]
print("max err over vehicles by recorded step (every 10th):", np.array2string(np.nanmax(e_t, axis=1)[::5], precision=2))
for v in bad[:5]:
    first = np.argmax(e_t[:, v] > 1e-9)
    print("vehicle", v, "first rec step with err>1e-9:", first * 10, "state there", tr_r[first, :4, v], "idx", r["traj_idx"][first, v],
          "mask g/r", g["traj_mask"][first, v], r["traj_mask"][first, v], "n_inf", r["n_infeasible"][v], "final", r["state"][:, v])
idx_same = (g["traj_idx"] == r["traj_idx"]).all(axis=0); m_same = (g["traj_mask"].view(np.uint32) == r["traj_mask"]).all(axis=0)
print("traj idx same frac", idx_same.mean(), "mask same frac", m_same.mean())
print("oracle: x range", r["state"][0].min(), r["state"][0].max(), "target_idx==last frac", (r["target_idx"] == 2033).mean(), "ninf>0 frac", (r["n_infeasible"] > 0).mean())
