"""N > 1 path on CPU: two ranks over gloo (127.0.0.1).  The hot path has no collective; what is
multi-rank is (1) the contiguous shard partition of the scenario axis, with inputs that depend only
on the GLOBAL scenario index, (2) the barrier / max / sum reductions bench.py uses around its timed
region, and (3) the final gather of per-shard summaries.  Each rank 'processes' its shard with the
CPU oracle (the checker -- this test never claims to exercise the CUDA path)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, T, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from oracle import c_oracle as co
    from sccav_cbf_b200 import scenarios as sc
    from sccav_cbf_b200.dist import Shards
    sh = Shards(backend="gloo")
    lo, hi = sc.shard_range(n_total, sh.rank, sh.world)
    b = sc.config2(n_total=n_total, M=4, T=T, seed=0, lo=lo, hi=hi)
    sh.barrier()
    r = co.rollout(co.default_params(**b.params), b.slot_desc, b.state, b.obst, b.course, T, nthreads=1)
    sh.barrier()
    tmax = sh.max(float(rank + 1))                      # stands for the per-rank device time
    solves = sh.sum(float(r["steps"].sum()) * b.M)
    full = sh.gather_summaries({"state": torch.from_numpy(r["state"]), "steps": torch.from_numpy(r["steps"]),
                                "n_active": torch.from_numpy(r["n_active"])})
    if rank == 0:
        q.put((tmax, solves, {k: v.numpy() for k, v in full.items()}, (lo, hi)))
    sh.close()


@pytest.mark.timeout(300)
def test_two_rank_shards_reproduce_the_single_process_batch():
    from oracle import c_oracle as co
    from sccav_cbf_b200 import scenarios as sc
    n_total, T, world = 96, 120, 2
    assert sc.shard_range(n_total, 0, 2) == (0, 48) and sc.shard_range(n_total, 1, 2) == (48, 96)
    assert sc.shard_range(10, 3, 4) == (9, 10) and sc.shard_range(2, 3, 4) == (2, 2)      # ragged / empty tail shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, solves, full, shard0 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = sc.config2(n_total=n_total, M=4, T=T, seed=0)
    ref = co.rollout(co.default_params(**b.params), b.slot_desc, b.state, b.obst, b.course, T, nthreads=2)
    assert shard0 == (0, 48) and tmax == 2.0
    assert solves == float(ref["steps"].sum()) * 4
    assert np.array_equal(full["steps"], ref["steps"]) and np.array_equal(full["n_active"], ref["n_active"])
    assert np.array_equal(full["state"], ref["state"])              # bit-identical: a vehicle's result does not depend on its shard


def test_scenarios_depend_only_on_the_global_index():
    from sccav_cbf_b200 import scenarios as sc
    whole = sc.config2(n_total=200000, M=3, T=10, lo=65000, hi=66000)       # straddles a 65,536 block boundary
    a = sc.config2(n_total=200000, M=3, T=10, lo=65000, hi=65536)
    b = sc.config2(n_total=200000, M=3, T=10, lo=65536, hi=66000)
    assert np.array_equal(whole.state, np.concatenate([a.state, b.state], axis=1))
    assert np.array_equal(whole.obst, np.concatenate([a.obst, b.obst], axis=2))
    c5 = sc.config5(n_total=1 << 20, T=5, lo=70000, hi=70010)
    c5b = sc.config5(n_total=1 << 24, T=5, lo=70000, hi=70010)
    assert np.array_equal(c5.alpha, c5b.alpha) and np.array_equal(c5.state, c5b.state)
