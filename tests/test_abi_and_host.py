"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/sccav_cbf.h
declares; host-side logic (scenario sharding, course generation, parameter marshalling)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from sccav_cbf_b200 import _native as nv
    from sccav_cbf_b200 import build
    build.build()
    L = nv.lib()
    hdr = open(os.path.join(ROOT, "include", "sccav_cbf.h")).read()
    declared = sorted(set(re.findall(r"\b(sccav_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), "libsccav_cbf.so does not export %s" % name
    assert sorted(nv.SYMBOLS) == declared
    assert L.sccav_version() == 100
    assert L.sccav_launch_count() == 0


def test_params_struct_layout_matches_header_defaults():
    from oracle import c_oracle as co
    from sccav_cbf_b200 import _native as nv
    from sccav_cbf_b200 import ops
    assert ctypes.sizeof(nv.Params) == ctypes.sizeof(co.Params) == 8 * 4 + 20 * 8
    p = nv.default_params()
    q = co.default_params()
    for name, _ in nv.Params._fields_:
        a, b = getattr(p, name), getattr(q, name)
        if name == "R":
            assert list(a) == list(b) == [1.0, 0.0, 0.0, 1.0]
        else:
            assert a == b, name
    assert p.max_steer == float(np.radians(30.0)) and p.lr == 1.45 and p.lf == 1.45
    r = ops.make_params(R=[[0.5, 0.0], [0.0, 20000.0]], alpha=2.0)
    assert list(r.R) == [0.5, 0.0, 0.0, 20000.0] and r.alpha == 2.0
    with pytest.raises(TypeError):
        ops.make_params(nonsense=1)


def test_no_cpu_fallback_without_device():
    """The product path must fail loudly when there is no CUDA device (never route to the oracle)."""
    import torch
    from sccav_cbf_b200 import _native as nv
    from sccav_cbf_b200 import ops
    if nv.lib().sccav_device_ok():
        pytest.skip("a CUDA device is present")
    with pytest.raises(nv.SccavError):
        ops.filter_step(ops.make_params(), [0], torch.zeros(4, 2, dtype=torch.float64), torch.ones(1, 8, 2, dtype=torch.float64),
                        torch.zeros(2, 2, dtype=torch.float64))
    src = "".join(open(os.path.join(ROOT, "sccav_cbf_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "sccav_cbf_b200")) if f.endswith(".py"))
    # no import of the checker anywhere in the product package, in any spelling
    assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M)
    assert "import_module(\"oracle" not in src and "__import__(\"oracle" not in src and "liboracle" not in src


def test_course_matches_reference_planner(refvec):
    from sccav_cbf_b200.course import config1_course
    cx, cy, cyaw = config1_course()
    assert len(cx) == 2034
    assert np.array_equal(cx, refvec["course"][0]) and np.array_equal(cy, refvec["course"][1])
    assert np.array_equal(cyaw, refvec["course"][2])


def test_scenarios_are_independent_of_sharding():
    from sccav_cbf_b200 import scenarios as sc
    whole = sc.config2(n_total=200000, M=8, T=10, lo=0, hi=200000)
    for world in (2, 4, 8):
        parts = []
        for rank in range(world):
            lo, hi = sc.shard_range(200000, rank, world)
            parts.append(sc.config2(n_total=200000, M=8, T=10, lo=lo, hi=hi))
        assert np.array_equal(np.concatenate([p.state for p in parts], axis=1), whole.state)
        assert np.array_equal(np.concatenate([p.obst for p in parts], axis=2), whole.obst)
    a = sc.config5(n_total=1 << 20, lo=70000, hi=70100)
    b = sc.config5(n_total=1 << 20, lo=70050, hi=70100)
    assert np.array_equal(a.alpha[50:], b.alpha) and np.array_equal(a.state[:, 50:], b.state)
    c3 = sc.config3(n_total=1000, M=16, T=5, lo=10, hi=20)
    assert c3.obst.shape == (16, 8, 10) and c3.params["seeker"] == 1
    c4 = sc.config4(n_total=1000, M=8, T=5, lo=0, hi=10)
    assert c4.M == 10 and c4.slot_desc[-1] == (2 | 0x80)


def test_header_constants_match_their_python_mirrors():
    """Slot types, slot flags, models, parameter flags and status codes of include/sccav_cbf.h against
    sccav_cbf_b200._native (the product's ctypes layer) and oracle.oracle (the checker)."""
    from oracle import oracle as o
    from sccav_cbf_b200 import _native as nv
    hdr = open(os.path.join(ROOT, "include", "sccav_cbf.h")).read()
    defs = {k: int(v, 0) for k, v in re.findall(r"#define\s+(SCCAV_[A-Z0-9_]+)\s+(0x[0-9a-fA-F]+|-?\d+)\b", hdr)}
    pairs = {
        "SCCAV_SLOT_ELLIPSE": "SLOT_ELLIPSE", "SCCAV_SLOT_CONE": "SLOT_CONE", "SCCAV_SLOT_LANE": "SLOT_LANE",
        "SCCAV_SLOT_RADIAL": "SLOT_RADIAL", "SCCAV_SLOT_DISTANCE": "SLOT_DISTANCE", "SCCAV_SLOT_ELLIPSE_PREP": "SLOT_ELLIPSE_PREP",
        "SCCAV_SLOT_LANE_SQRT": "SLOT_LANE_SQRT", "SCCAV_SLOT_STATIC": "SLOT_STATIC",
        "SCCAV_MODEL_DBM": "MODEL_DBM", "SCCAV_MODEL_KBM": "MODEL_KBM", "SCCAV_MODEL_NONE": "MODEL_NONE", "SCCAV_MODEL_DUM": "MODEL_DUM",
        "SCCAV_MODEL_SADBM": "MODEL_SADBM", "SCCAV_FLAG_PREPARED_ROWS": "FLAG_PREPARED_ROWS", "SCCAV_FLAG_QP_ENUMERATE": "FLAG_QP_ENUMERATE",
        "SCCAV_FLAG_FUSED_STEER": "FLAG_FUSED_STEER", "SCCAV_FLAG_BETA_IO": "FLAG_BETA_IO", "SCCAV_FLAG_SEEKER_DIRECT": "FLAG_SEEKER_DIRECT",
    }
    for c_name, py_name in pairs.items():
        assert c_name in defs, c_name
        assert getattr(nv, py_name) == defs[c_name], py_name
        assert getattr(o, py_name) == defs[c_name], py_name
    assert (o.STATUS_INACTIVE, o.STATUS_ACTIVE, o.STATUS_INFEASIBLE) == (defs["SCCAV_STATUS_INACTIVE"], defs["SCCAV_STATUS_ACTIVE"], defs["SCCAV_STATUS_INFEASIBLE"])
    assert nv.NFIELD == defs["SCCAV_NFIELD"] == o.NFIELD and nv.MAX_ROWS == defs["SCCAV_MAX_ROWS"]
