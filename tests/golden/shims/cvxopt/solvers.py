"""`cvxopt.solvers` stand-in: `cp(F)` evaluates the reference's closure and returns the exact
optimum of the 2-variable QP it encodes (see ../README.md).  Every call is logged in `LOG`."""
import os
import sys

import numpy as np

options = {}
LOG = []

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.abspath(os.path.join(_here, "..", "..", "..", ".."))
if _repo not in sys.path:
    sys.path.insert(0, _repo)


def cp(F, G=None, h=None, dims=None, A=None, b=None, kktsolver=None):
    from cvxopt import matrix
    from oracle import oracle as orc

    m, x0 = F()
    x0 = matrix(x0)
    n = len(x0)
    assert n == 2
    zero = matrix(0.0, (n, 1))
    f0, Df0 = F(zero)
    # constraint rows: f[1:] = -(A x - b) <= 0  ->  A = -Df[1:], b = f[1:](x = 0)
    Arows = -np.array(Df0.a[1:, :])
    brow = np.array(f0.a[1:, 0])
    # objective (x - r)^T R (x - r): H = z0 * 2R, gradient at 0 = -2 R r
    z = matrix(1.0, (m + 1, 1))
    _, _, H = F(zero, z)
    R = np.array(H.a) / 2.0
    g0 = np.array(Df0.a[0, :])
    r = np.linalg.solve(-2.0 * R.T, g0)
    u0, u1, mask, status = orc.qp2_exact(list(Arows[:, 0]), list(Arows[:, 1]), list(brow),
                                         float(r[0]), float(r[1]),
                                         (float(R[0, 0]), float(R[0, 1]), float(R[1, 0]), float(R[1, 1])))
    LOG.append(dict(A=Arows.copy(), b=brow.copy(), r=np.array(r), R=R.copy(), x0=np.array(x0.a).ravel(),
                    u=np.array([u0, u1]), mask=mask, status=status))
    return {"x": matrix([u0, u1]), "status": "optimal", "mask": mask, "qp_status": status}
