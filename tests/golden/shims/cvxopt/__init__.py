"""Minimal stand-in for `cvxopt` (golden-vector generation only; see ../README.md).

`matrix` follows cvxopt's conventions: column-major storage, a flat list is a column vector, a
list of lists is a list of COLUMNS (or of block-columns when the entries are matrices), `*`
between two matrices is the matrix product, a 1x1 matrix acts as a scalar, single-index
access is linear in column-major order.
"""
import numpy as np

from . import solvers  # noqa: F401  (re-exported: `from cvxopt import matrix, solvers`)


def _is_scalar(v):
    return isinstance(v, (int, float, np.floating, np.integer))


class matrix:
    def __init__(self, x=None, size=None, tc=None):
        if isinstance(x, matrix):
            a = x.a.copy()
        elif _is_scalar(x):
            a = np.full(size, float(x), dtype=np.float64)
            size = None
        elif isinstance(x, np.ndarray):
            a = np.array(x, dtype=np.float64)
            if a.ndim == 1:
                a = a.reshape(-1, 1)
        elif isinstance(x, (list, tuple)):
            if len(x) > 0 and isinstance(x[0], (list, tuple)):
                cols = []
                for col in x:
                    if len(col) > 0 and isinstance(col[0], matrix):
                        cols.append(np.vstack([c.a for c in col]))
                    else:
                        cols.append(np.array([_tofloat(v) for v in col], dtype=np.float64).reshape(-1, 1))
                a = np.hstack(cols)
            else:
                a = np.array([_tofloat(v) for v in x], dtype=np.float64).reshape(-1, 1)
        else:
            raise TypeError("matrix(): unsupported %r" % type(x))
        if size is not None:
            a = a.reshape(size, order="F")
        self.a = a

    # ---- shape
    @property
    def size(self):
        return self.a.shape

    @property
    def T(self):
        return matrix(self.a.T.copy())

    def __len__(self):
        return self.a.size

    def __iter__(self):
        return iter(self.a.ravel(order="F").tolist())

    def __float__(self):
        assert self.a.size == 1
        return float(self.a.ravel()[0])

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    # ---- indexing
    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            r = self.a[idx]
            if np.isscalar(r) or r.ndim == 0:
                return float(r)
            i, j = idx
            rr = np.atleast_2d(r)
            if isinstance(j, slice) and not isinstance(i, slice):
                rr = rr.reshape(1, -1)
            elif isinstance(i, slice) and not isinstance(j, slice):
                rr = rr.reshape(-1, 1)
            return matrix(rr.copy())
        flat = self.a.ravel(order="F")
        r = flat[idx]
        if isinstance(idx, slice):
            return matrix(r.copy().reshape(-1, 1))
        return float(r)

    def __setitem__(self, idx, val):
        if isinstance(val, matrix):
            v = val.a
        else:
            v = np.asarray(val, dtype=np.float64)
        if isinstance(idx, tuple):
            tgt = self.a[idx]
            self.a[idx] = v.reshape(np.shape(tgt)) if np.size(v) == np.size(tgt) else v
            return
        flat = self.a.ravel(order="F")
        if isinstance(idx, slice):
            flat[idx] = v.ravel(order="F")
        else:
            flat[idx] = float(v.ravel()[0])
        self.a = flat.reshape(self.a.shape, order="F")

    # ---- arithmetic
    def _scalar_like(self):
        return self.a.size == 1

    def __mul__(self, o):
        if isinstance(o, matrix):
            if self._scalar_like() and o.a.shape[0] != 1:
                return matrix(float(self) * o.a)
            if o._scalar_like() and self.a.shape[1] != 1:
                return matrix(self.a * float(o))
            return matrix(_matmul(self.a, o.a))
        if isinstance(o, np.ndarray) and o.size == 1:
            o = float(o.ravel()[0])
        return matrix(self.a * float(o))

    def __rmul__(self, o):
        if isinstance(o, np.ndarray) and o.size == 1:
            o = float(o.ravel()[0])
        return matrix(float(o) * self.a)

    def _other(self, o):
        if isinstance(o, matrix):
            return o.a if o.a.size != 1 or self.a.size == 1 else float(o)
        if isinstance(o, np.ndarray):
            return o.reshape(self.a.shape) if o.size == self.a.size else float(o.ravel()[0])
        return float(o)

    def __add__(self, o):
        if isinstance(o, matrix) and self.a.size == 1 and o.a.size != 1:
            return matrix(float(self) + o.a)
        return matrix(self.a + self._other(o))

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, matrix) and self.a.size == 1 and o.a.size != 1:
            return matrix(float(self) - o.a)
        return matrix(self.a - self._other(o))

    def __rsub__(self, o):
        return matrix(self._other(o) - self.a)

    def __neg__(self):
        return matrix(-self.a)

    def __truediv__(self, o):
        return matrix(self.a / float(o))

    def __pow__(self, k):
        return matrix(self.a ** k)

    def __repr__(self):
        return "matrix(%r)" % (self.a,)


def _tofloat(v):
    if isinstance(v, matrix):
        return float(v)
    if isinstance(v, np.ndarray):
        return float(v.ravel()[0])
    return float(v)


def _matmul(a, b):
    """Plain left-to-right dot products (no BLAS, no FMA): the summation order the oracle states."""
    n, k = a.shape
    k2, m = b.shape
    assert k == k2, (a.shape, b.shape)
    out = np.zeros((n, m), dtype=np.float64)
    for i in range(n):
        for j in range(m):
            acc = None
            for t in range(k):
                term = float(a[i, t]) * float(b[t, j])
                acc = term if acc is None else acc + term
            out[i, j] = acc
    return out


def sqrt(x):
    if isinstance(x, matrix):
        return matrix(np.sqrt(x.a))
    return float(np.sqrt(x))


def spdiag(x):
    return matrix(np.diag(np.asarray(list(x), dtype=np.float64)))
