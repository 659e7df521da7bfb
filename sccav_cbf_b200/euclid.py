"""Minimal 2-D / 3-D vector types with the surface of the ``euclid`` package the reference touches.

The reference does ``from euclid import *`` (cbf/obstacles.py:25, cbf/geometry.py:11,
cbf/controllers.py:27) and uses ``Vector2/Point2`` as ellipse centres and front-axle coordinates
(cbf/obstacles.py:146-156, cbf/controllers.py:105-110) and ``Vector3/Point3`` inside ``BoundingBox``
(cbf/obstacles.py:59-88).  ``euclid`` is neither vendored nor pinned there and is not installable
here, so this module provides the same names.  Components may be Python numbers or torch tensors
of shape [N] (one value per vehicle of a batch).
"""
from __future__ import annotations

import math

__all__ = ["Vector2", "Point2", "Vector3", "Point3", "Quaternion", "Matrix4"]


def _sqrt(v):
    return v.sqrt() if hasattr(v, "sqrt") else math.sqrt(v)


class Vector2:
    __slots__ = ("x", "y")

    def __init__(self, x=0, y=0):
        self.x = x
        self.y = y

    def copy(self):
        return self.__class__(self.x, self.y)

    __copy__ = copy

    def __eq__(self, o):
        if not hasattr(o, "x"):
            return False
        ex, ey = self.x == o.x, self.y == o.y
        return bool(ex.all() if hasattr(ex, "all") else ex) and bool(ey.all() if hasattr(ey, "all") else ey)

    def __ne__(self, o):
        return not self.__eq__(o)

    def __add__(self, o):
        return self.__class__(self.x + o.x, self.y + o.y)

    def __sub__(self, o):
        return Vector2(self.x - o.x, self.y - o.y)

    def __mul__(self, k):
        return self.__class__(self.x * k, self.y * k)

    __rmul__ = __mul__

    def __truediv__(self, k):
        return self.__class__(self.x / k, self.y / k)

    def __neg__(self):
        return self.__class__(-self.x, -self.y)

    def __iter__(self):
        return iter((self.x, self.y))

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def dot(self, o):
        return self.x * o.x + self.y * o.y

    def magnitude_squared(self):
        return self.x ** 2 + self.y ** 2

    def magnitude(self):
        return _sqrt(self.x ** 2 + self.y ** 2)

    __abs__ = magnitude

    def normalized(self):
        d = self.magnitude()
        return self.__class__(self.x / d, self.y / d)

    def __repr__(self):
        return "%s(%r, %r)" % (type(self).__name__, self.x, self.y)


class Point2(Vector2):
    pass


class Vector3:
    __slots__ = ("x", "y", "z")

    def __init__(self, x=0, y=0, z=0):
        self.x = x
        self.y = y
        self.z = z

    def copy(self):
        return self.__class__(self.x, self.y, self.z)

    __copy__ = copy

    def __eq__(self, o):
        return hasattr(o, "z") and self.x == o.x and self.y == o.y and self.z == o.z

    def __ne__(self, o):
        return not self.__eq__(o)

    def __add__(self, o):
        return self.__class__(self.x + o.x, self.y + o.y, self.z + o.z)

    def __sub__(self, o):
        return Vector3(self.x - o.x, self.y - o.y, self.z - o.z)

    def __mul__(self, k):
        return self.__class__(self.x * k, self.y * k, self.z * k)

    __rmul__ = __mul__

    def __truediv__(self, k):
        return self.__class__(self.x / k, self.y / k, self.z / k)

    def __neg__(self):
        return self.__class__(-self.x, -self.y, -self.z)

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __len__(self):
        return 3

    def __getitem__(self, i):
        return (self.x, self.y, self.z)[i]

    def dot(self, o):
        return self.x * o.x + self.y * o.y + self.z * o.z

    def magnitude(self):
        return _sqrt(self.x ** 2 + self.y ** 2 + self.z ** 2)

    __abs__ = magnitude

    def normalized(self):
        d = self.magnitude()
        return self.__class__(self.x / d, self.y / d, self.z / d) if d else self.copy()

    def __repr__(self):
        return "%s(%r, %r, %r)" % (type(self).__name__, self.x, self.y, self.z)


class Point3(Vector3):
    pass


class Quaternion:
    """Unit quaternion with euclid's conventions (heading about y, attitude about z, bank about x):
    the published algorithm of euclid's ``Quaternion.new_rotate_euler`` / ``__mul__`` /
    ``get_euler`` (euclideanspace.com formulas), restated because cbf/geometry.py:39,84-111 uses it."""
    __slots__ = ("w", "x", "y", "z")

    def __init__(self, w=1, x=0, y=0, z=0):
        self.w, self.x, self.y, self.z = w, x, y, z

    @classmethod
    def new_rotate_euler(cls, heading, attitude, bank):
        c1, s1 = math.cos(heading / 2), math.sin(heading / 2)
        c2, s2 = math.cos(attitude / 2), math.sin(attitude / 2)
        c3, s3 = math.cos(bank / 2), math.sin(bank / 2)
        return cls(c1 * c2 * c3 - s1 * s2 * s3, s1 * s2 * c3 + c1 * c2 * s3,
                   s1 * c2 * c3 + c1 * s2 * s3, c1 * s2 * c3 - s1 * c2 * s3)

    def get_euler(self):
        t = self.x * self.y + self.z * self.w
        if t > 0.4999:
            return 2 * math.atan2(self.x, self.w), math.pi / 2, 0
        if t < -0.4999:
            return -2 * math.atan2(self.x, self.w), -math.pi / 2, 0
        sqx, sqy, sqz = self.x ** 2, self.y ** 2, self.z ** 2
        heading = math.atan2(2 * self.y * self.w - 2 * self.x * self.z, 1 - 2 * sqy - 2 * sqz)
        attitude = math.asin(2 * t)
        bank = math.atan2(2 * self.x * self.w - 2 * self.y * self.z, 1 - 2 * sqx - 2 * sqz)
        return heading, attitude, bank

    def __mul__(self, v):
        w, x, y, z = self.w, self.x, self.y, self.z
        if isinstance(v, Quaternion):
            return Quaternion(w * v.w - x * v.x - y * v.y - z * v.z, w * v.x + x * v.w + y * v.z - z * v.y,
                              w * v.y + y * v.w + z * v.x - x * v.z, w * v.z + z * v.w + x * v.y - y * v.x)
        vx, vy, vz = v.x, v.y, v.z
        ww, xx, yy, zz = w * w, x * x, y * y, z * z
        wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
        return v.__class__(ww * vx + 2 * wy * vz - 2 * wz * vy + xx * vx + 2 * xy * vy + 2 * xz * vz - zz * vx - yy * vx,
                           2 * xy * vx + yy * vy + 2 * yz * vz + 2 * wz * vx - zz * vy + ww * vy - 2 * wx * vz - xx * vy,
                           2 * xz * vx + 2 * yz * vy + zz * vz - 2 * wy * vx - yy * vz + 2 * wx * vy - xx * vz + ww * vz)

    def __repr__(self):
        return "Quaternion(real=%.2f, imag=<%.2f, %.2f, %.2f>)" % (self.w, self.x, self.y, self.z)


class Matrix4:
    """Row-major 4x4 affine matrix with euclid's in-place post-multiplying ``rotate_euler`` /
    ``translate`` (so ``rotate_euler(...).translate(t)`` maps p to R (p + t), which is what
    cbf/geometry.py:118-122 builds)."""

    def __init__(self, rows=None):
        self.m = rows if rows is not None else [[1.0 if i == j else 0.0 for j in range(4)] for i in range(4)]

    def _imul(self, o):
        a, b = self.m, o
        self.m = [[sum(a[i][k] * b[k][j] for k in range(4)) for j in range(4)] for i in range(4)]
        return self

    def rotate_euler(self, heading, attitude, bank):
        ch, sh = math.cos(heading), math.sin(heading)
        ca, sa = math.cos(attitude), math.sin(attitude)
        cb, sb = math.cos(bank), math.sin(bank)
        r = [[ch * ca, sh * sb - ch * sa * cb, ch * sa * sb + sh * cb, 0.0],
             [sa, ca * cb, -ca * sb, 0.0],
             [-sh * ca, sh * sa * cb + ch * sb, -sh * sa * sb + ch * cb, 0.0],
             [0.0, 0.0, 0.0, 1.0]]
        return self._imul(r)

    def translate(self, x, y, z):
        t = [[1.0, 0.0, 0.0, x], [0.0, 1.0, 0.0, y], [0.0, 0.0, 1.0, z], [0.0, 0.0, 0.0, 1.0]]
        return self._imul(t)

    def transform(self, p):
        m = self.m
        return p.__class__(m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3],
                           m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3],
                           m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3])

    def inverse(self):
        import numpy as np
        return Matrix4(np.linalg.inv(np.array(self.m, dtype=np.float64)).tolist())

    def __repr__(self):
        return "Matrix4(%s)" % (self.m,)
