"""Minimal stand-in for the `euclid` package (golden-vector generation only; see ../README.md)."""
import math


class Vector2:
    __slots__ = ("x", "y")

    def __init__(self, x=0, y=0):
        self.x = x
        self.y = y

    def copy(self):
        return self.__class__(self.x, self.y)

    def __eq__(self, o):
        return self.x == o.x and self.y == o.y

    def __add__(self, o):
        return Vector2(self.x + o.x, self.y + o.y)

    def __sub__(self, o):
        return Vector2(self.x - o.x, self.y - o.y)

    def __mul__(self, k):
        return Vector2(self.x * k, self.y * k)

    __rmul__ = __mul__

    def __neg__(self):
        return Vector2(-self.x, -self.y)

    def magnitude(self):
        return math.sqrt(self.x ** 2 + self.y ** 2)

    __abs__ = magnitude

    def __repr__(self):
        return "Vector2(%.2f, %.2f)" % (self.x, self.y)


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

    def __eq__(self, o):
        return self.x == o.x and self.y == o.y and self.z == o.z

    def __add__(self, o):
        return Vector3(self.x + o.x, self.y + o.y, self.z + o.z)

    def __sub__(self, o):
        return Vector3(self.x - o.x, self.y - o.y, self.z - o.z)

    def __mul__(self, k):
        return Vector3(self.x * k, self.y * k, self.z * k)

    __rmul__ = __mul__

    def __neg__(self):
        return Vector3(-self.x, -self.y, -self.z)

    def magnitude(self):
        return math.sqrt(self.x ** 2 + self.y ** 2 + self.z ** 2)

    def normalized(self):
        d = self.magnitude()
        return Vector3(self.x / d, self.y / d, self.z / d) if d else self.copy()


class Point3(Vector3):
    pass


class Quaternion:
    def __init__(self, w=1, x=0, y=0, z=0):
        self.w, self.x, self.y, self.z = w, x, y, z

    @classmethod
    def new_rotate_euler(cls, heading, attitude, bank):
        c1, s1 = math.cos(heading / 2), math.sin(heading / 2)
        c2, s2 = math.cos(attitude / 2), math.sin(attitude / 2)
        c3, s3 = math.cos(bank / 2), math.sin(bank / 2)
        return cls(c1 * c2 * c3 - s1 * s2 * s3, s1 * s2 * c3 + c1 * c2 * s3,
                   s1 * c2 * c3 + c1 * s2 * s3, c1 * s2 * c3 - s1 * c2 * s3)

    def __mul__(self, v):
        w, x, y, z = self.w, self.x, self.y, self.z
        vx, vy, vz = v.x, v.y, v.z
        ww, xx, yy, zz = w * w, x * x, y * y, z * z
        wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
        return v.__class__(ww * vx + 2 * wy * vz - 2 * wz * vy + xx * vx + 2 * xy * vy + 2 * xz * vz - zz * vx - yy * vx,
                           2 * xy * vx + yy * vy + 2 * yz * vz + 2 * wz * vx - zz * vy + ww * vy - 2 * wx * vz - xx * vy,
                           2 * xz * vx + 2 * yz * vy + zz * vz - 2 * wy * vx - yy * vz + 2 * wx * vy - xx * vz + ww * vz)


class Matrix4:
    def __init__(self):
        self.m = [[1.0 if i == j else 0.0 for j in range(4)] for i in range(4)]

    def rotate_euler(self, heading, attitude, bank):
        return self

    def translate(self, x, y, z):
        return self

    def transform(self, p):
        return p

    def inverse(self):
        return self
