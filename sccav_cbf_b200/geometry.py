"""CARLA-style ``Rotation`` / ``Transform`` with the interface of the reference's cbf/geometry.py.

Host-side, per-object, plain Python floats: these types only describe obstacle bounding boxes on
their way INTO the path (``BoundingBox`` -> ``Ellipse2D`` / ``CollisionCone2D`` parameters,
cbf/obstacles.py:59-88,294-331,512-543); no arithmetic of the per-step solve runs through them.
The quaternion / matrix conventions are euclid's (heading about y, attitude about z, bank about
x; in-place post-multiplication), restated in ``sccav_cbf_b200.euclid``.
"""
from __future__ import annotations

from .euclid import Matrix4, Point3, Quaternion, Vector3


class Rotation:
    """cbf/geometry.py:13-112.  ``yaw`` is what the barrier constructors read (``theta`` of an
    ellipse, cbf/obstacles.py:303,330)."""

    def __init__(self, roll=0.0, pitch=0.0, yaw=0.0, right_handed=True):
        self.update(roll, pitch, yaw, right_handed)

    def update(self, roll=0.0, pitch=0.0, yaw=0.0, right_handed=True):
        self.pitch = pitch
        self.yaw = yaw
        self.roll = roll
        self.heading = self.yaw
        self.attitude = self.pitch
        self.bank = self.roll
        self._right_handed = right_handed
        self._quaternion = Quaternion.new_rotate_euler(yaw, pitch, roll)      # cbf/geometry.py:39

    def __eq__(self, other):
        return self.pitch == other.pitch and self.yaw == other.yaw and self.roll == other.roll

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return "%s(roll = %s, pitch = %s, yaw = %s, quaternion = %s)" % (type(self).__name__, self.roll, self.pitch,
                                                                         self.yaw, self._quaternion)

    __str__ = __repr__

    def set_right_handed_flag(self, _right_handed):
        self._right_handed = _right_handed

    def get_quaternion(self):
        return self._quaternion

    def get_up_vector(self):
        return self._quaternion * Vector3(0.0, 0.0, 1.0)

    def get_right_vector(self):
        return self._quaternion * (Vector3(0.0, -1.0, 0.0) if self._right_handed else Vector3(0.0, 1.0, 0.0))

    def get_forward_vector(self):
        return self._quaternion * Vector3(1.0, 0, 0)

    @classmethod
    def from_quaternion(cls, w=1.0, x=0.0, y=0.0, z=0.0):
        heading, attitude, bank = Quaternion(w=w, x=x, y=y, z=z).get_euler()
        return cls(roll=bank, pitch=attitude, yaw=heading)


class Transform:
    """cbf/geometry.py:114-146: ``Matrix4().rotate_euler(h, a, b).translate(x, y, z)``."""

    def __init__(self, location=None, rotation=None):
        self.location = Vector3() if location is None else location
        self.rotation = Rotation() if rotation is None else rotation
        self._matrix = Matrix4()
        self._matrix.rotate_euler(self.rotation.heading, self.rotation.attitude, self.rotation.bank)
        self._matrix.translate(self.location.x, self.location.y, self.location.z)

    def __eq__(self, other):
        return self.location == other.location and self.rotation == other.rotation

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return "Transform(%r)" % (self._matrix,)

    __str__ = __repr__

    def transform(self, p=None):
        return self._matrix.transform(Point3() if p is None else p)

    def transform_inverse(self, p=None):
        return self._matrix.inverse().transform(Point3() if p is None else p)

    def get_forward_vector(self):
        return self.rotation.get_forward_vector()

    def get_inverse_matrix(self):
        return self._matrix.inverse()

    def get_matrix(self):
        return self._matrix

    def get_right_vector(self):
        return self.rotation.get_right_vector()

    def get_up_vector(self):
        return self.rotation.get_up_vector()
