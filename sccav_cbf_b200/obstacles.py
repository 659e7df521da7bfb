"""Barrier definitions with the class API of the reference's cbf/obstacles.py, evaluated by the
CUDA path (``ops.barrier_partials`` -> K0, csrc/kernels.cuh) instead of per-object Python arithmetic.

Same names, argument meaning and error behaviour as the reference:

    Obstacle2DTypes, BoundingBox, Obstacle2DBase, Ellipse2D, CollisionCone2D, PolyLane, ObstacleList2D

Batched semantics (the extension SURVEY 8b describes): every scalar of the reference may be a
Python number (one scenario -- getters then return Python floats, like the reference) or a torch
tensor of shape [N] (N independent vehicles -- getters return CUDA tensors [N]).  The ego state
``s`` is a sequence of 4 numbers or a [4, N] tensor (x, y, theta, v).  An object owns ONE obstacle
slot; ``ObstacleList2D`` packs its values, in insertion order, into the structure-of-arrays buffer
``obst[M][8][N]`` of include/sccav_cbf.h (constraint index = dict order, as in the reference).

Reference defects that are NOT reproduced (DESIGN.md "Deviations"):
D1 ``Ellipse2D.dtheta`` raises TypeError there -> returns 0 here; D2 ``Ellipse2D.update(s=...)`` overwrote
the ellipse orientation / velocity with the ego's -> the ego state is stored separately; D3 ``update(b=)``
wrote ``a``; D4 ``from_bounding_box`` used an unbound ``id``; D5 ``ObstacleList2D.update_state`` with a
dict called ``dict.update``; D6 ``gradient()`` allocated m x 3 for 4-vectors.
There is no CPU evaluation path: without the CUDA library / a GPU every getter raises.
"""
from __future__ import annotations

import enum
import warnings
from collections.abc import MutableMapping
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nv
from . import ops
from ._batch import as_state, as_vec, batch_size, cuda_device, is_scalar, to_output
from .euclid import Point2, Vector2, Vector3
from .geometry import Rotation, Transform
from .utils import ZERO_TOL  # noqa: F401

H, HX, HY, HTH, HV, HT = range(6)


class IdentityObjects(enum.Enum):
    """cbf/obstacles.py:42-46."""
    DICT_EMPTY_UPDATE = ()


class Obstacle2DTypes(enum.Enum):
    """cbf/obstacles.py:49-55."""
    ELLIPSE2D = 0
    COLLISION_CONE2D = 1
    POLY_LANE = 2


class BoundingBox:
    """cbf/obstacles.py:59-88 (CARLA-style box: half extents, location, rotation, scalar speed)."""

    def __init__(self, extent=None, location=None, rotation=None, velocity: float = 0.0):
        self.extent = Vector3() if extent is None else extent
        self.location = Vector3() if location is None else location
        self.rotation = Rotation() if rotation is None else rotation
        self.velocity = velocity

    def __eq__(self, other):
        return self.location == other.location and self.extent == other.extent

    def __ne__(self, other):
        return not self.__eq__(other)

    def get_local_vertices(self):
        up = self.rotation.get_up_vector().normalized()
        right = self.rotation.get_right_vector().normalized()
        forward = self.rotation.get_forward_vector().normalized()
        e = self.extent
        out = []
        for sz in (-1, 1):
            for sx, sy in ((1, 1), (1, -1), (-1, -1), (-1, 1)):
                out.append(sz * e.z * up + sx * e.x * forward + sy * e.y * right)
        return out

    def get_world_vertices(self, transform: Transform = None):
        transform = Transform() if transform is None else transform
        return [transform.transform(v) for v in self.get_local_vertices()]


class Obstacle2DBase:
    """Interface every 2-D obstacle implements (cbf/obstacles.py:90-137).  Subclasses provide
    ``slot_type`` and ``fields()`` (the 8 values of their SoA slot); the getters below evaluate
    the slot on the GPU for the ego state last given to ``update`` / ``update_state``."""
    slot_type: Optional[int] = None

    def __init__(self):
        self.s = None
        self._cache = None

    # ---- SoA packing ------------------------------------------------------------------------
    def fields(self) -> List[object]:
        raise NotImplementedError

    def _invalidate(self):
        self._cache = None

    def _partials(self) -> Tuple[torch.Tensor, bool]:
        if self.slot_type is None:
            raise NotImplementedError("Obstacle2DBase has no barrier")
        if self.s is None:
            raise AttributeError("the ego state has not been set: call update(s=...) / update_state(s, s_obs) first")
        if self._cache is None:
            state, scalar = as_state(self.s)
            N = state.shape[1]
            f = self.fields()
            scalar = scalar and all(is_scalar(v) for v in f)
            N = max([N] + [batch_size(v) for v in f])
            if state.shape[1] != N:
                state = state.expand(4, N).contiguous()
            obst = torch.stack([as_vec(v, N, state.dtype, state.device) for v in f]).reshape(1, nv.NFIELD, N).contiguous()
            self._cache = (ops.barrier_partials([self.slot_type], state, obst)[0], scalar)
        return self._cache

    def _get(self, k: int):
        p, scalar = self._partials()
        return to_output(p[k], scalar)

    # ---- the getters of the reference ---------------------------------------------------------
    def evaluate(self, *args, **kwargs):
        return self._get(H) if self.slot_type is not None else 0

    def f(self, *args, **kwargs):
        """Alias of evaluate (the name cvxopt's F closure used)."""
        return self.evaluate(**kwargs)

    def dx(self, *args, **kwargs):
        return self._get(HX) if self.slot_type is not None else 0

    def dy(self, *args, **kwargs):
        return self._get(HY) if self.slot_type is not None else 0

    def dtheta(self, *args, **kwargs):
        return self._get(HTH) if self.slot_type is not None else 0

    def dv(self, *args, **kwargs):
        return self._get(HV) if self.slot_type is not None else 0

    def dt(self, *args, **kwargs):
        return self._get(HT) if self.slot_type is not None else 0

    def dbeta(self, *args, **kwargs):
        return 0

    def gradient(self, *args, **kwargs):
        """[h_x, h_y, h_theta, h_v] (cbf/obstacles.py:196-200): a list of floats, or a [4, N] tensor."""
        if self.slot_type is None:
            return [0.0, 0.0, 0.0, 0.0]
        p, scalar = self._partials()
        g = p[HX:HV + 1]
        return [float(v) for v in g[:, 0].tolist()] if scalar else g

    def update(self, *args, **kwargs):
        pass

    def update_state(self, s, s_obs=None, **kwargs):
        self.update(s=s, s_obs=s_obs)

    def update_coords(self, *args, **kwargs):
        pass

    def update_orientation(self, *args, **kwargs):
        pass


def _check_bbox(bbox):
    if not isinstance(bbox, BoundingBox):
        raise TypeError("Expected an object of type cbf.obstacles.BoundingBox as an input to fromBoundingBox() method, "
                        "but got " + type(bbox).__name__)


class Ellipse2D(Obstacle2DBase):
    """h = ((dx ct + dy st)/a)^2 + ((-dx st + dy ct)/b)^2 - 1 (cbf/obstacles.py:139-331)."""
    slot_type = nv.SLOT_ELLIPSE

    def __init__(self, a, b, center: Vector2 = None, theta=0, buffer=0, **kwargs):
        super().__init__()
        self.type = Obstacle2DTypes.ELLIPSE2D
        if "id" in kwargs:
            self.id = kwargs["id"]
        center = Vector2(0, 0) if center is None else center
        if not isinstance(center, Vector2):
            raise TypeError("Expected an object of type euclid.Vector2 for arg center, but got " + type(center).__name__ + ".")
        self.center = center
        self.theta = theta
        self.vel = Vector2()
        self.a = a + buffer                                            # cbf/obstacles.py:159-160
        self.b = b + buffer
        self.buffer = buffer
        self.BUFFER_FLAG = True

    def __repr__(self):
        return "%s(a = %s, b = %s, center = %s, theta = %s, buffer = %s, buffer_applied: %s )\n" % (
            type(self).__name__, self.a, self.b, self.center, self.theta, self.buffer, self.BUFFER_FLAG)

    def fields(self):
        return [self.center.x, self.center.y, self.a, self.b, self.theta, self.vel.x, self.vel.y, 0.0]

    def apply_buffer(self):
        if not self.BUFFER_FLAG:
            self.a = self.a + self.buffer
            self.b = self.b + self.buffer
            self.BUFFER_FLAG = True
            self._invalidate()
        else:
            warnings.warn("Warning: Buffer already applied. Call Ignored.")

    def remove_buffer(self):
        if self.BUFFER_FLAG:
            self.a = self.a - self.buffer
            self.b = self.b - self.buffer
            self.BUFFER_FLAG = False
            self._invalidate()
        else:
            warnings.warn("Warning: Buffer already removed. Call Ignored.")

    def update(self, s=None, s_obs=None, center=None, buffer=None, **kwargs):
        """cbf/obstacles.py:238-270 with the intended semantics (D2, D3): ``s`` is the EGO state and
        is stored apart from the obstacle's own orientation / velocity; ``b=`` sets ``b``."""
        if "a" in kwargs:
            self.a = kwargs["a"]
        if "b" in kwargs:
            self.b = kwargs["b"]
        if "theta" in kwargs:
            self.theta = kwargs["theta"]
        if center is not None:
            self.center = Point2(center.x, center.y)
        if s_obs is not None:
            self.center = Point2(s_obs[0], s_obs[1])
        if s is not None:
            self.s = s
        if buffer is not None:
            if self.BUFFER_FLAG:
                self.a = self.a - self.buffer + buffer
                self.b = self.b - self.buffer + buffer
            self.buffer = buffer
        self._invalidate()

    def update_coords(self, xy: Point2):
        self.center = xy
        self._invalidate()

    def update_velocity_by_magnitude(self, v):
        """Assumes theta is the heading (cbf/obstacles.py:272-277)."""
        if isinstance(v, torch.Tensor) or isinstance(self.theta, torch.Tensor):
            th = torch.as_tensor(self.theta)
            self.vel = Vector2(x=v * torch.cos(th), y=v * torch.sin(th))
        else:
            self.vel = Vector2(x=v * np.cos(self.theta), y=v * np.sin(self.theta))
        self._invalidate()

    def update_velocity(self, v: Vector2):
        self.vel = v.copy()
        self._invalidate()

    def update_orientation(self, yaw):
        self.theta = yaw
        self.update_velocity_by_magnitude(self.vel.magnitude())

    def update_by_bounding_box(self, bbox: BoundingBox):
        _check_bbox(bbox)
        self.update(a=bbox.extent.x, b=bbox.extent.y, center=Vector2(bbox.location.x, bbox.location.y), theta=bbox.rotation.yaw)

    @classmethod
    def from_bounding_box(cls, bbox: BoundingBox = None, buffer=0.5, **kwargs) -> "Ellipse2D":
        bbox = BoundingBox() if bbox is None else bbox
        _check_bbox(bbox)
        return cls(bbox.extent.x, bbox.extent.y, Vector2(bbox.location.x, bbox.location.y), bbox.rotation.yaw, buffer,
                   id=kwargs.get("id"))


class CollisionCone2D(Obstacle2DBase):
    """Collision-cone barrier h = p_rel . v_rel + |p_rel| |v_rel| cos(phi) (cbf/obstacles.py:333-543).
    ``s_obs`` = [cx, cy, theta_o, v_o]; the obstacle velocity direction is theta_o + beta as in
    ``update`` (obstacles.py:487-488; the constructor's beta-less variant, :367-368, is identical for
    the default beta = 0)."""
    slot_type = nv.SLOT_CONE

    def __init__(self, a=0.0, s=(0, 0, 0, 0), s_obs=(0, 0, 0, 0), buffer=1.50, **kwargs):
        super().__init__()
        self.type = Obstacle2DTypes.COLLISION_CONE2D
        if "id" in kwargs:
            self.id = kwargs["id"]
        self.beta = kwargs.get("beta", 0.0)
        self.s = s
        self.s_obs = s_obs
        self.a = a + buffer                                            # cbf/obstacles.py:357
        self.buffer = buffer
        self.BUFFER_FLAG = True

    def __repr__(self):
        return "%s(a = %s, buffer = %s, buffer_applied: %s )\n s_obs = %s" % (type(self).__name__, self.a, self.buffer,
                                                                             self.BUFFER_FLAG, self.s_obs)

    def fields(self):
        so = self.s_obs
        return [so[0], so[1], so[2], so[3], self.a, self.beta, 0.0, 0.0]

    def apply_buffer(self):
        if not self.BUFFER_FLAG:
            self.a = self.a + self.buffer
            self.BUFFER_FLAG = True
            self._invalidate()
        else:
            warnings.warn("Warning: Buffer already applied. Call Ignored.")

    def remove_buffer(self):
        if self.BUFFER_FLAG:
            self.a = self.a - self.buffer
            self.BUFFER_FLAG = False
            self._invalidate()
        else:
            warnings.warn("Warning: Buffer already removed. Call Ignored.")

    def dbeta(self, **kwargs):
        return self.dtheta(**kwargs)                                   # cbf/obstacles.py:465

    def update(self, s=None, s_obs=None, buffer=None, **kwargs):
        if "a" in kwargs:
            self.a = kwargs["a"]
        if s is not None:
            self.s = s
        if s_obs is not None:
            self.s_obs = s_obs
        if buffer is not None:
            if self.BUFFER_FLAG:
                self.a = self.a - self.buffer + buffer
            self.buffer = buffer
        if "beta" in kwargs:
            self.beta = kwargs["beta"]
        self._invalidate()

    def get_half_angle(self):
        """Apex half angle acos(cone_boundary / dist) (cbf/obstacles.py:507-510), from the evaluated cone."""
        state, scalar = as_state(self.s)
        so = [as_vec(v, state.shape[1], state.dtype, state.device) for v in self.s_obs[:2]]
        a = as_vec(self.a, state.shape[1], state.dtype, state.device)
        dist = torch.sqrt((state[0] - so[0]) ** 2 + (state[1] - so[1]) ** 2)
        cb = torch.where(dist.abs() > a.abs(), torch.sqrt((dist ** 2 - a ** 2).clamp_min(0)) + ZERO_TOL, torch.full_like(dist, ZERO_TOL))
        cos_phi = torch.where(dist > ZERO_TOL, cb / dist, torch.zeros_like(dist))
        return to_output(torch.acos(cos_phi), scalar)

    def update_by_bounding_box(self, bbox: BoundingBox):
        _check_bbox(bbox)
        self.a = np.hypot(bbox.extent.x, bbox.extent.y)
        self.update(s_obs=[bbox.location.x, bbox.location.y, 0.0, bbox.velocity])

    @classmethod
    def from_bounding_box(cls, s=(0.0, 0.0, 0.0, 0.0), bbox: BoundingBox = None, buffer=0.5, **kwargs) -> "CollisionCone2D":
        bbox = BoundingBox() if bbox is None else bbox
        _check_bbox(bbox)
        a = np.hypot(bbox.extent.x, bbox.extent.y)
        return cls(a=a, s=s, s_obs=[bbox.location.x, bbox.location.y, 0.0, bbox.velocity], buffer=buffer, id=kwargs.get("id"))


class PolyLane(Obstacle2DBase):
    """Polynomial lane boundary y = g(x) = sum c_i x^i; h = (cx - x)^2 + (g(cx) - y)^2 - buffer with cx
    the closest abscissa (cbf/obstacles.py:545-689).  Up to degree 5 (6 coefficients per slot)."""
    slot_type = nv.SLOT_LANE
    MAX_COEFFS = 6

    def __init__(self, coefficients, s=(0, 0, 0, 0), s_obs=(0, 0, 0, 0), buffer=1.50, **kwargs):
        super().__init__()
        if "id" in kwargs:
            self.id = kwargs["id"]
        self.beta = kwargs.get("beta", 0.0)
        self.type = Obstacle2DTypes.POLY_LANE
        # distance_form=True: h = sqrt(d^2) - buffer, the CBF_lane_sqrt / CBF_lane_cf_sqrt barrier of
        # test_scripts/stanley_controller_ellipse.py:465-512,546-579 (the class itself only has the squared form)
        if kwargs.get("distance_form", False):
            self.slot_type = nv.SLOT_LANE_SQRT
        self.update_coeffs(coefficients)
        self.s = s
        self.s_obs = s_obs
        self.buffer = buffer

    def update_coeffs(self, coefficients):
        """coefficients = [a0, a1, a2, ...] of f(x) = a0 + a1 x + a2 x^2 + ... (cbf/obstacles.py:578-592)."""
        c = coefficients if isinstance(coefficients, torch.Tensor) else np.asarray(coefficients, dtype=np.float64)
        n = c.shape[0]
        if n < 1 or n > self.MAX_COEFFS:
            raise ValueError("PolyLane supports 1..%d coefficients (degree <= 5), got %d" % (self.MAX_COEFFS, n))
        self.coeffs = c
        self.order = int(n - 1)
        self._invalidate()

    def evaluate_polynomial(self, x, **kwargs):
        c = self.coeffs
        g = 0.0 * x
        for i in range(c.shape[0] - 1, -1, -1):
            g = c[i] + g * x
        return g

    def fields(self):
        c = [self.coeffs[i] if i < self.coeffs.shape[0] else 0.0 for i in range(self.MAX_COEFFS)]
        c = [float(v) if not isinstance(v, torch.Tensor) else v for v in c]
        return [self.buffer] + c + [0.0]

    def update(self, s=None, s_obs=None, buffer=None, **kwargs):
        if s is not None:
            self.s = s
        if s_obs is not None:
            self.s_obs = s_obs
        if buffer is not None:
            self.buffer = buffer
        self._invalidate()

    def update_coeffs_by_curve_fit(self, x_pts, y_pts, n: int = 3, **kw):
        self.update_coeffs(self.fit_polynomial_curve(x_pts, y_pts, n=n, **kw).coeffs)

    @classmethod
    def fit_polynomial_curve(cls, x_pts, y_pts, n: int = 3, x_fixed_pts=None, y_fixed_pts=None, fixed_pts_idx=None,
                             alpha: float = 0.01, sigma=None, initial_coeffs=None) -> "PolyLane":
        """Weighted least-squares polynomial fit (cbf/obstacles.py:715-773: scipy ``curve_fit`` with
        ``sigma``; fixed points enter with the small sigma ``alpha``).  Host-side, once per lane --
        ``curve_fit`` on a model that is linear in its parameters converges to the weighted normal
        equations, solved here directly."""
        x_pts = np.asarray(x_pts, dtype=np.float64).flatten()
        y_pts = np.asarray(y_pts, dtype=np.float64).flatten()
        if x_pts.size != y_pts.size:
            raise ValueError("Incompatible array sizes for x points and y points. Received: %s and %s" % (x_pts.shape, y_pts.shape))
        sigma = np.full_like(x_pts, 10.0) if sigma is None else np.asarray(sigma, dtype=np.float64).copy()
        if x_fixed_pts is None and y_fixed_pts is not None:
            raise ValueError("Both fixed point arrays have to be specified. Received empty x fixed points.")
        if x_fixed_pts is not None and y_fixed_pts is None:
            raise ValueError("Both fixed point arrays have to be specified. Received empty y fixed points.")
        if x_fixed_pts is not None:
            x_fixed_pts = np.asarray(x_fixed_pts, dtype=np.float64).flatten()
            y_fixed_pts = np.asarray(y_fixed_pts, dtype=np.float64).flatten()
            x_pts = np.append(x_pts, x_fixed_pts)
            y_pts = np.append(y_pts, y_fixed_pts)
            sigma = np.append(sigma, alpha * np.ones_like(x_fixed_pts))
        if fixed_pts_idx is not None:
            sigma[fixed_pts_idx] = alpha
        V = np.vander(x_pts, n + 1, increasing=True) / sigma[:, None]
        coeffs, *_ = np.linalg.lstsq(V, y_pts / sigma, rcond=None)
        return cls(coeffs)

    @classmethod
    def fit_polynomial_curves(cls, x_pts: torch.Tensor, y_pts: torch.Tensor, n: int = 3, sigma: Optional[torch.Tensor] = None,
                              count: Optional[torch.Tensor] = None, **kwargs) -> "PolyLane":
        """Batched ``fit_polynomial_curve`` on the GPU (kernel KL, ``sccav_fit_lanes_*``): ``x_pts``, ``y_pts`` [K, C] CUDA
        tensors hold the points of C lanes (fixed points appended with their small sigma, as cbf/obstacles.py:749-756
        does), ``sigma`` [K, C] (default 10), ``count`` [C] points per lane.  Returns ONE PolyLane whose
        coefficients are [n + 1, C] tensors -- lane c belongs to vehicle c of a batch."""
        from . import ops
        coeffs, status = ops.fit_lanes(x_pts, y_pts, n=n, sigma=sigma, count=count)
        lane = cls(coeffs[: n + 1], **kwargs)
        lane.fit_status = status
        return lane


class ObstacleList2D(MutableMapping):
    """Insertion-ordered mapping id -> obstacle (cbf/obstacles.py:798-941); the constraint index of
    the QP is the dict order.  ``pack()`` produces the SoA buffer the kernels read."""

    def __init__(self, data=()):
        self.mapping: Dict[object, Obstacle2DBase] = {}
        self.update(data)
        self.timestamp = 0.0

    def __getitem__(self, key):
        return self.mapping[key]

    def __delitem__(self, key):
        del self.mapping[key]

    def __setitem__(self, key, value):
        if Obstacle2DBase not in value.__class__.__mro__:
            raise TypeError("Expected an object derived from Obstacle2DBase as value. Received " + type(value).__name__)
        self.mapping[key] = value

    def __iter__(self):
        return iter(self.mapping)

    def __len__(self):
        return len(self.mapping)

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, self.mapping)

    def set_timestamp(self, timestamp: float):
        self.timestamp = timestamp

    def update_by_bounding_box(self, bbox_dict=None, obs_type=Obstacle2DTypes.ELLIPSE2D, buffer=0.5):
        """Add the ids that entered the scene, update those that moved, remove those that left
        (cbf/obstacles.py:833-858)."""
        if bbox_dict is None:
            return
        for key, bbox in bbox_dict.items():
            if key in self.mapping:
                self.mapping[key].update_by_bounding_box(bbox)
            elif obs_type == Obstacle2DTypes.ELLIPSE2D:
                self[key] = Ellipse2D.from_bounding_box(bbox=bbox, buffer=buffer, id=key)
            elif obs_type == Obstacle2DTypes.COLLISION_CONE2D:
                self[key] = CollisionCone2D.from_bounding_box(bbox=bbox, buffer=buffer, id=key)
        for key in list(self.mapping.keys()):
            if key not in bbox_dict:
                self.pop(key)

    def update_state(self, s, s_obs_dict: dict = None, buffer: float = None, **kwargs):
        """Push the ego state into every obstacle; with ``s_obs_dict`` also the per-key obstacle
        states (intended semantics of cbf/obstacles.py:860-877, D5)."""
        if s_obs_dict is None:
            for obstacle in self.mapping.values():
                obstacle.update(s=s, s_obs=None, buffer=buffer, **kwargs)
            return
        if not isinstance(s_obs_dict, dict):
            raise ValueError("Expected dictionary for obstacle dictionary")
        for obstacle in self.mapping.values():
            obstacle.update(s=s, buffer=buffer, **kwargs)
        for key, s_obs in s_obs_dict.items():
            if key in self.mapping:
                self.mapping[key].update(s_obs=s_obs)
            else:
                warnings.warn("Unknown key provided in s_obs_dict. Corresponding key not found in the obstacle list. key: %r" % (key,))

    # ---- SoA packing + stacked getters ----------------------------------------------------------
    # ``solve_cbf`` on a list whose ELLIPSE obstacles have not changed since the previous call evaluates them in their
    # ingested form (ELLIPSE_PREP, include/sccav_cbf.h: everything of Ellipse2D that does not depend on the vehicle is
    # computed once, sccav_prepare_obstacles_*) -- the same functions as cbf/obstacles.py:193,218,229,316, a few ulp
    # apart, and the HBM-bound form of the operator.  Set to False for the reference's operation order on every call.
    use_prepared = True

    def _gather(self):
        """(descs, rows of 8 field values, all_scalar, cacheable, fingerprint) of the current list."""
        descs, rows, fp = [], [], []
        scalar, cacheable = True, True
        for obs in self.mapping.values():
            f = obs.fields()
            descs.append(obs.slot_type)
            rows.append(f)
            fp.append(obs.slot_type)
            for v in f:
                if isinstance(v, torch.Tensor):
                    if v.dim() > 0:
                        scalar = False
                    fp.append((id(v), v._version))
                elif isinstance(v, np.ndarray):
                    if not (v.ndim == 0 or v.size == 1):
                        scalar = False
                    cacheable = False                         # arrays change in place without a trace
                    fp.append(None)
                else:
                    fp.append(float(v))
        return descs, rows, scalar, cacheable, tuple(fp)

    def pack_host(self) -> Tuple[List[int], np.ndarray]:
        """(slot_desc, fields [M, 8] float64 numpy) of a list whose fields are all scalars -- the one-scenario call of
        the reference: no device work at all, the host entry point of the library takes it from here."""
        descs, rows, scalar, _, _ = self._gather()
        if not scalar:
            raise ValueError("pack_host needs scalar obstacle fields")
        arr = np.array([[float(v) for v in f] for f in rows], dtype=np.float64).reshape(len(rows), nv.NFIELD)
        return descs, arr

    def all_scalar(self) -> bool:
        for obs in self.mapping.values():
            for v in obs.fields():
                if isinstance(v, torch.Tensor) and v.dim() > 0:
                    return False
                if isinstance(v, np.ndarray) and not (v.ndim == 0 or v.size == 1):
                    return False
        return True

    def pack(self, state: torch.Tensor, prepared: bool = False) -> Tuple[List[int], torch.Tensor, bool]:
        """(slot_desc, obst [M, 8, N], all_scalar) for the ego batch ``state`` [4, N].

        The packed buffer is cached: as long as no field of any obstacle has changed (scalars by value, tensors by
        identity and in-place version) the same device tensor is returned and nothing is launched.  With
        ``prepared`` the ELLIPSE slots come back in their ingested form (see ``use_prepared``); ellipses whose
        velocity is exactly (0, 0) are declared STATIC."""
        descs, rows, scalar, cacheable, fp = self._gather()
        N = state.shape[1]
        for f in rows:
            for v in f:
                N = max(N, batch_size(v))
        key = (fp, N, state.dtype, state.device)
        c = getattr(self, "_pack_cache", None)
        if c is None or not cacheable or c["key"] != key:
            M = len(rows)
            if M == 0:
                obst = torch.empty((0, nv.NFIELD, N), dtype=state.dtype, device=state.device)
            else:
                # scalars go through ONE host array and one copy; tensor-valued fields are written over it
                host = np.zeros((M, nv.NFIELD), dtype=np.float64)
                tens = []
                for m, f in enumerate(rows):
                    for k, v in enumerate(f):
                        if isinstance(v, (torch.Tensor, np.ndarray)) and batch_size(v) > 1:
                            tens.append((m, k, v))
                        elif isinstance(v, torch.Tensor):
                            host[m, k] = float(v)
                        else:
                            host[m, k] = float(np.asarray(v).reshape(-1)[0])
                base = torch.from_numpy(host).to(device=state.device, dtype=state.dtype)
                obst = base.unsqueeze(2).expand(M, nv.NFIELD, N).contiguous()
                for m, k, v in tens:
                    obst[m, k].copy_(as_vec(v, N, state.dtype, state.device))
            c = {"key": key, "descs": descs, "obst": obst, "scalar": scalar, "prep": None,
                 "keep": [v for f in rows for v in f if isinstance(v, torch.Tensor)]}      # ids stay unique while cached
            self._pack_cache = c
        if prepared and c["obst"].shape[0] > 0:
            if c["prep"] is None:
                d2 = list(c["descs"])
                for m, f in enumerate(rows):
                    vx, vy = f[5], f[6]
                    if (d2[m] & nv.SLOT_TYPE_MASK) == nv.SLOT_ELLIPSE and not isinstance(vx, (torch.Tensor, np.ndarray)) \
                            and not isinstance(vy, (torch.Tensor, np.ndarray)) and float(vx) == 0.0 and float(vy) == 0.0:
                        d2[m] |= nv.SLOT_STATIC
                c["prep"] = ops.prepare_obstacles(d2, c["obst"])
            return c["prep"][0], c["prep"][1], c["scalar"]
        return c["descs"], c["obst"], c["scalar"]

    def _stacked(self, k: int):
        vals = [obs._get(k) for obs in self.mapping.values()]
        if all(not isinstance(v, torch.Tensor) for v in vals):
            return torch.tensor(vals, dtype=torch.float64).reshape(-1, 1)      # m x 1, like the cvxopt matrix
        return torch.stack([torch.as_tensor(v) for v in vals])

    def f(self, *args, **kwargs):
        return self._stacked(H)

    def dx(self, *args, **kwargs):
        return self._stacked(HX)

    def dy(self, *args, **kwargs):
        return self._stacked(HY)

    def dtheta(self, *args, **kwargs):
        return self._stacked(HTH)

    def dv(self, *args, **kwargs):
        return self._stacked(HV)

    def dt(self, *args, **kwargs):
        return self._stacked(HT)

    def dbeta(self, *args, **kwargs):
        vals = [obs.dbeta(**kwargs) for obs in self.mapping.values()]
        if all(not isinstance(v, torch.Tensor) for v in vals):
            return torch.tensor([float(v) for v in vals], dtype=torch.float64).reshape(-1, 1)
        ref = next(v for v in vals if isinstance(v, torch.Tensor))
        return torch.stack([v if isinstance(v, torch.Tensor) else torch.full_like(ref, float(v)) for v in vals])

    def gradient(self, *args, **kwargs):
        """m x 4 (D6: the reference allocates m x 3 for 4-vectors)."""
        g = [obs.gradient(**kwargs) for obs in self.mapping.values()]
        if all(not isinstance(v, torch.Tensor) for v in g):
            return torch.tensor(g, dtype=torch.float64).reshape(-1, 4)
        return torch.stack([torch.as_tensor(v) for v in g])


class BatchedObstacleList2D:
    """N per-vehicle ``ObstacleList2D``'s whose obstacles come and go (the CARLA deployment: every ego sees
    its own set of actors, cbf/obstacles.py:833-858, multi_obstacle_CBF_local_with_lanes.py:906-932), kept in
    the structure-of-arrays layout of the kernels: vehicle n holds ``count[n]`` obstacles of ONE type in its
    first slots -- ``ids`` [M, N] int32 (the dict keys, -1 = empty), ``obst`` [M, 8, N].

    ``update_by_bounding_box(box_id, box)`` is one launch of the ingest kernel (KB): ids already held are
    updated in place, ids that left the scene are removed, new ids are appended, exactly in the order the
    reference's dict would hold them.  Assign an instance to ``DBM_CBF_2DS.obstacle_list2d`` and call
    ``solve_cbf`` as usual; a vehicle with no obstacle gets u = u_ref."""

    def __init__(self, n_vehicles: int, capacity: int = 8, obs_type: Obstacle2DTypes = Obstacle2DTypes.ELLIPSE2D,
                 dtype: torch.dtype = torch.float64, device=None):
        if obs_type not in (Obstacle2DTypes.ELLIPSE2D, Obstacle2DTypes.COLLISION_CONE2D):
            raise TypeError("update_by_bounding_box makes Ellipse2D or CollisionCone2D obstacles")      # obstacles.py:843-846
        if not 1 <= capacity <= nv.MAX_ROWS:
            raise ValueError("capacity must be in [1, %d]" % nv.MAX_ROWS)
        self.device = cuda_device() if device is None else torch.device(device)
        self.obs_type = obs_type
        self.slot_type = nv.SLOT_ELLIPSE if obs_type == Obstacle2DTypes.ELLIPSE2D else nv.SLOT_CONE
        self.N, self.M = int(n_vehicles), int(capacity)
        self.ids = torch.full((self.M, self.N), -1, dtype=torch.int32, device=self.device)
        self.obst = torch.zeros((self.M, nv.NFIELD, self.N), dtype=dtype, device=self.device)
        self.count = torch.zeros((self.N,), dtype=torch.int32, device=self.device)
        self.dropped = torch.zeros((self.N,), dtype=torch.int32, device=self.device)

    def __len__(self):
        return self.M

    @property
    def slot_desc(self) -> List[int]:
        return [self.slot_type] * self.M

    def update_by_bounding_box(self, box_id: torch.Tensor, box: torch.Tensor, buffer: float = 0.5, rebuild: bool = False):
        """``box_id`` [K, N] int32 (< 0 = no box), ``box`` [K, 6, N] = extent.x, extent.y, location.x, location.y,
        rotation.yaw, velocity of ``BoundingBox`` (cbf/obstacles.py:59-64).  ``rebuild``: make every obstacle
        afresh as the CARLA driver does each tick (multi_obstacle_CBF_local_with_lanes.py:918-928).
        Returns ``dropped`` [N]: new ids that did not fit the capacity."""
        ops.ingest_boxes(self.slot_type, box_id.to(self.device), box.to(device=self.device, dtype=self.obst.dtype), self.ids,
                         self.obst, self.count, buffer=buffer, mode=nv.INGEST_REBUILD if rebuild else nv.INGEST_UPDATE,
                         dropped=self.dropped)
        return self.dropped

    def update_state(self, s=None, s_obs_dict=None, buffer=None, **kwargs):
        """The ego state reaches the kernels with ``solve_cbf``; nothing is stored per obstacle."""
        if s_obs_dict is not None or buffer is not None:
            raise ValueError("BatchedObstacleList2D takes obstacle updates through update_by_bounding_box")

    def pack(self, state: torch.Tensor):
        if state.shape[1] != self.N:
            raise ValueError("the ego batch has %d vehicles, the obstacle lists %d" % (state.shape[1], self.N))
        return self.slot_desc, self.obst if self.obst.dtype == state.dtype else self.obst.to(state.dtype), False
