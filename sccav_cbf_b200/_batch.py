"""Scalar <-> batch plumbing of the class API: Python numbers mean ONE scenario (results come
back as Python floats, like the reference); torch tensors / numpy arrays with a trailing axis N
mean N independent vehicles (results stay on the GPU).  No arithmetic happens here."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import _native as nv

DEFAULT_DTYPE = torch.float64


def cuda_device() -> torch.device:
    nv.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def is_scalar(v) -> bool:
    if isinstance(v, torch.Tensor):
        return v.dim() == 0
    if isinstance(v, np.ndarray):
        return v.ndim == 0 or v.size == 1
    return True


def batch_size(v) -> int:
    if isinstance(v, torch.Tensor):
        return 1 if v.dim() == 0 else int(v.shape[-1])
    if isinstance(v, np.ndarray):
        return 1 if v.ndim == 0 else int(v.shape[-1])
    return 1


def as_vec(v, N: int, dtype, device) -> torch.Tensor:
    """A contiguous [N] tensor on ``device``: scalars are broadcast, [N] tensors moved / cast."""
    if isinstance(v, torch.Tensor):
        t = v.to(device=device, dtype=dtype)
    elif isinstance(v, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(device=device, dtype=dtype)
    else:
        return torch.full((N,), float(v), dtype=dtype, device=device)
    t = t.reshape(-1)
    if t.numel() == 1:
        return t.expand(N).contiguous()
    if t.numel() != N:
        raise ValueError("expected a scalar or %d values, got %d" % (N, t.numel()))
    return t.contiguous()


def as_state(s, dtype=None, device=None) -> Tuple[torch.Tensor, bool]:
    """Ego state -> ([4, N] tensor on the GPU, scalar_mode).  Accepts a sequence of 4 numbers
    (also a 4x1 array, the cvxopt-matrix shape of the reference), a sequence of 4 [N] tensors, or
    a [4, N] tensor / array."""
    device = cuda_device() if device is None else device
    if isinstance(s, torch.Tensor):
        dt = dtype or (s.dtype if s.dtype in (torch.float32, torch.float64) else DEFAULT_DTYPE)
        t = s.to(device=device, dtype=dt)
        if t.dim() == 1:
            t = t.reshape(4, 1)
        if t.shape[0] != 4:
            raise ValueError("state must have 4 rows (x, y, theta, v), got shape %s" % (tuple(s.shape),))
        return t.contiguous(), False
    if isinstance(s, np.ndarray) and s.ndim == 2 and s.shape[1] > 1:
        return torch.from_numpy(np.ascontiguousarray(s, dtype=np.float64)).to(device=device, dtype=dtype or DEFAULT_DTYPE), False
    comps = [s[i] for i in range(4)]
    if any(isinstance(c, torch.Tensor) and c.dim() > 0 for c in comps):
        N = max(batch_size(c) for c in comps)
        ref = next(c for c in comps if isinstance(c, torch.Tensor) and c.dim() > 0)
        dt = dtype or (ref.dtype if ref.dtype in (torch.float32, torch.float64) else DEFAULT_DTYPE)
        return torch.stack([as_vec(c, N, dt, device) for c in comps]).contiguous(), False
    vals = [float(np.asarray(c).reshape(-1)[0]) if not isinstance(c, torch.Tensor) else float(c) for c in comps]
    return torch.tensor(vals, dtype=dtype or DEFAULT_DTYPE, device=device).reshape(4, 1), True


def state_scalar(s):
    """The ego state as a float64 numpy [4] when it is given as 4 plain numbers (one scenario), else None."""
    if isinstance(s, torch.Tensor):
        return None
    if isinstance(s, np.ndarray) and s.ndim == 2 and s.shape[1] > 1:
        return None
    try:
        comps = [s[i] for i in range(4)]
    except Exception:
        return None
    if any(isinstance(c, torch.Tensor) for c in comps):
        return None
    try:
        return np.array([float(np.asarray(c).reshape(-1)[0]) for c in comps], dtype=np.float64)
    except Exception:
        return None


def to_output(t: torch.Tensor, scalar: bool):
    """[N] tensor -> Python float in scalar mode (one D2H read), the tensor itself otherwise."""
    if scalar:
        return float(t.reshape(-1)[0].item())
    return t
