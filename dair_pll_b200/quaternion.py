"""Unit-quaternion helpers on torch tensors (w-first), device/dtype following the inputs.

Host-side mirror of the part of ``dair_pll/quaternion.py`` the hot path touches
(``multiply`` :89-104, ``inverse``, ``rotate`` :150-164, ``sinc`` :208-229, ``log`` :232-273,
``exp`` :276-309).  The sim-step kernels implement ``exp``/``multiply`` on-chip; these
versions serve state-space utilities (differences, errors) around the kernels.
"""
import torch
from torch import Tensor


def multiply(q: Tensor, r: Tensor) -> Tensor:
    """Hamilton product q (x) r."""
    qw, qv = q[..., :1], q[..., 1:]
    rw, rv = r[..., :1], r[..., 1:]
    w = qw * rw - (qv * rv).sum(-1, keepdim=True)
    v = qw * rv + rw * qv + torch.linalg.cross(qv, rv, dim=-1)
    return torch.cat((w, v), -1)


def inverse(q: Tensor) -> Tensor:
    """Conjugate (= inverse for unit quaternions)."""
    return torch.cat((q[..., :1], -q[..., 1:]), -1)


def rotate(q: Tensor, p: Tensor) -> Tensor:
    """Rotate vectors p (*,3) by unit quaternions q (*,4)."""
    w, v = q[..., :1], q[..., 1:]
    t = 2 * torch.linalg.cross(v, p, dim=-1)
    return p + w * t + torch.linalg.cross(v, t, dim=-1)


def sinc(x: Tensor) -> Tensor:
    """sin(x)/x with the removable singularity filled in (sinc(0) = 1)."""
    nz = x != 0
    safe = torch.where(nz, x, torch.ones_like(x))
    return torch.where(nz, torch.sin(safe) / safe, torch.ones_like(x))


def exp(r: Tensor) -> Tensor:
    """Rotation vector (*,3) -> quaternion [cos(|r|/2), r sinc(|r|/2)/2]."""
    half = r.norm(dim=-1, keepdim=True) / 2
    return torch.cat((torch.cos(half), r * sinc(half) / 2), -1)


def log(q: Tensor) -> Tensor:
    """Quaternion -> rotation vector; inverse of :func:`exp`."""
    v = q[..., 1:]
    s = v.norm(dim=-1, keepdim=True)
    theta = 2 * torch.atan2(s, q[..., :1])
    nz = s > 0
    return v * torch.where(nz, theta / torch.where(nz, s, torch.ones_like(s)), torch.zeros_like(s))
