"""Seeded synthetic state batches for tests and ``bench.py``.

Pure torch on a CPU generator (reproducible everywhere), then moved to the requested device.
"""
import math
from typing import Tuple

import torch
from torch import Tensor

CUBE_HALF = 0.0524
CUBE_NOMINAL = dict(m=0.37, inertia=0.00081, half=CUBE_HALF, mu_box=0.15, mu_ground=1.0)


def _rot_row2(quat: Tensor) -> Tensor:
    """Third row of the rotation matrix of unit quaternions (*,4) -> (*,3)."""
    w, x, y, z = quat.unbind(-1)
    return torch.stack((2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)), -1)


def cube_states(batch: int, seed: int = 0, half=(CUBE_HALF,) * 3, dtype=torch.float64,
                device='cpu') -> Tensor:
    """(batch, 13) cube states covering all contact regimes: random orientation; height =
    support height of the box for that orientation + delta, delta ~ 50% U(-5 mm, 10 mm)
    (contact-rich) / 50% U(10 mm, 200 mm) (flight); w_body ~ N(0, 5^2), v ~ N(0, 1)."""
    g = torch.Generator().manual_seed(seed)
    quat = torch.randn(batch, 4, generator=g, dtype=torch.float64)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    xy = torch.rand(batch, 2, generator=g, dtype=torch.float64) - 0.5
    support = (_rot_row2(quat).abs() * torch.tensor(half, dtype=torch.float64)).sum(-1)
    near = torch.rand(batch, generator=g, dtype=torch.float64) < 0.5
    u = torch.rand(batch, generator=g, dtype=torch.float64)
    delta = torch.where(near, -0.005 + 0.015 * u, 0.01 + 0.19 * u)
    z = support + delta
    omega = 5.0 * torch.randn(batch, 3, generator=g, dtype=torch.float64)
    vel = torch.randn(batch, 3, generator=g, dtype=torch.float64)
    x = torch.cat((quat, xy, z[:, None], omega, vel), -1)
    return x.to(dtype=dtype, device=device)


def perturb_next_state(x_next: Tensor, seed: int = 1, sigma_q: float = 1e-3, sigma_v: float = 1e-2,
                       n_q: int = 7) -> Tensor:
    """Measurement-like noise on a simulated next state; the quaternion is re-normalised."""
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(x_next.shape, generator=g, dtype=torch.float64)
    scale = torch.cat((torch.full((n_q,), sigma_q), torch.full((x_next.shape[-1] - n_q,), sigma_v))).double()
    out = x_next.double().cpu() + noise * scale
    out[..., :4] = out[..., :4] / out[..., :4].norm(dim=-1, keepdim=True)
    return out.to(dtype=x_next.dtype, device=x_next.device)


def cube_pi_cm_perturbed(seed: int = 0) -> Tensor:
    """(1,10) pi_cm: URDF nominal values perturbed by 10% with a non-zero CoM offset (mm scale)."""
    g = torch.Generator().manual_seed(1000 + seed)
    r = lambda n: 1 + 0.1 * (2 * torch.rand(n, generator=g, dtype=torch.float64) - 1)  # noqa: E731
    m = CUBE_NOMINAL['m'] * r(1)
    c = 0.005 * (2 * torch.rand(3, generator=g, dtype=torch.float64) - 1)
    diag = CUBE_NOMINAL['inertia'] * r(3)
    off = 2e-5 * (2 * torch.rand(3, generator=g, dtype=torch.float64) - 1)
    return torch.cat((m, m * c, diag, off)).reshape(1, 10)


def cube_learnables_perturbed(seed: int = 0) -> Tuple[Tensor, Tensor, Tensor]:
    """pi_cm (1,10), friction_params (2,) [box, ground], half lengths (3,)."""
    g = torch.Generator().manual_seed(2000 + seed)
    r = lambda n: 1 + 0.1 * (2 * torch.rand(n, generator=g, dtype=torch.float64) - 1)  # noqa: E731
    friction = torch.tensor([CUBE_NOMINAL['mu_box'], CUBE_NOMINAL['mu_ground']], dtype=torch.float64) * r(2)
    half = CUBE_HALF * r(3)
    return cube_pi_cm_perturbed(seed), friction, half


# ---------------------------------------------------------------------------
# elbow (two 0.1 x 0.05 x 0.05 boxes, y-axis hinge at (-0.035, 0.06, 0))
# ---------------------------------------------------------------------------
ELBOW_HALF = (0.05, 0.025, 0.025)
ELBOW_JOINT_ORIGIN = (-0.035, 0.06, 0.0)
ELBOW_BOX2_OFFSET = (0.035, 0.0, 0.0)
ELBOW_NOMINAL = dict(m=0.37, inertia=0.0006167, mu_box=0.3, mu_ground=1.0)


def _quat_to_rot(quat: Tensor) -> Tensor:
    w, x, y, z = quat.unbind(-1)
    return torch.stack((
        torch.stack((1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)), -1),
        torch.stack((2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)), -1),
        torch.stack((2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)), -1)), -2)


def elbow_states(batch: int, seed: int = 0, dtype=torch.float64, device='cpu') -> Tensor:
    """(batch, 15) elbow states [quat | pos | hinge angle | w_body | v_world | hinge rate]: random
    orientation and hinge angle; height = (lowest corner of either box at z = 0) + delta with the same
    contact-rich / flight mixture as :func:`cube_states`; hinge rate ~ N(0, 6^2)."""
    g = torch.Generator().manual_seed(seed)
    quat = torch.randn(batch, 4, generator=g, dtype=torch.float64)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    xy = torch.rand(batch, 2, generator=g, dtype=torch.float64) - 0.5
    theta = (2 * torch.rand(batch, generator=g, dtype=torch.float64) - 1) * math.pi
    R1 = _quat_to_rot(quat)
    c, s_ = torch.cos(theta), torch.sin(theta)
    zero, one = torch.zeros_like(c), torch.ones_like(c)
    Ry = torch.stack((torch.stack((c, zero, s_), -1), torch.stack((zero, one, zero), -1),
                      torch.stack((-s_, zero, c), -1)), -2)
    R2 = R1 @ Ry
    half = torch.tensor(ELBOW_HALF, dtype=torch.float64)
    pj = torch.tensor(ELBOW_JOINT_ORIGIN, dtype=torch.float64)
    off = torch.tensor(ELBOW_BOX2_OFFSET, dtype=torch.float64)
    low1 = -(R1[:, 2, :].abs() * half).sum(-1)
    low2 = (R1 @ pj)[:, 2] + (R2 @ off)[:, 2] - (R2[:, 2, :].abs() * half).sum(-1)
    lowest = torch.minimum(low1, low2)
    near = torch.rand(batch, generator=g, dtype=torch.float64) < 0.5
    u = torch.rand(batch, generator=g, dtype=torch.float64)
    delta = torch.where(near, -0.005 + 0.015 * u, 0.01 + 0.19 * u)
    z = -lowest + delta
    omega = 5.0 * torch.randn(batch, 3, generator=g, dtype=torch.float64)
    vel = torch.randn(batch, 3, generator=g, dtype=torch.float64)
    rate = 6.0 * torch.randn(batch, 1, generator=g, dtype=torch.float64)
    x = torch.cat((quat, xy, z[:, None], theta[:, None], omega, vel, rate), -1)
    return x.to(dtype=dtype, device=device)


def elbow_learnables_perturbed(seed: int = 0) -> Tuple[Tensor, Tensor, Tensor]:
    """pi_cm (2,10), friction_params (3,) [box1, box2, ground], half lengths (2,3): nominal URDF values
    perturbed by 10%, plus mm-scale centre-of-mass offsets."""
    g = torch.Generator().manual_seed(3000 + seed)
    r = lambda n: 1 + 0.1 * (2 * torch.rand(n, generator=g, dtype=torch.float64) - 1)  # noqa: E731
    rows = []
    for com0 in ((0., 0., 0.), ELBOW_BOX2_OFFSET):
        m = ELBOW_NOMINAL['m'] * r(1)
        c = torch.tensor(com0, dtype=torch.float64) + 0.004 * (2 * torch.rand(3, generator=g, dtype=torch.float64) - 1)
        diag = ELBOW_NOMINAL['inertia'] * r(3)
        offd = 2e-5 * (2 * torch.rand(3, generator=g, dtype=torch.float64) - 1)
        rows.append(torch.cat((m, m * c, diag, offd)))
    friction = torch.tensor([ELBOW_NOMINAL['mu_box'], ELBOW_NOMINAL['mu_box'], ELBOW_NOMINAL['mu_ground']],
                            dtype=torch.float64) * r(3)
    half = torch.tensor([ELBOW_HALF, ELBOW_HALF], dtype=torch.float64) * r(6).reshape(2, 3)
    return torch.stack(rows), friction, half
