"""Lie-group state spaces: the slice of ``dair_pll/state_space.py`` on the hot path
(``StateSpace.q/v/x/q_v/euler_step`` :171-311, ``FixedBaseSpace`` :560-640,
``FloatingBaseSpace`` :400-500, ``ProductSpace`` :650-741).  Samplers, noisers and
error metrics are data-generation/evaluation code and stay with the reference.

Same coordinates as the reference: q = [quat (w first), base position, joints],
v = [body-frame angular velocity, world-frame linear velocity, joint rates].
"""
from typing import List, Tuple

import torch
from torch import Tensor

from dair_pll_b200 import quaternion

N_QUAT, N_POS, N_ANG = 4, 3, 3


class StateSpace:
    """Configuration manifold G (dimension n_q coordinates, n_v tangent) x velocities."""

    def __init__(self, n_q: int, n_v: int) -> None:
        assert n_q >= 0 and n_v >= 0
        self.n_q, self.n_v, self.n_x = n_q, n_v, n_q + n_v

    # -- slicing ---------------------------------------------------------
    def q(self, x: Tensor) -> Tensor:
        assert x.shape[-1] == self.n_x
        return x[..., :self.n_q]

    def v(self, x: Tensor) -> Tensor:
        assert x.shape[-1] == self.n_x
        return x[..., self.n_q:]

    def q_v(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        return self.q(x), self.v(x)

    def x(self, q: Tensor, v: Tensor) -> Tensor:
        assert q.shape[-1] == self.n_q and v.shape[-1] == self.n_v
        return torch.cat((q, v), -1)

    # -- group structure -------------------------------------------------
    def exponential(self, q: Tensor, dq: Tensor) -> Tensor:
        raise NotImplementedError

    def configuration_difference(self, q_1: Tensor, q_2: Tensor) -> Tensor:
        raise NotImplementedError

    def project_configuration(self, q: Tensor) -> Tensor:
        return q

    def zero_state(self) -> Tensor:
        raise NotImplementedError

    def euler_step(self, q: Tensor, v: Tensor, dt: float) -> Tensor:
        """q (+) v dt -- geodesic forward Euler (state_space.py:295-311)."""
        assert q.shape[-1] == self.n_q and v.shape[-1] == self.n_v
        return self.exponential(q, v * dt)

    def finite_difference(self, q: Tensor, q_plus: Tensor, dt: float) -> Tensor:
        return self.configuration_difference(q, q_plus) / dt

    def project_state(self, x: Tensor) -> Tensor:
        return self.x(self.project_configuration(self.q(x)), self.v(x))


class FixedBaseSpace(StateSpace):
    """R^n joints on a fixed base; the world model instance is ``FixedBaseSpace(0)``."""

    def __init__(self, n_joints: int) -> None:
        super().__init__(n_joints, n_joints)

    def exponential(self, q: Tensor, dq: Tensor) -> Tensor:
        return q + dq

    def configuration_difference(self, q_1: Tensor, q_2: Tensor) -> Tensor:
        return q_2 - q_1

    def zero_state(self) -> Tensor:
        return torch.zeros(self.n_x)


class FloatingBaseSpace(StateSpace):
    """SE(3) x R^n_joints with quaternion orientation (7 + n, 6 + n)."""

    def __init__(self, n_joints: int) -> None:
        assert n_joints >= 0
        super().__init__(N_QUAT + N_POS + n_joints, N_ANG + N_POS + n_joints)

    def quat(self, q_or_x: Tensor) -> Tensor:
        return q_or_x[..., :N_QUAT]

    def base(self, q_or_x: Tensor) -> Tensor:
        return q_or_x[..., N_QUAT:N_QUAT + N_POS]

    def exponential(self, q: Tensor, dq: Tensor) -> Tensor:
        # body-frame rotation vector => right multiplication (state_space.py:466-486)
        quat = quaternion.multiply(q[..., :N_QUAT], quaternion.exp(dq[..., :N_ANG]))
        return torch.cat((quat, q[..., N_QUAT:] + dq[..., N_ANG:]), -1)

    def configuration_difference(self, q_1: Tensor, q_2: Tensor) -> Tensor:
        rel = quaternion.multiply(quaternion.inverse(q_1[..., :N_QUAT]), q_2[..., :N_QUAT])
        return torch.cat((quaternion.log(rel), q_2[..., N_QUAT:] - q_1[..., N_QUAT:]), -1)

    def zero_state(self) -> Tensor:
        # the reference returns all zeros here (state_space.py:495-499); kept for parity
        return torch.zeros(self.n_x)


class ProductSpace(StateSpace):
    """Cartesian product; coordinates of the factors are concatenated (q's, then v's)."""

    def __init__(self, spaces: List[StateSpace]) -> None:
        super().__init__(sum(s.n_q for s in spaces), sum(s.n_v for s in spaces))
        self.spaces = spaces
        self._q_sizes = [s.n_q for s in spaces]
        self._v_sizes = [s.n_v for s in spaces]

    def q_split(self, q: Tensor) -> List[Tensor]:
        return list(torch.split(q, self._q_sizes, -1))

    def v_split(self, v: Tensor) -> List[Tensor]:
        return list(torch.split(v, self._v_sizes, -1))

    def exponential(self, q: Tensor, dq: Tensor) -> Tensor:
        parts = [s.exponential(qi, dqi) for s, qi, dqi in zip(self.spaces, self.q_split(q), self.v_split(dq))]
        return torch.cat(parts, -1)

    def configuration_difference(self, q_1: Tensor, q_2: Tensor) -> Tensor:
        parts = [s.configuration_difference(a, b)
                 for s, a, b in zip(self.spaces, self.q_split(q_1), self.q_split(q_2))]
        return torch.cat(parts, -1)

    def project_configuration(self, q: Tensor) -> Tensor:
        return torch.cat([s.project_configuration(qi) for s, qi in zip(self.spaces, self.q_split(q))], -1)

    def zero_state(self) -> Tensor:
        zeros = [s.zero_state() for s in self.spaces]
        return torch.cat([s.q(z) for s, z in zip(self.spaces, zeros)] +
                         [s.v(z) for s, z in zip(self.spaces, zeros)], -1)
