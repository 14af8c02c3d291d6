"""Time integration wrappers on the hot path: ``Integrator.simulate``
(dair_pll/integrator.py:75-99) and ``VelocityIntegrator.step`` (:153-162).
The other integrators of the reference serve the deep/MuJoCo systems (out of scope).
"""
from typing import Callable, Optional, Tuple

import torch
from torch import Tensor
from torch.nn import Module

from dair_pll_b200.state_space import StateSpace

PartialStepCallback = Callable[[Tensor, Tensor], Tuple[Tensor, Tensor]]


class Integrator(Module):
    """Integrates dynamics given as a ``partial_step`` callback."""

    def __init__(self, space: StateSpace, partial_step_callback: PartialStepCallback, dt: float) -> None:
        super().__init__()
        self.partial_step_callback: Optional[PartialStepCallback] = partial_step_callback
        self.space = space
        self.dt = dt
        self.out_size = type(self).calc_out_size(space)

    def partial_step(self, x: Tensor, carry: Tensor) -> Tuple[Tensor, Tensor]:
        assert self.partial_step_callback is not None
        return self.partial_step_callback(x, carry)

    def step(self, x: Tensor, carry: Tensor) -> Tuple[Tensor, Tensor]:
        raise NotImplementedError

    def simulate(self, x_0: Tensor, carry_0: Tensor, steps: int) -> Tuple[Tensor, Tensor]:
        """(*, n_x) -> (*, steps + 1, n_x) state and carry trajectories."""
        assert steps >= 0 and x_0.shape[-1] == self.space.n_x
        xs, carries = [x_0], [carry_0]
        x, carry = x_0, carry_0
        for _ in range(steps):
            x, carry = self.step(x, carry)
            xs.append(x)
            carries.append(carry)
        return torch.stack(xs, -2), torch.stack(carries, -2)

    @staticmethod
    def calc_out_size(space: StateSpace) -> int:
        return space.n_x


class VelocityIntegrator(Integrator):
    """``partial_step`` returns the next velocity; the configuration follows by an implicit
    Euler step on the group: q+ = q (+) v+ dt."""

    def step(self, x: Tensor, carry: Tensor) -> Tuple[Tensor, Tensor]:
        q = self.space.q(x)
        v_next, carry = self.partial_step(x, carry)
        return self.space.x(self.space.euler_step(q, v_next, self.dt), v_next), carry

    @staticmethod
    def calc_out_size(space: StateSpace) -> int:
        return space.n_v
