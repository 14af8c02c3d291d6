"""``System`` base class: the part of ``dair_pll/system.py`` (:47-191) the learnable
multibody system relies on -- ``simulate`` with the explicit time axis on the initial
condition, the outer-batch loop above ``max_batch_dim``, and the carry plumbing.
"""
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Tuple

import torch
from torch import Tensor
from torch.nn import Module

from dair_pll_b200.integrator import Integrator
from dair_pll_b200.state_space import StateSpace


@dataclass
class MeshSummary:
    vertices: Tensor = field(default_factory=lambda: torch.zeros(0, 3))
    faces: Tensor = field(default_factory=lambda: torch.zeros(0, 3, dtype=torch.long))


@dataclass
class SystemSummary:
    scalars: Dict[str, float] = field(default_factory=dict)
    videos: Dict[str, tuple] = field(default_factory=dict)
    meshes: Dict[str, MeshSummary] = field(default_factory=dict)


class System(Module):
    """Dynamical system = state space + integrator (+ samplers owned by the harness)."""

    def __init__(self, space: StateSpace, integrator: Integrator, max_batch_dim: Optional[int] = None) -> None:
        super().__init__()
        self.space = space
        self.integrator = integrator
        self.carry_callback: Optional[Callable[[], Tensor]] = lambda: torch.zeros((1, 1))
        self.max_batch_dim = max_batch_dim

    def set_carry_sampler(self, callback: Callable[[], Tensor]) -> None:
        self.carry_callback = callback

    def preprocess_initial_condition(self, x_0: Tensor, carry_0: Tensor) -> Tuple[Tensor, Tensor]:
        """(*, T_0, n_x) -> (*, n_x): the most recent state starts the integration."""
        assert x_0.dim() >= 2 and carry_0.dim() >= 1 and x_0.shape[-1] == self.space.n_x
        if self.max_batch_dim is not None:
            assert x_0.dim() <= 2 + self.max_batch_dim
        return x_0[..., -1, :], carry_0

    def simulate(self, x_0: Tensor, carry_0: Tensor, steps: int = 1) -> Tuple[Tensor, Tensor]:
        """(*, T_0, n_x) initial sequence -> (*, steps + 1, n_x) trajectory."""
        if self.max_batch_dim is not None and (x_0.dim() - 2) > self.max_batch_dim:
            # the reference loops the outermost batch dimension here (system.py:115-124),
            # calling simulate() with its default ``steps``; mirrored literally.
            outs = [self.simulate(x, c) for x, c in zip(x_0, carry_0)]
            return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])
        x, carry = self.preprocess_initial_condition(x_0, carry_0)
        return self.integrator.simulate(x, carry, steps)

    def summary(self, statistics: Dict) -> SystemSummary:
        del statistics
        return SystemSummary()
