"""ORACLE (test infrastructure, not product code) -- runs only where
``/root/reference`` exists (the build container), never on the GPU box.

Imports the reference's OWN modules (``dair_pll.multibody_learnable_system``,
``multibody_terms``, ``geometry``, ``inertia``, ``integrator``, ``state_space``,
``system``) with the third-party packages that are absent from this image
(pydrake, drake_pytorch, sappy, fcl, pywavefront, optuna, wandb, ...) replaced by
inert stubs, then assembles a ``MultibodyLearnableSystem`` without calling the
pydrake-dependent constructors: the five symbolic callables come from
``oracle/callables.py`` and the solver from ``oracle/cone_qp.py`` (the two
un-vendored boundaries, restated); everything else -- the ContactNets loss
assembly (multibody_learnable_system.py:104-197), forward dynamics (:199-304),
ContactTerms.forward (multibody_terms.py:428-521), LagrangianTerms.forward
(:214-237), Box/Plane/GeometryCollider (geometry.py), the inertia conversions
(inertia.py), VelocityIntegrator and the state spaces -- is the reference's code
executing unmodified.  ``oracle/gen_golden.py`` uses this to write tests/golden/.
"""
import os
import sys
import types
from unittest.mock import MagicMock

import torch

REFERENCE_ROOT = os.environ.get('DAIR_PLL_REFERENCE', '/root/reference')

_STUBS = [
    'pydrake', 'pydrake.all', 'pydrake.geometry', 'pydrake.math', 'pydrake.multibody',
    'pydrake.multibody.plant', 'pydrake.multibody.tree', 'pydrake.multibody.parsing',
    'pydrake.symbolic', 'pydrake.systems', 'pydrake.systems.framework', 'pydrake.systems.analysis',
    'pydrake.systems.sensors', 'pydrake.visualization', 'pydrake.autodiffutils', 'pydrake.common',
    'pydrake.common.eigen_geometry', 'pydrake.common.value', 'pydrake.systems.primitives',
    'pydrake.systems.rendering', 'pydrake.systems.planar_scenegraph_visualizer',
    'drake_pytorch', 'sappy', 'fcl', 'pywavefront', 'optuna', 'optuna.trial', 'optuna.study', 'wandb',
    'mujoco_py', 'matplotlib', 'matplotlib.pyplot', 'moviepy', 'moviepy.editor', 'PIL', 'PIL.Image',
    'PIL.ImageDraw', 'PIL.ImageFont', 'git', 'psutil_stub',
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'dair_pll'))


class _Mock(MagicMock):
    # let ``class Foo(pydrake.Something)`` / ``Type_[float]`` in annotations work
    def __getitem__(self, item):
        return MagicMock()


def import_reference():
    """Returns the reference ``dair_pll`` package with stubs installed."""
    if not available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # pylint: disable=broad-except
                sys.modules[name] = _Mock(name=name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import dair_pll  # noqa  pylint: disable=import-outside-toplevel
    from dair_pll import (multibody_learnable_system, multibody_terms, geometry, inertia,  # noqa
                          integrator, state_space, system, tensor_utils, quaternion)
    return dair_pll


def build_reference_system(kind: str, dt: float, pi_cm: torch.Tensor, friction: torch.Tensor,
                           half_lengths, solver=None, mesh_vertices=None, mesh_width: int = 256, mesh_seed: int = 0,
                           body_geometries=None):
    """Reference ``MultibodyLearnableSystem`` for ``kind`` in {'cube','elbow'}.

    Args:
        pi_cm: (n_bodies, 10) inertia in the reference's pi_cm format (inertia.py:21-24).
        friction: (n_geometries,) friction_params, body geometries first, ground last.
        half_lengths: list of (3,) box half-lengths, one per body geometry.
    """
    import_reference()
    from dair_pll.multibody_learnable_system import MultibodyLearnableSystem
    from dair_pll.multibody_terms import MultibodyTerms, LagrangianTerms, ContactTerms
    from dair_pll.geometry import Box, Plane
    from dair_pll.inertia import InertialParameterConverter
    from dair_pll.integrator import VelocityIntegrator
    from dair_pll.state_space import ProductSpace, FixedBaseSpace, FloatingBaseSpace
    from dair_pll.system import System
    from torch.nn import Module, ModuleList, Parameter

    from oracle.callables import TreeCallables, CUBE_TREE, ELBOW_TREE
    from oracle.cone_qp import OracleSAPSolver

    from oracle.callables import CHAIN3_TREE
    tree = kind if not isinstance(kind, str) else {'cube': CUBE_TREE, 'elbow': ELBOW_TREE, 'chain3': CHAIN3_TREE}[kind]
    calls = TreeCallables(tree)

    lt = LagrangianTerms.__new__(LagrangianTerms)
    Module.__init__(lt)
    lt.mass_matrix = calls.mass_matrix
    lt.lagrangian_forces = calls.lagrangian_forces
    lt.inertial_parameters = Parameter(InertialParameterConverter.pi_cm_to_theta(pi_cm.double()))

    ct = ContactTerms.__new__(ContactTerms)
    Module.__init__(ct)
    ct.geometry_rotations = calls.geometry_rotations
    ct.geometry_translations = calls.geometry_translations
    ct.geometry_spatial_jacobians = calls.geometry_spatial_jacobians
    n_body_geoms = len(tree.geometry_body) - 1
    if body_geometries is not None:
        # caller-built reference geometry objects (Sphere, Polygon, ...), one per body geometry
        geoms = list(body_geometries) + [Plane()]
    elif mesh_vertices is not None:
        # the reference's own DeepSupportConvex (geometry.py:255-325): random ICNN initialisation and
        # random direction perturbations, made reproducible by seeding
        from dair_pll.geometry import DeepSupportConvex
        torch.manual_seed(mesh_seed)
        geoms = [DeepSupportConvex(torch.as_tensor(v, dtype=torch.float64), width=mesh_width) for v in mesh_vertices] \
            + [Plane()]
        for geo in geoms[:-1]:
            geo.perturbations = geo.perturbations.double()
    else:
        geoms = [Box(torch.as_tensor(h, dtype=torch.float64), 4) for h in half_lengths] + [Plane()]
    assert len(geoms) == n_body_geoms + 1
    ct.geometries = ModuleList(geoms)
    ct.friction_params = Parameter(friction.double().clone())
    # (a = ground (Plane sorts first, geometry.py:46), b = body geometry)
    ct.collision_candidates = torch.tensor([[n_body_geoms] * n_body_geoms,
                                            list(range(n_body_geoms))]).long()

    terms = MultibodyTerms.__new__(MultibodyTerms)
    Module.__init__(terms)
    terms.lagrangian_terms = lt
    terms.contact_terms = ct

    space = ProductSpace([FixedBaseSpace(0), FloatingBaseSpace(tree.n_joints)])
    sysm = MultibodyLearnableSystem.__new__(MultibodyLearnableSystem)
    integrator = VelocityIntegrator(space, sysm.sim_step, dt)
    System.__init__(sysm, space, integrator)
    sysm.multibody_terms = terms
    sysm.solver = solver if solver is not None else OracleSAPSolver()
    sysm.dt = dt
    sysm.set_carry_sampler(lambda: torch.Tensor([False]))
    sysm.max_batch_dim = 1
    return sysm
