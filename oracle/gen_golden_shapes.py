"""ORACLE (test infrastructure).  Writes tests/golden/shape_{sphere,polygon,framed_box}.npz by running the REFERENCE's own
``contactnets_loss`` / ``sim_step`` (through oracle/ref_shim.py) for a single floating body whose collision geometry
is the reference's ``Sphere`` (dair_pll/geometry.py:415-456) or ``Polygon`` (:220-252) against the ground plane.
Needs /root/reference: build container only; the fixtures are committed.

    python -m oracle.gen_golden_shapes

Note: the reference's ``Sphere.__init__`` cannot run as written (``assert radius.numel == 1`` compares a bound method
with 1, geometry.py:430); the object is therefore built with ``__new__`` and given its ``length_param`` directly --
every method that is exercised (``get_radius``, ``support_points``) is the reference's own.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

from oracle import ref_shim  # noqa: E402

DT = 0.0068
POLY = np.array([[0.06, 0.0, -0.03], [-0.05, 0.04, -0.035], [-0.04, -0.05, -0.03], [0.0, 0.0, 0.07], [0.05, 0.05, 0.02],
                 [-0.06, 0.01, 0.03], [0.02, -0.06, 0.025], [0.055, -0.03, -0.01], [-0.02, 0.06, 0.0], [0.0, -0.02, -0.06]])


def states(n, seed, support):
    """random orientation; height = support height of the shape + delta as in synthetic.cube_states"""
    g = torch.Generator().manual_seed(seed)
    quat = torch.randn(n, 4, generator=g, dtype=torch.float64)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    xy = torch.rand(n, 2, generator=g, dtype=torch.float64) - 0.5
    near = torch.rand(n, generator=g, dtype=torch.float64) < 0.7
    u = torch.rand(n, generator=g, dtype=torch.float64)
    delta = torch.where(near, -0.004 + 0.01 * u, 0.01 + 0.1 * u)
    z = support(quat) + delta
    omega = 4.0 * torch.randn(n, 3, generator=g, dtype=torch.float64)
    vel = 0.7 * torch.randn(n, 3, generator=g, dtype=torch.float64)
    return torch.cat((quat, xy, z[:, None], omega, vel), -1)


def rot(quat):
    from dair_pll_b200.synthetic import _quat_to_rot
    return _quat_to_rot(quat)


def case(name, geom, learn_name, support, seed, tree='cube'):
    pi_cm = torch.tensor([[0.31, 0.31 * 0.002, -0.31 * 0.001, 0.31 * 0.0015, 6.1e-4, 7.3e-4, 6.6e-4, 1e-5, -2e-5, 1.5e-5]],
                         dtype=torch.float64)
    friction = torch.tensor([0.4, 0.9], dtype=torch.float64)
    system = ref_shim.build_reference_system(tree, DT, pi_cm, friction, None, body_geometries=[geom])
    n = 320
    x = states(n, seed, support)
    with torch.no_grad():
        nxt, _ = system.integrator.step(x, torch.zeros(n, 1))
    gnoise = torch.Generator().manual_seed(seed + 1)
    x_plus = nxt.clone()
    x_plus[:, 7:] += 0.02 * torch.randn(n, 6, generator=gnoise, dtype=torch.float64)
    loss = system.contactnets_loss(x, torch.zeros(n, 0), x_plus)
    loss.mean().backward()
    mt = system.multibody_terms
    out = dict(dt=np.array(DT), x=x.numpy(), x_plus=x_plus.numpy(), x_next=nxt.numpy(), pi_cm=pi_cm.numpy(),
               theta=mt.lagrangian_terms.inertial_parameters.detach().numpy(), friction_params=friction.numpy(),
               loss=loss.detach().numpy(), grad_theta=mt.lagrangian_terms.inertial_parameters.grad.numpy(),
               grad_friction=mt.contact_terms.friction_params.grad.numpy())
    p = getattr(geom, learn_name)
    out['shape_param'] = p.detach().numpy()
    out['grad_shape_param'] = p.grad.numpy()
    # prediction-loss path: a 3-step rollout by the reference's own integrator, weighted sum of the states, autograd
    # through every step's QP -> gradients of theta, friction, the shape parameter and the initial state
    for q in system.parameters():
        q.grad = None
    m, steps = 40, 3
    x0 = x[:m].clone().requires_grad_()
    gw = torch.Generator().manual_seed(seed + 2)
    w = torch.randn(m, steps, 13, generator=gw, dtype=torch.float64)
    xs, cur = [], x0
    for _ in range(steps):
        cur, _ = system.integrator.step(cur, torch.zeros(m, 1))
        xs.append(cur)
    traj = torch.stack(xs, 1)
    (traj * w).sum().backward()
    out.update(roll_x0=x0.detach().numpy(), roll_w=w.numpy(), roll_traj=traj.detach().numpy(),
               roll_grad_x0=x0.grad.numpy(), roll_grad_theta=mt.lagrangian_terms.inertial_parameters.grad.numpy(),
               roll_grad_friction=mt.contact_terms.friction_params.grad.numpy(), roll_grad_shape_param=p.grad.numpy())
    path = os.path.join(ROOT, 'tests', 'golden', f'shape_{name}.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, 'mean loss', float(loss.mean()), 'in contact', float((loss.detach() > 1e-12).float().mean()))


def main():
    ref_shim.import_reference()
    from dair_pll.geometry import Polygon, Sphere
    from torch.nn import Module, Parameter
    sphere = Sphere.__new__(Sphere)
    Module.__init__(sphere)
    sphere.length_param = Parameter(torch.tensor(0.05, dtype=torch.float64))
    case('sphere', sphere, 'length_param', lambda q: torch.full((q.shape[0],), 0.05, dtype=torch.float64), 101)
    verts = torch.from_numpy(POLY)
    poly = Polygon(verts, 4)

    def support(q):
        row = rot(q)[:, 2, :]                       # R[2, :]: height of vertex v is row . v
        return -(row @ verts.t()).min(dim=1).values
    case('polygon', poly, 'vertices', support, 202)
    # a Box (geometry.py:359-413) in a collision frame that is offset and rotated in the link (FRAMED_BODY_TREE): the
    # reference gets R_WG, p_WG and the geometry Jacobian per collision frame (multibody_terms.py:299-310)
    from dair_pll.geometry import Box
    from oracle.callables import FRAMED_BODY_TREE, TreeCallables
    half = torch.tensor([0.05, 0.035, 0.025], dtype=torch.float64)
    box = Box(half, 4)
    off = torch.tensor(FRAMED_BODY_TREE.geometry_offset[0], dtype=torch.float64)
    Rg = FRAMED_BODY_TREE.geometry_rotation(0, torch.float64)
    signs = torch.tensor([[(i >> 2) & 1, (i >> 1) & 1, i & 1] for i in range(8)], dtype=torch.float64) * 2 - 1
    corners = off + (signs * half) @ Rg.t()                     # link frame

    def support_framed(q):
        row = rot(q)[:, 2, :]
        return -(row @ corners.t()).min(dim=1).values
    case('framed_box', box, 'length_params', support_framed, 303, tree=FRAMED_BODY_TREE)


if __name__ == '__main__':
    main()
