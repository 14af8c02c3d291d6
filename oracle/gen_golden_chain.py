"""ORACLE (test infrastructure).  Writes tests/golden/chain3.npz, chain3r.npz (rotated collision frames), slider3.npz (a prismatic joint), tree4.npz, tree4g.npz (boxes spread unevenly over the links) and tree6.npz by running the REFERENCE's own
``contactnets_loss`` / ``sim_step`` (through oracle/ref_shim.py) for a three-link floating chain (oracle/callables.py:
CHAIN3_TREE: off-axis second joint with a rotated joint frame, one box per link) and for a BRANCHING four-link tree
(TREE4_TREE: two links off the root, a third off one of them) and a six-link tree (TREE6_TREE) -- the fixtures of the generic tree kernels (SURVEY.md 8(f)
N2).  Needs /root/reference: build container only; the fixture is committed.

    python -m oracle.gen_golden_chain
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
from oracle.callables import (CHAIN3_TREE, CHAIN3R_TREE, SLIDER3_TREE, TREE4_TREE, TREE4G_TREE, TREE6_TREE,  # noqa: E402
                              TreeCallables)

DT = 0.0068
HALF = np.array([[0.05, 0.025, 0.025], [0.045, 0.03, 0.02], [0.03, 0.02, 0.035]])


HALF6 = np.array([[0.05, 0.03, 0.02], [0.04, 0.02, 0.02], [0.04, 0.02, 0.02], [0.02, 0.04, 0.02], [0.03, 0.02, 0.03],
                  [0.02, 0.035, 0.02]])
HALF4 = np.array([[0.05, 0.025, 0.025], [0.045, 0.03, 0.02], [0.03, 0.035, 0.02], [0.03, 0.02, 0.035]])


def lowest_corner(calls, q, half=HALF):
    """z of the lowest box corner of any link at the configuration q (with its stored position)."""
    R = calls.geometry_rotations(q)
    p = calls.geometry_translations(q)
    low = []
    for g in range(len(half)):
        ext = (R[:, g, 2, :].abs() * torch.from_numpy(half[g])).sum(-1)
        low.append(p[:, g, 2] - ext)
    return torch.stack(low, -1).min(-1).values


def states(n, seed, calls, half=HALF):
    nj = calls.tree.n_bodies - 1
    g = torch.Generator().manual_seed(seed)
    quat = torch.randn(n, 4, generator=g, dtype=torch.float64)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    xy = torch.rand(n, 2, generator=g, dtype=torch.float64) - 0.5
    th = (2 * torch.rand(n, nj, generator=g, dtype=torch.float64) - 1) * 2.5
    for j in range(nj):
        if calls.tree.is_prismatic(j + 1):
            th[:, j] *= 0.02                     # a sliding joint's coordinate is a length (m)
    q = torch.cat((quat, xy, torch.zeros(n, 1, dtype=torch.float64), th), -1)
    near = torch.rand(n, generator=g, dtype=torch.float64) < 0.7
    u = torch.rand(n, generator=g, dtype=torch.float64)
    delta = torch.where(near, -0.004 + 0.01 * u, 0.01 + 0.1 * u)
    q[:, 6] = -lowest_corner(calls, q, half) + delta
    omega = 4.0 * torch.randn(n, 3, generator=g, dtype=torch.float64)
    vel = 0.7 * torch.randn(n, 3, generator=g, dtype=torch.float64)
    rate = 5.0 * torch.randn(n, nj, generator=g, dtype=torch.float64)
    return torch.cat((q, omega, vel, rate), -1)


def make(tree, half, coms, friction, name, n, m):
    calls = TreeCallables(tree)
    nb = tree.n_bodies
    ng = len(half)                         # box geometries (any distribution over the links)
    gen = torch.Generator().manual_seed(7)
    rows = []
    for com0 in coms:
        mass = 0.3 + 0.1 * torch.rand(1, generator=gen, dtype=torch.float64)
        c = torch.tensor(com0, dtype=torch.float64) + 0.004 * (2 * torch.rand(3, generator=gen, dtype=torch.float64) - 1)
        diag = 6e-4 * (1 + 0.2 * (2 * torch.rand(3, generator=gen, dtype=torch.float64) - 1))
        offd = 2e-5 * (2 * torch.rand(3, generator=gen, dtype=torch.float64) - 1)
        rows.append(torch.cat((mass, mass * c, diag, offd)))
    pi_cm = torch.stack(rows)
    friction = torch.tensor(friction, dtype=torch.float64)
    system = ref_shim.build_reference_system(tree, DT, pi_cm, friction, [h.tolist() for h in half])
    x = states(n, 11, calls, half)
    nx = x.shape[1]
    nq = 7 + nb - 1
    with torch.no_grad():
        nxt, _ = system.integrator.step(x, torch.zeros(n, 1))
    x_plus = nxt.clone()
    x_plus[:, nq:] += 0.02 * torch.randn(n, nx - nq, generator=torch.Generator().manual_seed(12), dtype=torch.float64)
    loss = system.contactnets_loss(x, torch.zeros(n, 0), x_plus)
    loss.mean().backward()
    mt = system.multibody_terms
    out = dict(dt=np.array(DT), x=x.numpy(), x_plus=x_plus.numpy(), x_next=nxt.numpy(), pi_cm=pi_cm.numpy(),
               theta=mt.lagrangian_terms.inertial_parameters.detach().numpy(), friction_params=friction.numpy(),
               half_lengths=half, loss=loss.detach().numpy(),
               grad_theta=mt.lagrangian_terms.inertial_parameters.grad.numpy(),
               grad_friction=mt.contact_terms.friction_params.grad.numpy(),
               grad_length=np.stack([mt.contact_terms.geometries[i].length_params.grad.numpy().reshape(3) for i in range(ng)]))
    # prediction-loss path: 3-step rollout by the reference's own integrator, weighted sum of the states, autograd through
    # every step's QP -> gradients of theta, friction, box lengths and the initial state
    for q in system.parameters():
        q.grad = None
    steps = 3
    x0 = x[:m].clone().requires_grad_()
    w = torch.randn(m, steps, nx, generator=torch.Generator().manual_seed(13), dtype=torch.float64)
    xs, cur = [], x0
    for _ in range(steps):
        cur, _ = system.integrator.step(cur, torch.zeros(m, 1))
        xs.append(cur)
    traj = torch.stack(xs, 1)
    (traj * w).sum().backward()
    out.update(roll_x0=x0.detach().numpy(), roll_w=w.numpy(), roll_traj=traj.detach().numpy(), roll_grad_x0=x0.grad.numpy(),
               roll_grad_theta=mt.lagrangian_terms.inertial_parameters.grad.numpy(),
               roll_grad_friction=mt.contact_terms.friction_params.grad.numpy(),
               roll_grad_length=np.stack([mt.contact_terms.geometries[i].length_params.grad.numpy().reshape(3)
                                          for i in range(ng)]))
    path = os.path.join(ROOT, 'tests', 'golden', f'{name}.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, 'mean loss', float(loss.mean()))


def make_shapes():
    """tests/golden/chain3s.npz: CHAIN3_TREE's kinematics with the reference's Box on link 0, Sphere on link 1 and Polygon
    on link 2 (geometry.py:359-413, 415-456, 220-252) -- the fixture of the witness-point form of the tree kernels."""
    from dair_pll.geometry import Box, Polygon, Sphere
    from torch.nn import Module, Parameter
    from oracle.gen_golden_shapes import POLY
    tree = CHAIN3_TREE
    calls = TreeCallables(tree)
    box = Box(torch.from_numpy(HALF[0]), 4)
    sphere = Sphere.__new__(Sphere)                 # (Sphere.__init__ cannot run as written: see gen_golden_shapes.py)
    Module.__init__(sphere)
    sphere.length_param = Parameter(torch.tensor(0.03, dtype=torch.float64))
    poly = Polygon(torch.from_numpy(POLY) * 0.6, 4)
    gen = torch.Generator().manual_seed(7)
    rows = []
    for com0 in ((0., 0., 0.), (0.035, 0., 0.), (0.03, -0.01, 0.)):
        mass = 0.3 + 0.1 * torch.rand(1, generator=gen, dtype=torch.float64)
        c = torch.tensor(com0, dtype=torch.float64) + 0.004 * (2 * torch.rand(3, generator=gen, dtype=torch.float64) - 1)
        diag = 6e-4 * (1 + 0.2 * (2 * torch.rand(3, generator=gen, dtype=torch.float64) - 1))
        offd = 2e-5 * (2 * torch.rand(3, generator=gen, dtype=torch.float64) - 1)
        rows.append(torch.cat((mass, mass * c, diag, offd)))
    pi_cm = torch.stack(rows)
    friction = torch.tensor([0.3, 0.45, 0.25, 0.9], dtype=torch.float64)
    system = ref_shim.build_reference_system(tree, DT, pi_cm, friction, None, body_geometries=[box, sphere, poly])
    n = 160
    x = states(n, 21, calls, np.array([HALF[0], [0.03, 0.03, 0.03], np.abs(POLY * 0.6).max(0)]))
    with torch.no_grad():
        nxt, _ = system.integrator.step(x, torch.zeros(n, 1))
    x_plus = nxt.clone()
    x_plus[:, 9:] += 0.02 * torch.randn(n, 8, generator=torch.Generator().manual_seed(22), dtype=torch.float64)
    loss = system.contactnets_loss(x, torch.zeros(n, 0), x_plus)
    loss.mean().backward()
    mt = system.multibody_terms
    out = dict(dt=np.array(DT), x=x.numpy(), x_plus=x_plus.numpy(), x_next=nxt.numpy(), pi_cm=pi_cm.numpy(),
               theta=mt.lagrangian_terms.inertial_parameters.detach().numpy(), friction_params=friction.numpy(),
               loss=loss.detach().numpy(), grad_theta=mt.lagrangian_terms.inertial_parameters.grad.numpy(),
               grad_friction=mt.contact_terms.friction_params.grad.numpy(),
               box_length_params=box.length_params.detach().numpy(), grad_box_length_params=box.length_params.grad.numpy(),
               sphere_radius=sphere.length_param.detach().numpy(), grad_sphere_radius=sphere.length_param.grad.numpy(),
               polygon_vertices=poly.vertices.detach().numpy(), grad_polygon_vertices=poly.vertices.grad.numpy())
    path = os.path.join(ROOT, 'tests', 'golden', 'chain3s.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, 'mean loss', float(loss.mean()))


def main():
    ref_shim.import_reference()
    make_shapes()
    make(CHAIN3_TREE, HALF, ((0., 0., 0.), (0.035, 0., 0.), (0.03, -0.01, 0.)), [0.3, 0.45, 0.25, 0.9], 'chain3', 256, 24)
    make(CHAIN3R_TREE, HALF, ((0., 0., 0.), (0.035, 0., 0.), (0.03, -0.01, 0.)), [0.3, 0.45, 0.25, 0.9], 'chain3r', 128, 8)
    make(SLIDER3_TREE, HALF, ((0., 0., 0.), (0.035, 0., 0.), (0.02, -0.01, 0.)), [0.3, 0.45, 0.25, 0.9], 'slider3', 128, 8)
    make(TREE4_TREE, HALF4, ((0., 0., 0.), (0.035, 0., 0.), (0.0, -0.03, 0.), (0.03, -0.01, 0.)),
         [0.3, 0.45, 0.35, 0.25, 0.9], 'tree4', 192, 16)
    make(TREE4G_TREE, HALF4[:3], ((0., 0., 0.), (0.035, 0., 0.), (0.0, -0.03, 0.), (0.03, -0.01, 0.)),
         [0.3, 0.45, 0.25, 0.9], 'tree4g', 128, 8)
    make(TREE6_TREE, HALF6, ((0., 0., 0.), (0.03, 0., 0.), (-0.03, 0., 0.), (0., -0.03, 0.), (0.03, -0.01, 0.), (0., -0.03, 0.01)),
         [0.3, 0.45, 0.35, 0.25, 0.5, 0.4, 0.9], 'tree6', 128, 8)


if __name__ == '__main__':
    main()
