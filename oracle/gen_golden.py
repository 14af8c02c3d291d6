"""ORACLE (test infrastructure).  Generates tests/golden/*.npz by running the REFERENCE's
own Python (through oracle/ref_shim.py) on seeded inputs.  Needs /root/reference, so it
runs in the build container only; the fixtures it writes are committed.

    python -m oracle.gen_golden
"""
import glob
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

DT = 0.0068  # examples/contactnets_simple.py:52
ASSETS = os.path.join(ref_shim.REFERENCE_ROOT, 'assets')


def real_cube_pairs(n_files: int, stride: int):
    files = sorted(glob.glob(os.path.join(ASSETS, 'contactnets_cube', '*.pt')),
                   key=lambda p: int(os.path.basename(p)[:-3]))[:n_files]
    xs, xps = [], []
    for f in files:
        t = torch.load(f)
        xs.append(t[:-1][::stride])
        xps.append(t[1:][::stride])
    return torch.cat(xs), torch.cat(xps)


def canonical_contact_permutation(system, q_plus):
    """Reference contact order (topk, unspecified) -> ascending box-vertex index."""
    from dair_pll.geometry import _UNIT_BOX_VERTICES
    ct = system.multibody_terms.contact_terms
    R_WC = ct.geometry_rotations(q_plus)
    perms = []
    n_geom = len(ct.geometries) - 1
    for g in range(n_geom):
        R_AB = R_WC[..., n_geom, :, :].transpose(-1, -2) @ R_WC[..., g, :, :]
        d_B = -R_AB.transpose(-1, -2)[..., 2]
        pts = ct.geometries[g].support_points(d_B).detach()               # (B,4,3) reference order
        signs = (pts > 0).long()
        idx = signs[..., 0] * 4 + signs[..., 1] * 2 + signs[..., 2]        # vertex index (geometry.py:39-41)
        perms.append(torch.argsort(idx, dim=-1) + 4 * g)
    return torch.cat(perms, -1)                                            # (B, n_c)


def reorder_force(force, perm):
    """force (B, 3 n_c) reference layout [n ; (tx,ty)...] with contacts permuted by perm."""
    n_c = perm.shape[-1]
    fn = torch.gather(force[:, :n_c], 1, perm)
    ft = force[:, n_c:].reshape(-1, n_c, 2)
    ft = torch.gather(ft, 1, perm[..., None].expand(-1, -1, 2)).reshape(-1, 2 * n_c)
    return torch.cat((fn, ft), -1)


class RecordingSolver:
    """Wraps the oracle solver to capture what the reference passed to / got from it."""

    def __init__(self):
        from oracle.cone_qp import OracleSAPSolver
        self.inner = OracleSAPSolver()
        self.calls = []

    def apply(self, J, q, eps):
        f = self.inner.apply(J, q, eps)
        self.calls.append((J.detach().clone(), q.detach().clone(), eps, f.detach().clone()))
        return f


def cube_case(name, x, x_plus, pi_cm, friction, half, sim_states, rollout_steps):
    from dair_pll import tensor_utils
    solver = RecordingSolver()
    system = ref_shim.build_reference_system('cube', DT, pi_cm, friction, [half.tolist()], solver=solver)
    mt = system.multibody_terms
    u = torch.zeros(x.shape[:-1] + (0,))
    loss = system.contactnets_loss(x, u, x_plus)
    loss.mean().backward()
    J_s, q_s, eps, f_s = solver.calls[-1]
    n_c = 4
    P = tensor_utils.sappy_reorder_mat(n_c)
    force_ref = (P @ f_s[..., None])[..., 0]
    perm = canonical_contact_permutation(system, x_plus[:, :7])
    out = dict(
        dt=np.float64(DT), x=x.numpy(), x_plus=x_plus.numpy(), pi_cm=pi_cm.numpy(),
        friction_params=friction.numpy(), half_lengths=half.numpy(),
        theta=mt.lagrangian_terms.inertial_parameters.detach().numpy(),
        loss=loss.detach().numpy(), force=reorder_force(force_ref, perm).numpy(),
        qp_J=J_s.numpy(), qp_q=q_s.numpy(), qp_eps=np.float64(eps), qp_f=f_s.numpy(),
        grad_theta=mt.lagrangian_terms.inertial_parameters.grad.numpy(),
        grad_friction=mt.contact_terms.friction_params.grad.numpy(),
        grad_length=mt.contact_terms.geometries[0].length_params.grad.numpy())
    # terms at (q+, v+) for term-level parity
    with torch.no_grad():
        D, M, J, phi, acc = mt(x_plus[:, :7], x_plus[:, 7:], u)
    out.update(terms_M=M.numpy(), terms_acc=acc.numpy(), terms_phi_sorted=np.sort(phi.numpy(), -1))
    # one step and a short rollout through the reference's System.simulate
    with torch.no_grad():
        x0 = sim_states
        traj, _ = system.simulate(x0.unsqueeze(-2), torch.zeros(x0.shape[0], 1), rollout_steps)
    out.update(sim_x0=x0.numpy(), sim_traj=traj.numpy())
    path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    np.savez_compressed(path, **out)
    print(name, 'B', x.shape[0], 'mean loss %.12e' % loss.mean().item(), 'size %.1f KB' % (os.path.getsize(path) / 1e3))


def elbow_case(name, x, x_plus, pi_cm, friction, half, sim_states, rollout_steps):
    from dair_pll import tensor_utils
    solver = RecordingSolver()
    system = ref_shim.build_reference_system('elbow', DT, pi_cm, friction, [h.tolist() for h in half], solver=solver)
    mt = system.multibody_terms
    u = torch.zeros(x.shape[:-1] + (0,))
    loss = system.contactnets_loss(x, u, x_plus)
    loss.mean().backward()
    J_s, q_s, eps, f_s = solver.calls[-1]
    n_c = 8
    P = tensor_utils.sappy_reorder_mat(n_c)
    force_ref = (P @ f_s[..., None])[..., 0]
    perm = canonical_contact_permutation(system, x_plus[:, :8])
    out = dict(
        dt=np.float64(DT), x=x.numpy(), x_plus=x_plus.numpy(), pi_cm=pi_cm.numpy(),
        friction_params=friction.numpy(), half_lengths=half.numpy(),
        theta=mt.lagrangian_terms.inertial_parameters.detach().numpy(),
        loss=loss.detach().numpy(), force=reorder_force(force_ref, perm).numpy(),
        grad_theta=mt.lagrangian_terms.inertial_parameters.grad.numpy(),
        grad_friction=mt.contact_terms.friction_params.grad.numpy(),
        grad_length=np.stack([mt.contact_terms.geometries[g].length_params.grad.numpy().reshape(3) for g in range(2)]))
    with torch.no_grad():
        D, M, J, phi, acc = mt(x_plus[:, :8], x_plus[:, 8:], u)
    out.update(terms_M=M.numpy(), terms_acc=acc.numpy(), terms_phi_sorted=np.sort(phi.numpy(), -1))
    with torch.no_grad():
        traj, _ = system.simulate(sim_states.unsqueeze(-2), torch.zeros(sim_states.shape[0], 1), rollout_steps)
    out.update(sim_x0=sim_states.numpy(), sim_traj=traj.numpy())
    path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    np.savez_compressed(path, **out)
    print(name, 'B', x.shape[0], 'mean loss %.12e' % loss.mean().item(), 'size %.1f KB' % (os.path.getsize(path) / 1e3))


def box_vertices(half):
    return torch.tensor([[sx * half[0], sy * half[1], sz * half[2]] for sx in (-1., 1.) for sy in (-1., 1.)
                         for sz in (-1., 1.)], dtype=torch.float64)


def elbow_mesh_case(name, x, pi_cm, friction, width, sim_n, rollout_steps):
    """contactnets_elbow_mesh.urdf: both links carry a DeepSupportConvex (learned support function)."""
    verts = [box_vertices(synthetic_half()) for _ in range(2)]
    system = ref_shim.build_reference_system('elbow', DT, pi_cm, friction, None, mesh_vertices=verts,
                                             mesh_width=width, mesh_seed=5)
    mt = system.multibody_terms
    u = torch.zeros(x.shape[:-1] + (0,))
    with torch.no_grad():
        xn, _ = system.integrator.step(x, torch.zeros(x.shape[0], 1))
    from dair_pll_b200 import synthetic as syn
    x_plus = syn.perturb_next_state(xn, seed=31, n_q=8)
    loss = system.contactnets_loss(x, u, x_plus)
    loss.mean().backward()
    out = dict(dt=np.float64(DT), x=x.numpy(), x_plus=x_plus.numpy(), pi_cm=pi_cm.numpy(),
               friction_params=friction.numpy(), theta=mt.lagrangian_terms.inertial_parameters.detach().numpy(),
               loss=loss.detach().numpy(),
               grad_theta=mt.lagrangian_terms.inertial_parameters.grad.numpy(),
               grad_friction=mt.contact_terms.friction_params.grad.numpy())
    for gi in range(2):
        net = mt.contact_terms.geometries[gi].network
        out[f'net{gi}_Wd0'] = net.input_weights[0].detach().numpy()
        out[f'net{gi}_Wd1'] = net.input_weights[1].detach().numpy()
        out[f'net{gi}_Wh'] = net.hidden_weights[0].detach().numpy()
        out[f'net{gi}_wout'] = net.output_weight.detach().numpy()
        out[f'net{gi}_perturbations'] = mt.contact_terms.geometries[gi].perturbations.numpy()
        out[f'net{gi}_grad_Wd0'] = net.input_weights[0].grad.numpy()
        out[f'net{gi}_grad_Wd1'] = net.input_weights[1].grad.numpy()
        out[f'net{gi}_grad_Wh'] = net.hidden_weights[0].grad.numpy()
        out[f'net{gi}_grad_wout'] = net.output_weight.grad.numpy()
    with torch.no_grad():
        traj, _ = system.simulate(x[:sim_n].unsqueeze(-2), torch.zeros(sim_n, 1), rollout_steps)
    out.update(sim_x0=x[:sim_n].numpy(), sim_traj=traj.numpy())
    path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    np.savez_compressed(path, **out)
    print(name, 'B', x.shape[0], 'mean loss %.12e' % loss.mean().item(), 'size %.1f KB' % (os.path.getsize(path) / 1e3))


def synthetic_half():
    from dair_pll_b200 import synthetic as syn
    return syn.ELBOW_HALF


def main():
    assert ref_shim.available(), 'needs the reference tree'
    ref_shim.import_reference()
    torch.set_default_dtype(torch.float64)
    sys.path.insert(0, ROOT)
    from dair_pll_b200 import synthetic

    # (1) nominal URDF parameters on real tosses (the survey's Appendix C regime)
    x, xp = real_cube_pairs(40, 9)
    pi_nom = torch.tensor([[0.37, 0, 0, 0, .00081, .00081, .00081, 0, 0, 0]])
    cube_case('cube_real_nominal', x, xp, pi_nom, torch.tensor([0.15, 1.0]), torch.tensor([.0524] * 3),
              x[::8][:48], 6)
    # (2) perturbed parameters with CoM offset on real tosses
    pi, fr, half = synthetic.cube_learnables_perturbed(0)
    x, xp = real_cube_pairs(60, 13)
    cube_case('cube_real_perturbed', x, xp, pi, fr, half, x[::6][:48], 6)
    # (3) synthetic states (all contact regimes); x+ = reference step + noise
    xs = synthetic.cube_states(384, seed=7)
    system = ref_shim.build_reference_system('cube', DT, pi, fr, [half.tolist()])
    with torch.no_grad():
        xn, _ = system.integrator.step(xs, torch.zeros(xs.shape[0], 1))
    xp = synthetic.perturb_next_state(xn, seed=8)
    cube_case('cube_synthetic', xs, xp, pi, fr, half, xs[:64], 4)
    # (4) elbow (two boxes + hinge), nominal URDF parameters and perturbed ones, synthetic states
    for name, seed, learn in (('elbow_nominal', 11, None), ('elbow_perturbed', 13, synthetic.elbow_learnables_perturbed(0))):
        if learn is None:
            m, i0 = synthetic.ELBOW_NOMINAL['m'], synthetic.ELBOW_NOMINAL['inertia']
            pi_e = torch.tensor([[m, 0, 0, 0, i0, i0, i0, 0, 0, 0], [m, m * 0.035, 0, 0, i0, i0, i0, 0, 0, 0]])
            fr_e = torch.tensor([0.3, 0.3, 1.0])
            half_e = torch.tensor([synthetic.ELBOW_HALF, synthetic.ELBOW_HALF])
        else:
            pi_e, fr_e, half_e = learn
        xe = synthetic.elbow_states(320, seed=seed)
        system = ref_shim.build_reference_system('elbow', DT, pi_e, fr_e, [h.tolist() for h in half_e])
        with torch.no_grad():
            xn, _ = system.integrator.step(xe, torch.zeros(xe.shape[0], 1))
        xpe = synthetic.perturb_next_state(xn, seed=seed + 1, n_q=8)
        elbow_case(name, xe, xpe, pi_e, fr_e, half_e, xe[:48], 4)
    # (5) elbow with learned (ICNN) geometry, width 64 to keep the fixture small
    pi_e, fr_e, _ = synthetic.elbow_learnables_perturbed(1)
    elbow_mesh_case('elbow_mesh_w64', synthetic.elbow_states(192, seed=17), pi_e, fr_e, 64, 32, 3)


if __name__ == '__main__':
    main()
