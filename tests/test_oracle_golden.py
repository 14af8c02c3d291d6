"""CPU: pins the portable oracle (oracle/contactnets_oracle.py + oracle/cone_qp.c) against the
golden vectors produced by the reference's own Python (oracle/gen_golden.py), and certifies the
QP solutions independently (KKT + a second algorithm)."""
import numpy as np
import pytest
import torch

from oracle import cone_qp
from oracle import contactnets_oracle as co
from oracle.callables import CUBE_TREE, ELBOW_TREE, SLIDER3_TREE, TREE4_TREE, TREE6_TREE, TreeCallables
from tests.util import load_golden, max_rel_to_scale, oracle_params_from_golden, rel_err

CASES = ['cube_real_nominal', 'cube_real_perturbed', 'cube_synthetic']
CALLS = TreeCallables(CUBE_TREE)
ELBOW_CASES = ['elbow_nominal', 'elbow_perturbed']
ELBOW_CALLS = TreeCallables(ELBOW_TREE)


@pytest.mark.parametrize('name', CASES)
def test_loss_force_and_gradients_match_reference_python(name):
    g = load_golden(name)
    P = oracle_params_from_golden(g)
    x, xp = torch.from_numpy(g['x']), torch.from_numpy(g['x_plus'])
    loss, force = co.contactnets_loss(CALLS, P, x, xp, float(g['dt']), return_force=True)
    loss.mean().backward()
    assert rel_err(loss.detach().numpy(), g['loss'], 1e-9).max() < 1e-9
    assert np.abs(loss.detach().numpy() - g['loss']).max() < 1e-13
    # forces: 1e-9 of the per-sample force scale (weakly determined components, cond(Q) ~ 5e5)
    scale = np.maximum(np.abs(g['force']).max(axis=1, keepdims=True), 1e-6)
    assert (np.abs(force.numpy() - g['force']) / scale).max() < 1e-8
    assert max_rel_to_scale(P.inertial_parameters.grad.numpy(), g['grad_theta']) < 1e-9
    assert max_rel_to_scale(P.friction_params.grad.numpy(), g['grad_friction']) < 1e-9
    assert max_rel_to_scale(P.length_params[0].grad.numpy(), g['grad_length']) < 1e-9


@pytest.mark.parametrize('name', CASES)
def test_terms_match_reference_python(name):
    g = load_golden(name)
    P = oracle_params_from_golden(g, False)
    xp = torch.from_numpy(g['x_plus'])
    M, J, phi, acc = co.multibody_terms(CALLS, P, xp[:, :7], xp[:, 7:])
    assert np.abs(M.numpy() - g['terms_M']).max() < 1e-15
    assert np.abs(acc.numpy() - g['terms_acc']).max() < 1e-11 * max(1.0, np.abs(g['terms_acc']).max())
    assert np.abs(np.sort(phi.numpy(), -1) - g['terms_phi_sorted']).max() < 1e-15


@pytest.mark.parametrize('name', CASES)
def test_simulation_matches_reference_python(name):
    g = load_golden(name)
    P = oracle_params_from_golden(g, False)
    x0 = torch.from_numpy(g['sim_x0'])
    steps = g['sim_traj'].shape[1] - 1
    traj = co.simulate(CALLS, P, x0, float(g['dt']), steps).numpy()
    # first step tight; later steps compound contact sensitivity
    assert np.abs(traj[:, 1] - g['sim_traj'][:, 1]).max() < 1e-9
    assert np.abs(traj - g['sim_traj']).max() < 1e-7


@pytest.mark.parametrize('name', CASES)
def test_qp_solutions_are_kkt_certified(name):
    """Solver-independent certificate: f in K, Qf+q in K, complementarity."""
    g = load_golden(name)
    A, q, eps = torch.from_numpy(g['qp_J']), torch.from_numpy(g['qp_q']), float(g['qp_eps'])
    kkt_golden = cone_qp.kkt_residual(A, q, eps, torch.from_numpy(g['qp_f']))
    assert kkt_golden.max().item() < 1e-11
    f, _, iters, _ = cone_qp.solve(g['qp_J'], g['qp_q'], eps)
    assert cone_qp.kkt_residual(A, q, eps, torch.from_numpy(f)).max().item() < 1e-11
    assert iters.max() < 100


def test_second_algorithm_agrees():
    """Accelerated projected gradient on the dual (no shared logic beyond the projection)."""
    g = load_golden('cube_real_perturbed')
    sel = np.argsort(-np.abs(g['qp_f']).max(axis=1))[:64]
    A, q, eps = g['qp_J'][sel], g['qp_q'][sel], float(g['qp_eps'])
    f_newton, _, _, _ = cone_qp.solve(A, q, eps)
    f_apg, _ = cone_qp.solve_apg(A, q, eps)
    scale = np.abs(f_newton).max(axis=1, keepdims=True)
    assert (np.abs(f_newton - f_apg) / scale).max() < 1e-9


def test_survey_anchor_values():
    """SURVEY.md Appendix C regression anchors (nominal URDF parameters, real tosses):
    theta0 of the nominal cube and the free-flight step of 0.pt[0]."""
    pi_cm = torch.tensor([[0.37, 0, 0, 0, .00081, .00081, .00081, 0, 0, 0]], dtype=torch.float64)
    theta = co.pi_cm_to_theta(pi_cm)[0].numpy()
    assert np.allclose(theta[:4], [-0.4971, -3.4087, -3.4087, -3.4087], atol=5e-5)
    assert np.allclose(theta[4:], 0, atol=1e-15)
    P = co.OracleParams(co.pi_cm_to_theta(pi_cm), torch.tensor([0.15, 1.0], dtype=torch.float64),
                        [torch.tensor([[.0524] * 3], dtype=torch.float64)])
    g = load_golden('cube_real_nominal')
    x0 = torch.from_numpy(g['x'][:1])        # trajectory 0, first state: free flight
    xn = co.sim_step(CALLS, P, x0, 0.0068)[0].numpy()
    assert abs(xn[12] - (x0[0, 12].item() - 9.81 * 0.0068)) < 1e-12
    assert np.allclose(xn[7:12], x0[0, 7:12].numpy(), atol=1e-12)
    assert np.allclose(xn[:4], [-0.15705441058, -0.86646730033, 0.46472603915, 0.09272560657], atol=1e-10)


def test_icnn_support_equals_autograd_of_support_function():
    """deep_support_function.py:213-266: forward() is the input-Jacobian of the network output."""
    torch.manual_seed(0)
    W = 32
    net = dict(Wd0=torch.randn(3, W, dtype=torch.float64), Wd1=torch.randn(3, W, dtype=torch.float64),
               Wh=torch.randn(W, W, dtype=torch.float64) / W, wout=torch.randn(W, dtype=torch.float64),
               perturbations=torch.zeros(4, 3, dtype=torch.float64))
    d = torch.randn(17, 3, dtype=torch.float64)
    d = (d / d.norm(dim=-1, keepdim=True)).requires_grad_()
    lrelu = torch.nn.functional.leaky_relu
    h0 = lrelu(d @ net['Wd0'], 0.5)
    h1 = lrelu(h0 @ net['Wh'].abs() + d @ net['Wd1'], 0.5)
    f = (h1 @ net['wout'].abs()).sum()
    (jac,) = torch.autograd.grad(f, d)
    assert torch.allclose(co.icnn_support(net, d.detach()), jac, atol=1e-12)


@pytest.mark.parametrize('name', ELBOW_CASES)
def test_elbow_matches_reference_python(name):
    """Two-body articulated system (contactnets_elbow.urdf): loss, impulses, gradients, terms, rollout."""
    g = load_golden(name)
    P = oracle_params_from_golden(g)
    x, xp = torch.from_numpy(g['x']), torch.from_numpy(g['x_plus'])
    loss, force = co.contactnets_loss(ELBOW_CALLS, P, x, xp, float(g['dt']), return_force=True)
    loss.mean().backward()
    assert np.abs(loss.detach().numpy() - g['loss']).max() < 1e-13
    assert rel_err(loss.detach().numpy(), g['loss'], 1e-9).max() < 1e-9
    scale = np.maximum(np.abs(g['force']).max(axis=1, keepdims=True), 1e-6)
    assert (np.abs(force.numpy() - g['force']) / scale).max() < 1e-8
    assert max_rel_to_scale(P.inertial_parameters.grad.numpy(), g['grad_theta']) < 1e-9
    assert max_rel_to_scale(P.friction_params.grad.numpy(), g['grad_friction']) < 1e-9
    gl = np.stack([p.grad.numpy().reshape(3) for p in P.length_params])
    assert max_rel_to_scale(gl, g['grad_length']) < 1e-9
    with torch.no_grad():
        M, J, phi, acc = co.multibody_terms(ELBOW_CALLS, P, xp[:, :8], xp[:, 8:])
        assert np.abs(M.numpy() - g['terms_M']).max() < 1e-15
        assert np.abs(acc.numpy() - g['terms_acc']).max() < 1e-10 * max(1.0, np.abs(g['terms_acc']).max())
        steps = g['sim_traj'].shape[1] - 1
        traj = co.simulate(ELBOW_CALLS, P, torch.from_numpy(g['sim_x0']), float(g['dt']), steps).numpy()
    assert np.abs(traj[:, 1] - g['sim_traj'][:, 1]).max() < 1e-9
    assert np.abs(traj - g['sim_traj']).max() < 1e-6


@pytest.mark.parametrize('tree,n_q,seed', [(CUBE_TREE, 7, 1), (ELBOW_TREE, 8, 2), (TREE4_TREE, 10, 3), (TREE6_TREE, 12, 4),
                                           (SLIDER3_TREE, 9, 5)])
def test_restated_callables_satisfy_the_power_balance(tree, n_q, seed):
    """Physics certificate for the un-vendored symbolic callables (parity unpinned): along the
    contact-free flow q' = q (+) v dt the generated M(q) and F(q, v) must conserve energy,
        d/dt (1/2 v^T M v + V) = v^T F + 1/2 v^T Mdot v + Vdot = 0,
    which ties the Coriolis/centrifugal and gravity terms of F to M independently of how either was
    derived (Mdot, Vdot by central differences along the flow)."""
    from dair_pll_b200 import synthetic
    calls = TreeCallables(tree)
    if tree is CUBE_TREE:
        pi, _, _ = synthetic.cube_learnables_perturbed(0)
        x = synthetic.cube_states(64, seed=seed)
    elif tree is ELBOW_TREE:
        pi, _, _ = synthetic.elbow_learnables_perturbed(0)
        x = synthetic.elbow_states(64, seed=seed)
    else:
        # branching trees and the sliding joint: random inertias (off-centre, non-diagonal) and random states
        g = torch.Generator().manual_seed(seed)
        nb = tree.n_bodies
        m = 0.3 + 0.1 * torch.rand(nb, 1, generator=g, dtype=torch.float64)
        c = 0.02 * (2 * torch.rand(nb, 3, generator=g, dtype=torch.float64) - 1)
        diag = 6e-4 * (1 + 0.3 * (2 * torch.rand(nb, 3, generator=g, dtype=torch.float64) - 1))
        offd = 5e-5 * (2 * torch.rand(nb, 3, generator=g, dtype=torch.float64) - 1)
        pi = torch.cat((m, m * c, diag, offd), -1)
        quat = torch.randn(64, 4, generator=g, dtype=torch.float64)
        quat = quat / quat.norm(dim=-1, keepdim=True)
        pos = torch.rand(64, 3, generator=g, dtype=torch.float64)
        joints = 0.8 * (2 * torch.rand(64, nb - 1, generator=g, dtype=torch.float64) - 1)
        vel = torch.cat((4 * torch.randn(64, 3, generator=g, dtype=torch.float64), torch.randn(64, 3, generator=g, dtype=torch.float64),
                         3 * torch.randn(64, nb - 1, generator=g, dtype=torch.float64)), -1)
        x = torch.cat((quat, pos, joints, vel), -1)
    inertia = co.theta_to_inertia_vector(co.pi_cm_to_theta(pi))
    q, v = x[:, :n_q], x[:, n_q:]
    ine = inertia.expand(q.shape[:-1] + inertia.shape)

    def flow(h):
        quat = co.quat_mul(q[:, :4], co.quat_exp(v[:, :3] * h))
        return torch.cat((quat, q[:, 4:] + v[:, 3:] * h), -1)

    def potential(qq):
        R, o, *_ = calls.kinematics(qq)
        return sum(inertia[i, 0] * 9.81 * (o[i][:, 2] + (R[i] @ inertia[i, 1:4])[:, 2]) for i in range(tree.n_bodies))

    h = 1e-6
    M = calls.mass_matrix(q, ine)
    F = calls.lagrangian_forces(q, v, None, ine)
    assert torch.allclose(M, M.transpose(-1, -2), atol=1e-15)
    assert (torch.linalg.eigvalsh(M) > 0).all()
    Mdot = (calls.mass_matrix(flow(h), ine) - calls.mass_matrix(flow(-h), ine)) / (2 * h)
    Vdot = (potential(flow(h)) - potential(flow(-h))) / (2 * h)
    power = (v * F).sum(-1) + 0.5 * (v * (Mdot @ v[..., None])[..., 0]).sum(-1) + Vdot
    assert power.abs().max().item() < 1e-7 * (v * F).sum(-1).abs().max().item()


def _mesh_oracle_params(g, requires_grad=True):
    nets = []
    for gi in range(2):
        nets.append({k: torch.from_numpy(g[f'net{gi}_{k}']).clone() for k in ('Wd0', 'Wd1', 'Wh', 'wout', 'perturbations')})
    P = co.OracleParams(torch.from_numpy(g['theta']).clone(), torch.from_numpy(g['friction_params']).clone(), [], nets)
    if requires_grad:
        P.requires_grad_()
    return P


def test_elbow_with_learned_geometry_matches_reference_python():
    """contactnets_elbow_mesh: DeepSupportConvex + HomogeneousICNN witness points (reference code ran the
    networks); loss, inertia/friction gradients and the gradients of all four weight tensors of both nets."""
    g = load_golden('elbow_mesh_w64')
    P = _mesh_oracle_params(g)
    x, xp = torch.from_numpy(g['x']), torch.from_numpy(g['x_plus'])
    loss = co.contactnets_loss(ELBOW_CALLS, P, x, xp, float(g['dt']))
    loss.mean().backward()
    assert np.abs(loss.detach().numpy() - g['loss']).max() < 1e-13
    assert max_rel_to_scale(P.inertial_parameters.grad.numpy(), g['grad_theta']) < 1e-9
    assert max_rel_to_scale(P.friction_params.grad.numpy(), g['grad_friction']) < 1e-9
    for gi in range(2):
        for k in ('Wd0', 'Wd1', 'Wh', 'wout'):
            assert max_rel_to_scale(P.icnn[gi][k].grad.numpy(), g[f'net{gi}_grad_{k}']) < 1e-9, (gi, k)
    with torch.no_grad():
        traj = co.simulate(ELBOW_CALLS, P, torch.from_numpy(g['sim_x0']), float(g['dt']), g['sim_traj'].shape[1] - 1)
    assert np.abs(traj.numpy()[:, 1] - g['sim_traj'][:, 1]).max() < 1e-9


@pytest.mark.parametrize('tree,seed', [(ELBOW_TREE, 2), (TREE4_TREE, 3), (TREE6_TREE, 4), (SLIDER3_TREE, 5)])
def test_restated_geometry_jacobians_equal_the_derivative_of_the_kinematics(tree, seed):
    """Second certificate for the un-vendored callables (parity unpinned): along the flow q' = q (+) v dt every collision
    frame's origin moves with J_v v and turns with J_w v -- d/dt p_WG = J_v v, d/dt R_WG = S(J_w v) R_WG -- central
    differences of geometry_translations / geometry_rotations against geometry_spatial_jacobians (multibody_terms.py:
    299-310), for hinges, branching trees and the sliding joint."""
    calls = TreeCallables(tree)
    g = torch.Generator().manual_seed(seed)
    nb = tree.n_bodies
    quat = torch.randn(32, 4, generator=g, dtype=torch.float64)
    quat = quat / quat.norm(dim=-1, keepdim=True)
    q = torch.cat((quat, torch.rand(32, 3, generator=g, dtype=torch.float64),
                   0.8 * (2 * torch.rand(32, nb - 1, generator=g, dtype=torch.float64) - 1)), -1)
    v = torch.randn(32, 6 + nb - 1, generator=g, dtype=torch.float64)

    def flow(h):
        return torch.cat((co.quat_mul(q[:, :4], co.quat_exp(v[:, :3] * h)), q[:, 4:] + v[:, 3:] * h), -1)
    h = 1e-6
    J = calls.geometry_spatial_jacobians(q)                       # (32, n_g, 6, n_v)
    twist = (J @ v[:, None, :, None])[..., 0]                      # (32, n_g, 6) = [omega_W ; v_W]
    pdot = (calls.geometry_translations(flow(h)) - calls.geometry_translations(flow(-h))) / (2 * h)
    assert (pdot - twist[..., 3:]).abs().max().item() < 1e-8
    R = calls.geometry_rotations(q)
    Rdot = (calls.geometry_rotations(flow(h)) - calls.geometry_rotations(flow(-h))) / (2 * h)
    W = Rdot @ R.transpose(-1, -2)                                 # = S(omega_W)
    omega = torch.stack((W[..., 2, 1], W[..., 0, 2], W[..., 1, 0]), -1)
    assert (omega - twist[..., :3]).abs().max().item() < 1e-8
