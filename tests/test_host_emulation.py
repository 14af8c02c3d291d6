"""CPU: the per-sample DEVICE math (dair_pll_b200/csrc/*.cuh), compiled for the host by g++,
against the golden vectors from the reference's Python and against the oracle.  This checks the
arithmetic the kernels execute (cone projection, Newton solve, loss, hand-derived envelope
backward, time step) in a container without a GPU; the GPU tests then check the kernels
themselves through the C ABI."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import contactnets_oracle as co
from dair_pll_b200.geometry import place_in_link_frame
from oracle.callables import CUBE_TREE, TreeCallables
from tests.util import dptr, host_emulation_lib, kernel_level_params, load_golden, max_rel_to_scale, rel_err

CASES = ['cube_real_nominal', 'cube_real_perturbed', 'cube_synthetic']


def emul_loss(g, dtype=np.float64):
    lib = host_emulation_lib()
    inertia, mu, half = kernel_level_params(g)
    x = np.ascontiguousarray(g['x'], dtype=dtype)
    xp = np.ascontiguousarray(g['x_plus'], dtype=dtype)
    B = x.shape[0]
    loss, force = np.zeros(B, dtype), np.zeros((B, 12), dtype)
    iters, grad = np.zeros(B, np.int32), np.zeros(14, dtype)
    ct = ctypes.c_double if dtype == np.float64 else ctypes.c_float
    fn = lib.emul_cube_loss_f64 if dtype == np.float64 else lib.emul_cube_loss_f32
    fn(dptr(x), dptr(xp), dptr(inertia.astype(dtype)), dptr(mu.astype(dtype)), dptr(half.astype(dtype)),
       ct(float(g['dt'])), ct(1e-3), ctypes.c_int64(B), dptr(loss), dptr(force), dptr(iters), dptr(grad))
    return loss, force, iters, grad


def chain_to_leaves(g, grad14):
    """Push callable-level gradients through theta->inertia, |.|, mu-combination with autograd."""
    theta = torch.from_numpy(g['theta']).clone().requires_grad_()
    fr = torch.from_numpy(g['friction_params']).clone().requires_grad_()
    ln = torch.from_numpy(g['half_lengths']).clone().reshape(1, 3).requires_grad_()
    inertia = co.theta_to_inertia_vector(theta).reshape(10)
    mu = fr.abs()
    mu_pair = 2 * mu[0] * mu[1] / (mu[0] + mu[1])
    half = ln.abs().reshape(3)
    flat = torch.cat((inertia, mu_pair.reshape(1), half))
    flat.backward(torch.from_numpy(np.asarray(grad14, dtype=np.float64)))
    return theta.grad.numpy(), fr.grad.numpy(), ln.grad.numpy()


@pytest.mark.parametrize('name', CASES)
def test_device_math_fp64_matches_reference_golden(name):
    g = load_golden(name)
    loss, force, iters, grad = emul_loss(g)
    B = loss.shape[0]
    assert np.abs(loss - g['loss']).max() < 1e-13
    assert rel_err(loss, g['loss'], 1e-9).max() < 1e-9            # north_star: 1e-9 relative, fp64
    scale = np.maximum(np.abs(g['force']).max(axis=1, keepdims=True), 1e-6)
    assert (np.abs(force - g['force']) / scale).max() < 1e-8
    gt, gf, gl = chain_to_leaves(g, grad / B)                      # golden grads are of loss.mean()
    assert max_rel_to_scale(gt, g['grad_theta']) < 1e-9
    assert max_rel_to_scale(gf, g['grad_friction']) < 1e-9
    assert max_rel_to_scale(gl, g['grad_length']) < 1e-9
    assert iters.max() <= 60


@pytest.mark.parametrize('name', CASES)
def test_device_math_fp32_variant(name):
    """fp32 variant (fp32 storage, fp64 arithmetic): 1e-4 on losses and gradients (north_star).  The
    comparison is against the fp64 reference results on the unrounded inputs, so it includes the
    effect of rounding states and parameters to fp32."""
    g = load_golden(name)
    loss, _, _, grad = emul_loss(g, np.float32)
    B = loss.shape[0]
    assert np.abs(loss - g['loss']).max() < 1e-4 * max(np.abs(g['loss']).max(), 1e-3)
    assert abs(loss.mean() - g['loss'].mean()) < 1e-4 * abs(g['loss'].mean())
    gt, gf, gl = chain_to_leaves(g, grad.astype(np.float64) / B)
    assert max_rel_to_scale(gt, g['grad_theta']) < 1e-4
    assert max_rel_to_scale(gf, g['grad_friction']) < 1e-4
    assert max_rel_to_scale(gl, g['grad_length']) < 1e-4


@pytest.mark.parametrize('name', CASES)
def test_device_step_matches_reference_golden(name):
    g = load_golden(name)
    lib = host_emulation_lib()
    inertia, mu, half = kernel_level_params(g)
    x0 = np.ascontiguousarray(g['sim_x0'])
    B = x0.shape[0]
    xn, iters = np.zeros((B, 13)), np.zeros(B, np.int32)
    lib.emul_cube_step_f64(dptr(x0), dptr(inertia), dptr(mu), dptr(half), ctypes.c_double(float(g['dt'])),
                           ctypes.c_double(1e-4), ctypes.c_int64(B), dptr(xn), None, dptr(iters))
    assert np.abs(xn - g['sim_traj'][:, 1]).max() < 1e-9


def test_device_math_matches_oracle_on_random_states():
    from dair_pll_b200 import synthetic
    pi, fr, half = synthetic.cube_learnables_perturbed(3)
    P = co.OracleParams(co.pi_cm_to_theta(pi), fr, [half.reshape(1, 3)])
    calls = TreeCallables(CUBE_TREE)
    x = synthetic.cube_states(512, seed=11)
    xp = synthetic.perturb_next_state(co.sim_step(calls, P, x, 0.0068), seed=12)
    loss_o = co.contactnets_loss(calls, P, x, xp, 0.0068).numpy()
    g = dict(x=x.numpy(), x_plus=xp.numpy(), theta=P.inertial_parameters.numpy(), friction_params=fr.numpy(),
             half_lengths=half.numpy(), dt=0.0068)
    loss, _, _, _ = emul_loss(g)
    assert np.abs(loss - loss_o).max() < 1e-12
    assert rel_err(loss, loss_o, 1e-9).max() < 1e-9


def test_device_parameter_preparation_and_chain_rule():
    """cn_params.cuh: theta -> [m, c, I_cm/m] and its dual-number chain rule vs the oracle + autograd."""
    lib = host_emulation_lib()
    torch.manual_seed(3)
    pi_cm = torch.tensor([[0.37, 0.37 * 0.002, -0.37 * 0.001, 0.37 * 0.003, 8.1e-4, 8.5e-4, 7.9e-4, 1e-5, -2e-5, 3e-5]],
                         dtype=torch.float64)
    theta = co.pi_cm_to_theta(pi_cm).clone().requires_grad_()
    vec = co.theta_to_inertia_vector(theta).reshape(10)
    g = torch.randn(10, dtype=torch.float64)
    vec.backward(g)
    th = theta.detach().numpy().reshape(10).copy()
    inertia, grad = np.zeros(10), np.zeros(10)
    lib.emul_theta_chain_f64(dptr(th), dptr(g.numpy().copy()), dptr(inertia), dptr(grad))
    assert np.allclose(inertia, vec.detach().numpy(), rtol=1e-13, atol=1e-18)
    assert np.allclose(grad, theta.grad.numpy().reshape(10), rtol=1e-11, atol=1e-16)


ELBOW_KIN = np.array([-0.035, 0.06, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.035, 0.0, 0.0])   # joint origin, axis, box offsets


def elbow_kernel_level_params(g):
    inertia = co.theta_to_inertia_vector(torch.from_numpy(g['theta'])).reshape(20).numpy()
    mu = np.abs(g['friction_params'])
    mu_pair = np.array([2 * mu[2] * mu[0] / (mu[2] + mu[0]), 2 * mu[2] * mu[1] / (mu[2] + mu[1])])
    return inertia, mu_pair, np.abs(g['half_lengths']).reshape(6).copy()


def elbow_chain_to_leaves(g, grad28):
    theta = torch.from_numpy(g['theta']).clone().requires_grad_()
    fr = torch.from_numpy(g['friction_params']).clone().requires_grad_()
    ln = torch.from_numpy(g['half_lengths']).clone().requires_grad_()
    mu = fr.abs()
    flat = torch.cat((co.theta_to_inertia_vector(theta).reshape(20),
                      (2 * mu[2] * mu[0] / (mu[2] + mu[0])).reshape(1), (2 * mu[2] * mu[1] / (mu[2] + mu[1])).reshape(1),
                      ln.abs().reshape(6)))
    flat.backward(torch.from_numpy(np.asarray(grad28, dtype=np.float64)))
    return theta.grad.numpy(), fr.grad.numpy(), ln.grad.numpy()


@pytest.mark.parametrize('name', ['elbow_nominal', 'elbow_perturbed'])
def test_elbow_device_math_matches_reference_golden(name):
    g = load_golden(name)
    lib = host_emulation_lib()
    inertia, mu, half = elbow_kernel_level_params(g)
    x, xp = np.ascontiguousarray(g['x']), np.ascontiguousarray(g['x_plus'])
    B = x.shape[0]
    loss, force, iters, grad = np.zeros(B), np.zeros((B, 24)), np.zeros(B, np.int32), np.zeros(28)
    lib.emul_elbow_loss_f64(dptr(x), dptr(xp), dptr(inertia), dptr(mu), dptr(half), dptr(ELBOW_KIN),
                            ctypes.c_double(float(g['dt'])), ctypes.c_double(1e-3), ctypes.c_int64(B), dptr(loss),
                            dptr(force), dptr(iters), dptr(grad))
    assert np.abs(loss - g['loss']).max() < 1e-12
    assert rel_err(loss, g['loss'], 1e-9).max() < 1e-9
    scale = np.maximum(np.abs(g['force']).max(axis=1, keepdims=True), 1e-6)
    assert (np.abs(force - g['force']) / scale).max() < 1e-7
    gt, gf, gl = elbow_chain_to_leaves(g, grad / B)
    assert max_rel_to_scale(gt, g['grad_theta']) < 1e-9
    assert max_rel_to_scale(gf, g['grad_friction']) < 1e-9
    assert max_rel_to_scale(gl, g['grad_length']) < 1e-9
    x0 = np.ascontiguousarray(g['sim_x0'])
    xn = np.zeros_like(x0)
    lib.emul_elbow_step_f64(dptr(x0), dptr(inertia), dptr(mu), dptr(half), dptr(ELBOW_KIN),
                            ctypes.c_double(float(g['dt'])), ctypes.c_double(1e-4), ctypes.c_int64(x0.shape[0]),
                            dptr(xn), None, None)
    assert np.abs(xn - g['sim_traj'][:, 1]).max() < 1e-9


def test_step_tangents_match_oracle_autograd():
    """Backward of the learnable time step (K7): dual-number tangents of a 3-step rollout, contracted
    with a random upstream gradient, against autograd through the oracle (implicit differentiation of
    the QP as sappy's backward would provide, multibody_learnable_system.py:293-304)."""
    lib = host_emulation_lib()
    g = load_golden('cube_real_perturbed')
    inertia, mu, half = kernel_level_params(g)
    # states near / in contact so the QP is active
    x0 = np.ascontiguousarray(g['sim_x0'][:24])
    B, steps = x0.shape[0], 3
    rng = np.random.default_rng(0)
    xbar = rng.standard_normal((B, steps, 13))
    gparams, gx0 = np.zeros((B, 14)), np.zeros((B, 13))
    lib.emul_cube_rollout_grad_f64(dptr(x0), dptr(inertia), dptr(mu), dptr(half), ctypes.c_double(float(g['dt'])),
                                   ctypes.c_double(1e-4), ctypes.c_int64(B), ctypes.c_int(steps), dptr(xbar),
                                   dptr(gparams), dptr(gx0))
    # oracle: autograd through the same rollout, gradients w.r.t. the callable-level parameters and x0
    calls = TreeCallables(CUBE_TREE)
    inertia_t = torch.from_numpy(inertia.copy()).reshape(1, 10).requires_grad_()
    mu_t = torch.from_numpy(mu.copy()).requires_grad_()
    half_t = torch.from_numpy(half.copy()).requires_grad_()
    x0_t = torch.from_numpy(x0.copy()).requires_grad_()
    orig = co.theta_to_inertia_vector
    co.theta_to_inertia_vector = lambda th: inertia_t
    try:
        P = co.OracleParams(torch.zeros(1, 10, dtype=torch.float64), torch.stack((mu_t[0], mu_t[0])), [half_t.reshape(1, 3)])
        tr = co.simulate(calls, P, x0_t, float(g['dt']), steps)
    finally:
        co.theta_to_inertia_vector = orig
    (tr[:, 1:] * torch.from_numpy(xbar)).sum().backward()
    ref_params = np.concatenate((inertia_t.grad.numpy().reshape(10), mu_t.grad.numpy(), half_t.grad.numpy()))
    assert max_rel_to_scale(gparams.sum(0), ref_params) < 1e-9       # north_star: parameter gradients at 1e-9
    assert max_rel_to_scale(gx0, x0_t.grad.numpy()) < 1e-9


@pytest.mark.parametrize('name', CASES)
def test_device_dense_terms_match_reference_golden(name):
    """MultibodyTerms.forward outputs (M, M^-1 F, phi from the reference's code; J, Delassus from the oracle)."""
    g = load_golden(name)
    lib = host_emulation_lib()
    inertia, mu, half = kernel_level_params(g)
    xp = g['x_plus']
    B = xp.shape[0]
    q, v = np.ascontiguousarray(xp[:, :7]), np.ascontiguousarray(xp[:, 7:])
    M, J, phi = np.zeros((B, 6, 6)), np.zeros((B, 12, 6)), np.zeros((B, 4))
    acc, D = np.zeros((B, 6)), np.zeros((B, 12, 12))
    lib.emul_cube_terms_f64(dptr(q), dptr(v), dptr(inertia), dptr(mu), dptr(half), ctypes.c_int64(B), dptr(M),
                            dptr(J), dptr(phi), dptr(acc), dptr(D))
    assert np.abs(M - g['terms_M']).max() < 1e-15
    assert np.abs(acc - g['terms_acc']).max() < 1e-10 * max(1.0, np.abs(g['terms_acc']).max())
    assert np.abs(np.sort(phi, -1) - g['terms_phi_sorted']).max() < 1e-15
    P = co.OracleParams(torch.from_numpy(g['theta']), torch.from_numpy(g['friction_params']),
                        [torch.from_numpy(g['half_lengths']).reshape(1, 3)])
    with torch.no_grad():
        Mo, Jo, phio, _ = co.multibody_terms(TreeCallables(CUBE_TREE), P, torch.from_numpy(q), torch.from_numpy(v))
        Do = Jo @ torch.linalg.solve(Mo, Jo.transpose(-1, -2))
    assert np.abs(J - Jo.numpy()).max() < 1e-14
    assert np.abs(phi - phio.numpy()).max() < 1e-15
    assert np.abs(D - Do.numpy()).max() < 1e-10 * np.abs(Do.numpy()).max()


def test_elbow_step_tangents_match_oracle_autograd():
    """Backward of the learnable time step of the two-body system: dual-number tangents of a 2-step rollout
    (43 directions), chained to the leaf parameters, against autograd through the oracle."""
    from oracle.callables import ELBOW_TREE
    from tests.util import oracle_params_from_golden
    lib = host_emulation_lib()
    g = load_golden('elbow_perturbed')
    inertia, mu, half = elbow_kernel_level_params(g)
    x0 = np.ascontiguousarray(g['sim_x0'][:12])
    B, steps = x0.shape[0], 2
    rng = np.random.default_rng(1)
    xbar = rng.standard_normal((B, steps, 15))
    gparams, gx0 = np.zeros((B, 28)), np.zeros((B, 15))
    lib.emul_elbow_rollout_grad_f64(dptr(x0), dptr(inertia), dptr(mu), dptr(half), dptr(ELBOW_KIN),
                                    ctypes.c_double(float(g['dt'])), ctypes.c_double(1e-4), ctypes.c_int64(B),
                                    ctypes.c_int(steps), dptr(xbar), dptr(gparams), dptr(gx0))
    P = oracle_params_from_golden(g)
    x0_t = torch.from_numpy(x0.copy()).requires_grad_()
    tr = co.simulate(TreeCallables(ELBOW_TREE), P, x0_t, float(g['dt']), steps)
    (tr[:, 1:] * torch.from_numpy(xbar)).sum().backward()
    gt, gf, gl = elbow_chain_to_leaves(g, gparams.sum(0))
    assert max_rel_to_scale(gx0, x0_t.grad.numpy()) < 1e-9
    assert max_rel_to_scale(gt, P.inertial_parameters.grad.numpy()) < 1e-9
    assert max_rel_to_scale(gf, P.friction_params.grad.numpy()) < 1e-9
    assert max_rel_to_scale(gl, np.stack([p.grad.numpy().reshape(3) for p in P.length_params])) < 1e-9


def test_elbow_closed_form_step_equals_the_dense_step_and_the_reference_golden():
    """The rollout kernels' step of the two-body system (elbow_step_sample_wf: closed-form mass terms and inverse, packed
    Hessian) against the reference-code golden (one step, 1e-9) and against the dense formulation that the backward
    differentiates, over a multi-step rollout on random states (next states and contact forces)."""
    from dair_pll_b200 import synthetic
    lib = host_emulation_lib()
    g = load_golden('elbow_perturbed')
    inertia, mu, half = elbow_kernel_level_params(g)
    dt = ctypes.c_double(float(g['dt']))
    x0 = np.ascontiguousarray(g['sim_x0'])
    xn, us = np.zeros_like(x0), np.zeros((x0.shape[0], 7))
    lib.emul_elbow_step_wf_f64(dptr(x0), dptr(inertia), dptr(mu), dptr(half), dptr(ELBOW_KIN), None, dt, ctypes.c_double(1e-4),
                               ctypes.c_int64(x0.shape[0]), dptr(xn), None, dptr(us))
    assert np.abs(xn - g['sim_traj'][:, 1]).max() < 1e-9
    x = synthetic.elbow_states(600, seed=9).numpy().copy()
    worst_x = worst_f = 0.0
    active = 0
    for _ in range(4):
        a, b = np.zeros_like(x), np.zeros_like(x)
        fa, fb = np.zeros((x.shape[0], 24)), np.zeros((x.shape[0], 24))
        lib.emul_elbow_step_wf_f64(dptr(x), dptr(inertia), dptr(mu), dptr(half), dptr(ELBOW_KIN), None, dt,
                                   ctypes.c_double(1e-4), ctypes.c_int64(x.shape[0]), dptr(a), dptr(fa), None)
        lib.emul_elbow_step_f64(dptr(x), dptr(inertia), dptr(mu), dptr(half), dptr(ELBOW_KIN), dt, ctypes.c_double(1e-4),
                                ctypes.c_int64(x.shape[0]), dptr(b), dptr(fb), None)
        worst_x = max(worst_x, np.abs(a - b).max())
        worst_f = max(worst_f, np.abs(fa - fb).max() / max(np.abs(fb).max(), 1e-300))
        active += int((np.abs(fb).max(-1) > 0).sum())
        x = b
    assert active > 200                       # the comparison does exercise the solver
    assert worst_x < 1e-9 and worst_f < 1e-7, (worst_x, worst_f)


def test_elbow_tangents_with_kept_optima_equal_the_dual_number_solves():
    """dpll_elbow_rollout_grad_saved_f64's arithmetic: with every step's QP optimum kept by the forward rollout, a
    dual-number step is ONE evaluation + one 7x7 solve at the optimum; its tangents must equal those of the full
    dual-number Newton solves (the implicit-function derivative either way)."""
    lib = host_emulation_lib()
    g = load_golden('elbow_perturbed')
    inertia, mu, half = elbow_kernel_level_params(g)
    x0 = np.ascontiguousarray(g['sim_x0'][:10])
    B, steps = x0.shape[0], 4
    rng = np.random.default_rng(2)
    xbar = rng.standard_normal((B, steps, 15))
    out = []
    for fn in (lib.emul_elbow_rollout_grad_f64, lib.emul_elbow_rollout_grad_saved_f64):
        gparams, gx0 = np.zeros((B, 28)), np.zeros((B, 15))
        fn(dptr(x0), dptr(inertia), dptr(mu), dptr(half), dptr(ELBOW_KIN), ctypes.c_double(float(g['dt'])),
           ctypes.c_double(1e-4), ctypes.c_int64(B), ctypes.c_int(steps), dptr(xbar), dptr(gparams), dptr(gx0))
        out.append((gparams, gx0))
    assert np.abs(out[0][0]).max() > 0
    assert max_rel_to_scale(out[1][0], out[0][0]) < 1e-9
    assert max_rel_to_scale(out[1][1], out[0][1]) < 1e-9


def test_free_flight_fast_path_equals_the_generic_path():
    """The triage phase's register-only evaluation of free-flight samples: it fires exactly on the samples
    whose solve is trivial (0 iterations, zero forces) and gives the generic path's loss and gradient."""
    from dair_pll_b200 import synthetic
    lib = host_emulation_lib()
    g = load_golden('cube_synthetic')
    inertia, mu, half = kernel_level_params(g)
    x = synthetic.cube_states(4000, seed=3).numpy().copy()
    xp = np.zeros_like(x)
    lib.emul_cube_step_f64(dptr(x), dptr(inertia), dptr(mu), dptr(half), ctypes.c_double(float(g['dt'])),
                           ctypes.c_double(1e-4), ctypes.c_int64(x.shape[0]), dptr(xp), None, None)
    xp = synthetic.perturb_next_state(torch.from_numpy(xp), seed=4).numpy().copy()
    xp[:50, 6] -= 0.2                                     # some deep penetrations: flight test must still be exact
    B = x.shape[0]
    loss_f, flags, grad_f = np.zeros(B), np.zeros(B, np.int32), np.zeros(14)
    lib.emul_cube_loss_free_f64(dptr(x), dptr(xp), dptr(inertia), dptr(mu), dptr(half), ctypes.c_double(float(g['dt'])),
                                ctypes.c_double(1e-3), ctypes.c_int64(B), dptr(loss_f), dptr(flags), dptr(grad_f))
    free = flags == 1
    assert 0.3 < free.mean() < 0.95
    # generic path on the same samples
    loss, force, iters, grad = np.zeros(B), np.zeros((B, 12)), np.zeros(B, np.int32), np.zeros(14)
    lib.emul_cube_loss_f64(dptr(x), dptr(xp), dptr(inertia), dptr(mu), dptr(half), ctypes.c_double(float(g['dt'])),
                           ctypes.c_double(1e-3), ctypes.c_int64(B), dptr(loss), dptr(force), dptr(iters), dptr(grad))
    assert np.array_equal(free, (iters == 0) & (np.abs(force).max(1) == 0))
    assert np.abs(loss_f[free] - loss[free]).max() <= 1e-14 * np.abs(loss[free]).max()
    xs, xps = np.ascontiguousarray(x[free]), np.ascontiguousarray(xp[free])
    n = xs.shape[0]
    l2, f2, i2, grad_g = np.zeros(n), np.zeros((n, 12)), np.zeros(n, np.int32), np.zeros(14)
    lib.emul_cube_loss_f64(dptr(xs), dptr(xps), dptr(inertia), dptr(mu), dptr(half), ctypes.c_double(float(g['dt'])),
                           ctypes.c_double(1e-3), ctypes.c_int64(n), dptr(l2), dptr(f2), dptr(i2), dptr(grad_g))
    assert np.abs(grad_f - grad_g).max() <= 1e-12 * np.abs(grad_g).max()
    assert np.abs(grad_g[11:]).max() > 0                  # the penetration term is exercised


@pytest.mark.parametrize('name', ['elbow_nominal', 'elbow_perturbed'])
def test_elbow_dense_terms_match_oracle(name):
    """MultibodyTerms.forward for the two-body system (dpll_elbow_terms_f64's per-sample code): M, J, phi, the
    contact-free acceleration and the Delassus operator against the oracle's tree code at the golden states."""
    from oracle.callables import ELBOW_TREE
    from tests.util import oracle_params_from_golden
    g = load_golden(name)
    lib = host_emulation_lib()
    inertia, mu, half = elbow_kernel_level_params(g)
    xp = g['x_plus']
    B = xp.shape[0]
    q, v = np.ascontiguousarray(xp[:, :8]), np.ascontiguousarray(xp[:, 8:])
    M, J, phi = np.zeros((B, 7, 7)), np.zeros((B, 24, 7)), np.zeros((B, 8))
    acc, D = np.zeros((B, 7)), np.zeros((B, 24, 24))
    lib.emul_elbow_terms_f64(dptr(q), dptr(v), dptr(inertia), dptr(mu), dptr(half), dptr(ELBOW_KIN), ctypes.c_int64(B),
                             dptr(M), dptr(J), dptr(phi), dptr(acc), dptr(D))
    P = oracle_params_from_golden(g, requires_grad=False)
    with torch.no_grad():
        Mo, Jo, phio, acco = co.multibody_terms(TreeCallables(ELBOW_TREE), P, torch.from_numpy(q), torch.from_numpy(v))
        Do = Jo @ torch.linalg.solve(Mo, Jo.transpose(-1, -2))
    assert np.abs(M - Mo.numpy()).max() < 1e-14 * max(1.0, np.abs(Mo.numpy()).max())
    assert np.abs(J - Jo.numpy()).max() < 1e-13
    assert np.abs(phi - phio.numpy()).max() < 1e-14
    assert np.abs(acc - acco.numpy()).max() < 1e-9 * max(1.0, np.abs(acco.numpy()).max())
    assert np.abs(D - Do.numpy()).max() < 1e-9 * np.abs(Do.numpy()).max()


@pytest.mark.parametrize('steps', [1, 12])
def test_reverse_mode_step_adjoint_matches_forward_mode_tangents(steps):
    """K7 (cn_cube_adjoint.cuh): the hand-derived reverse-mode backward of the learnable step -- implicit
    differentiation of the QP by one 6x6 Cholesky solve per step plus the adjoints of the step's other lines --
    against the dual-number tangents of the same step code (27 forward-mode rollouts per toss), on real tosses
    in and around contact: parameter and initial-state gradients to 1e-10."""
    lib = host_emulation_lib()
    g = load_golden('cube_real_perturbed')
    inertia, mu, half = kernel_level_params(g)
    x0 = np.ascontiguousarray(g['sim_x0'])
    n = x0.shape[0]
    xbar = np.random.default_rng(steps).standard_normal((n, steps, 13))
    out = []
    for fn in (lib.emul_cube_rollout_grad_f64, lib.emul_cube_rollout_backward_f64):
        gp, gx = np.zeros((n, 14)), np.zeros((n, 13))
        fn(dptr(x0), dptr(inertia), dptr(mu), dptr(half), ctypes.c_double(float(g['dt'])), ctypes.c_double(1e-4),
           ctypes.c_int64(n), ctypes.c_int(steps), dptr(xbar), dptr(gp), dptr(gx))
        out.append((gp, gx))
    (gp_f, gx_f), (gp_r, gx_r) = out
    assert np.abs(gp_f).max() > 0 and np.abs(gx_f).max() > 0
    for a, b in ((gp_r, gp_f), (gx_r, gx_f)):
        per_sample = np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-300)
        assert per_sample.max() < 1e-9                            # every toss on its own
    assert max_rel_to_scale(gp_r.sum(0), gp_f.sum(0)) < 1e-10     # the parameter gradient of the batch
    assert max_rel_to_scale(gx_r, gx_f) < 1e-10


def _shape_case(name):
    """golden + the product's geometry object and the witness points it yields on the CPU (torch ops)"""
    from dair_pll_b200.geometry import Box, Polygon, Sphere
    g = load_golden(name)
    p = torch.from_numpy(g['shape_param'])
    if name == 'shape_framed_box':
        # the Box sits in a collision frame offset and rotated in the link (oracle.callables.FRAMED_BODY_TREE)
        from oracle.callables import FRAMED_BODY_TREE as tree
        geom = Box(p.reshape(3), 4)
        geom.set_frame(torch.tensor(tree.geometry_offset[0], dtype=torch.float64), tree.geometry_rotation(0, torch.float64))
    else:
        geom = Sphere(p) if name == 'shape_sphere' else Polygon(p, 4)
    return g, geom


def _support_direction(quat):
    w, x, y, z = quat.unbind(-1)
    s = 2.0 / (w * w + x * x + y * y + z * z)
    return -torch.stack((s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)), -1)


@pytest.mark.parametrize('name', ['shape_sphere', 'shape_polygon', 'shape_framed_box'])
def test_witness_point_device_math_matches_reference_golden(name):
    """N4 (plane-convex contacts beyond the box): the witness-point loss / step code against goldens produced by the
    REFERENCE's own Sphere and Polygon classes (oracle/gen_golden_shapes.py): losses, gradients of theta, friction
    and the shape parameter (radius / vertices), and the one-step next state, at 1e-9."""
    lib = host_emulation_lib()
    g, geom = _shape_case(name)
    theta = torch.from_numpy(g['theta']).clone().requires_grad_()
    fr = torch.from_numpy(g['friction_params']).clone().requires_grad_()
    inertia_t = co.theta_to_inertia_vector(theta).reshape(10)
    m = fr.abs()
    mu_t = (2 * m[0] * m[1] / (m[0] + m[1])).reshape(1)
    x, xp = np.ascontiguousarray(g['x']), np.ascontiguousarray(g['x_plus'])
    B = x.shape[0]

    def witness(states):
        p = place_in_link_frame(geom, _support_direction(torch.from_numpy(states[:, :4])))
        n_c = p.shape[-2]
        full = torch.cat((p, p.new_zeros(B, 4 - n_c, 3)), -2) if n_c < 4 else p
        return p, full, n_c
    p, full, n_c = witness(xp)
    pts = np.ascontiguousarray(full.detach().numpy())
    loss, g11, gp = np.zeros(B), np.zeros(11), np.zeros((B, 4, 3))
    inertia, mu = inertia_t.detach().numpy().copy(), mu_t.detach().numpy().copy()
    lib.emul_body_loss_pts_f64(dptr(x), dptr(xp), dptr(inertia), dptr(mu), dptr(pts), ctypes.c_int(n_c),
                               ctypes.c_double(float(g['dt'])), ctypes.c_double(1e-3), ctypes.c_int64(B), dptr(loss),
                               dptr(g11), dptr(gp))
    assert np.abs(loss - g['loss']).max() < 1e-12
    assert rel_err(loss, g['loss'], 1e-9).max() < 1e-9
    # chain rule to the leaves (golden gradients are of loss.mean())
    torch.cat((inertia_t, mu_t)).backward(torch.from_numpy(g11 / B))
    p.backward(torch.from_numpy(gp[:, :n_c] / B))
    shape_leaf = {'shape_sphere': lambda: geom.length_param, 'shape_polygon': lambda: geom.vertices,
                  'shape_framed_box': lambda: geom.length_params}[name]()
    assert max_rel_to_scale(theta.grad.numpy(), g['grad_theta']) < 1e-9
    assert max_rel_to_scale(fr.grad.numpy(), g['grad_friction']) < 1e-9
    assert max_rel_to_scale(shape_leaf.grad.numpy(), g['grad_shape_param']) < 1e-9
    # one learnable time step
    _, full0, _ = witness(x)
    pts0 = np.ascontiguousarray(full0.detach().numpy())
    xn = np.zeros((B, 13))
    lib.emul_body_step_pts_f64(dptr(x), dptr(inertia), dptr(mu), dptr(pts0), ctypes.c_int(n_c),
                               ctypes.c_double(float(g['dt'])), ctypes.c_double(1e-4), ctypes.c_int64(B), dptr(xn))
    assert np.abs(xn - g['x_next']).max() < 1e-9


# ---- generic serial chain (cn_chain.cuh) -----------------------------------------------------------

def chain_kin_rows(tree):
    """(n, 31) kinematic table of dpll_chain_*.  Row b, link part: [joint origin | fixed rotation row-major | axis | . |
    parent | . | joint type (1 = prismatic)]; row b, box-slot part: [box offset (15:18) | rotation link <- collision frame
    row-major (19:28) | link the box sits on (29) | slot in use (30)] -- the tree's body geometries fill the first slots."""
    n = tree.n_bodies
    n_geoms = len(tree.geometry_body) - 1
    rows = []
    for b in range(n):
        Rfix = tree.joint_rotation(b, torch.float64).numpy().reshape(-1) if b > 0 else np.eye(3).reshape(-1)
        if b < n_geoms:
            off, Rg, link, used = tree.geometry_offset[b], tree.geometry_rotation(b, torch.float64).numpy().reshape(-1), \
                tree.geometry_body[b], 1.0
        else:
            off, Rg, link, used = (0., 0., 0.), np.eye(3).reshape(-1), 0, 0.0
        rows.append(np.concatenate((tree.joint_origin[b], Rfix, tree.axis[b], off, [float(max(tree.parent[b], 0))], Rg,
                                    [float(tree.is_prismatic(b))], [float(link)], [used])))
    return np.ascontiguousarray(np.stack(rows))


def chain_kernel_level_params(g, n):
    """inertia (10 n), pair friction and half lengths of the n box slots (empty slots: 1 and 0)."""
    inertia = co.theta_to_inertia_vector(torch.from_numpy(g['theta'])).reshape(10 * n).numpy()
    mu = np.abs(g['friction_params'])
    ng = len(mu) - 1                                        # body geometries; the ground is last
    mu_pair = np.array([2 * mu[ng] * mu[i] / (mu[ng] + mu[i]) for i in range(ng)] + [1.0] * (n - ng))
    half = np.zeros((n, 3))
    half[:ng] = np.abs(g['half_lengths'])
    return inertia, mu_pair, half.reshape(3 * n).copy()


def chain_grad_to_leaves(g, grad, n):
    theta = torch.from_numpy(g['theta']).clone().requires_grad_()
    fr = torch.from_numpy(g['friction_params']).clone().requires_grad_()
    ln = torch.from_numpy(g['half_lengths']).clone().requires_grad_()
    mu = fr.abs()
    ng = mu.shape[0] - 1
    pad_mu = [torch.ones((), dtype=torch.float64)] * (n - ng)
    flat = torch.cat((co.theta_to_inertia_vector(theta).reshape(10 * n),
                      torch.stack([2 * mu[ng] * mu[i] / (mu[ng] + mu[i]) for i in range(ng)] + pad_mu),
                      ln.abs().reshape(3 * ng), torch.zeros(3 * (n - ng), dtype=torch.float64)))
    flat.backward(torch.from_numpy(np.asarray(grad, dtype=np.float64)))
    return theta.grad.numpy(), fr.grad.numpy(), ln.grad.numpy()


def emul_chain_loss(n, g, kin, x, xp, eps=1e-3):
    lib = host_emulation_lib()
    inertia, mu, half = chain_kernel_level_params(g, n)
    B = x.shape[0]
    loss, force, iters, grad = np.zeros(B), np.zeros((B, 12 * n)), np.zeros(B, np.int32), np.zeros(14 * n)
    rc = lib.emul_chain_loss_f64(ctypes.c_int(n), dptr(x), dptr(xp), dptr(inertia), dptr(mu), dptr(half), dptr(kin),
                                 ctypes.c_double(float(g['dt'])), ctypes.c_double(eps), ctypes.c_int64(B), dptr(loss),
                                 dptr(force), dptr(iters), dptr(grad))
    assert rc == 0
    return loss, force, iters, grad


@pytest.mark.parametrize('name', ['chain3', 'chain3r', 'slider3', 'tree4', 'tree4g', 'tree6'])
def test_chain_and_tree_device_math_matches_reference_golden(name):
    """Three links in series with a rotated off-axis second joint (CHAIN3_TREE), the same with every box in a ROTATED
    collision frame (CHAIN3R_TREE), a hinge followed by a SLIDING joint (SLIDER3_TREE), a BRANCHING four-link tree (TREE4_TREE:
    two links off the root, a third off one of them), the same tree with its boxes spread UNEVENLY over the links (TREE4G_TREE:
    two on the root, none on two links) and a six-link tree (TREE6_TREE, the largest instantiation): loss,
    parameter gradients and one time step against the REFERENCE's
    own contactnets_loss / sim_step run on the oracle's tree callables (oracle/gen_golden_chain.py)."""
    from oracle.callables import CHAIN3_TREE, CHAIN3R_TREE, SLIDER3_TREE, TREE4_TREE, TREE4G_TREE, TREE6_TREE
    tree = {'chain3': CHAIN3_TREE, 'chain3r': CHAIN3R_TREE, 'slider3': SLIDER3_TREE, 'tree4': TREE4_TREE,
            'tree4g': TREE4G_TREE, 'tree6': TREE6_TREE}[name]
    n = tree.n_bodies
    g = load_golden(name)
    kin = chain_kin_rows(tree)
    x, xp = np.ascontiguousarray(g['x']), np.ascontiguousarray(g['x_plus'])
    B = x.shape[0]
    loss, _, iters, grad = emul_chain_loss(n, g, kin, x, xp)
    assert np.abs(loss - g['loss']).max() < 1e-12
    assert rel_err(loss, g['loss'], 1e-9).max() < 1e-9
    gt, gf, gl = chain_grad_to_leaves(g, grad / B, n)
    assert max_rel_to_scale(gt, g['grad_theta']) < 1e-9
    assert max_rel_to_scale(gf, g['grad_friction']) < 1e-9
    assert max_rel_to_scale(gl, g['grad_length']) < 1e-9
    assert iters.max() <= 60
    inertia, mu, half = chain_kernel_level_params(g, n)
    xn = np.zeros_like(x)
    rc = host_emulation_lib().emul_chain_step_f64(ctypes.c_int(n), dptr(x), dptr(inertia), dptr(mu), dptr(half), dptr(kin),
                                                  ctypes.c_double(float(g['dt'])), ctypes.c_double(1e-4),
                                                  ctypes.c_int64(B), dptr(xn))
    assert rc == 0
    assert np.abs(xn - g['x_next']).max() < 1e-9


@pytest.mark.parametrize('name', ['elbow_nominal', 'elbow_perturbed'])
def test_chain2_reproduces_the_elbow_goldens(name):
    """The hand-derived two-body kernels are the N = 2 instance of the generic recursion: same reference goldens."""
    from oracle.callables import ELBOW_TREE
    g = load_golden(name)
    kin = chain_kin_rows(ELBOW_TREE)
    x, xp = np.ascontiguousarray(g['x']), np.ascontiguousarray(g['x_plus'])
    B = x.shape[0]
    loss, force, _, grad = emul_chain_loss(2, g, kin, x, xp)
    assert np.abs(loss - g['loss']).max() < 1e-12
    assert rel_err(loss, g['loss'], 1e-9).max() < 1e-9
    scale = np.maximum(np.abs(g['force']).max(axis=1, keepdims=True), 1e-6)
    assert (np.abs(force - g['force']) / scale).max() < 1e-7
    gt, gf, gl = chain_grad_to_leaves(g, grad / B, 2)
    assert max_rel_to_scale(gt, g['grad_theta']) < 1e-9
    assert max_rel_to_scale(gf, g['grad_friction']) < 1e-9
    assert max_rel_to_scale(gl, g['grad_length']) < 1e-9
    inertia, mu, half = chain_kernel_level_params(g, 2)
    x0 = np.ascontiguousarray(g['sim_x0'])
    xn = np.zeros_like(x0)
    host_emulation_lib().emul_chain_step_f64(ctypes.c_int(2), dptr(x0), dptr(inertia), dptr(mu), dptr(half), dptr(kin),
                                             ctypes.c_double(float(g['dt'])), ctypes.c_double(1e-4),
                                             ctypes.c_int64(x0.shape[0]), dptr(xn))
    assert np.abs(xn - g['sim_traj'][:, 1]).max() < 1e-9


def test_tensor_core_digit_planes_are_exact_to_42_bits():
    """cn_icnn_tc.cuh: a weight column splits into TC_NS balanced base-128 int8 digits of a power-of-two column scale;
    the planes stand for the weight to 2^-(7 NS) of that scale, the stored digits fit int8 with the -128 pairing
    (|digit| <= 64), and the image offsets of all (k, plane, j, i) are a permutation of the image bytes."""
    lib = host_emulation_lib()
    lib.emul_tc_digits.restype = ctypes.c_int
    rng = np.random.default_rng(0)
    for scale in (1.0, 3.7e-5, 8.1e6):
        q = rng.standard_normal(256) * scale
        q[7] = 0.0
        q[11] = np.abs(q).max() * (1 - 1e-16)        # right at the column maximum
        dig = np.zeros(256 * 8, np.int8)
        rec, sigma = np.zeros(256), np.zeros(1)
        ns = lib.emul_tc_digits(dptr(q), 256, dptr(dig), dptr(rec), dptr(sigma))
        assert ns == 6
        s = sigma[0]
        assert np.abs(q).max() < s <= 2 * np.abs(q).max() and np.log2(s) == np.round(np.log2(s))
        assert np.abs(rec - q).max() <= 2.0 ** -(7 * ns) * s
        assert np.abs(dig[:256 * ns].astype(np.int32)).max() <= 64
    assert np.array_equal(lib.emul_tc_digits(dptr(np.zeros(4)), 4, dptr(dig), dptr(rec), dptr(sigma)) and rec[:4], np.zeros(4))
    lib.emul_tc_image_offset.restype = ctypes.c_int
    seen = np.zeros(lib.emul_tc_image_bytes(), np.int32)
    for k in range(3):
        for s_ in range(6):
            for j in (0, 1, 15, 16, 17, 255):
                for i in (0, 7, 8, 31, 32, 255):
                    seen[lib.emul_tc_image_offset(k, s_, j, i)] += 1
    assert seen.max() == 1 and seen.sum() == 3 * 6 * 36


@pytest.mark.parametrize('name', ['chain3r', 'slider3', 'tree4g'])
def test_tree_dense_terms_match_oracle(name):
    """MultibodyTerms.forward for the generic trees (dpll_chain_terms_f64's per-sample code): M, J, phi, the contact-free
    acceleration and the Delassus operator against the oracle's tree callables at the golden states -- rotated collision
    frames, the sliding joint, boxes spread unevenly over the links (fewer boxes than links)."""
    from oracle.callables import CHAIN3R_TREE, SLIDER3_TREE, TREE4G_TREE
    from tests.util import oracle_params_from_golden
    tree = {'chain3r': CHAIN3R_TREE, 'slider3': SLIDER3_TREE, 'tree4g': TREE4G_TREE}[name]
    n, n_boxes = tree.n_bodies, len(tree.geometry_body) - 1
    g = load_golden(name)
    lib = host_emulation_lib()
    inertia, mu, half = chain_kernel_level_params(g, n)
    kin = chain_kin_rows(tree)
    xp = g['x_plus']
    B, nq, nv, k = xp.shape[0], 7 + n - 1, 6 + n - 1, 12 * n_boxes
    q, v = np.ascontiguousarray(xp[:, :nq]), np.ascontiguousarray(xp[:, nq:])
    M, J, phi = np.zeros((B, nv, nv)), np.zeros((B, k, nv)), np.zeros((B, 4 * n_boxes))
    acc, D = np.zeros((B, nv)), np.zeros((B, k, k))
    rc = lib.emul_chain_terms_f64(ctypes.c_int(n), ctypes.c_int(n_boxes), dptr(q), dptr(v), dptr(inertia), dptr(mu), dptr(half),
                                  dptr(kin), ctypes.c_int64(B), dptr(M), dptr(J), dptr(phi), dptr(acc), dptr(D))
    assert rc == 0
    P = oracle_params_from_golden(g, requires_grad=False)
    with torch.no_grad():
        Mo, Jo, phio, acco = co.multibody_terms(TreeCallables(tree), P, torch.from_numpy(q), torch.from_numpy(v))
        Do = Jo @ torch.linalg.solve(Mo, Jo.transpose(-1, -2))
    assert np.abs(M - Mo.numpy()).max() < 1e-14 * max(1.0, np.abs(Mo.numpy()).max())
    assert np.abs(J - Jo.numpy()).max() < 1e-13
    assert np.abs(phi - phio.numpy()).max() < 1e-14
    assert np.abs(acc - acco.numpy()).max() < 1e-9 * max(1.0, np.abs(acco.numpy()).max())
    assert np.abs(D - Do.numpy()).max() < 1e-9 * np.abs(Do.numpy()).max()


def test_tree_witness_point_device_math_matches_reference_golden(assets_dir):
    """The witness-point form of the tree kernels (chain_loss_sample / chain_step_sample with pts): CHAIN3_TREE's kinematics
    with the reference's Box on link 0, Sphere on link 1 and Polygon on link 2 (oracle/gen_golden_chain.py:make_shapes) --
    the shapes' support points come from the product's own host code (MultibodyTerms.chain_witness_points: torch forward
    kinematics + the geometry classes), the loss, its gradients w.r.t. theta, friction, the box lengths, the sphere's radius
    and the polygon's vertices (chain rule through the points) and one time step from the device math, at 1e-9."""
    import os
    from dair_pll_b200.geometry import Polygon, Sphere
    from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem
    g = load_golden('chain3s')
    s = MultibodyLearnableSystem({'chain3': os.path.join(assets_dir, 'chain3.urdf')}, float(g['dt']))
    mt = s.multibody_terms
    ct = mt.contact_terms
    ct.geometries[1] = Sphere(torch.from_numpy(g['sphere_radius']))
    ct.geometries[2] = Polygon(torch.from_numpy(g['polygon_vertices']), 4)
    with torch.no_grad():
        mt.lagrangian_terms.inertial_parameters.copy_(torch.from_numpy(g['theta']))
        ct.friction_params.copy_(torch.from_numpy(g['friction_params']))
        ct.geometries[0].length_params.copy_(torch.from_numpy(g['box_length_params']))
    n = 3
    x, xp = np.ascontiguousarray(g['x']), np.ascontiguousarray(g['x_plus'])
    B = x.shape[0]
    pts_t, packed = mt.chain_witness_points(torch.from_numpy(xp[:, :9]))
    assert packed == (4 | (1 << 3) | (4 << 6)) and pts_t.shape == (B, 3, 4, 3)
    kin = np.ascontiguousarray(mt.chain_kinematic_table(torch.device('cpu'), witness=True).numpy())
    inertia_t = mt.lagrangian_terms.inertia_vector().reshape(10 * n)
    mu_t = ct.pair_friction().reshape(n)
    inertia, mu = inertia_t.detach().numpy().copy(), mu_t.detach().numpy().copy()
    pts = np.ascontiguousarray(pts_t.detach().numpy())
    lib = host_emulation_lib()
    loss, grad, gp = np.zeros(B), np.zeros(14 * n), np.zeros((B, n, 4, 3))
    rc = lib.emul_chain_loss_pts_f64(ctypes.c_int(n), dptr(x), dptr(xp), dptr(inertia), dptr(mu), dptr(kin), dptr(pts),
                                     ctypes.c_uint(packed), ctypes.c_double(float(g['dt'])), ctypes.c_double(1e-3),
                                     ctypes.c_int64(B), dptr(loss), dptr(grad), dptr(gp))
    assert rc == 0
    assert np.abs(loss - g['loss']).max() < 1e-12
    assert rel_err(loss, g['loss'], 1e-9).max() < 1e-9
    # chain rule to the leaves (golden gradients are of loss.mean())
    torch.cat((inertia_t, mu_t)).backward(torch.from_numpy(grad[:11 * n] / B))
    pts_t.backward(torch.from_numpy(gp / B))
    assert max_rel_to_scale(mt.lagrangian_terms.inertial_parameters.grad.numpy(), g['grad_theta']) < 1e-9
    assert max_rel_to_scale(ct.friction_params.grad.numpy(), g['grad_friction']) < 1e-9
    assert max_rel_to_scale(ct.geometries[0].length_params.grad.numpy(), g['grad_box_length_params']) < 1e-9
    assert max_rel_to_scale(ct.geometries[1].length_param.grad.numpy(), g['grad_sphere_radius']) < 1e-9
    assert max_rel_to_scale(ct.geometries[2].vertices.grad.numpy(), g['grad_polygon_vertices']) < 1e-9
    # one learnable time step
    pts0, packed0 = mt.chain_witness_points(torch.from_numpy(x[:, :9]))
    p0 = np.ascontiguousarray(pts0.detach().numpy())
    xn = np.zeros_like(x)
    rc = lib.emul_chain_step_pts_f64(ctypes.c_int(n), dptr(x), dptr(inertia), dptr(mu), dptr(kin), dptr(p0), ctypes.c_uint(packed0),
                                     ctypes.c_double(float(g['dt'])), ctypes.c_double(1e-4), ctypes.c_int64(B), dptr(xn))
    assert rc == 0
    assert np.abs(xn - g['x_next']).max() < 1e-9
