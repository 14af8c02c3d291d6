"""GPU: the CUDA path, called through the C ABI (dair_pll_b200.ops -> libdair_pll_b200.so) and the
reference-shaped module API, against (a) golden vectors from the reference's own Python,
(b) the CPU oracle on seeded random inputs, (c) size-independent properties at full batch size."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from dair_pll_b200 import ops, synthetic  # noqa: E402
from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem  # noqa: E402
from tests.util import kernel_level_params, load_golden, max_rel_to_scale, rel_err  # noqa: E402

CASES = ['cube_real_nominal', 'cube_real_perturbed', 'cube_synthetic']
DEV = 'cuda:0'


def _system(g, assets_dir, dtype=torch.float64):
    """Module with the golden file's learnable parameters loaded through state_dict()."""
    s = MultibodyLearnableSystem({'cube': os.path.join(assets_dir, 'cube.urdf')}, float(g['dt']))
    s.load_state_dict({
        'multibody_terms.lagrangian_terms.inertial_parameters': torch.from_numpy(g['theta']),
        'multibody_terms.contact_terms.friction_params': torch.from_numpy(g['friction_params']),
        'multibody_terms.contact_terms.geometries.0.length_params': torch.from_numpy(g['half_lengths']).reshape(1, 3)})
    return s.to(DEV)


def test_native_library_is_loaded():
    from dair_pll_b200 import _lib
    _lib.load()
    assert any('libdair_pll_b200.so' in line for line in open('/proc/self/maps'))


@pytest.mark.parametrize('name', CASES)
def test_loss_and_gradients_match_reference_golden_fp64(name, assets_dir):
    g = load_golden(name)
    s = _system(g, assets_dir)
    x = torch.from_numpy(g['x']).to(DEV)
    xp = torch.from_numpy(g['x_plus']).to(DEV)
    loss = s.contactnets_loss(x, torch.zeros(x.shape[0], 0, device=DEV), xp)
    assert loss.shape == (x.shape[0],)
    loss.mean().backward()
    l = loss.detach().cpu().numpy()
    assert np.abs(l - g['loss']).max() < 1e-13
    assert rel_err(l, g['loss'], 1e-9).max() < 1e-9          # north_star tolerance (fp64)
    mt = s.multibody_terms
    assert max_rel_to_scale(mt.lagrangian_terms.inertial_parameters.grad.cpu().numpy(), g['grad_theta']) < 1e-9
    assert max_rel_to_scale(mt.contact_terms.friction_params.grad.cpu().numpy(), g['grad_friction']) < 1e-9
    assert max_rel_to_scale(mt.contact_terms.geometries[0].length_params.grad.cpu().numpy(), g['grad_length']) < 1e-9


@pytest.mark.parametrize('name', CASES)
def test_impulses_match_reference_golden(name):
    g = load_golden(name)
    inertia, mu, half = (torch.from_numpy(a).to(DEV) for a in kernel_level_params(g))
    x, xp = torch.from_numpy(g['x']).to(DEV), torch.from_numpy(g['x_plus']).to(DEV)
    _, _, _, force, iters = ops.cube_loss_raw(x, xp, inertia, mu, half, float(g['dt']), 1e-3, want_force=True,
                                              want_iters=True)
    scale = np.maximum(np.abs(g['force']).max(axis=1, keepdims=True), 1e-6)
    assert (np.abs(force.cpu().numpy() - g['force']) / scale).max() < 1e-8
    assert int(iters.max()) <= 60


@pytest.mark.parametrize('name', CASES)
def test_rollout_matches_reference_golden(name, assets_dir):
    g = load_golden(name)
    s = _system(g, assets_dir)
    x0 = torch.from_numpy(g['sim_x0']).to(DEV)
    steps = g['sim_traj'].shape[1] - 1
    with torch.no_grad():
        traj, carry = s.simulate(x0.unsqueeze(-2), torch.zeros(x0.shape[0], 1, device=DEV), steps)
    assert traj.shape == g['sim_traj'].shape and carry.shape == (x0.shape[0], steps + 1, 1)
    t = traj.cpu().numpy()
    assert np.abs(t[:, 1] - g['sim_traj'][:, 1]).max() < 1e-9     # one step: next states at 1e-9
    assert np.abs(t - g['sim_traj']).max() < 1e-7                 # compounding contact sensitivity
    # forward_dynamics / integrator.step agree with the rollout's first step
    v_next = s.forward_dynamics(x0[:, :7], x0[:, 7:], torch.zeros(x0.shape[0], 0, device=DEV))
    assert torch.equal(v_next, traj[:, 1, 7:])
    x1, _ = s.integrator.step(x0, torch.zeros(x0.shape[0], 1, device=DEV))
    assert torch.allclose(x1, traj[:, 1], atol=1e-14, rtol=0)


@pytest.mark.parametrize('n', [2048, 150000])        # a partly filled pool per warp / every warp in steady state
def test_matches_cpu_oracle_on_random_inputs(n):
    from oracle import contactnets_oracle as co
    from oracle.callables import CUBE_TREE, TreeCallables
    calls = TreeCallables(CUBE_TREE)
    pi, fr, half = synthetic.cube_learnables_perturbed(5)
    P = co.OracleParams(co.pi_cm_to_theta(pi), fr, [half.reshape(1, 3)]).requires_grad_()
    x = synthetic.cube_states(n, seed=21)
    with torch.no_grad():
        xp = synthetic.perturb_next_state(co.sim_step(calls, P, x, 0.0068), seed=22)
    loss_o = co.contactnets_loss(calls, P, x, xp, 0.0068)
    loss_o.sum().backward()
    g = dict(theta=P.inertial_parameters.detach().numpy(), friction_params=fr.detach().numpy(),
             half_lengths=half.detach().numpy())
    inertia, mu, hl = (torch.from_numpy(a).to(DEV) for a in kernel_level_params(g))
    loss, grad, loss_sum, _, _ = ops.cube_loss_raw(x.to(DEV), xp.to(DEV), inertia, mu, hl, 0.0068, 1e-3)
    assert rel_err(loss.cpu().numpy(), loss_o.detach().numpy(), 1e-9).max() < 1e-9
    assert abs(loss_sum.item() - loss_o.sum().item()) < 1e-9 * abs(loss_o.sum().item())
    # chain the kernel's callable-level gradient to the leaves and compare with oracle autograd
    theta = P.inertial_parameters.detach().clone().requires_grad_()
    frl = fr.detach().clone().requires_grad_()
    ln = half.detach().clone().reshape(1, 3).requires_grad_()
    m = frl.abs()
    flat = torch.cat((co.theta_to_inertia_vector(theta).reshape(10), (2 * m[0] * m[1] / (m[0] + m[1])).reshape(1),
                      ln.abs().reshape(3)))
    flat.backward(grad.cpu())
    assert max_rel_to_scale(theta.grad.numpy(), P.inertial_parameters.grad.numpy()) < 1e-9
    assert max_rel_to_scale(frl.grad.numpy(), P.friction_params.grad.numpy()) < 1e-9
    assert max_rel_to_scale(ln.grad.numpy(), P.length_params[0].grad.numpy()) < 1e-9


def test_weighted_backward_equals_sum_of_per_sample_gradients():
    g = load_golden('cube_synthetic')
    inertia, mu, half = (torch.from_numpy(a).to(DEV).requires_grad_() for a in kernel_level_params(g))
    x, xp = torch.from_numpy(g['x']).to(DEV), torch.from_numpy(g['x_plus']).to(DEV)
    w = torch.rand(x.shape[0], dtype=torch.float64, device=DEV)
    loss = ops.CubeContactNetsLoss.apply(x, xp, inertia, mu, half, float(g['dt']), 1e-3)
    (loss * w).sum().backward()                      # non-uniform upstream gradient -> weighted kernel pass
    got = torch.cat((inertia.grad, mu.grad, half.grad)).cpu().numpy()
    ref = np.zeros(14)
    for lo in range(0, x.shape[0], 64):              # per-chunk uniform passes, weights applied outside
        for i in range(lo, min(lo + 64, x.shape[0])):
            _, gi, _, _, _ = ops.cube_loss_raw(x[i:i + 1], xp[i:i + 1], inertia.detach(), mu.detach(), half.detach(),
                                               float(g['dt']), 1e-3)
            ref += w[i].item() * gi.cpu().numpy()
    assert max_rel_to_scale(got, ref) < 1e-11


def test_wavefront_and_simple_kernels_agree():
    """Both kernel variants run the same per-sample algorithm (the compiler may contract FMAs
    differently in the two kernels, so agreement is to rounding, not bitwise)."""
    g = load_golden('cube_synthetic')
    inertia, mu, half = (torch.from_numpy(a).to(DEV) for a in kernel_level_params(g))
    x = synthetic.cube_states(100003, seed=31, device=DEV)
    traj, _ = ops.cube_rollout(x, inertia, mu, half, 0.0068, 1)
    xp = synthetic.perturb_next_state(traj[:, 1], seed=32)
    try:
        ops.set_loss_variant(1)
        a = ops.cube_loss_raw(x, xp, inertia, mu, half, 0.0068, 1e-3, want_force=True, want_iters=True)
    finally:
        ops.set_loss_variant(0)
    b = ops.cube_loss_raw(x, xp, inertia, mu, half, 0.0068, 1e-3, want_force=True, want_iters=True)
    assert rel_err(a[0].cpu().numpy(), b[0].cpu().numpy(), 1e-9).max() < 1e-10
    fscale = b[3].abs().amax(dim=1, keepdim=True).clamp(min=1e-6)
    assert ((a[3] - b[3]).abs() / fscale).max().item() < 1e-8
    assert (a[4] != b[4]).double().mean().item() < 0.01          # iteration counts agree for > 99%
    assert max_rel_to_scale(a[1].cpu().numpy(), b[1].cpu().numpy()) < 1e-10
    assert abs(a[2].item() - b[2].item()) < 1e-10 * abs(b[2].item())


def test_empty_and_ragged_batches(assets_dir):
    g = load_golden('cube_synthetic')
    s = _system(g, assets_dir)
    x, xp = torch.from_numpy(g['x']).to(DEV), torch.from_numpy(g['x_plus']).to(DEV)
    full = s.contactnets_loss(x, None, xp)
    empty = s.contactnets_loss(x[:0], None, xp[:0])
    assert empty.shape == (0,)
    for n in (1, 31, 33, 129):
        assert torch.equal(s.contactnets_loss(x[:n], None, xp[:n]), full[:n])
    # extra leading batch dimensions (*, n_x) -> (*,)
    two = s.contactnets_loss(x[:24].reshape(4, 6, 13), None, xp[:24].reshape(4, 6, 13))
    assert two.shape == (4, 6) and torch.equal(two.reshape(-1), full[:24])
    # strided (non-contiguous) views are accepted
    assert torch.equal(s.contactnets_loss(x[::2], None, xp[::2]), full[::2])


def test_bitwise_determinism_and_shard_additivity():
    """Same inputs -> identical bits; gradient of the whole batch = sum over sample shards (the
    property the multi-GPU path relies on)."""
    pi, fr, half = synthetic.cube_learnables_perturbed(1)
    from oracle import contactnets_oracle as co
    g = dict(theta=co.pi_cm_to_theta(pi).numpy(), friction_params=fr.numpy(), half_lengths=half.numpy())
    inertia, mu, hl = (torch.from_numpy(a).to(DEV) for a in kernel_level_params(g))
    B = 1 << 20                                       # BASELINE.json full size (config 5)
    x = synthetic.cube_states(B, seed=3, device=DEV)
    traj, _ = ops.cube_rollout(x, inertia, mu, hl, 0.0068, 1)
    xp = synthetic.perturb_next_state(traj[:, 1], seed=4)
    l1, g1, s1, _, it = ops.cube_loss_raw(x, xp, inertia, mu, hl, 0.0068, 1e-3, want_iters=True)
    l2, g2, s2, _, _ = ops.cube_loss_raw(x, xp, inertia, mu, hl, 0.0068, 1e-3)
    assert torch.equal(l1, l2) and torch.equal(g1, g2) and torch.equal(s1, s2)
    assert torch.isfinite(l1).all() and (l1 >= -1e-12).all()
    assert int(it.max()) <= 60
    parts = [ops.cube_loss_raw(x[i::8].contiguous(), xp[i::8].contiguous(), inertia, mu, hl, 0.0068, 1e-3)
             for i in range(8)]
    gsum = sum(p[1] for p in parts)
    assert max_rel_to_scale(gsum.cpu().numpy(), g1.cpu().numpy()) < 1e-11
    assert abs(sum(p[2] for p in parts).item() - s1.item()) < 1e-11 * abs(s1.item())
    assert abs(l1.sum().item() - s1.item()) < 1e-11 * abs(s1.item())


def test_symmetry_properties_full_size():
    """Size-independent properties: the loss is invariant under world yaw and horizontal translation."""
    pi, fr, half = synthetic.cube_learnables_perturbed(2)
    from oracle import contactnets_oracle as co
    g = dict(theta=co.pi_cm_to_theta(pi).numpy(), friction_params=fr.numpy(), half_lengths=half.numpy())
    inertia, mu, hl = (torch.from_numpy(a).to(DEV) for a in kernel_level_params(g))
    B = 65536
    x = synthetic.cube_states(B, seed=5, device=DEV)
    traj, _ = ops.cube_rollout(x, inertia, mu, hl, 0.0068, 1)
    xp = synthetic.perturb_next_state(traj[:, 1], seed=6)
    base = ops.cube_loss_raw(x, xp, inertia, mu, hl, 0.0068, 1e-3)[0]
    ang = 0.7
    qz = torch.tensor([np.cos(ang / 2), 0, 0, np.sin(ang / 2)], dtype=torch.float64, device=DEV)
    Rz = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]],
                      dtype=torch.float64, device=DEV)
    from dair_pll_b200 import quaternion

    def yaw(s):
        out = s.clone()
        out[:, :4] = quaternion.multiply(qz.expand(s.shape[0], 4), s[:, :4])
        out[:, 4:7] = s[:, 4:7] @ Rz.t() + torch.tensor([0.3, -1.2, 0.0], dtype=torch.float64, device=DEV)
        out[:, 10:13] = s[:, 10:13] @ Rz.t()          # world-frame linear velocity rotates; body w does not
        return out
    moved = ops.cube_loss_raw(yaw(x), yaw(xp), inertia, mu, hl, 0.0068, 1e-3)[0]
    err = (moved - base).abs() / base.abs().clamp(min=1e-9)
    assert err.max().item() < 1e-7 and err.median().item() < 1e-12


@pytest.mark.parametrize('name', CASES)
def test_fp32_variant_matches_reference_golden(name, assets_dir):
    """fp32 variant (fp32 states/parameters/outputs, fp64 arithmetic inside): 1e-4 (north_star) on
    losses, parameter gradients and next states, against the fp64 reference results."""
    g = load_golden(name)
    s = _system(g, assets_dir)
    x, xp = torch.from_numpy(g['x']).float().to(DEV), torch.from_numpy(g['x_plus']).float().to(DEV)
    loss = s.contactnets_loss(x, None, xp)
    assert loss.dtype == torch.float32
    loss.mean().backward()
    l = loss.detach().cpu().numpy().astype(np.float64)
    assert np.abs(l - g['loss']).max() < 1e-4 * max(np.abs(g['loss']).max(), 1e-3)
    assert abs(l.mean() - g['loss'].mean()) < 1e-4 * abs(g['loss'].mean())
    mt = s.multibody_terms
    assert max_rel_to_scale(mt.lagrangian_terms.inertial_parameters.grad.cpu().numpy(), g['grad_theta']) < 1e-4
    assert max_rel_to_scale(mt.contact_terms.friction_params.grad.cpu().numpy(), g['grad_friction']) < 1e-4
    assert max_rel_to_scale(mt.contact_terms.geometries[0].length_params.grad.cpu().numpy(), g['grad_length']) < 1e-4
    x0 = torch.from_numpy(g['sim_x0']).float().to(DEV)
    with torch.no_grad():
        traj, _ = s.simulate(x0.unsqueeze(-2), torch.zeros(x0.shape[0], 1, device=DEV), 1)
    assert traj.dtype == torch.float32
    ref = g['sim_traj'][:, 1]
    assert np.abs(traj[:, 1].cpu().numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


def test_invalid_arguments_raise():
    x = torch.zeros(4, 12, dtype=torch.float64, device=DEV)
    p = torch.ones(10, dtype=torch.float64, device=DEV)
    with pytest.raises(ValueError):
        ops.cube_loss_raw(x, x, p, p[:1], p[:3], 0.0068, 1e-3)
    with pytest.raises(TypeError):
        ops.cube_loss_raw(x.half(), x.half(), p.half(), p[:1].half(), p[:3].half(), 0.0068, 1e-3)


ELBOW_CASES = ['elbow_nominal', 'elbow_perturbed']


def _elbow_system(g, assets_dir):
    s = MultibodyLearnableSystem({'elbow': os.path.join(assets_dir, 'elbow.urdf')}, float(g['dt']))
    s.load_state_dict({
        'multibody_terms.lagrangian_terms.inertial_parameters': torch.from_numpy(g['theta']),
        'multibody_terms.contact_terms.friction_params': torch.from_numpy(g['friction_params']),
        'multibody_terms.contact_terms.geometries.0.length_params': torch.from_numpy(g['half_lengths'][0]).reshape(1, 3),
        'multibody_terms.contact_terms.geometries.1.length_params': torch.from_numpy(g['half_lengths'][1]).reshape(1, 3)})
    return s.to(DEV)


@pytest.mark.parametrize('name', ELBOW_CASES)
def test_elbow_loss_gradients_and_rollout_match_reference_golden(name, assets_dir):
    """Two-body articulated asset through the module API: losses, parameter gradients (theta of both
    bodies, three friction parameters, both boxes' lengths), impulses and rollouts at 1e-9."""
    g = load_golden(name)
    s = _elbow_system(g, assets_dir)
    assert s.space.n_x == 15
    x, xp = torch.from_numpy(g['x']).to(DEV), torch.from_numpy(g['x_plus']).to(DEV)
    loss = s.contactnets_loss(x, None, xp)
    loss.mean().backward()
    l = loss.detach().cpu().numpy()
    assert np.abs(l - g['loss']).max() < 1e-12
    assert rel_err(l, g['loss'], 1e-9).max() < 1e-9
    mt = s.multibody_terms
    assert max_rel_to_scale(mt.lagrangian_terms.inertial_parameters.grad.cpu().numpy(), g['grad_theta']) < 1e-9
    assert max_rel_to_scale(mt.contact_terms.friction_params.grad.cpu().numpy(), g['grad_friction']) < 1e-9
    gl = np.stack([mt.contact_terms.geometries[i].length_params.grad.cpu().numpy().reshape(3) for i in range(2)])
    assert max_rel_to_scale(gl, g['grad_length']) < 1e-9
    inertia, mu, half, kin = (t.detach() for t in s._elbow_params(torch.float64, torch.device(DEV)))
    _, _, _, force, iters = ops.elbow_loss_raw(x, xp, inertia, mu, half, kin, float(g['dt']), 1e-3, want_force=True,
                                               want_iters=True)
    scale = np.maximum(np.abs(g['force']).max(axis=1, keepdims=True), 1e-6)
    assert (np.abs(force.cpu().numpy() - g['force']) / scale).max() < 1e-7
    assert int(iters.max()) <= 60
    x0 = torch.from_numpy(g['sim_x0']).to(DEV)
    steps = g['sim_traj'].shape[1] - 1
    with torch.no_grad():
        traj, _ = s.simulate(x0.unsqueeze(-2), torch.zeros(x0.shape[0], 1, device=DEV), steps)
    t = traj.cpu().numpy()
    assert np.abs(t[:, 1] - g['sim_traj'][:, 1]).max() < 1e-9
    assert np.abs(t - g['sim_traj']).max() < 1e-6


def _mesh_system(g, assets_dir, width):
    from dair_pll_b200.deep_support_function import HomogeneousICNN
    s = MultibodyLearnableSystem({'elbow': os.path.join(assets_dir, 'elbow_mesh.urdf')}, float(g['dt']))
    geoms = s.multibody_terms.contact_terms.geometries
    sd = {'multibody_terms.lagrangian_terms.inertial_parameters': torch.from_numpy(g['theta']),
          'multibody_terms.contact_terms.friction_params': torch.from_numpy(g['friction_params'])}
    for gi in range(2):
        geoms[gi].network = HomogeneousICNN(2, width)
        geoms[gi].perturbations = torch.from_numpy(g[f'net{gi}_perturbations'])
        pre = f'multibody_terms.contact_terms.geometries.{gi}.network.'
        sd.update({pre + 'input_weights.0': torch.from_numpy(g[f'net{gi}_Wd0']),
                   pre + 'input_weights.1': torch.from_numpy(g[f'net{gi}_Wd1']),
                   pre + 'hidden_weights.0': torch.from_numpy(g[f'net{gi}_Wh']),
                   pre + 'output_weight': torch.from_numpy(g[f'net{gi}_wout'])})
    s.load_state_dict(sd)
    return s.to(DEV)


def test_elbow_learned_geometry_matches_reference_golden(assets_dir):
    """Elbow with DeepSupportConvex geometry (config 3's system, network width 64 in the fixture): loss,
    inertia / friction gradients, the gradients of every ICNN weight tensor, and one-step rollouts."""
    g = load_golden('elbow_mesh_w64')
    s = _mesh_system(g, assets_dir, 64)
    x, xp = torch.from_numpy(g['x']).to(DEV), torch.from_numpy(g['x_plus']).to(DEV)
    loss = s.contactnets_loss(x, None, xp)
    loss.mean().backward()
    l = loss.detach().cpu().numpy()
    assert np.abs(l - g['loss']).max() < 1e-12
    assert rel_err(l, g['loss'], 1e-9).max() < 1e-9
    mt = s.multibody_terms
    assert max_rel_to_scale(mt.lagrangian_terms.inertial_parameters.grad.cpu().numpy(), g['grad_theta']) < 1e-9
    assert max_rel_to_scale(mt.contact_terms.friction_params.grad.cpu().numpy(), g['grad_friction']) < 1e-9
    for gi in range(2):
        net = mt.contact_terms.geometries[gi].network
        for k, p in (('Wd0', net.input_weights[0]), ('Wd1', net.input_weights[1]), ('Wh', net.hidden_weights[0]),
                     ('wout', net.output_weight)):
            assert max_rel_to_scale(p.grad.cpu().numpy(), g[f'net{gi}_grad_{k}']) < 1e-9, (gi, k)
    x0 = torch.from_numpy(g['sim_x0']).to(DEV)
    steps = g['sim_traj'].shape[1] - 1
    with torch.no_grad():
        traj, _ = s.simulate(x0.unsqueeze(-2), torch.zeros(x0.shape[0], 1, device=DEV), steps)
    t = traj.cpu().numpy()
    assert np.abs(t[:, 1] - g['sim_traj'][:, 1]).max() < 1e-9
    assert np.abs(t - g['sim_traj']).max() < 1e-6


def test_rollout_backward_matches_oracle_autograd(assets_dir):
    """Prediction-loss path: gradients of a multi-step rollout w.r.t. every learnable
    parameter and the initial state, through the module API, against autograd through the CPU oracle
    (implicit differentiation of each step's QP)."""
    from oracle import contactnets_oracle as co
    from oracle.callables import CUBE_TREE, TreeCallables
    g = load_golden('cube_real_perturbed')
    s = _system(g, assets_dir)
    x0 = torch.from_numpy(g['sim_x0'][:32]).to(DEV).requires_grad_()
    steps = 3
    target = torch.from_numpy(g['sim_traj'][:32, 1:steps + 1]).to(DEV) + 0.01
    traj, _ = s.simulate(x0.unsqueeze(-2), torch.zeros(32, 1, device=DEV), steps)
    loss = ((traj[:, 1:] - target) ** 2).sum()
    loss.backward()
    # oracle, in two forms of the same update (contactnets_oracle.forward_dynamics): the reference's
    # v+ = v- + M^-1 J^T f with f = Pi(-(A w + q)/eps) -- whose DERIVATIVES carry ~cond(Q) eps ~ 1e-9 of rounding
    # because f amplifies the rounding of w by 1/eps = 1e4 -- and the identical, better conditioned v- + L^-T w.
    got = [mt_p.grad.cpu().numpy() for mt_p in (s.multibody_terms.lagrangian_terms.inertial_parameters,
                                                 s.multibody_terms.contact_terms.friction_params,
                                                 s.multibody_terms.contact_terms.geometries[0].length_params)]
    got.append(x0.grad.cpu().numpy())
    refs = {}
    for primal in (False, True):
        co.PRIMAL_UPDATE = primal
        try:
            P = co.OracleParams(torch.from_numpy(g['theta']).clone(), torch.from_numpy(g['friction_params']).clone(),
                                [torch.from_numpy(g['half_lengths']).clone().reshape(1, 3)]).requires_grad_()
            x0o = torch.from_numpy(g['sim_x0'][:32]).clone().requires_grad_()
            tro = co.simulate(TreeCallables(CUBE_TREE), P, x0o, float(g['dt']), steps)
            ((tro[:, 1:] - target.cpu()) ** 2).sum().backward()
        finally:
            co.PRIMAL_UPDATE = False
        assert np.abs(traj.detach().cpu().numpy() - tro.detach().numpy()).max() < 1e-8
        refs[primal] = [P.inertial_parameters.grad.numpy(), P.friction_params.grad.numpy(), P.length_params[0].grad.numpy(),
                        x0o.grad.numpy()]
    # north_star: parameter gradients within 1e-9 (every solve ends with a polishing Newton step in dual arithmetic,
    # whose tangent is the implicit-function derivative at the converged point) -- held against the well-conditioned
    # form; the reference's own form agrees to its rounding, and the two oracle forms differ from each other by
    # as much as the kernels differ from the reference form
    for a, b in zip(got, refs[True]):
        assert max_rel_to_scale(a, b) < 1e-9
    for a, b, c in zip(got, refs[False], refs[True]):
        assert max_rel_to_scale(a, b) < max(1e-9, 3 * max_rel_to_scale(c, b))
        assert max_rel_to_scale(a, b) < 2e-8


def test_edge_cases_match_oracle():
    """Edge cases of the domain: axis-aligned cube resting flat (ties between the four bottom corners
    and degenerate top-k), zero velocities (zero sliding speed: the |.| and sqrt branches), deep
    penetration, far-away flight, and a non-unit quaternion (Drake normalises the rotation)."""
    from oracle import contactnets_oracle as co
    from oracle.callables import CUBE_TREE, TreeCallables
    calls = TreeCallables(CUBE_TREE)
    pi, fr, half = synthetic.cube_learnables_perturbed(4)
    P = co.OracleParams(co.pi_cm_to_theta(pi), fr, [half.reshape(1, 3)])
    h = half.numpy()
    rows = [
        [1, 0, 0, 0, 0, 0, h[2], 0, 0, 0, 0, 0, 0],                 # flat on the ground, at rest
        [1, 0, 0, 0, 0, 0, h[2] - 0.004, 0, 0, 0, 0.3, 0, -0.5],    # flat, penetrating, sliding in x
        [1, 0, 0, 0, 0, 0, 5.0, 1, 2, 3, 0, 0, 0],                  # far away, spinning
        [2, 0, 0, 0, 0, 0, h[2] + 1e-4, 0, 0, 0, 0, 0, -0.2],       # non-unit quaternion (norm 2)
        [0.7071067811865476, 0.7071067811865476, 0, 0, 0, 0, h[1] + 1e-3, 0, 0, 0.5, 0.2, -0.1, -0.3],  # on a side face
        [0.9, 0.1, 0.3, 0.2, 0.1, -0.2, 0.03, 3, -2, 1, 0.5, 0.5, -1.5],   # deep penetration, tumbling
    ]
    x = torch.tensor(rows, dtype=torch.float64)
    with torch.no_grad():
        xp = co.sim_step(calls, P, x, 0.0068)
        xp2 = xp.clone()
        xp2[:, 7:] += 0.05                                           # perturbed observations
    g = dict(theta=P.inertial_parameters.numpy(), friction_params=fr.numpy(), half_lengths=half.numpy())
    inertia, mu, hl = (torch.from_numpy(a).to(DEV) for a in kernel_level_params(g))
    for target in (xp, xp2):
        lo = co.contactnets_loss(calls, P, x, target, 0.0068).numpy()
        l, _, _, _, it = ops.cube_loss_raw(x.to(DEV), target.to(DEV), inertia, mu, hl, 0.0068, 1e-3, want_iters=True)
        assert np.abs(l.cpu().numpy() - lo).max() < 1e-11 * max(1.0, np.abs(lo).max())
    traj, _ = ops.cube_rollout(x.to(DEV), inertia, mu, hl, 0.0068, 1)
    assert np.abs(traj[:, 1].cpu().numpy() - xp.numpy()).max() < 1e-8


def test_solver_failure_mask_zeroes_loss_and_gradient():
    """multibody_learnable_system.py:186-192: samples whose impulse exceeds 1e3 (or is not finite) get
    force := 0 and constant := 0, i.e. loss 0 and no gradient.  A 1e5 m/s velocity jump forces that."""
    g = load_golden('cube_synthetic')
    inertia, mu, half = (torch.from_numpy(a).to(DEV) for a in kernel_level_params(g))
    x = torch.from_numpy(g['x'][:64]).to(DEV)
    xp = torch.from_numpy(g['x_plus'][:64]).to(DEV).clone()
    ref = ops.cube_loss_raw(x, xp, inertia, mu, half, float(g['dt']), 1e-3, want_force=True)
    bad = xp.clone()
    bad[::4, 12] = 1e5                        # absurd upward velocity with the corners on the ground
    bad[::4, 6] = 0.05
    bad[1::8, 9] = float('nan')               # non-finite observation
    out = ops.cube_loss_raw(x, bad, inertia, mu, half, float(g['dt']), 1e-3, want_force=True)
    masked = torch.zeros(64, dtype=torch.bool, device=DEV)
    masked[::4] = True
    masked[1::8] = True
    assert (out[0][masked] == 0).all() and (out[3][masked] == 0).all()
    assert torch.equal(out[0][~masked], ref[0][~masked])
    assert torch.isfinite(out[1]).all()
    # gradient = sum over the unmasked samples only
    keep = ops.cube_loss_raw(x[~masked], xp[~masked], inertia, mu, half, float(g['dt']), 1e-3)
    assert max_rel_to_scale(out[1].cpu().numpy(), keep[1].cpu().numpy()) < 1e-11


@pytest.mark.parametrize('name', CASES)
def test_dense_terms_match_reference_golden(name, assets_dir):
    """MultibodyTerms.forward: M, M^-1 F and phi against the reference-code goldens; J and the Delassus
    operator against the oracle (same canonical contact order)."""
    from oracle import contactnets_oracle as co
    from oracle.callables import CUBE_TREE, TreeCallables
    g = load_golden(name)
    s = _system(g, assets_dir)
    xp = torch.from_numpy(g['x_plus']).to(DEV)
    D, M, J, phi, acc = s.multibody_terms(xp[:, :7], xp[:, 7:], None)
    assert np.abs(M.cpu().numpy() - g['terms_M']).max() < 1e-15
    assert np.abs(acc.cpu().numpy() - g['terms_acc']).max() < 1e-10 * max(1.0, np.abs(g['terms_acc']).max())
    assert np.abs(np.sort(phi.cpu().numpy(), -1) - g['terms_phi_sorted']).max() < 1e-15
    P = co.OracleParams(torch.from_numpy(g['theta']), torch.from_numpy(g['friction_params']),
                        [torch.from_numpy(g['half_lengths']).reshape(1, 3)])
    with torch.no_grad():
        Mo, Jo, phio, _ = co.multibody_terms(TreeCallables(CUBE_TREE), P, xp.cpu()[:, :7], xp.cpu()[:, 7:])
        Do = Jo @ torch.linalg.solve(Mo, Jo.transpose(-1, -2))
    assert np.abs(J.cpu().numpy() - Jo.numpy()).max() < 1e-14
    assert np.abs(phi.cpu().numpy() - phio.numpy()).max() < 1e-15
    assert np.abs(D.cpu().numpy() - Do.numpy()).max() < 1e-10 * np.abs(Do.numpy()).max()


@pytest.mark.parametrize('width,rows', [(64, 1000), (256, 4099)])
def test_support_network_kernels_match_oracle_autograd(width, rows):
    """The fused support-network layers (dpll_icnn_*: input layer, slope masks, output contraction, weight-gradient
    reductions) + library GEMMs against autograd through the oracle's restatement of
    deep_support_function.py:238-266: support points and the gradients of all four weights."""
    from dair_pll_b200.deep_support_function import ICNNSupport
    from oracle import contactnets_oracle as co
    torch.manual_seed(width)
    ws = [torch.randn(3, width, dtype=torch.float64), torch.randn(3, width, dtype=torch.float64),
          torch.randn(width, width, dtype=torch.float64) / width, torch.randn(width, dtype=torch.float64)]
    d = torch.randn(rows, 3, dtype=torch.float64)
    d = d / d.norm(dim=-1, keepdim=True)
    gp = torch.randn(rows, 3, dtype=torch.float64)
    a = [w.clone().to(DEV).requires_grad_() for w in ws]
    p = ICNNSupport.apply(d.to(DEV), a[0], a[1], a[2], a[3], 0.5)
    (p * gp.to(DEV)).sum().backward()
    b = [w.clone().requires_grad_() for w in ws]
    po = co.icnn_support(dict(Wd0=b[0], Wd1=b[1], Wh=b[2], wout=b[3]), d)
    (po * gp).sum().backward()
    assert torch.allclose(p.cpu(), po, rtol=1e-11, atol=1e-12)
    for x, y in zip(a, b):
        scale = y.grad.abs().max()
        assert (x.grad.cpu() - y.grad).abs().max() <= 1e-10 * scale


def test_graphed_step_replays_the_eager_step(assets_dir):
    """parallel.GraphedStep: the captured forward + backward + staging of the flat gradient buffer gives the same
    numbers as eager launches, replay after replay, also after the inputs are overwritten in place."""
    from dair_pll_b200 import parallel
    g = load_golden('cube_synthetic')
    s = _system(g, assets_dir)
    x, xp = torch.from_numpy(g['x']).to(DEV), torch.from_numpy(g['x_plus']).to(DEV)
    reducer = parallel.GradientAllReduce(list(s.parameters()), DEV, 1)

    def step():
        reducer.zero()
        mean = s.contactnets_loss(x, None, xp).mean()
        mean.backward()
        return reducer.stage(mean)
    eager = step().clone()
    graphed = parallel.GraphedStep(step, DEV)
    for _ in range(3):
        assert torch.equal(graphed(), eager)
    x.copy_(x.flip(0)), xp.copy_(xp.flip(0))             # same multiset of pairs, new contents at the captured addresses
    flipped = graphed().clone()
    assert torch.allclose(flipped, eager, rtol=1e-12, atol=1e-18) and torch.equal(step(), flipped)


def test_elbow_rollout_backward_matches_oracle_autograd(assets_dir):
    """Prediction-loss path for the two-body system: gradients of a 2-step rollout w.r.t. theta of both bodies,
    the three friction parameters, both boxes' lengths and the initial state, through the module API, against
    autograd through the CPU oracle."""
    from oracle import contactnets_oracle as co
    from oracle.callables import ELBOW_TREE, TreeCallables
    from tests.util import oracle_params_from_golden
    g = load_golden('elbow_perturbed')
    s = _elbow_system(g, assets_dir)
    n, steps = 16, 2
    x0 = torch.from_numpy(g['sim_x0'][:n]).to(DEV).requires_grad_()
    target = torch.from_numpy(g['sim_traj'][:n, 1:steps + 1]).to(DEV) + 0.01
    traj, _ = s.simulate(x0.unsqueeze(-2), torch.zeros(n, 1, device=DEV), steps)
    ((traj[:, 1:] - target) ** 2).sum().backward()
    P = oracle_params_from_golden(g)
    x0o = torch.from_numpy(g['sim_x0'][:n]).clone().requires_grad_()
    tro = co.simulate(TreeCallables(ELBOW_TREE), P, x0o, float(g['dt']), steps)
    ((tro[:, 1:] - target.cpu()) ** 2).sum().backward()
    assert np.abs(traj.detach().cpu().numpy() - tro.detach().numpy()).max() < 1e-8
    mt = s.multibody_terms
    assert max_rel_to_scale(mt.lagrangian_terms.inertial_parameters.grad.cpu().numpy(), P.inertial_parameters.grad.numpy()) < 1e-9
    assert max_rel_to_scale(mt.contact_terms.friction_params.grad.cpu().numpy(), P.friction_params.grad.numpy()) < 1e-9
    gl = np.stack([mt.contact_terms.geometries[i].length_params.grad.cpu().numpy().reshape(3) for i in range(2)])
    assert max_rel_to_scale(gl, np.stack([p.grad.numpy().reshape(3) for p in P.length_params])) < 1e-9
    assert max_rel_to_scale(x0.grad.cpu().numpy(), x0o.grad.numpy()) < 1e-9


def test_batch_loss_mean_and_sum_shortcuts_equal_the_generic_reductions(assets_dir):
    """``contactnets_loss(...).mean()`` / ``.sum()`` come from the launch's own reduction and fused gradient
    (ops.BatchLoss); they must equal the generic torch reductions over the per-sample loss and their autograd
    backward (second, weighted launch), and in-place edits must switch the shortcut off."""
    g = load_golden('cube_real_perturbed')
    s = _system(g, assets_dir)
    x, xp = torch.from_numpy(g['x']).to(DEV), torch.from_numpy(g['x_plus']).to(DEV)
    params = list(s.parameters())

    def grads_of(reduce):
        for p in params:
            p.grad = None
        out = reduce(s.contactnets_loss(x, None, xp))
        out.backward()
        return out.detach().clone(), [p.grad.clone() for p in params]
    for fused, generic in ((lambda l: l.mean(), lambda l: torch.mean(l.as_subclass(torch.Tensor))),
                           (lambda l: l.sum(), lambda l: (l * 1.0).sum())):
        vf, gf = grads_of(fused)
        vg, gg = grads_of(generic)
        assert abs(vf.item() - vg.item()) <= 1e-13 * abs(vg.item())
        for a, b in zip(gf, gg):
            assert torch.allclose(a, b, rtol=1e-11, atol=1e-16)
    loss = s.contactnets_loss(x, None, xp)
    assert type(loss * 2.0) is torch.Tensor and loss.shape == (x.shape[0],)
    with torch.no_grad():
        ref = loss.mean().item()
        loss.mul_(3.0)
        assert abs(loss.mean().item() - 3.0 * ref) <= 1e-12 * abs(ref)


def test_scheduler_boundary_batch_sizes_agree_with_the_simple_kernel():
    """Batch sizes at the seams of the wavefront scheduler -- fewer samples than warps, exactly / one off a full
    triage visit per warp, one off a full slot pool per warp -- give the per-sample results of the
    one-sample-per-thread kernel: every sample is finalised exactly once and none is lost in a queue."""
    g = load_golden('cube_synthetic')
    inertia, mu, half = (torch.from_numpy(a).to(DEV) for a in kernel_level_params(g))
    warps = 148 * 2 * 4
    sizes = [1, 5, 32, 63, warps - 1, warps + 7, 32 * warps - 1, 32 * warps, 32 * warps + 1, 64 * warps + 33, 96 * warps - 5]
    nmax = max(sizes)
    x = synthetic.cube_states(nmax, seed=41, device=DEV)
    traj, _ = ops.cube_rollout(x, inertia, mu, half, 0.0068, 1)
    xp = synthetic.perturb_next_state(traj[:, 1], seed=42)
    try:
        ops.set_loss_variant(1)
        ref = ops.cube_loss_raw(x, xp, inertia, mu, half, 0.0068, 1e-3, want_iters=True)
    finally:
        ops.set_loss_variant(0)
    ref_loss = ref[0].cpu().numpy()
    for n in sizes:
        loss, grad, total, _, iters = ops.cube_loss_raw(x[:n], xp[:n], inertia, mu, half, 0.0068, 1e-3, want_iters=True)
        l = loss.cpu().numpy()
        assert rel_err(l, ref_loss[:n], 1e-9).max() < 1e-10, n
        assert abs(total.item() - l.sum()) <= 1e-11 * abs(l.sum()), n
        assert int(iters.min()) >= 0 and int(iters.max()) <= 100, n


def test_contactnets_training_example_recovers_geometry():
    """examples/contactnets_simple.py (BASELINE config 1) end to end on the GPU: tosses simulated at the URDF's
    parameters, device-resident slices, Adam on the ContactNets loss from 30%-perturbed parameters -- the training
    loss falls and the box size moves to the truth (what ContactNets identifies best from toss data)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('contactnets_simple', os.path.join(root, 'examples', 'contactnets_simple.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.run(epochs=40, n_pop=64, batch_size=1024, lr=3e-3, seed=0, perturbation=0.3, verbose=False)
    h = out['history']
    assert h[-1] < 0.5 * h[0]
    # the prediction loss (experiment.py:292-320) through the differentiable 2-step rollout learns as well
    pred = mod.run(epochs=15, n_pop=32, batch_size=1024, lr=3e-3, seed=0, perturbation=0.3, verbose=False,
                   contactnets=False, t_prediction=2)
    assert pred['history'][-1] < 0.5 * pred['history'][0]
    truth = out['truth']['half_lengths']
    err0 = max(abs(a - b) for a, b in zip(out['initial']['half_lengths'], truth))
    err = max(abs(a - b) for a, b in zip(out['learned']['half_lengths'], truth))
    assert err0 > 2e-3 and err < 0.4 * err0


def test_dynamic_distribution_variant_agrees_to_rounding():
    """Variant 2 (warps draw their triage chunks from a global counter): same per-sample results, gradient sums equal
    up to the rounding of a different summation order; sizes around the chunk and pool seams."""
    g = load_golden('cube_synthetic')
    inertia, mu, half = (torch.from_numpy(a).to(DEV) for a in kernel_level_params(g))
    x = synthetic.cube_states(200001, seed=51, device=DEV)
    traj, _ = ops.cube_rollout(x, inertia, mu, half, 0.0068, 1)
    xp = synthetic.perturb_next_state(traj[:, 1], seed=52)
    for n in (1, 31, 32, 33, 4097, 200001):
        a = ops.cube_loss_raw(x[:n], xp[:n], inertia, mu, half, 0.0068, 1e-3, want_force=True, want_iters=True)
        try:
            ops.set_loss_variant(2)
            b = ops.cube_loss_raw(x[:n], xp[:n], inertia, mu, half, 0.0068, 1e-3, want_force=True, want_iters=True)
        finally:
            ops.set_loss_variant(0)
        assert torch.equal(a[0], b[0]) and torch.equal(a[3], b[3]) and torch.equal(a[4], b[4]), n
        assert max_rel_to_scale(a[1].cpu().numpy(), b[1].cpu().numpy()) < 1e-12, n
        assert abs(a[2].item() - b[2].item()) <= 1e-12 * abs(a[2].item()), n
