"""ORACLE (test infrastructure).  Writes tests/golden/elbow_mixed_w64.npz by running the REFERENCE's own
``contactnets_loss`` / ``simulate`` (through oracle/ref_shim.py) for the two-body elbow with MIXED collision geometry: the
reference's ``Box`` (geometry.py:359-413) on the first link and its ``DeepSupportConvex`` (:255-325, network width 64 to keep
the fixture small) on the second.  Needs /root/reference: build container only; the fixture is committed.

    python -m oracle.gen_golden_mixed
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


def main():
    ref_shim.import_reference()
    torch.set_default_dtype(torch.float64)
    from dair_pll.geometry import Box, DeepSupportConvex
    from dair_pll_b200 import synthetic
    pi_e, fr_e, half_e = synthetic.elbow_learnables_perturbed(1)
    half = torch.as_tensor(synthetic.ELBOW_HALF, dtype=torch.float64)
    signs = torch.tensor([[(i >> 2) & 1, (i >> 1) & 1, i & 1] for i in range(8)], dtype=torch.float64) * 2 - 1
    torch.manual_seed(6)
    box = Box(half_e[0].double(), 4)
    mesh = DeepSupportConvex(signs * half, width=64)
    mesh.perturbations = mesh.perturbations.double()
    system = ref_shim.build_reference_system('elbow', DT, pi_e, fr_e, None, body_geometries=[box, mesh])
    mt = system.multibody_terms
    x = synthetic.elbow_states(160, seed=23)
    with torch.no_grad():
        xn, _ = system.integrator.step(x, torch.zeros(x.shape[0], 1))
    x_plus = synthetic.perturb_next_state(xn, seed=24, n_q=8)
    loss = system.contactnets_loss(x, torch.zeros(x.shape[0], 0), x_plus)
    loss.mean().backward()
    net = mesh.network
    out = dict(dt=np.float64(DT), x=x.numpy(), x_plus=x_plus.numpy(), pi_cm=pi_e.numpy(), friction_params=fr_e.numpy(),
               theta=mt.lagrangian_terms.inertial_parameters.detach().numpy(), loss=loss.detach().numpy(),
               grad_theta=mt.lagrangian_terms.inertial_parameters.grad.numpy(),
               grad_friction=mt.contact_terms.friction_params.grad.numpy(),
               box_length_params=box.length_params.detach().numpy(), grad_box_length_params=box.length_params.grad.numpy(),
               net_Wd0=net.input_weights[0].detach().numpy(), net_Wd1=net.input_weights[1].detach().numpy(),
               net_Wh=net.hidden_weights[0].detach().numpy(), net_wout=net.output_weight.detach().numpy(),
               net_perturbations=mesh.perturbations.numpy(),
               net_grad_Wd0=net.input_weights[0].grad.numpy(), net_grad_Wd1=net.input_weights[1].grad.numpy(),
               net_grad_Wh=net.hidden_weights[0].grad.numpy(), net_grad_wout=net.output_weight.grad.numpy())
    with torch.no_grad():
        traj, _ = system.simulate(x[:24].unsqueeze(-2), torch.zeros(24, 1), 3)
    out.update(sim_x0=x[:24].numpy(), sim_traj=traj.numpy())
    path = os.path.join(ROOT, 'tests', 'golden', 'elbow_mixed_w64.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, 'mean loss %.12e' % loss.mean().item(), 'size %.1f KB' % (os.path.getsize(path) / 1e3))


if __name__ == '__main__':
    main()
