"""Kernel-only timing of the elbow loss kernel variants (0 = wavefront, 1 = one sample per thread, 2 = triage/solve passes)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dair_pll_b200 import ops, synthetic  # noqa: E402
from dair_pll_b200.inertia import InertialParameterConverter as IPC  # noqa: E402
from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem  # noqa: E402

dev = torch.device('cuda', 0)
s = MultibodyLearnableSystem({'elbow': os.path.join(ROOT, 'dair_pll_b200', 'assets', 'elbow.urdf')}, 0.0068)
pi, fr, half = synthetic.elbow_learnables_perturbed(0)
s.load_state_dict({'multibody_terms.lagrangian_terms.inertial_parameters': IPC.pi_cm_to_theta(pi),
                   'multibody_terms.contact_terms.friction_params': fr,
                   'multibody_terms.contact_terms.geometries.0.length_params': half[0].reshape(1, 3),
                   'multibody_terms.contact_terms.geometries.1.length_params': half[1].reshape(1, 3)})
s = s.to(dev)
for B in (262144, 65536, 1048576):
    x = synthetic.elbow_states(B, seed=0, device=dev)
    with torch.no_grad():
        traj, _ = s.simulate(x.unsqueeze(-2), torch.zeros(B, 1, device=dev), 1)
    xp = synthetic.perturb_next_state(traj[:, 1], seed=1, n_q=8)
    inertia, mu, hl, kin = (t.detach() for t in s._elbow_params(torch.float64, dev))
    ref = None
    for variant in (2, 0):
        ops.set_loss_variant(variant)
        for _ in range(2):
            out = ops.elbow_loss_raw(x, xp, inertia, mu, hl, kin, 0.0068, 1e-3, want_iters=True)
        torch.cuda.synchronize()
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        for _ in range(5):
            ops.elbow_loss_raw(x, xp, inertia, mu, hl, kin, 0.0068, 1e-3)
        en.record()
        torch.cuda.synchronize()
        ms = st.elapsed_time(en) / 5
        it = out[4].double()
        msg = f'variant {variant} B={B}: {ms:.3f} ms  {B / ms / 1e3:.2f} M samples/s  mean iters {it.mean().item():.3f} loss_sum {out[2].item():.12e}'
        if ref is not None:
            msg += f'  max |dloss| vs variant 2: {(out[0] - ref[0]).abs().max().item():.2e}  grad rel {((out[1] - ref[1]).abs().max() / ref[1].abs().max()).item():.2e}'
        ref = out
        print(msg, flush=True)
    ops.set_loss_variant(0)
    order = torch.argsort(ref[4], descending=True, stable=True)
    xo, xpo = x[order].contiguous(), xp[order].contiguous()
    for _ in range(2):
        o2 = ops.elbow_loss_raw(xo, xpo, inertia, mu, hl, kin, 0.0068, 1e-3, flags=ops.LOSS_DYNAMIC)
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(5):
        ops.elbow_loss_raw(xo, xpo, inertia, mu, hl, kin, 0.0068, 1e-3, flags=ops.LOSS_DYNAMIC)
    en.record()
    torch.cuda.synchronize()
    ms = st.elapsed_time(en) / 5
    print(f'variant 0 cost order + dynamic B={B}: {ms:.3f} ms  {B / ms / 1e3:.2f} M samples/s  '
          f'per-sample equal {torch.equal(o2[0], ref[0][order])}  grad rel {((o2[1] - ref[1]).abs().max() / ref[1].abs().max()).item():.2e}', flush=True)
