"""A/B timing of elbow loss kernel builds: DAIR_PLL_B200_LIB=<variant .so> python tools/time_elbow_ab.py -- 262,144 pairs in
natural order (static ranges) and in cost order (dynamic chunks), kernel only, CUDA events."""
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
B = 262144
x = synthetic.elbow_states(B, seed=0, device=dev)
with torch.no_grad():
    traj, _ = s.simulate(x.unsqueeze(-2), torch.zeros(B, 1, device=dev), 1)
xp = synthetic.perturb_next_state(traj[:, 1], seed=1, n_q=8)
inertia, mu, hl, kin = (t.detach() for t in s._elbow_params(torch.float64, dev))


def t_ms(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(reps):
        fn()
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / reps


out = ops.elbow_loss_raw(x, xp, inertia, mu, hl, kin, 0.0068, 1e-3, want_iters=True)
nat = t_ms(lambda: ops.elbow_loss_raw(x, xp, inertia, mu, hl, kin, 0.0068, 1e-3))
order = torch.argsort(out[4], descending=True, stable=True)
xo, xpo = x[order].contiguous(), xp[order].contiguous()
cost = t_ms(lambda: ops.elbow_loss_raw(xo, xpo, inertia, mu, hl, kin, 0.0068, 1e-3, flags=ops.LOSS_DYNAMIC))
print(f'{os.environ.get("DAIR_PLL_B200_LIB", "default")}: natural {nat:.4f} ms  cost order {cost:.4f} ms  loss_sum {out[2].item():.12e}',
      flush=True)
