"""Config 3: elbow with learned (ICNN, width 256) geometry, loss + backward at B = 262,144; CUDA events."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dair_pll_b200 import synthetic  # noqa: E402
from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=262144)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--graph', action='store_true', help='also time the step replayed as one CUDA graph')
a = ap.parse_args()
dev = torch.device('cuda', 0)
torch.manual_seed(0)
s = MultibodyLearnableSystem({'elbow': os.path.join(ROOT, 'dair_pll_b200', 'assets', 'elbow_mesh.urdf')}, 0.0068).to(dev)
x = synthetic.elbow_states(a.batch, seed=0, device=dev)
with torch.no_grad():
    traj, _ = s.simulate(x.unsqueeze(-2), torch.zeros(a.batch, 1, device=dev), 1)
xp = synthetic.perturb_next_state(traj[:, 1], seed=1, n_q=8)


def step():
    for p in s.parameters():
        p.grad = None
    loss = s.contactnets_loss(x, None, xp)
    loss.mean().backward()
    return loss.detach()     # (a live autograd graph from an eager step would pin the parameters' gradient
    # accumulators to the stream it ran on and break a later capture)


for _ in range(2):
    loss = step()
torch.cuda.synchronize()
st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st.record()
for _ in range(a.reps):
    step()
en.record()
torch.cuda.synchronize()
ms = st.elapsed_time(en) / a.reps
flops = a.batch * 8 * 4 * 2 * 256 * 256     # 4 DGEMMs of (D x 256 x 256), D = 8 directions per sample
print(f'elbow-mesh loss+backward B={a.batch}: {ms:.2f} ms  {a.batch / ms / 1e3:.2f} M samples/s  '
      f'ICNN GEMM {flops / ms / 1e9:.1f} TFLOP/s-equivalent  loss mean {loss.mean().item():.6e}')

if a.graph:
    from dair_pll_b200 import parallel
    g = parallel.GraphedStep(step, dev)
    for _ in range(2):
        g()
    torch.cuda.synchronize()
    st.record()
    for _ in range(10):
        g()
    en.record()
    torch.cuda.synchronize()
    msg = st.elapsed_time(en) / 10
    grads = [p.grad.clone() for p in s.parameters()]
    step()
    torch.cuda.synchronize()
    err = max(((p.grad - q).abs().max() / p.grad.abs().max().clamp(min=1e-300)).item() for p, q in zip(s.parameters(), grads))
    print(f'one CUDA graph: {msg:.2f} ms  {a.batch / msg / 1e3:.2f} M samples/s   max rel grad diff graph vs eager {err:.2e}')
