"""Where the time of the support-network step goes (config 3 sizes): the width x width products alone (cuBLAS
DGEMM), and the whole ICNNSupport forward + backward; CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dair_pll_b200.deep_support_function import ICNNSupport  # noqa: E402

dev = torch.device('cuda', 0)
D, W = 262144 * 4, 256
torch.manual_seed(0)
A = torch.randn(D, W, dtype=torch.float64, device=dev)
Bm = torch.randn(W, W, dtype=torch.float64, device=dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


fl = 2.0 * D * W * W
ms = timeit(lambda: A @ Bm)
print(f'(D x 256) @ (256 x 256), D = {D}: {ms:.2f} ms  {fl / ms / 1e9:.1f} TFLOP/s')
ms = timeit(lambda: A.t() @ A)
print(f'(256 x D) @ (D x 256): {ms:.2f} ms  {fl / ms / 1e9:.1f} TFLOP/s')
ms = timeit(lambda: torch.where(A > 0, 1.0, 0.5).to(A.dtype))
print(f'mask pass over (D x 256): {ms:.2f} ms  {3 * D * W * 8 / ms / 1e6:.0f} GB/s-equivalent')
ws = [torch.randn(3, W, dtype=torch.float64, device=dev).requires_grad_(), torch.randn(3, W, dtype=torch.float64, device=dev).requires_grad_(),
      (torch.randn(W, W, dtype=torch.float64, device=dev) / W).requires_grad_(), torch.randn(W, dtype=torch.float64, device=dev).requires_grad_()]
d = torch.randn(D, 3, dtype=torch.float64, device=dev)
d = d / d.norm(dim=-1, keepdim=True)
gp = torch.randn(D, 3, dtype=torch.float64, device=dev)


def fwd_bwd():
    for w in ws:
        w.grad = None
    p = ICNNSupport.apply(d, ws[0], ws[1], ws[2], ws[3], 0.5)
    p.backward(gp)


ms = timeit(fwd_bwd, 3)
print(f'ICNNSupport forward + backward, one network, D = {D}: {ms:.2f} ms  ({3 * fl / ms / 1e9:.1f} TFLOP/s on the three products)')
