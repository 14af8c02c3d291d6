"""Support points on the tensor cores (csrc/cn_icnn_tc.cu) against the FP64 layer path (dpll_icnn_* + library GEMMs):
agreement and kernel time.  CUDA events; not the bench."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dair_pll_b200 import ops  # noqa: E402
from dair_pll_b200.deep_support_function import HomogeneousICNN  # noqa: E402

dev = torch.device('cuda', 0)
torch.manual_seed(0)
net = HomogeneousICNN(2, 256, scale=0.06).to(dev)
ws = [net.input_weights[0].detach(), net.input_weights[1].detach(), net.hidden_weights[0].detach(), net.output_weight.detach()]


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


sizes = [int(a) for a in sys.argv[1:]] or [100, 128, 1000, 20000, 1 << 20]
for D in sizes:
    d = torch.randn(D, 3, dtype=torch.float64, device=dev)
    d = d / d.norm(dim=-1, keepdim=True)
    ref = ops.icnn_support_forward(d, *ws, 0.5)[0]
    got = ops.icnn_support_points_tc(d, *ws, 0.5)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item() / scale
    bad = ((got - ref).abs().amax(-1) > 1e-9 * scale).sum().item()
    print(f'D = {D}: max |p_tc - p_fp64| / max|p| = {err:.3e}  rows off by more than 1e-9: {bad}', flush=True)
    if D >= 20000:
        ms_ref = timeit(lambda: ops.icnn_support_forward(d, *ws, 0.5))
        ms_tc = timeit(lambda: ops.icnn_support_points_tc(d, *ws, 0.5))
        print(f'   forward: FP64 layer path {ms_ref:.3f} ms   tensor-core kernel {ms_tc:.3f} ms')
