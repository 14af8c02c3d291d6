"""Pinned host -> device copy rates on this box: contiguous rows against cudaMemcpy2DAsync of a column band (the velocity half
of a state batch, 48 of every 104 bytes).  Decides whether the end-to-end step should skip the columns the loss never reads."""
import torch
from cuda.bindings import runtime as rt

B = 1 << 20
dev = torch.device('cuda', 0)
h = torch.randn(B, 13, dtype=torch.float64).pin_memory()
d = torch.empty(B, 13, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
kind = rt.cudaMemcpyKind.cudaMemcpyHostToDevice


def t_ms(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def copy2d(col0, ncol, rows=B, row0=0):
    err, = rt.cudaMemcpy2DAsync(d.data_ptr() + row0 * 104 + col0 * 8, 104, h.data_ptr() + row0 * 104 + col0 * 8, 104, ncol * 8,
                                rows, kind, stream)
    assert err == rt.cudaError_t.cudaSuccess, err


full = t_ms(lambda: d.copy_(h, non_blocking=True))
print(f'contiguous {B} x 104 B: {full:.3f} ms  {B * 104 / full / 1e6:.1f} GB/s')
for col0, ncol in ((0, 13), (7, 6), (0, 7), (4, 9)):
    ms = t_ms(lambda: copy2d(col0, ncol))
    print(f'2D columns [{col0}, {col0 + ncol}) = {ncol * 8} of 104 B per row: {ms:.3f} ms  {B * ncol * 8 / ms / 1e6:.1f} GB/s of payload')
ms = t_ms(lambda: [copy2d(7, 6, B // 8, c * (B // 8)) for c in range(8)])
print(f'2D columns [7, 13) in 8 row chunks: {ms:.3f} ms')
d.zero_()
copy2d(7, 6)
torch.cuda.synchronize()
assert torch.equal(d[:, 7:].cpu(), h[:, 7:]) and (d[:, :7] == 0).all()
print('2D copy lands where it should')
