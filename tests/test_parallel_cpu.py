"""CPU, world_size 2 over gloo: the host-side data-parallel plumbing (shard bounds, the flat
gradient buffer aliased by ``.grad``, mean all-reduce).  The kernels themselves need a GPU; here a
linear stand-in loss exercises exactly the code path bench.py uses around them."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dair_pll_b200.parallel import GradientAllReduce, shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    theta = torch.nn.Parameter(torch.arange(10, dtype=torch.float64).reshape(1, 10) / 10)
    fric = torch.nn.Parameter(torch.tensor([0.15, 1.0], dtype=torch.float64))
    red = GradientAllReduce([theta, fric], torch.device('cpu'), world)
    data = torch.randn(64, 12, dtype=torch.float64)
    lo, hi = shard_bounds(64, world, rank)
    for _ in range(2):                                   # two steps: buffer reuse + zeroing
        red.zero()
        w = torch.cat((theta.reshape(-1), fric))
        local = ((data[lo:hi] @ w) ** 2).mean()
        local.backward()
        flat = red(local).clone()
    assert theta.grad.data_ptr() == red.flat.data_ptr()  # .grad aliases the all-reduce buffer
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'flat.pt')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    flat = torch.load(out)
    torch.manual_seed(0)
    theta = (torch.arange(10, dtype=torch.float64).reshape(1, 10) / 10).requires_grad_()
    fric = torch.tensor([0.15, 1.0], dtype=torch.float64).requires_grad_()
    data = torch.randn(64, 12, dtype=torch.float64)
    loss = ((data @ torch.cat((theta.reshape(-1), fric))) ** 2).mean()   # equal shards: mean of means
    loss.backward()
    ref = torch.cat((theta.grad.reshape(-1), fric.grad, loss.detach().reshape(1)))
    assert torch.allclose(flat, ref, rtol=1e-13, atol=1e-15)
