"""GPU, two or more devices (skipped on a one-GPU box): the in-kernel peer-memory exchange of the data-parallel
step against torch.distributed, one process per GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = 0.0068


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from dair_pll_b200 import ops, parallel, synthetic
    from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem
    comm = parallel.PeerComm(dev)
    # (1) the stand-alone exchange, many epochs, ranks drifting
    for k in range(200):
        v = torch.arange(17, dtype=torch.float64, device=dev) * (rank + 1) + k
        got = comm.all_reduce_sum(v)
        want = sum(torch.arange(17, dtype=torch.float64) * (r + 1) + k for r in range(world))
        assert torch.equal(got.cpu(), want), (rank, k)
        if k % 17 == rank:
            torch.cuda._sleep(2_000_000)
    # (2) the sharded step: global mean and gradients identical on every rank and equal to the one-GPU result
    s = MultibodyLearnableSystem({'cube': os.path.join(ROOT, 'dair_pll_b200', 'assets', 'cube.urdf')}, DT).to(dev)
    n = 40001
    x = synthetic.cube_states(n, seed=71, device=dev)
    with torch.no_grad():
        traj, _ = s.simulate(x.unsqueeze(-2), torch.zeros(n, 1, device=dev), 1)
    xp = synthetic.perturb_next_state(traj[:, 1], seed=72)
    full = s.contactnets_loss(x, None, xp)
    full.mean().backward()
    ref = [p.grad.clone() for p in s.parameters()]
    ref_mean = full.mean().item()
    del full                                   # no autograd graph may outlive its step (CUDA graph capture below)
    lo, hi = parallel.shard_bounds(n, world, rank)
    s.data_parallel = comm
    params = list(s.parameters())

    def step():
        for p in params:
            p.grad = None
        mean = s.contactnets_loss(x[lo:hi], None, xp[lo:hi]).mean()
        mean.backward()
        return mean
    mean = step().item()
    assert abs(mean - ref_mean) <= 1e-13 * abs(ref_mean)
    for p, r in zip(params, ref):
        assert torch.allclose(p.grad, r, rtol=1e-12, atol=1e-18)
    # general upstream gradient: summed over ranks by the stand-alone exchange
    for p in params:
        p.grad = None
    w = torch.linspace(0.5, 1.5, n, dtype=torch.float64, device=dev)
    (s.contactnets_loss(x[lo:hi], None, xp[lo:hi]) * w[lo:hi]).sum().backward()
    gw = [p.grad.clone() for p in params]
    s.data_parallel = None
    for p in params:
        p.grad = None
    (s.contactnets_loss(x, None, xp) * w).sum().backward()
    for a, p in zip(gw, params):
        assert torch.allclose(a, p.grad, rtol=1e-11, atol=1e-18)
    s.data_parallel = comm
    torch.cuda.synchronize(dev)
    # (3) the whole step, exchange included, as ONE CUDA graph
    graphed = parallel.GraphedStep(step, dev)
    for _ in range(20):
        m = graphed()
    assert abs(m.item() - ref_mean) <= 1e-13 * abs(ref_mean)
    flat = torch.cat([p.grad.reshape(-1) for p in params] + [m.detach().reshape(1)]).cpu()
    gathered = [None] * world
    dist.all_gather_object(gathered, flat)
    assert all(torch.equal(g, gathered[0]) for g in gathered)      # identical bits on every rank
    comm.check()
    comm.close()
    if rank == 0:
        torch.save(flat, out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_peer_memory_exchange_two_ranks(tmp_path):
    with socket.socket() as sock:
        sock.bind(('127.0.0.1', 0))
        port = sock.getsockname()[1]
    out = str(tmp_path / 'flat.pt')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert os.path.exists(out)
