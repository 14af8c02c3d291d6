"""Sample-level data parallelism and host->device pipelining around the loss kernels.

The ContactNets loss is independent per (x, x_plus) pair (multibody_learnable_system.py:118-124);
the only coupling is ``loss.mean()`` (drake_experiment.py:222-223) and the summed parameter
gradient.  So: one process per GPU, each rank evaluates its shard of the batch, and ONE NCCL
all-reduce moves a flat buffer [d/d theta (n_b x 10), d/d friction (n_g), d/d lengths ..., loss]
(15 doubles for the cube).  Parameter ``.grad`` tensors are views into that buffer, so autograd
accumulates straight into the NCCL send buffer and nothing is packed or copied per step.
"""
from typing import Callable, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous shard [lo, hi) of n samples for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bind_host_to_gpu(device_index: int) -> Optional[str]:
    """Pins the calling process to the CPUs of the NUMA node its GPU hangs off (NVML's ideal affinity), so that pinned
    staging buffers allocated afterwards are first-touched in the memory next to that GPU's PCIe root.  With one process
    per GPU the host->device copies of all ranks otherwise share whichever socket the processes happened to start on.
    Returns a short description, or None if NVML is unavailable or the affinity cannot be set (containers with a
    restricted cpuset): the step then simply runs unbound."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        import os
        cpus = sorted(os.sched_getaffinity(0))
        return f'{len(cpus)} cpus [{cpus[0]}..{cpus[-1]}]'
    except Exception:      # noqa: BLE001 -- best effort by design
        return None


class PeerComm:
    """Communicator of the in-kernel gradient exchange (``dpll_comm_*``, csrc/cn_comm.cuh): every rank owns a
    small buffer in its HBM that all peers map through CUDA IPC, and the reduction kernel of the loss launch
    pushes its [gradient | loss sum | count] row into every peer's buffer over NVLink and sums the world's
    rows -- the data-parallel step issues no NCCL collective.  ``torch.distributed`` (any backend) is used
    once, here, to hand the IPC handles around.

    Attach it with ``system.data_parallel = PeerComm(device)``: ``loss.mean()`` / ``loss.sum()`` of the returned
    batch loss and the parameter gradients then cover the samples of all ranks (identical bits on every rank).
    One communicator serves one stream at a time; every rank must make the same sequence of calls."""

    def __init__(self, device: torch.device, group: Optional[dist.ProcessGroup] = None) -> None:
        import ctypes
        from dair_pll_b200 import _lib
        self._lib = _lib
        lib = _lib.load()
        self.device = device
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        nbytes = lib.dpll_comm_handle_bytes()
        mine = ctypes.create_string_buffer(nbytes)
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.dpll_comm_create(self.rank, self.world, ctypes.byref(handle), mine), 'dpll_comm_create')
            self.handle = handle
            if self.world > 1:
                gathered = [None] * self.world
                dist.all_gather_object(gathered, bytes(mine.raw), group=group)
                blob = ctypes.create_string_buffer(b''.join(gathered), nbytes * self.world)
            else:
                blob = mine
            _lib.check(lib.dpll_comm_connect(self.handle, blob), 'dpll_comm_connect')
        if self.world > 1:
            dist.barrier(group=group)        # every rank has mapped every buffer before the first exchange

    def all_reduce_sum(self, t: Tensor, scale: float = 1.0) -> Tensor:
        """Stand-alone exchange of up to 32 doubles (one tiny kernel, no NCCL): returns scale * sum over ranks."""
        buf = t.detach().to(torch.float64).contiguous().clone()
        with torch.cuda.device(self.device):
            rc = self._lib.load().dpll_comm_allreduce_f64(self.handle, buf.data_ptr(), buf.numel(), float(scale),
                                                          torch.cuda.current_stream().cuda_stream)
        self._lib.check(rc, 'dpll_comm_allreduce')
        return buf.to(t.dtype).view_as(t)

    def check(self) -> None:
        """Host-side check (synchronises): raises if an exchange timed out waiting for a peer."""
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.load().dpll_comm_error(self.handle), 'dpll_comm')

    def close(self) -> None:
        if self.handle is not None:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                self._lib.load().dpll_comm_destroy(self.handle)
            self.handle = None


class GradientAllReduce:
    """Flat gradient buffer shared with the parameters' ``.grad`` + mean all-reduce over ranks."""

    def __init__(self, params: List[Tensor], device: torch.device, world: int,
                 group: Optional[dist.ProcessGroup] = None) -> None:
        self.params = list(params)
        self.world = world
        self.group = group
        self.numel = sum(p.numel() for p in self.params) + 1          # + 1 slot for the loss
        dtype = self.params[0].dtype
        self.flat = torch.zeros(self.numel, dtype=dtype, device=device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)        # autograd accumulates in place
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def rebind(self) -> None:
        """Re-attach ``.grad`` views (after something set ``p.grad = None``)."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat[off:].data_ptr():
                g = self.flat[off:off + p.numel()].view_as(p)
                if p.grad is not None:
                    g.copy_(p.grad)
                else:
                    g.zero_()
                p.grad = g
            off += p.numel()

    def stage(self, local_mean_loss: Tensor) -> Tensor:
        """Local half: makes sure the gradients live in ``flat`` and appends this rank's mean loss."""
        self.rebind()
        self.flat[-1:].copy_(local_mean_loss.detach().reshape(1))
        return self.flat

    def reduce(self) -> Tensor:
        """Collective half: ONE all-reduce of ``flat`` (sum / world) over NCCL (gloo in the CPU tests)."""
        if self.world > 1:
            if dist.get_backend(self.group) == 'nccl':
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)     # one kernel: sum / world
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.flat.mul_(1.0 / self.world)
        return self.flat

    def __call__(self, local_mean_loss: Tensor) -> Tensor:
        """Averages gradients (already in ``flat``) and the loss over ranks; returns ``flat``."""
        self.stage(local_mean_loss)
        return self.reduce()


class GraphedStep:
    """The rank-local part of a training step (forward, backward, staging of the flat gradient buffer)
    captured into a CUDA graph and replayed.

    A step over device-resident data is one long kernel plus a dozen microsecond-sized launches (parameter
    preparation, fixed-order reduction, ``mean`` and its backward); replaying them as a graph removes the
    Python/launch overhead between them, which is what bounds small batches.  ``fn`` must be free of host
    synchronisation and of collectives (the NCCL all-reduce is issued after the replay, on the same stream)
    and must write its results into the same tensors on every call (``GradientAllReduce.stage`` does).  The
    kernels take the capturing stream through the C ABI.  Drop every result of earlier EAGER steps that still carries
    an autograd graph before capturing: a live graph keeps the parameters' gradient accumulators, which stay bound to the
    stream that step ran on, and autograd's end-of-backward stream synchronisation then fails inside the capture."""

    def __init__(self, fn: Callable[[], Tensor], device: torch.device, warmup: int = 3) -> None:
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def __call__(self) -> Tensor:
        self.graph.replay()
        return self.out


class HostBatchPipeline:
    """``loss.mean()`` of a batch that lives in (pinned) HOST memory: the batch is cut into
    chunks whose host->device copies run on a side stream and overlap the loss kernels of the
    previous chunks.  The result is differentiable w.r.t. the system's parameters."""

    def __init__(self, system, device: torch.device, dtype: torch.dtype, max_batch: int, chunks: Optional[int] = None) -> None:
        self.system = system
        self.device = device
        # a chunk's launch lasts at least its longest Newton chain (~0.1 ms), so small batches are cut into fewer pieces
        self.chunks = chunks if chunks is not None else max(1, min(8, max_batch // 65536))
        chunks = self.chunks
        n_x = system.space.n_x
        self.x_dev = torch.empty((max_batch, n_x), dtype=dtype, device=device)
        self.xp_dev = torch.empty((max_batch, n_x), dtype=dtype, device=device)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.events = [torch.cuda.Event() for _ in range(chunks)]

    def loss_mean_from_host(self, x_host: Tensor, xp_host: Tensor) -> Tensor:
        return self.loss_sum_from_host(x_host, xp_host) / x_host.shape[0]

    def loss_sum_from_host(self, x_host: Tensor, xp_host: Tensor) -> Tensor:
        B = x_host.shape[0]
        assert B <= self.x_dev.shape[0]
        main = torch.cuda.current_stream(self.device)
        self.copy_stream.wait_stream(main)            # previous step's kernels are done with the buffers
        bounds = [shard_bounds(B, self.chunks, c) for c in range(self.chunks)]
        with torch.cuda.stream(self.copy_stream):
            for c, (lo, hi) in enumerate(bounds):
                self.x_dev[lo:hi].copy_(x_host[lo:hi], non_blocking=True)
                self.xp_dev[lo:hi].copy_(xp_host[lo:hi], non_blocking=True)
                self.events[c].record(self.copy_stream)
        total = None
        for c, (lo, hi) in enumerate(bounds):
            if hi == lo:
                continue
            main.wait_event(self.events[c])
            part = self.system.contactnets_loss(self.x_dev[lo:hi], None, self.xp_dev[lo:hi]).sum()
            total = part if total is None else total + part
        return total

    def capture_step(self, x_host: Tensor, xp_host: Tensor, params: List[Tensor], denom: float, out_host: Tensor,
                     comm: Optional[PeerComm] = None, warmup: int = 3) -> 'GraphedStep':
        """The whole host-to-host training step as ONE CUDA graph: the chunked copies out of the pinned staging buffers
        ``x_host`` / ``xp_host`` (copy stream), each chunk's loss launch as soon as its rows have landed, the backward, the
        exchange of [parameter gradients | loss / denom] over ``comm`` (in-kernel, peer memory) and the copy of that vector
        into pinned ``out_host``.  A step is then ``graph(); stream.synchronize()`` -- no per-chunk Python or launch cost, which
        is what bounded the multi-GPU end-to-end step (8 eager chunk launches against 27 MB of copies per rank).  The staging
        buffers are re-filled in place by the caller between steps."""
        assert x_host.is_pinned() and xp_host.is_pinned() and out_host.is_pinned()
        params = list(params)

        def fn():
            for p in params:
                p.grad = None
            mean = self.loss_sum_from_host(x_host, xp_host) / denom
            mean.backward()
            flat = torch.cat([p.grad.reshape(-1) for p in params] + [mean.detach().reshape(1)])
            if comm is not None:
                flat = comm.all_reduce_sum(flat)
            out_host.copy_(flat, non_blocking=True)
            return flat
        return GraphedStep(fn, self.device, warmup)
