#!/usr/bin/env python
"""Benchmark of the hot path: ContactNets loss + backward, cube, fp64 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--dtype f64|f32] [--impl reference]

One step = one pass of ``system.contactnets_loss(x, u, x_plus).mean().backward()`` over one
synthetic batch of B state pairs per GPU.  Prints ONE JSON line (rank 0).
For N > 1 launch with ``python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N``:
one process per GPU, the batch sharded by sample (each rank its own B samples: weak scaling),
NCCL all-reduce of the 15-double gradient/loss buffer inside every step.

``--impl reference`` times the reference's CPU path instead: the oracle port
(oracle/contactnets_oracle.py: the reference's algorithm restated in batched fp64 PyTorch +
the C cone-QP solver standing in for sappy) on all host cores, on a bounded sample of the
same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

DT = 0.0068
# Algorithmic FLOPs per sample (FMA = 2): F = F0 + F_IT * newton_iters
F0_CUBE, FIT_CUBE = 4600.0, 2100.0
BYTES_PER_SAMPLE = {torch.float64: 2 * 13 * 8 + 8, torch.float32: 2 * 13 * 4 + 4}
METRIC = 'ContactNets loss+backward samples/s, cube'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=1 << 20, help='samples per GPU per step')
    ap.add_argument('--dtype', choices=['f64', 'f32'], default='f64')
    ap.add_argument('--impl', choices=['b200', 'reference'], default='b200')
    ap.add_argument('--cpu-sample', type=int, default=65536, help='samples in the CPU baseline step')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-secondary', action='store_true', help='skip the other BASELINE.json configs')
    ap.add_argument('--no-graph', action='store_true', help='time eager launches instead of the CUDA-graph replay')
    ap.add_argument('--variant', type=int, default=0, help='0 = wavefront loss kernel, 1 = one sample per thread')
    return ap.parse_args()


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        clocks, reasons, smax, power = [], set(), None, []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                clocks.append(float(r[1]))
                smax = float(r[2])
                power.append(float(r[3]))
            except ValueError:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        clocks.sort()
        med = clocks[len(clocks) // 2] if clocks else None
        return {'sm_mhz': med, 'sm_max_mhz': smax, 'reasons': sorted(reasons), 'samples': len(clocks),
                'power_w_max': max(power) if power else None}


def make_system(device, dtype):
    from dair_pll_b200 import synthetic
    from dair_pll_b200.inertia import InertialParameterConverter as IPC
    from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem
    system = MultibodyLearnableSystem({'cube': os.path.join(ROOT, 'dair_pll_b200', 'assets', 'cube.urdf')}, DT)
    pi, fr, half = synthetic.cube_learnables_perturbed(0)
    system.load_state_dict({
        'multibody_terms.lagrangian_terms.inertial_parameters': IPC.pi_cm_to_theta(pi),
        'multibody_terms.contact_terms.friction_params': fr,
        'multibody_terms.contact_terms.geometries.0.length_params': half.reshape(1, 3)})
    return system.to(device)


def make_batch(system, batch, seed, device, dtype):
    """x: seeded synthetic states; x_plus: one learnable-system step + measurement noise."""
    from dair_pll_b200 import synthetic
    x = synthetic.cube_states(batch, seed=seed, device=device, dtype=torch.float64)
    with torch.no_grad():
        traj, _ = system.simulate(x.unsqueeze(-2), torch.zeros(batch, 1, device=device), 1)
    xp = synthetic.perturb_next_state(traj[:, 1], seed=seed + 7919)
    return x.to(dtype).contiguous(), xp.to(dtype).contiguous()


def _time_gpu(fn, device, reps, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(device)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize(device)
    return s.elapsed_time(e) / reps


def secondary_configs(device):
    """The other BASELINE.json configs, measured through the public API on device-resident inputs
    (CUDA events; context for the headline number, not part of it)."""
    from dair_pll_b200 import synthetic
    from dair_pll_b200.multibody_learnable_system import MultibodyLearnableSystem
    out = {}
    # config 1: cube loss + backward at B = 65,536, fp64 and the fp32 variant
    system = make_system(device, torch.float64)
    x, xp = make_batch(system, 65536, 11, device, torch.float64)

    from dair_pll_b200 import parallel
    reducer = parallel.GradientAllReduce(list(system.parameters()), device, 1)

    def cube_step(xx, xxp):
        reducer.zero()
        mean = system.contactnets_loss(xx, None, xxp).mean()
        mean.backward()
        return reducer(mean)
    for name, (xx, xxp) in {'f64': (x, xp), 'f32_storage': (x.float(), xp.float())}.items():
        ms_eager = _time_gpu(lambda: cube_step(xx, xxp), device, 10)
        graphed = parallel.GraphedStep(lambda: cube_step(xx, xxp), device)
        ms = _time_gpu(graphed, device, 20)
        out[f'cube_loss_backward_B65536_{name}'] = {'ms': ms, 'samples_per_s': 65536 / ms * 1e3, 'eager_ms': ms_eager}
    # config 4: rollout, 4,096 cube tosses x 80 steps
    x0 = synthetic.cube_states(4096, seed=5, device=device)
    carry = torch.zeros(4096, 1, device=device)

    def roll():
        with torch.no_grad():
            system.simulate(x0.unsqueeze(-2), carry, 80)
    ms = _time_gpu(roll, device, 5)
    out['cube_rollout_4096x80_f64'] = {'ms': ms, 'steps_per_s': 4096 * 80 / ms * 1e3}
    # config 3: elbow with learned (ICNN, width 256) geometry, loss + backward at B = 262,144
    torch.manual_seed(0)
    elbow = MultibodyLearnableSystem({'elbow': os.path.join(ROOT, 'dair_pll_b200', 'assets', 'elbow_mesh.urdf')}, DT).to(device)
    Be = 262144
    xe = synthetic.elbow_states(Be, seed=3, device=device)
    with torch.no_grad():
        te, _ = elbow.simulate(xe.unsqueeze(-2), torch.zeros(Be, 1, device=device), 1)
    xpe = synthetic.perturb_next_state(te[:, 1], seed=4, n_q=8)

    def elbow_step():
        for p in elbow.parameters():
            p.grad = None
        elbow.contactnets_loss(xe, None, xpe).mean().backward()
    ms = _time_gpu(elbow_step, device, 3, warmup=1)
    out['elbow_mesh_loss_backward_B262144_f64'] = {
        'ms': ms, 'samples_per_s': Be / ms * 1e3,
        'note': 'support-function networks: three (1.05M x 256 x 256) FP64 products per network on cuBLAS (torch.matmul) with the memory-bound layers fused in the dpll_icnn_* kernels; elbow loss kernel with warp-level triage / solve passes'}
    return out


def cpu_reference_step(batch, threads):
    """Builds the oracle-port step (loss + backward) on the host; returns a callable."""
    from dair_pll_b200 import synthetic
    from oracle import contactnets_oracle as co
    from oracle.callables import CUBE_TREE, TreeCallables
    from oracle.cone_qp import OracleSAPSolver
    torch.set_num_threads(threads)
    calls = TreeCallables(CUBE_TREE)
    pi, fr, half = synthetic.cube_learnables_perturbed(0)
    P = co.OracleParams(co.pi_cm_to_theta(pi), fr.clone(), [half.reshape(1, 3).clone()]).requires_grad_()
    solver = OracleSAPSolver(nthreads=threads)
    x = synthetic.cube_states(batch, seed=0)
    with torch.no_grad():
        xp = synthetic.perturb_next_state(co.sim_step(calls, P, x, DT, solver), seed=7919)

    def step():
        for t in P.leaves():
            t.grad = None
        loss = co.contactnets_loss(calls, P, x, xp, DT, solver)
        loss.mean().backward()
        return loss
    return step


def time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.cpu_sample
    step = cpu_reference_step(B, threads)
    # bounded: keep the whole run within a few minutes
    steps = max(1, min(args.steps, 10))
    warmup = max(1, min(args.warmup, 2))
    sec = time_cpu(step, steps, warmup)
    value = B / sec
    sample = f'{B} cube state pairs per step (same generator/parameters as the GPU arm), {steps} steps'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC + ' fp64', 'value': value, 'unit': 'samples/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'cube_contactnets_loss_backward_B{args.batch}_per_gpu', 'batch_per_gpu': args.batch,
                   'dt': DT, 'eps': 1e-3, 'cpu_sample_per_step': B,
                   'note': 'same workload as the GPU arm; each CPU step is a bounded sample of it'},
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': threads, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
        return
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch N>1 with torch.distributed.run --nproc-per-node N')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # keep stdout to the single JSON line
        dist.init_process_group('nccl', device_id=device)
    dtype = torch.float64 if args.dtype == 'f64' else torch.float32
    from dair_pll_b200 import ops, parallel
    ops.set_loss_variant(args.variant)

    system = make_system(device, dtype)
    B = args.batch
    x, xp = make_batch(system, B, seed=rank, device=device, dtype=dtype)
    u = torch.zeros(B, 0, device=device, dtype=dtype)
    params = [p for p in system.parameters()]
    reducer = parallel.GradientAllReduce(params, device, world)

    def step_local():
        reducer.zero()
        loss = system.contactnets_loss(x, u, xp)
        mean = loss.mean()
        mean.backward()
        return reducer.stage(mean)    # flat device buffer [grads..., loss] of this rank

    def step_resident():
        step_local()
        return reducer.reduce()       # ONE all-reduce(sum)/world over NCCL

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for _ in range(steps):
            fn()
        stop.record()
        barrier()
        ms = torch.tensor([start.elapsed_time(stop)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    # ---- device-resident throughput (value) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_eager = timed(step_resident, args.steps, max(args.warmup, 3))
    if args.no_graph:
        ms_step = ms_eager
    else:
        # the same step, captured once into a CUDA graph and replayed (parallel.GraphedStep)
        graphed = parallel.GraphedStep(step_local, device)

        def step_graphed():
            graphed()
            return reducer.reduce()
        ms_step = timed(step_graphed, args.steps, max(args.warmup, 3))
    # the timed region lasts ~10 ms, shorter than one nvidia-smi query: keep the same step running (untimed) for
    # about half a second so that the clock / throttle samples are taken under this load
    load_fn = step_resident if args.no_graph else step_graphed
    for _ in range(max(1, min(5000, int(500.0 / ms_step)))):      # same count on every rank (ms_step is the max over ranks)
        load_fn()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    value = world * B / (ms_step * 1e-3)

    # ---- end to end through the public API with HOST buffers ----
    xh, xph = x.cpu().pin_memory(), xp.cpu().pin_memory()
    loader = parallel.HostBatchPipeline(system, device, dtype, B)
    out_host = torch.empty(reducer.numel, dtype=torch.float64).pin_memory()

    def step_e2e():
        reducer.zero()
        mean = loader.loss_mean_from_host(xh, xph)
        mean.backward()
        flat = reducer(mean)
        out_host.copy_(flat, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the user reads the loss / gradients on the host
    ms_e2e = timed(step_e2e, max(3, args.steps // 2), 3)
    e2e_value = world * B / (ms_e2e * 1e-3)
    h2d = 2 * B * 13 * x.element_size()
    d2h = out_host.numel() * 8

    # ---- kernel-only timing for the roofline (CUDA events around the raw launch) ----
    inertia, mu, half = system._cube_params(dtype)
    inertia, mu, half = inertia.detach(), mu.detach(), half.detach()
    _, _, _, _, iters = ops.cube_loss_raw(x, xp, inertia, mu, half, DT, 1e-3, want_iters=True)
    mean_iters = iters.double().mean().item()

    def kernel_only():
        ops.cube_loss_raw(x, xp, inertia, mu, half, DT, 1e-3)
    ms_kernel = timed(kernel_only, args.steps, 3)
    flops_per_sample = F0_CUBE + FIT_CUBE * mean_iters
    achieved_tflops = B * flops_per_sample / (ms_kernel * 1e-3) / 1e12
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    # the arithmetic is fp64 for both storage types (the fp32 variant is fp32 STORAGE), so is the roofline
    peak_flops = ops.fma_peak(torch.float64, device, sms * 8, 200000)
    hbm_gbs = B * BYTES_PER_SAMPLE[dtype] / (ms_kernel * 1e-3) / 1e9
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')))
        key = f"cube_loss_wf_kernel_{args.dtype}_B{B}"
        if args.variant == 0 and key in tr:
            traffic = tr[key]['dram_bytes_read'] + tr[key]['dram_bytes_write']     # one ncu --set full capture, per launch
    except (OSError, ValueError, KeyError):
        pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        Bc = args.cpu_sample
        sec = time_cpu(cpu_reference_step(Bc, threads), 3, 1)
        cpu_baseline = {'value': Bc / sec, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                        'sample': f'{Bc} cube state pairs per step x 3 steps, oracle port (batched fp64 torch + C cone-QP '
                                  f'solver, {threads} threads), same generator/parameters'}

    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = secondary_configs(device)

    line = {
        'metric': METRIC + (' fp64' if dtype == torch.float64 else ' fp32'),
        'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'cube_contactnets_loss_backward_B{B}_per_gpu', 'batch_per_gpu': B,
                   'global_batch': world * B, 'dt': DT, 'eps': 1e-3, 'parallelism': f'dp{world}',
                   'l2_policy': f'inputs larger than L2 ({2 * B * 13 * x.element_size() / 1e6:.0f} MB per step vs 126 MB)',
                   'mean_newton_iters': mean_iters, 'storage': args.dtype,
                   'step': 'eager launches' if args.no_graph else 'CUDA graph replay of the public-API step (rank-local part) + NCCL all-reduce',
                   'eager_ms_per_step': ms_eager},
        'e2e': {'value': e2e_value, 'unit': 'samples/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h},
        'gpu_launches': 3 * args.steps,   # per step: parameter preparation + loss/backward + reduce/chain rule
                                          # (loss.mean() and its backward reuse that launch, ops.BatchLoss)
        'roofline': {'bound': 'fp64_cuda_core',
                     'achieved': achieved_tflops, 'peak': peak_flops / 1e12, 'unit': 'TFLOP/s',
                     'frac': achieved_tflops / (peak_flops / 1e12), 'traffic': traffic,
                     'algorithmic_bytes': B * BYTES_PER_SAMPLE[dtype],
                     'kernel': 'cube_loss_wf_kernel' if args.variant == 0 else 'cube_loss_kernel', 'kernel_ms': ms_kernel,
                     'flops_per_sample': flops_per_sample,
                     'peak_source': 'measured in this run: dpll_fma_peak (dependent-free FMA chains, all SMs)',
                     'hbm': {'achieved_gbs': hbm_gbs, 'peak_gbs': hbm_peak, 'frac': hbm_gbs / hbm_peak,
                             'bytes_per_sample': BYTES_PER_SAMPLE[dtype]}},
        'cpu_baseline': cpu_baseline,
        'clocks': clocks,
        'secondary': secondary,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
