#!/usr/bin/env python
"""Benchmark: PDs embedded per second (BASELINE.json metric), B200 arm and CPU reference arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one projection direction (PD) through the hot path in ONE library call (esper_pd_embed_fused): pairwise RMS
distances (K0+K1) -> epsilon sweep (K2) -> tanh fit (MINPACK lmdif restated in C) -> Gaussian kernel / degree
normalisation (K3) -> leading eigenpairs (K4).  Workload at every N: BASELINE config 2, one synthetic PD of N=2000 images,
250x250 px, tau=5, SNR 0.1, 1001-point eps sweep, top-10 eigenpairs; with N GPUs every rank embeds
its own PDs (config 3 sharding, no data-path collective) -> weak scaling.

`value`  : inputs resident in HBM when the timed region starts (each PD stack is 500 MB > the 126 MB L2).
`e2e`    : the same steps with HOST (pinned) image stacks: H2D of every stack and D2H of the distance matrix / sweep /
           eigenpairs inside the timed region.
After the measurement (outside every timed region): `c5_rot` at N=1 (BASELINE config 5 shape: 126 manifolds [2000, 15], 8-D
rotation sweep, evaluations/s) and `c4_rowblock` at N>1 (BASELINE config 4: one PD of N=50 000 x 128^2 split by row blocks).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_IMG, BOX, TAU, SNR, K_EIG = 2000, 250, 5, 0.1, 10
P = BOX * BOX
LOG_EPS = np.arange(-100, 100.2, 0.2)
METRIC = 'PDs embedded per second (N=2000 images, 250x250 px; dist + eps sweep + top-10 eigenpairs)'
UNIT = 'PD/s'
WORKLOAD = 'C2: single synthetic PD, N=2000 images of 250x250 px (tau=5, SNR 0.1), eps sweep (1001) + top-10 eigenvectors'


def env_rank():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_sample(stack, n_loop=64, n_eps_sample=8, full_eig=True):
    """Time the reference algorithm (oracle port, as written) on a bounded sample and scale to one PD.

    distances : the per-pair Python loop of Distances.py:96-104 on the first `n_loop` images, scaled by
                the pair count to N=2000;
    eps sweep : GaussianBandwidth.py:36-44 on the full N=2000 distance matrix for `n_eps_sample` of the
                1001 epsilons spread over the grid (sequential float64 sum), scaled by 1001/n_eps_sample;
    kernel + degree + normalisation + full SVD (DiffusionMaps.py:109-169) at full size, not scaled.
    Returns (seconds per PD, detail dict).
    """
    from oracle import esper_oracle as O
    n = stack.shape[0]
    t0 = time.perf_counter()
    X = O.standardize_stack(stack[:n_loop])
    O.distances_as_written(X)
    t_loop = time.perf_counter() - t0
    pairs_s, pairs_full = n_loop * (n_loop - 1) // 2, n * (n - 1) // 2
    t_dist = t_loop * pairs_full / pairs_s
    t0 = time.perf_counter()
    D = O.distances_gram_f64(O.standardize_stack(stack))            # only to obtain D for the later stages
    t_vec_dist = time.perf_counter() - t0
    ks = np.linspace(0, len(LOG_EPS) - 1, n_eps_sample).astype(int)
    t0 = time.perf_counter()
    O.eps_sweep_logsum(D, LOG_EPS[ks], sequential=True)
    t_sweep = (time.perf_counter() - t0) * len(LOG_EPS) / len(ks)
    t_eig = 0.0
    if full_eig:
        t0 = time.perf_counter()
        O.dmap_embed(D, 0.2834)
        t_eig = time.perf_counter() - t0
    total = t_dist + t_sweep + t_eig
    detail = dict(dist_loop_s=t_dist, sweep_s=t_sweep, kernel_svd_s=t_eig, vectorised_dist_s=t_vec_dist,
                  vectorised_total_s=t_vec_dist + t_sweep + t_eig)
    return total, detail


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([i.get('num_threads', 1) for i in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every host core for its BLAS calls."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def run_reference(args):
    rank, _, world = env_rank()
    if rank != 0:
        return 0
    from manifoldem_esper_b200 import synth
    use_all_host_threads()
    stack = synth.synthetic_pd(pd=1, box=BOX, n_side=20, tau=TAU, snr=SNR)
    times, detail, walls = [], {}, []
    for it in range(args.warmup + args.steps):
        w0 = time.perf_counter()
        t, detail = cpu_sample(stack, n_loop=48, n_eps_sample=4, full_eig=True)
        if it >= args.warmup:
            times.append(t); walls.append(time.perf_counter() - w0)
    sec = float(np.mean(times))
    wall_per_step = float(np.mean(walls))
    sample = ('oracle port of the reference as written, EXTRAPOLATED per step: pair loop on 48 images scaled x%d by pair count to N=2000, '
              'eps sweep on 4 of 1001 eps scaled x250, kernel+degree+full SVD at full size; %.1f s of CPU work per step' % (
                  N_IMG * (N_IMG - 1) // 2 // (48 * 47 // 2), wall_per_step))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': 1.0 / sec, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'host': 'cpu', 'extrapolated': True, 'same_config': False,
                   'measured_s_per_step': wall_per_step,
                   'note': 'a full reference step takes ~3.5 minutes of Python pair loop and element-wise sums; every step here times a bounded '
                           'sample of it and scales by pair / epsilon count (ms_per_step and value are the scaled figures)'},
        'cpu_baseline': {'value': 1.0 / sec, 'unit': UNIT, 'cores': blas_threads(), 'kind': 'port', 'sample': sample,
                         'detail_s_per_pd': detail, 'vectorised_value': 1.0 / detail['vectorised_total_s'],
                         'vectorised_note': 'same math with a float64 dgemm Gram instead of the pair loop (not the reference as written)'},
        'e2e': {'value': 1.0 / sec, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.proc, self.lines, self.idx = None, [], gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Statistics over the samples that arrived inside [t_begin, t_end + 0.1 s] (the timed region); the sampler is started
        earlier (nvidia-smi needs ~0.1 s to deliver its first line), and when a short timed region caught no sample the
        samples of the serial pass directly before it are used and the fact is noted."""
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        if not self.lines:                     # nothing yet at all: take one late sample rather than none (marked in `window`)
            t_wait = time.time()
            while not self.lines and time.time() - t_wait < 3.0:
                time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        inside = [ln for (t, ln) in self.lines if t_begin is None or (t_begin <= t <= (t_end or t) + 0.1)]
        window = 'timed region'
        if not inside:
            inside, window = [ln for (_, ln) in self.lines], 'serial pass + timed region (no sample arrived inside the short timed region)'
        self.window = window
        for ln in inside:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


NATIVE_FIT = True      # stage 2b through esper_tanh_fit (MINPACK lmdif restated in C, no interpreter lock); --scipy-fit: curve_fit


def embed_step(ctx, GB, coef, a0, n, p, imgs=None, download=False, out=None):
    """One PD through stages 1-3 on `ctx`.  imgs=None -> the ctx-resident stack; out = caller buffer for D."""
    if NATIVE_FIT:                                           # ONE library call: esper_pd_embed_fused
        r = ctx.pd_embed_fused(imgs, LOG_EPS, a0, k=K_EIG, n=n, P=p, out=out if download else None)
        return r.get('D'), r['eps'], r['val'], r['vec'], r['iters']
    D = ctx.dist_rms(imgs, n=n, P=p, download=download, out=out)   # --scipy-fit: the staged calls with scipy's curve_fit in between
    L = ctx.eps_sweep(None, coef, 10.0, m=n)
    popt, resnorm, R2 = GB.fit(LOG_EPS, L, a0, native=False)
    eps = float(np.exp(-(popt[1] / popt[0])))
    val, vec, iters = ctx.dmap_embed(None, eps, k=K_EIG, m=n)
    return D, eps, val, vec, iters


def c5_rotation_sweep(ctx):
    """BASELINE config 5 shape on one GPU: 126 manifolds [2000, 15], 8-D subspace -> device preparation (normalisation + 28 conic
    fits per PD), host partner selection, device sweep (4 dependent (evaluate all, first arg-max) launch pairs)."""
    from manifoldem_esper_b200 import Rotation_Histograms as RH, synth
    n_pd, dim = 126, 8
    U0s = [synth.random_manifold(m=2000, ncol=15, seed=s) for s in range(n_pd)]
    RH.run_many(U0s[:4], dim=dim, PCA=True, CTF=True, ctx=ctx)                 # warm-up
    ctx.set_profiling(True); ctx.reset_profiling()
    t0 = time.perf_counter()
    preps = RH.prepare_many(U0s, dim=dim, PCA=True, CTF=True, ctx=ctx)
    t_prep = time.perf_counter() - t0
    t0 = time.perf_counter()
    RH.sweep_many(preps, dim, ctx=ctx)
    t_sweep = time.perf_counter() - t0
    k_rot, n_rot = ctx.kernel_ms('rot')
    k_con, _ = ctx.kernel_ms('conic')
    ctx.set_profiling(False)
    evals = sum(len(p_['comboList']) * len(p_['Rij_final']) * 160 for p_ in preps)
    return {'workload': 'C5 shape: 126 synthetic manifolds [2000, 15], dim 8 (2_Manifold_Rotations sweep, 160 angles x 4 Rij x subspaces)',
            'evaluations': int(evals), 'sweep_wall_s': t_sweep, 'sweep_kernel_ms': k_rot, 'sweep_launches': n_rot,
            'prepare_wall_s': t_prep, 'prepare_kernel_ms': k_con,
            'evaluations_per_s': evals / t_sweep if t_sweep > 0 else None,
            'evaluations_per_s_kernels': evals / (k_rot * 1e-3) if k_rot > 0 else None,
            'reference_evaluations_per_s_per_core': 1140.0, 'reference_note': 'SURVEY section 6: 0.88 ms per evaluation in the reference loop'}


def c4_rowblock(ctx, local_rank, rank, world, dev, stream):
    """BASELINE config 4: ONE PD of N=50 000 images of 128x128 px split by distance-matrix row blocks over the ranks
    (rowblock.py; all-reduce of the sweep tables, all-gather of the degrees, one all-gather of the Lanczos vector per step).
    The stack is generated on every GPU by esper_tau_snr_stack (same seed -> same bits)."""
    import torch
    import torch.distributed as dist
    from manifoldem_esper_b200 import rowblock, synth
    n, P, k = 50000, 128 * 128, 10
    out = {'workload': 'C4: one PD, N=50000 images of 128x128 px, row blocks over %d GPUs' % world}
    with torch.cuda.stream(stream):
        ctx.dist_config(split_k=8)
        ctx.tau_snr_stack(synth.pristine_states(box=128, n_side=20, pd=1), 125, 0.1, seed=1000, download=False)
        ctx.sync()
        res = None
        for fused in (False, True):
            be = rowblock.CudaBackend(ctx, None, n, P, dev, fused=fused)
            comm = rowblock.TorchComm(dev, stream=be.stream)
            rowblock.embed_oversized(be, comm, n, k=k, moments=True)                  # warm-up (allocations, NCCL channels, IPC mappings)
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            r = rowblock.embed_oversized(be, comm, n, k=k, moments=True)
            torch.cuda.synchronize(); dist.barrier()
            dt = time.perf_counter() - t0
            tm = torch.tensor([dt] + [r['timings'][q] for q in ('dist_s', 'sweep_fit_s', 'markov_s', 'lanczos_s')], dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            key = 'fused_symv_allgather' if fused else 'nccl_allgather'
            out[key] = {'total_s': float(tm[0]), 'dist_s': float(tm[1]), 'sweep_fit_s': float(tm[2]), 'markov_s': float(tm[3]),
                        'lanczos_s': float(tm[4]), 'lanczos_steps': int(r['steps']), 'eps': float(r['eps']), 'sweep_path': r.get('sweep_path')}
            if not fused:
                res, be0 = r, be
            else:
                be.comm_close()
        # residuals |Ms v - lambda v| against the resident row blocks and orthonormality of the returned vectors
        vec = torch.from_numpy(res['vec']).to(dev)
        w_loc = torch.zeros(be0.w_local.shape[0], dtype=torch.float64, device=dev)
        lam = np.sqrt(res['val'])
        resid = []
        for i in range(k):
            q = vec[:, i].contiguous()
            ctx.symv_rows_d(be0.nrows, n, q.data_ptr(), w_loc.data_ptr()); ctx.sync()
            rr = w_loc[:be0.nrows] - lam[i] * q[be0.row0:be0.row0 + be0.nrows]
            s_ = (rr * rr).sum()
            dist.all_reduce(s_)
            resid.append(float(s_.sqrt()))
        out['max_residual'] = max(resid)
        out['max_orthogonality_error'] = float((vec.T @ vec - torch.eye(k, dtype=torch.float64, device=dev)).abs().max())
        out['lambda'] = [float(x) for x in lam]
        ctx.dist_config(split_k=0)
    return out


def run_ours(args):
    import torch
    from manifoldem_esper_b200 import GaussianBandwidth as GB, _capi, synth
    rank, local_rank, world = env_rank()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)
    sys.setswitchinterval(2e-4)          # pipeline threads hand the interpreter lock over quickly (default 5 ms)
    # pinned stacks and host threads on the NUMA node of this rank's GPU (matters for the e2e figure at 8 GPUs)
    from manifoldem_esper_b200 import hostbind
    numa = hostbind.bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else {'bound': False}
    stream = torch.cuda.Stream()
    ctx = _capi.Context(local_rank, stream=stream.cuda_stream)

    # two distinct PD stacks per rank in pinned host memory (inputs of successive steps differ)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                        # started before the stacks are generated: nvidia-smi can take a second for its first line on a
                                               # fresh box (a default run once ended with no sample at all); samples are selected by time later
    n_stacks = 2
    ms_serial = 0.0
    pinned = []
    for s in range(n_stacks):
        st = synth.synthetic_pd(pd=1 + rank * n_stacks + s, box=BOX, n_side=20, tau=TAU, snr=SNR).reshape(N_IMG, P)
        t = torch.from_numpy(st).pin_memory()
        pinned.append(t)
    coef = 1. / (2. * np.exp(LOG_EPS))
    a0 = np.array([[-0.2], [0.3], [-0.4], [0.1]])          # fixed tanh-fit start (the reference draws it at random)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident arm: stacks already in HBM.  Two PDs are in flight per GPU (one host thread + ctx +
    # stream each) so that the host-side tanh fit / Lanczos convergence tests of one PD overlap the
    # kernels of the other; a serial pass on one stream provides clean per-kernel timings. ----
    NF = max(1, args.inflight)
    # mode 0 (one persistent kernel for the whole solve, convergence test on the device) for both arms since the merged schedule:
    # its 120 CTAs (15 clusters of 8) leave 28 SMs to the kernels of the other PDs in flight (measured: 391 PD/s with 3 in flight
    # against 380 PD/s with the streamed barrier kernel and 6 in flight).
    eig_mode = args.eig_mode if args.eig_mode >= 0 else 0
    ctxs, streams = [ctx], [stream]
    for _ in range(NF - 1):
        st_ = torch.cuda.Stream()
        ctxs.append(_capi.Context(local_rank, stream=st_.cuda_stream)); streams.append(st_)
    for i, c in enumerate(ctxs):
        c.upload_images(pinned[i % n_stacks].numpy())
        c.dmap_config(eig_mode)
        c.gram_config(1 if args.gram_f16 else 0)       # 3xFP16 operand planes (library default) / --gram-tf32: 3xTF32
        c.sweep_config(1 if args.sweep_moments else 0) # moment sweep (library default) / --sweep-general
        c.sync()
    iters_used = []

    def pipelined(n_steps, e2e, nf=None):
        """n_steps PDs over nf ctxs (thread t takes steps t, t + nf, ...); returns device ms."""
        errs = []
        nf = NF if nf is None else nf

        def worker(t):
            try:
                torch.cuda.set_device(local_rank)
                for s in range(t, n_steps, nf):
                    if e2e:
                        embed_step(ctxs[t], GB, coef, a0, N_IMG, P, imgs=pinned[t % n_stacks].numpy(), download=True,
                                   out=d_out[t].numpy())
                    else:
                        out = embed_step(ctxs[t], GB, coef, a0, N_IMG, P)
                        iters_used.append(out[4])
            except Exception as e:                     # noqa: BLE001
                errs.append(e)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for st_ in streams[1:]:
            st_.wait_event(e0)
        th = [threading.Thread(target=worker, args=(t,)) for t in range(nf)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        for st_ in streams[1:]:
            eb = torch.cuda.Event()
            eb.record(st_)
            streams[0].wait_event(eb)
        e1.record(streams[0])
        barrier()
        if errs:
            raise errs[0]
        return e0.elapsed_time(e1)

    for w in range(args.warmup):
        embed_step(ctxs[w % NF], GB, coef, a0, N_IMG, P)
    pipelined(NF, False)
    # serial profiled pass (per-kernel CUDA-event durations; single stream, nothing overlapping)
    for c in ctxs:
        c.set_profiling(True); c.reset_profiling()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for s in range(args.steps):
        embed_step(ctx, GB, coef, a0, N_IMG, P)
    ev1.record(stream)
    barrier()
    ms_serial = ev0.elapsed_time(ev1)
    kern = {}
    for name in ('prep', 'gram', 'dist_finalize', 'refine', 'sweep', 'markov', 'lanczos'):
        tot, cnt = 0.0, 0
        for c in ctxs:
            a, b_ = c.kernel_ms(name)
            tot += a; cnt += b_
        kern[name] = (tot, cnt)
    for c in ctxs:
        c.set_profiling(False)

    # the timed region: K steps, two in flight
    iters_used.clear()
    launches0 = sum(c.launch_count() for c in ctxs)
    for c in ctxs:
        c.set_profiling(2); c.reset_profiling()          # CUDA events around the Gram launches of the timed region only
    t_begin = time.time()
    ms = pipelined(args.steps, False)
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    gram_live = [c.kernel_ms('gram') for c in ctxs]
    gram_live_ms = sum(a for a, _ in gram_live) / max(sum(b_ for _, b_ in gram_live), 1)
    for c in ctxs:
        c.set_profiling(False)
    launches = sum(c.launch_count() for c in ctxs) - launches0

    # ---- end-to-end arm: host stacks -> H2D -> stages -> D2H of D, sweep, eigenpairs (NF in flight) ----
    # The caller owns the output buffers; here they are page-locked like the input stacks.
    d_out = [torch.empty((N_IMG, N_IMG), dtype=torch.float64).pin_memory() for _ in range(NF)]
    # The end-to-end arm is bound by the 500 MB host-to-device copy of every stack: the eigensolver that needs no host round
    # trips while those copies run (mode 0: one persistent kernel with the convergence test on the device) is the better fit
    # there (measured: 106 PD/s against 91 PD/s with the streamed barrier kernel whose batches the host paces).
    eig_mode_e2e = args.eig_mode if args.eig_mode >= 0 else 0
    for c in ctxs:
        c.dmap_config(eig_mode_e2e)
    # two in flight: the copy of one stack runs under the kernels of the other; more only split the link (measured: 108 PD/s with 2,
    # 94 - 108 with 3, 105 with 4)
    NF_E2E = min(NF, max(1, args.inflight_e2e))
    pipelined(NF_E2E, True, NF_E2E)
    ms_e2e = pipelined(args.steps, True, NF_E2E)

    # raw host->device rate of this box for one 500 MB stack (what bounds the e2e figure); all ranks copy at the same time, so at
    # N > 1 this is the per-GPU share of the box's aggregate pinned-copy ceiling, not the link rate of one GPU
    barrier()
    probe = torch.empty((N_IMG, P), dtype=torch.float32, device='cuda')
    probe.copy_(pinned[0], non_blocking=True)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(4):
        probe.copy_(pinned[0], non_blocking=True)
    p1.record()
    torch.cuda.synchronize()
    h2d_gbs = 4 * N_IMG * P * 4 / (p0.elapsed_time(p1) * 1e-3) / 1e9
    del probe

    # dense TF32 library GEMM on this box (the peak the Gram kernel is compared with; reported next to the
    # MEASURED_PEAKS-based denominator, never instead of it)
    tf32_probe = None
    if rank == 0:
        prev = torch.backends.cuda.matmul.allow_tf32
        try:
            torch.backends.cuda.matmul.allow_tf32 = True
            a_ = torch.randn((8192, 8192), dtype=torch.float32, device='cuda')
            b_ = torch.randn((8192, 8192), dtype=torch.float32, device='cuda')
            for _ in range(3):
                torch.matmul(a_, b_)
            torch.cuda.synchronize()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            for _ in range(10):
                torch.matmul(a_, b_)
            q1.record()
            torch.cuda.synchronize()
            tf32_probe = 10 * 2.0 * 8192 ** 3 / (q0.elapsed_time(q1) * 1e-3) / 1e12
            del a_, b_
        except Exception:                                  # noqa: BLE001 -- the probe is informational
            tf32_probe = None
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    # ---- after the measurement: the other BASELINE configurations, once (never inside a timed region of the metric) ----
    extra = {}
    if not args.no_extras:
        for c in ctxs[1:]:
            c.close()
        ctxs = ctxs[:1]
        try:
            if world == 1:
                extra['c5_rot'] = c5_rotation_sweep(ctx)
            else:
                extra['c4_rowblock'] = c4_rowblock(ctx, local_rank, rank, world, torch.device('cuda', local_rank), stream)
        except Exception as e:                                 # noqa: BLE001 -- reported, never fatal for the metric line
            extra['c5_rot' if world == 1 else 'c4_rowblock'] = {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    tms = torch.tensor([ms, ms_e2e, float(launches), ms_serial], dtype=torch.float64, device='cuda')
    if world > 1:
        import torch.distributed as dist
        mx = tms.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tms.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, ms_e2e, launches, ms_serial = float(mx[0]), float(mx[1]), int(sm[2]), float(mx[3])
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return 0

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    # The kernel's duration comes from the serial pass (one PD at a time, the kernel alone on the GPU for ~0.6 ms): the
    # denominator is the BURST figure of MEASURED_PEAKS.json (round-1 verdict); the in-step duration is reported against the
    # sustained figure next to it.
    bf16_burst, bf16_sus = peaks.get('bf16_tflops'), peaks.get('bf16_tflops_sustained')
    peak_src = 'measured (MEASURED_PEAKS.json bf16_tflops, burst: the kernel is timed alone in a serial pass; kind::f16 runs at the bf16 rate)'
    if not bf16_burst:
        bf16_burst, bf16_sus = 1590.0, 1400.0
        peak_src = 'fallback (B200_PROFILING.md: 1.59 PFLOP/s bf16 burst; kind::f16 runs at the bf16 rate)'
    gram_kernel_name = 'gram_wide_kernel<F16> (tcgen05 kind::f16, 3xFP16 split of 22 bits, 128x256 tiles on the upper triangle)'
    div = 1.0
    if not args.gram_f16:
        gram_kernel_name = 'gram_wide_kernel (tcgen05 kind::tf32, 3xTF32, 128x256 tiles on the upper triangle)'
        div = 2.0
        peak_src = peak_src.replace('kind::f16 runs at the bf16 rate', 'dense TF32 is half the bf16 rate on tcgen05')
    tf32_peak = bf16_burst / div
    sus_peak = (bf16_sus or bf16_burst) / div
    gram_ms = kern['gram'][0] / max(kern['gram'][1], 1)               # serial pass of K steps on one stream: CUDA events on that stream
    alg_flops = 2.0 * N_IMG * N_IMG * P                           # GEMM convention, SURVEY 8(d)
    n_tiles = (N_IMG + 127) // 128
    n_wide = sum((n_tiles + 1) // 2 - i // 2 for i in range(n_tiles))      # 128x256 tiles covering the upper triangle
    kpad = 64 if args.gram_f16 else 32
    issued_flops = 3.0 * n_wide * 128 * 256 * ((P + kpad - 1) // kpad * kpad) * 2.0
    achieved = alg_flops / (gram_ms * 1e-3) / 1e12 if gram_ms > 0 else 0.0
    traffic, traffic_note = None, 'no ncu capture of this kernel in profiles/gram_traffic.json'
    try:
        gt = json.load(open(os.path.join(ROOT, 'profiles', 'gram_traffic.json')))
        key = 'f16' if args.gram_f16 else 'tf32'
        if key in gt:
            traffic, traffic_note = gt[key].get('dram_bytes_per_launch'), gt[key].get('note')
    except Exception:
        pass
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    lan_ms_per_step = kern['lanczos'][0] / max(args.steps, 1)
    per_step = {k: round(v[0] / max(args.steps, 1), 4) for k, v in kern.items()}

    pds = world * args.steps
    line = {
        'metric': METRIC, 'value': pds / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16x3+f64' if args.gram_f16 else 'tf32x3+f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'pd_per_rank_per_step': 1, 'sharding': 'independent PDs per rank, no data-path collective',
                   'pipelining': '%d PDs in flight per GPU (one host thread / ctx / stream each)' % NF, 'serial_ms_per_step': ms_serial / args.steps,
                   'api': 'esper_pd_embed_fused (one C call per PD)' if NATIVE_FIT else 'staged calls with scipy curve_fit',
                   'eigensolver': {0: 'mode 0: one persistent cluster kernel, one exchange per Lanczos step, on-device convergence test', 1: 'mode 1: streamed barrier kernel, two grids per SM',
                                   2: 'mode 2: streamed, no speculative batch', 3: 'mode 3: resident barrier kernel'}.get(eig_mode),
                   'l2': 'inputs larger than L2 (500 MB stack per PD, two alternating stacks per rank)',
                   'lanczos_steps': iters_used[-1] if iters_used else None, 'numa_bind_rank0': numa,
                   'tanh_fit': 'esper_tanh_fit (MINPACK lmdif in C)' if NATIVE_FIT else 'scipy.optimize.curve_fit',
                   'kernels': {'gram_f16': bool(args.gram_f16), 'sweep_moments': bool(args.sweep_moments)}},
        'e2e': {'value': pds / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(N_IMG * P * 4),
                'd2h_bytes_per_step': int(N_IMG * N_IMG * 8 + len(LOG_EPS) * 8 + K_EIG * 8 + N_IMG * K_EIG * 8),
                'ms_per_step': ms_e2e / args.steps, 'h2d_probe_gbs': round(h2d_gbs, 2),
                'h2d_floor_ms_per_step': round(N_IMG * P * 4 / (h2d_gbs * 1e9) * 1e3, 3),
                'eigensolver_mode': eig_mode_e2e, 'pds_in_flight': NF_E2E,
                'note': 'bounded by the host->device copy of the 500 MB stack; h2d_probe_gbs = bare pinned cudaMemcpyAsync rate per GPU with all %d '
                        'ranks copying at once (the box ceiling at this N), h2d_floor_ms_per_step = 500 MB at that rate, frac_of_h2d_floor = floor / measured' % world},
        'gpu_launches': launches,
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'kernel': gram_kernel_name,
                     'achieved': achieved, 'peak': tf32_peak, 'unit': 'TFLOP/s', 'frac': achieved / tf32_peak if tf32_peak else None,
                     'traffic': traffic, 'peak_source': peak_src, 'flops_per_launch': alg_flops,
                     'flops_convention': '2*N^2*P (full GEMM); the kernel issues 3 MMAs (hi*hi + hi*lo + lo*hi) per product on 72 tiles of 128x256 covering the upper triangle',
                     'traffic_note': traffic_note,
                     'issued_tflops': issued_flops / (gram_ms * 1e-3) / 1e12 if gram_ms > 0 else 0.0,
                     'frac_issued': (issued_flops / (gram_ms * 1e-3) / 1e12) / tf32_peak if gram_ms > 0 and tf32_peak else None,
                     'frac_in_step_vs_sustained': (alg_flops / (gram_live_ms * 1e-3) / 1e12) / sus_peak if gram_live_ms > 0 and sus_peak else None,
                     'peak_sustained': sus_peak,
                     'avg_launch_ms': gram_ms, 'launches_timed': kern['gram'][1],
                     'tf32_gemm_probe_tflops': tf32_probe,
                     'avg_launch_ms_pipelined': gram_live_ms, 'launches_pipelined': sum(b_ for _, b_ in gram_live),
                     'timing': 'avg_launch_ms: CUDA events on the launching stream in a serial pass of the same K steps (one PD at a time); '
                               'avg_launch_ms_pipelined: the same events inside the timed region with several PDs in flight -- an event pair '
                               'there also spans the time the kernel waits for SMs held by other PDs\' kernels'},
        'kernel_ms_per_step': per_step,
        'roofline_hbm': {'bound': 'hbm', 'kernel': 'lanczos (symv + reorthogonalisation), Ms L2-resident at N=2000',
                         'achieved': (np.mean(iters_used) * N_IMG * N_IMG * 8 / (lan_ms_per_step * 1e-3) / 1e9) if lan_ms_per_step > 0 else 0.0,
                         'peak': hbm_peak, 'unit': 'GB/s'},
    }
    line['roofline_hbm']['frac'] = line['roofline_hbm']['achieved'] / hbm_peak if hbm_peak else None
    line['e2e']['frac_of_h2d_floor'] = round(line['e2e']['h2d_floor_ms_per_step'] / line['e2e']['ms_per_step'], 3) if ms_e2e > 0 else None
    if extra:
        line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        t0 = time.time()
        use_all_host_threads()
        sec, detail = cpu_sample(pinned[0].numpy().reshape(N_IMG, BOX, BOX), n_loop=400, n_eps_sample=250, full_eig=True)
        line['cpu_baseline'] = {
            'value': 1.0 / sec, 'unit': UNIT, 'cores': blas_threads(), 'kind': 'port',
            'sample': 'oracle port of the reference as written: pair loop on 400 of 2000 images (79800 pairs) scaled x25.05 by pair '
                      'count, eps sweep on 250 of 1001 eps scaled x4.0, kernel+degree+full SVD at full size; %.0f s of CPU work' % (time.time() - t0),
            'detail_s_per_pd': detail, 'vectorised_value': 1.0 / detail['vectorised_total_s'],
            'vectorised_note': 'same math restated with a float64 dgemm Gram instead of the pair loop (not the reference)'}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=48)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-numa-bind', action='store_true', help='do not restrict ranks to the CPUs of their GPU\'s NUMA node')
    ap.add_argument('--eig-mode', type=int, default=-1, help='esper_dmap_config mode (-1 = the library default, mode 0)')
    ap.add_argument('--gram-tf32', action='store_true', help='3xTF32 operand planes for the Gram kernel (esper_gram_config 0) instead of the default 3xFP16')
    ap.add_argument('--scipy-fit', action='store_true', help='tanh fit with scipy.optimize.curve_fit (the reference\'s call) instead of esper_tanh_fit')
    ap.add_argument('--sweep-general', action='store_true', help='general sweep kernel (esper_sweep_config 0) instead of the default moment sweep')
    ap.add_argument('--inflight', type=int, default=3, help='PDs in flight per GPU (host threads / ctxs / streams)')
    ap.add_argument('--inflight-e2e', type=int, default=2, help='PDs in flight per GPU in the end-to-end arm (copy-bound)')
    ap.add_argument('--no-extras', action='store_true', help='skip the post-measurement c5_rot (N=1) / c4_rowblock (N>1) runs')
    args = ap.parse_args()
    args.gram_f16, args.sweep_moments = not args.gram_tf32, not args.sweep_general
    global NATIVE_FIT
    NATIVE_FIT = not args.scipy_fit
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
