"""One oversized PD split by distance-matrix row blocks over the GPUs of a node (BASELINE config 4).

The reference cannot run this case (a 50 000 x 50 000 float64 distance matrix and an O(m^3) SVD,
DiffusionMaps.py:166); the arithmetic per element is the same as in the single-GPU path.  Rank g owns the rows
[row0, row0 + nrows) of D, A and Ms (128-row aligned).  Data-path collectives (the only ones):

    all-reduce(sum)  of the 1001 partial epsilon-sweep sums           once (twice with the threshold pass)
    all-gather       of the degrees (n float64)                        once
    all-gather       of the Lanczos vector w = Ms q (n float64)        once per Lanczos step

After the all-gather every rank orthogonalises the same replicated vector with the same deterministic
kernels, so no further reduction is needed.  On GPUs the per-step SYMV and its all-gather can run as ONE exchange
(`esper_symv_allgather_d` + `esper_comm_wait_d`): the SYMV kernel writes its rows of w as {value, step} packets (flag in the
data, no fences) into the rank's OWN buffer, CTA by CTA as they finish, and a gather kernel on every rank PULLS all n
packets -- its own from local memory, the others over NVLink through CUDA IPC peer mappings; `CudaBackend(fused=True)` selects
it, the NCCL all-gather remains as the reference implementation of the same exchange (`fused=False`; identical bits).

Stream contract: `CudaBackend` allocates its torch tensors and runs the NCCL collectives on the ctx's own stream
(`esper_stream`, wrapped as a `torch.cuda.ExternalStream`), so library kernels, torch fills and collectives are ordered
by stream order alone; `TorchComm(device, stream=backend.stream)` must be given the same stream.  `embed_oversized` is written against a small backend interface
so that the orchestration is exercised on CPU (gloo, numpy backend in tests/) and on GPUs (NCCL, CudaBackend).
"""
from __future__ import annotations

import numpy as np

from . import GaussianBandwidth

TILE = 128


def row_partition(n: int, rank: int, world: int):
    """(row0, nrows, rows_per_rank): contiguous blocks of whole 128-row tiles, equal padded size."""
    tiles = (n + TILE - 1) // TILE
    tpr = (tiles + world - 1) // world
    row0 = rank * tpr * TILE
    nrows = min(n, (rank + 1) * tpr * TILE) - row0
    if nrows < 1:
        raise ValueError('row_partition: %d images over %d ranks leaves rank %d without rows' % (n, world, rank))
    return row0, nrows, tpr * TILE


class TorchComm:
    """torch.distributed collectives used by the row-block path (backend nccl on GPUs, gloo on CPU)."""

    def __init__(self, device=None, stream=None):
        import contextlib
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device if device is not None else torch.device('cpu')
        # GPU: every collective and staging copy is enqueued under the backend's stream (see the stream contract above)
        self._on = (lambda: torch.cuda.stream(stream)) if stream is not None else contextlib.nullcontext

    def allreduce_sum(self, arr):
        with self._on():
            t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            return t.cpu().numpy()

    def allreduce_minmax(self, lo, hi):
        """(global min of lo, global max of hi) of two scalars."""
        with self._on():
            t = self.torch.tensor([-float(lo), float(hi)], dtype=self.torch.float64, device=self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            t = t.cpu().numpy()
        return -float(t[0]), float(t[1])

    def allgather_concat(self, arr, chunk):
        """Gather equal-size (zero padded to `chunk`) float64 blocks from all ranks into one array."""
        buf = np.zeros(chunk, dtype=np.float64)
        buf[:len(arr)] = arr
        with self._on():
            t = self.torch.from_numpy(buf).to(self.device)
            out = self.torch.empty(chunk * self.world, dtype=self.torch.float64, device=self.device)
            self.dist.all_gather_into_tensor(out, t)
            return out.cpu().numpy()

    def allgather_vec(self, out_tensor, in_tensor):
        with self._on():
            self.dist.all_gather_into_tensor(out_tensor, in_tensor)


class CudaBackend:
    """Row-block kernels of libesper_b200 on one GPU; vectors are torch CUDA tensors (NCCL-ready).

    All torch work of the backend (allocations, fills, collectives) is enqueued on `self.stream`, the ctx's own stream, so it
    is ordered with the library kernels; pass `stream=backend.stream` to `TorchComm`."""

    def __init__(self, ctx, stack, n, P, device, fused=False):
        import torch
        self.torch, self.ctx, self.device = torch, ctx, device
        self.stack, self.n, self.P = stack, n, P
        self.fused = bool(fused)
        self.stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=device)
        self._step = 0                                      # packet flags: monotone over the life of the backend

    def dist_rows(self, row0, nrows):
        self.row0, self.nrows = row0, nrows
        self.ctx.dist_rms_rows(self.stack, self.n, self.P, row0, nrows)

    def sweep_rows(self, coef, thr):
        return self.ctx.eps_sweep_rows(self.nrows, self.n, coef, thr)

    def moment_stats_rows(self):
        return self.ctx.moment_stats_rows(self.nrows, self.n)

    def moment_accum_rows(self, coef, thr, plan):
        return self.ctx.moment_accum_rows(self.nrows, self.n, coef, thr, plan)

    def degree_rows(self, eps):
        return self.ctx.degree_rows(self.nrows, self.n, eps)

    def markov_rows(self, eps, degree_all):
        self.ctx.markov_rows(self.nrows, self.n, eps, degree_all)

    def lanczos_alloc(self, max_steps, rows_per, world):
        t = self.torch
        with t.cuda.stream(self.stream):
            self.V = t.zeros((max_steps + 2, self.n), dtype=t.float64, device=self.device)
            self.w_local = t.zeros(rows_per, dtype=t.float64, device=self.device)
            self.w_full = t.zeros(rows_per * world, dtype=t.float64, device=self.device)
            self.ab = t.zeros((max_steps + 2, 2), dtype=t.float64, device=self.device)
        if self.fused and getattr(self, '_comm_key', None) != (rows_per, world):
            # one gather buffer per rank, mapped by every peer through CUDA IPC; set up once per backend and reused by
            # later embeddings (IPC mapping costs milliseconds per peer)
            import torch.distributed as dist
            # the exchange spins on flags written by the peers: every rank needs a GPU of its own (ranks sharing one
            # GPU are not guaranteed to run at the same time)
            devs = [None] * world
            dist.all_gather_object(devs, int(self.ctx.device))
            if len(set(devs)) != world or world > t.cuda.device_count():
                raise RuntimeError('fused SYMV + all-gather needs one GPU per rank (devices of the ranks: %s)' % devs)
            self.comm_close()
            handle = self.ctx.comm_alloc(rows_per * world, world)
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            self.ctx.comm_open(handles, dist.get_rank(), world)
            dist.barrier()                                  # nobody reads a buffer that is not mapped everywhere yet
            self._comm_key = (rows_per, world)               # a fresh buffer is zeroed; flags stay monotone anyway

    def lanczos_start(self, degree_all):
        self.ctx.lanczos_start_d(self.n, degree_all, self.V.data_ptr())

    def symv(self, j):
        self.ctx.symv_rows_d(self.nrows, self.n, self.V[j].data_ptr(), self.w_local.data_ptr())

    def symv_allgather(self, j):
        """Fused: w = Ms[rows, :] q_j stored into every rank's gather buffer; returns after enqueueing the wait."""
        self._step += 1                                   # advanced at the point of use: flags never repeat, also after a
        step = self._step                                 # failed embedding (the same sequence on every rank)
        self.ctx.symv_allgather_d(self.nrows, self.n, self.row0, self.V[j].data_ptr(), step)
        self._w_ptr = self.ctx.comm_wait_d(step, self.n)

    def fused_steps(self, j0, n_steps):
        """Steps j0 .. j0 + n_steps - 1 of the fused exchange enqueued by ONE library call (esper_lanczos_fused_steps_d): the host
        paces the iteration once per convergence test, not once per step."""
        step0 = self._step + 1
        self._step += n_steps                             # the same sequence of tags on every rank
        self.ctx.lanczos_fused_steps_d(self.nrows, self.n, self.row0, self.V.data_ptr(), self.ab.data_ptr(), j0, n_steps, step0)

    def ortho(self, j):
        w_ptr = self._w_ptr if self.fused else self.w_full.data_ptr()
        self.ctx.lanczos_ortho_d(self.n, j, self.V.data_ptr(), w_ptr, self.ab[j - 1].data_ptr())

    def comm_check(self):
        if self.fused:
            st = self.ctx.comm_status()
            if st:
                raise RuntimeError('peer-memory all-gather: an element of the gathered vector never arrived (watchdog)')

    def comm_close(self):
        """Unmap the peers' buffers (collective: every rank must call it)."""
        if getattr(self, '_comm_key', None) is not None:
            import torch.distributed as dist
            self.ctx.sync()
            dist.barrier()                                  # every rank is done reading its peers
            self.ctx.comm_close()
            self._comm_key = None

    def read_ab(self, steps):
        with self.torch.cuda.stream(self.stream):
            ab = self.ab[:steps].cpu().numpy()
        return ab[:, 0].copy(), ab[:, 1].copy()

    def ritz_vectors(self, steps, k, S):
        return self.ctx.ritz_vectors_d(self.n, steps, k, self.V.data_ptr(), S)

    def tridiag_ritz(self, alpha, beta, k, tol):
        from . import _capi
        return _capi.tridiag_ritz(alpha, beta, k, tol)

    def sync(self):
        self.ctx.sync()
        self.torch.cuda.synchronize()


def sweep_sums(backend, comm, coef, thr, moments=False):
    """sum_{ij : t < thr} exp(-t), t = coef[k] D_ij^2, over all ranks' row blocks -> (sums[k], 'general' | 'moments').

    General: every rank evaluates its block (`esper_eps_sweep_rows`), one all-reduce of the 1001 sums.  `moments=True`
    (opt-in; moment_sweep.py): all-reduce of the block statistics (min / max of the non-zero D^2, number of zeros), one
    plan for all ranks, all-reduce of the [32, 21] moment tables, evaluation on the host; falls back to the general sweep
    when D^2 spans more than 32 bins or the coefficients are not a non-increasing grid."""
    from . import moment_sweep as MSW
    if moments and hasattr(backend, 'moment_stats_rows') and MSW.usable(coef, thr):
        st = np.asarray(backend.moment_stats_rows(), dtype=np.float64)
        mn, mx = comm.allreduce_minmax(st[0], st[1])
        nz, nv = comm.allreduce_sum(st[2:4])
        status, pl = MSW.plan(mn, mx, nz, nv, coef, thr)
        if status == 0:
            tot = comm.allreduce_sum(np.asarray(backend.moment_accum_rows(coef, thr, pl), dtype=np.float64).reshape(-1))
            return MSW.finalize(pl, tot, coef, thr), 'moments'
    return comm.allreduce_sum(backend.sweep_rows(coef, thr)), 'general'


def embed_oversized(backend, comm, n, k=16, eps=None, a0=None, tol=0.0, max_steps=1024, check_every=16,
                    log_eps=None, verbose=False, moments=False):
    """Distances -> epsilon -> Markov rows -> leading k eigenpairs of one PD split over comm.world ranks.

    Returns dict(eps, val [k] = lambda^2, vec [n, k], steps, logSumWij, popt) (identical on every rank).
    `a0` (tanh-fit start, 4 values) must be the same on all ranks; default is a fixed start.
    `moments=True` selects the moment sweep (see sweep_sums).
    """
    import time
    rank, world = comm.rank, comm.world
    row0, nrows, rows_per = row_partition(n, rank, world)
    sync = getattr(backend, 'sync', lambda: None)
    timings = {}
    t0 = time.perf_counter()
    backend.dist_rows(row0, nrows)
    sync(); timings['dist_s'] = time.perf_counter() - t0; t0 = time.perf_counter()
    out = {}
    if eps is None:
        log_eps = np.arange(-100, 100.2, 0.2) if log_eps is None else np.asarray(log_eps, dtype=np.float64)
        e = np.exp(log_eps)
        # find_threshold (GaussianBandwidth.py:26-32) needs the global sum at the largest epsilon
        ss = comm.allreduce_sum(backend.sweep_rows(np.array([1. / (2. * np.max(e))]), np.inf))[0]
        thr = max(-np.log(0.01 * ss / n), 10)
        sums, sweep_path = sweep_sums(backend, comm, 1. / (2. * e), thr, moments=moments)
        with np.errstate(divide='ignore'):
            logSumWij = np.log(sums)
        out['sweep_path'] = sweep_path
        if a0 is None:
            a0 = np.array([[-0.2], [0.3], [-0.4], [0.1]])
        state = np.random.get_state()
        np.random.seed(12345)                     # the retry loop redraws a0 from the global RNG: same on all ranks
        popt, resnorm, R2 = GaussianBandwidth.fit(log_eps, logSumWij, a0)
        np.random.set_state(state)
        eps = float(np.exp(-(popt[1] / popt[0])))
        out.update(logSumWij=logSumWij, popt=popt, R2=R2)
    sync(); timings['sweep_fit_s'] = time.perf_counter() - t0; t0 = time.perf_counter()
    deg_all = comm.allgather_concat(backend.degree_rows(eps), rows_per)[:n]
    backend.markov_rows(eps, deg_all)
    sync(); timings['markov_s'] = time.perf_counter() - t0; t0 = time.perf_counter()

    max_steps = int(min(max_steps, n - 1))
    kw = k - 1
    backend.lanczos_alloc(max_steps, rows_per, world)
    backend.lanczos_start(deg_all)
    steps, converged, theta, S = 0, kw == 0, np.zeros(0), np.zeros((0, 0))
    next_check = min(max_steps, max(3 * (kw + 2) + 8, 32))
    while not converged and steps < max_steps:
        j = steps + 1
        if getattr(backend, 'fused', False) and hasattr(backend, 'fused_steps'):
            nb = max(1, min(next_check, max_steps) - steps)       # all steps up to the next convergence test in one library call
            backend.fused_steps(j, nb)
            steps += nb
        elif getattr(backend, 'fused', False):
            backend.symv_allgather(j)                             # SYMV writing wire-format packets + the pull
            backend.ortho(j)
            steps += 1
        else:
            backend.symv(j)                                       # w_local = Ms[rows, :] q_j
            comm.allgather_vec(backend.w_full, backend.w_local)   # the per-step collective
            backend.ortho(j)                                      # replicated, deterministic
            steps += 1
        if steps >= next_check or steps == max_steps:
            alpha, beta = backend.read_ab(steps)
            getattr(backend, 'comm_check', lambda: None)()
            converged, theta, S = backend.tridiag_ritz(alpha, beta, k, tol)
            breakdown = not (beta[-1] > 1e-14 * max(abs(alpha[0]), 1e-300))
            converged = converged or breakdown or steps >= n - 1
            next_check = steps + check_every
            if verbose and rank == 0:
                print('lanczos step %d converged=%s' % (steps, converged))
    if not converged:
        raise RuntimeError('row-block Lanczos did not converge in %d steps' % steps)
    vec = backend.ritz_vectors(steps, k, S) if kw > 0 else backend.ritz_vectors(0, 1, np.zeros((0, 0)))
    sync(); timings['lanczos_s'] = time.perf_counter() - t0
    val = np.concatenate(([1.0], theta[:kw] ** 2))
    out.update(eps=eps, val=val, vec=vec, steps=steps, row0=row0, nrows=nrows, timings=timings)
    return out
