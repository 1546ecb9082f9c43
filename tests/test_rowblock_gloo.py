"""CPU, world_size 2 over gloo: the oversized-PD row-block orchestration (manifoldem_esper_b200/rowblock.py)
with a numpy stand-in for the CUDA kernels.  The collectives (sweep-sum all-reduce, degree all-gather, per-step
Lanczos vector all-gather), the 128-row partition, the replicated deterministic orthogonalisation and the host
tridiagonal solver of libesper_b200 (esper_tridiag_ritz, CPU code) are the real ones."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from manifoldem_esper_b200 import _capi, rowblock

WORKER = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch
import torch.distributed as dist
from manifoldem_esper_b200 import rowblock, _capi
from oracle import esper_oracle as O

class NumpyBackend:
    """Reference arithmetic (oracle) on the owned row block; vectors are torch CPU tensors."""
    def __init__(self, D_full):
        self.Dfull, self.n = D_full, D_full.shape[0]
    def dist_rows(self, row0, nrows):
        self.row0, self.nrows = row0, nrows
        self.D = self.Dfull[row0:row0 + nrows]
    def sweep_rows(self, coef, thr):
        D2 = self.D * self.D
        return np.array([np.sum(np.exp(-(c * D2)[(c * D2) < thr])) for c in coef])
    def moment_stats_rows(self):
        q = (self.D * self.D).ravel()
        nzm = q == 0
        return np.array([q[~nzm].min() if np.any(~nzm) else np.inf, q.max(), float(nzm.sum()), float(q.size)])
    def moment_accum_rows(self, coef, thr, plan):
        import importlib.util
        spec = importlib.util.spec_from_file_location('sweep_moments_proto', os.path.join(%(root)r, 'tools', 'sweep_moments_proto.py'))
        proto = importlib.util.module_from_spec(spec); spec.loader.exec_module(proto)
        q = (self.D * self.D).ravel()
        return proto.accumulate_numpy(q, np.ones_like(q), plan, coef, thr)
    def degree_rows(self, eps):
        self.A = np.exp(-1. * ((self.D ** 2.) / (2. * eps)))
        return self.A.sum(axis=1)
    def markov_rows(self, eps, d):
        di = d[self.row0:self.row0 + self.nrows]
        self.Ms = (np.sqrt(1. / di)[:, None] * (self.A * (1. / d)[None, :])) * np.sqrt(d)[None, :]
    def lanczos_alloc(self, max_steps, rows_per, world):
        self.V = torch.zeros((max_steps + 2, self.n), dtype=torch.float64)
        self.w_local = torch.zeros(rows_per, dtype=torch.float64)
        self.w_full = torch.zeros(rows_per * world, dtype=torch.float64)
        self.ab = np.zeros((max_steps + 2, 2))
    def lanczos_start(self, d):
        v0 = np.sqrt(d); v0 /= np.linalg.norm(v0)
        q = np.cos(np.arange(self.n) * 0.7 + 0.3)
        for _ in range(2):
            q -= v0 * (v0 @ q)
        self.V[0] = torch.from_numpy(v0); self.V[1] = torch.from_numpy(q / np.linalg.norm(q))
    def symv(self, j):
        self.w_local[:self.nrows] = torch.from_numpy(self.Ms @ self.V[j].numpy())
    def ortho(self, j):
        w = self.w_full[:self.n].numpy().copy()
        V = self.V[:j + 1].numpy()
        alpha = 0.0
        for _ in range(2):
            h = V @ w
            w -= V.T @ h
            alpha += h[j]
        beta = np.linalg.norm(w)
        self.V[j + 1] = torch.from_numpy(w / beta)
        self.ab[j - 1] = (alpha, beta)
    def read_ab(self, steps):
        return self.ab[:steps, 0].copy(), self.ab[:steps, 1].copy()
    def tridiag_ritz(self, alpha, beta, k, tol):
        return _capi.tridiag_ritz(alpha, beta, k, tol)
    def ritz_vectors(self, steps, k, S):
        U = np.empty((self.n, k))
        U[:, 0] = self.V[0].numpy()
        if k > 1:
            U[:, 1:] = self.V[1:steps + 1].numpy().T @ S
        U /= np.linalg.norm(U, axis=0)
        return U

dist.init_process_group('gloo')
g = np.load(os.path.join(%(root)r, 'tests', 'golden', 'stage23_medium.npz'))
D = g['D']
comm = rowblock.TorchComm()
res = rowblock.embed_oversized(NumpyBackend(D), comm, D.shape[0], k=8)
res_m = rowblock.embed_oversized(NumpyBackend(D), comm, D.shape[0], k=8, moments=True)
wide = np.exp(np.random.default_rng(4).uniform(-12, 12, D.shape))          # more than 32 bins: returns to the general sweep
bk = NumpyBackend(wide); bk.dist_rows(*rowblock.row_partition(D.shape[0], comm.rank, comm.world)[:2])
_, path_wide = rowblock.sweep_sums(bk, comm, 1. / (2. * np.exp(np.arange(-100, 100.2, 0.2))), 10.0, moments=True)
val_r, U_r = O.dmap_embed(D, res['eps'])
err_val = float(np.max(np.abs(res['val'] - val_r[:8]) / val_r[:8]))
err_vec = [float(min(np.abs(res['vec'][:, i] - U_r[:, i]).max(), np.abs(res['vec'][:, i] + U_r[:, i]).max())) for i in range(5)]
L_r = O.eps_sweep_logsum(D, np.arange(-100, 100.2, 0.2), sequential=False)
out = dict(rank=comm.rank, row0=res['row0'], nrows=res['nrows'], eps=res['eps'], eps_ref=float(g['eps']), steps=res['steps'],
           err_val=err_val, err_vec=err_vec, err_log=float(np.max(np.abs(res['logSumWij'] - L_r))),
           val_hash=float(np.sum(res['val'])), vec_hash=float(np.sum(np.abs(res['vec']))),
           sweep_path=res['sweep_path'], sweep_path_m=res_m['sweep_path'], path_wide=path_wide,
           err_log_m=float(np.max(np.abs(res_m['logSumWij'] - L_r))), eps_m=res_m['eps'],
           err_val_m=float(np.max(np.abs(res_m['val'] - res['val']) / res['val'])))
print('RESULT ' + json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
'''


def free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_row_partition():
    assert rowblock.row_partition(2000, 0, 2) == (0, 1024, 1024)
    assert rowblock.row_partition(2000, 1, 2) == (1024, 976, 1024)
    parts = [rowblock.row_partition(50000, r, 8) for r in range(8)]
    assert sum(p[1] for p in parts) == 50000 and all(p[0] % 128 == 0 for p in parts)
    assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(7))
    with pytest.raises(ValueError):
        rowblock.row_partition(200, 3, 4)


def test_tridiag_ritz_matches_numpy(lib_built):
    rng = np.random.default_rng(1)
    n = 90
    a, b = rng.uniform(0.03, 0.05, n), np.abs(rng.normal(0.006, 0.001, n))
    b[-1] = 1e-13
    w, v = np.linalg.eigh(np.diag(a) + np.diag(b[:-1], 1) + np.diag(b[:-1], -1))
    conv, th, S = _capi.tridiag_ritz(a, b, 8)
    assert conv and np.max(np.abs(th - w[::-1][:7])) < 1e-15
    for i in range(7):
        assert min(np.abs(S[:, i] - v[:, -1 - i]).max(), np.abs(S[:, i] + v[:, -1 - i]).max()) < 1e-12
    b[-1] = 0.01
    assert not _capi.tridiag_ritz(a, b, 8)[0]                    # a large residual coupling -> not converged


def test_two_rank_rowblock_embedding(tmp_path, lib_built):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(free_port()), str(script)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS='2'))
    assert res.returncode == 0, res.stderr[-3000:]
    outs = sorted((json.loads(l[7:]) for l in res.stdout.splitlines() if l.startswith('RESULT ')), key=lambda o: o['rank'])
    assert len(outs) == 2
    assert (outs[0]['row0'], outs[0]['nrows'], outs[1]['row0'], outs[1]['nrows']) == (0, 256, 256, 144)
    for o in outs:
        assert abs(o['eps'] - o['eps_ref']) / o['eps_ref'] < 2e-5      # different tanh-fit start than the golden run
        assert o['err_log'] < 1e-9 and o['err_val'] < 1e-9 and max(o['err_vec']) < 1e-6
        # the moment sweep over row blocks (plan from all-reduced statistics, all-reduced tables) gives the same curve
        assert o['sweep_path'] == 'general' and o['sweep_path_m'] == 'moments' and o['path_wide'] == 'general'
        assert o['err_log_m'] < 1e-12                                  # the curve itself: 2e-15
        assert abs(o['eps_m'] - o['eps']) / o['eps'] < 1e-7 and o['err_val_m'] < 1e-7   # the fit amplifies 1e-15 to ~1e-9
    # replicated state: both ranks end with the same numbers
    assert outs[0]['val_hash'] == outs[1]['val_hash'] and outs[0]['steps'] == outs[1]['steps']
    assert abs(outs[0]['vec_hash'] - outs[1]['vec_hash']) < 1e-9


def test_cuda_backend_implements_the_backend_interface():
    """Every method embed_oversized calls on the numpy test backend exists on the GPU backend (no GPU needed to check)."""
    import re
    names = set(re.findall(r'^    def ([a-z][a-z_]*)\(', WORKER.split('class NumpyBackend')[1].split('\ndef ')[0], flags=re.M))
    assert {'dist_rows', 'symv', 'ortho', 'read_ab', 'ritz_vectors'} <= names
    for name in sorted(names) + ['symv_allgather', 'comm_check', 'comm_close', 'sync']:
        assert callable(getattr(rowblock.CudaBackend, name, None)), name
