#!/usr/bin/env python
"""First hardware run of the kernels written after round 1's GPU budget was spent (never executed so far):

  * 3xFP16 operand planes for the Gram kernel     esper_gram_config(ctx, 1)    gram_wide_kernel<true>, stats_split_kernel<true>
  * cluster standardise-and-split pass            esper_prep_config(ctx, 1)    stats_split_cluster_kernel<HALF>
  * moment sweep                                  esper_sweep_config(ctx, 1)   moment_{stats,plan,accum,final}_kernel
  * eigensolver schedule without speculation      esper_dmap_config(ctx, 2)    host control flow of lanczos_attempt only

    gpurun --timeout 900 -- 'for w in gram prep sweep eig; do timeout 200 python tools/experimental_check.py $w; done'

(one process per stage: a device fault in one kernel does not take the later stages with it)

Every stage prints before the next starts (flush), small shapes first, so a hang or fault is attributable.  Then:

    ESPER_TEST_EXPERIMENTAL=1 python -m pytest tests -m gpu -q -k "f16_operands or moment_sweep or cluster_prep"
    python bench.py --gram-f16 --sweep-moments --prep-cluster --eig-mode 2

When both are green and faster, make them the defaults (ctx->gram_mode / sweep_mode = 1 in esper_create), drop the
`experimental` marker in tests/test_gpu_parity.py and re-capture profiles/ (launch list + ncu --set full of the Gram kernel).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import _capi, synth  # noqa: E402
from oracle import esper_oracle as O  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
ctx = _capi.Context(0)
rng = np.random.default_rng(0)
ok = True
LOG_EPS = np.arange(-100, 100.2, 0.2)
coef = 1. / (2. * np.exp(LOG_EPS))


def rel(D, Dr):
    return float(np.max(np.abs(D - Dr) / np.maximum(Dr, 1e-2)))


if what in ('all', 'gram'):
    print('== 3xFP16 operand planes ==', flush=True)
    for n, P in [(129, 64), (256, 1000), (300, 4096), (640, 128 * 128), (2000, 62500)]:
        if n == 2000:
            X = synth.synthetic_pd(pd=3, box=250, n_side=20, tau=5, snr=0.1).reshape(2000, 62500)
        else:
            X = rng.standard_normal((n, P), dtype=np.float32)
        Dr = O.distances_gram_f64(O.standardize_stack(X.reshape(n, 1, P)))
        ctx.gram_config(0)
        D0 = ctx.dist_rms(X)
        ctx.gram_config(1)
        D1 = ctx.dist_rms(X)
        used = ctx.gram_info()
        e0, e1 = rel(D0, Dr), rel(D1, Dr)
        good = e1 < 1e-5 and used == 1 and np.array_equal(D1, D1.T) and not np.diag(D1).any()
        ok &= good
        print('n=%4d P=%5d  rel err tf32x3 %.2e  f16x3 %.2e  (f16 planes used: %d)  %s' % (n, P, e0, e1, used, 'ok' if good else 'FAILED'), flush=True)
    # accumulation bias and eigenvalues at the benchmark shape (last X)
    d2, d2r = D1 * D1, Dr * Dr
    off = ~np.eye(2000, dtype=bool)
    bias = float(np.mean((d2 - d2r)[off] / d2r[off]))
    val_r, _ = O.dmap_embed(Dr, 0.2834)
    val, _, steps = ctx.dmap_embed(None, 0.2834, k=10, m=2000)
    ev = float(np.max(np.abs(val - val_r[:10]) / val_r[:10]))
    ok &= abs(bias) < 1e-7 and ev < 1e-6
    print('mean relative error of D^2 %.2e (bound 1e-7), eigenvalue error %.2e (bound 1e-6), %d Lanczos steps' % (bias, ev, steps), flush=True)
    ctx.set_profiling(True)
    for mode in (0, 1):
        ctx.gram_config(mode)
        ctx.dist_rms(X)
        ctx.reset_profiling()
        for _ in range(5):
            ctx.dist_rms(None, n=2000, P=62500, download=False)
        ctx.sync()
        print('%s: prep %.3f ms  gram %.3f ms  finalize %.3f ms' % ('3xFP16' if mode else '3xTF32', ctx.kernel_ms('prep')[0] / 5,
                                                                   ctx.kernel_ms('gram')[0] / 5, ctx.kernel_ms('dist_finalize')[0] / 5), flush=True)
    ctx.set_profiling(False)
    ctx.gram_config(0)
if what in ('all', 'prep'):
    print('== cluster standardise-and-split pass ==', flush=True)
    X = synth.synthetic_pd(pd=3, box=250, n_side=20, tau=5, snr=0.1).reshape(2000, 62500)
    for n, P in [(70, 999), (300, 16384), (2000, 62500)]:
        Xc = X if n == 2000 else rng.standard_normal((n, P), dtype=np.float32)
        ctx.prep_config(0)
        D0 = ctx.dist_rms(Xc)
        ctx.prep_config(1)
        D1 = ctx.dist_rms(Xc)
        good = rel(D1, D0) < 1e-6
        ok &= good
        print('n=%4d P=%5d  cluster size %d  |cluster - default| %.2e  %s' % (n, P, ctx.prep_info(), rel(D1, D0), 'ok' if good else 'FAILED'), flush=True)
    ctx.set_profiling(True)
    for gm in (0, 1):
        for pm in (0, 1):
            ctx.gram_config(gm); ctx.prep_config(pm)
            ctx.dist_rms(X)
            ctx.reset_profiling()
            for _ in range(5):
                ctx.dist_rms(None, n=2000, P=62500, download=False)
            ctx.sync()
            print('%s planes, %s prep: %.3f ms' % ('FP16' if gm else 'TF32', 'cluster' if pm else 'one-CTA', ctx.kernel_ms('prep')[0] / 5), flush=True)
    ctx.set_profiling(False)
    ctx.gram_config(0); ctx.prep_config(0)

if what in ('all', 'sweep'):
    print('== moment sweep ==', flush=True)
    gold = os.path.join(ROOT, 'tests', 'golden')
    cases = []
    for name in ['stage123_small', 'stage123_pristine', 'stage23_medium', 'stage1_dupes']:
        g = np.load(os.path.join(gold, name + '.npz'))
        cases.append((name, g['D'], g['logSumWij'] if 'logSumWij' in g else None))
    Xs = rng.standard_normal((2000, 8192))
    G = Xs @ Xs.T
    nrm = np.diag(G)
    Db = np.sqrt(np.maximum(nrm[:, None] + nrm[None, :] - 2 * G, 0) / 8192) * (1.38 / np.sqrt(2))
    Db = 0.5 * (Db + Db.T)
    np.fill_diagonal(Db, 0.0)
    cases.append(('benchmark shape', Db, None))
    for name, D, ref in cases:
        ctx.sweep_config(0)
        L0 = ctx.eps_sweep(D, coef, 10.0)
        ctx.sweep_config(1)
        L1 = ctx.eps_sweep(None, coef, 10.0, m=D.shape[0])
        path, bins = ctx.sweep_info()
        fin = np.isfinite(L0)
        d01 = float(np.max(np.abs(L1[fin] - L0[fin]))) if np.array_equal(np.isfinite(L1), fin) else float('inf')
        dref = float(np.max(np.abs(L1[fin] - ref[fin]))) if ref is not None else float('nan')
        good = path == 1 and d01 < 1e-12
        ok &= good
        print('%-18s m=%4d  path %d bins %2d  |moments - general| %.2e  |moments - reference| %.2e  %s'
              % (name, D.shape[0], path, bins, d01, dref, 'ok' if good else 'FAILED'), flush=True)
    ctx.set_profiling(True)
    for mode in (0, 1):
        ctx.sweep_config(mode)
        ctx.eps_sweep(None, coef, 10.0, m=2000)
        ctx.reset_profiling()
        for _ in range(5):
            ctx.eps_sweep(None, coef, 10.0, m=2000)
        print('%s sweep: %.3f ms of kernels' % ('moment' if mode else 'general', ctx.kernel_ms('sweep')[0] / 5), flush=True)
    ctx.set_profiling(False)
    ctx.sweep_config(0)

if what in ('all', 'eig'):
    print('== eigensolver schedule without the speculative batch (esper_dmap_config 2) ==', flush=True)
    Xe = synth.synthetic_pd(pd=3, box=250, n_side=20, tau=5, snr=0.1).reshape(2000, 62500)
    ctx.dist_rms(Xe, download=False)
    ctx.set_profiling(True)
    res = {}
    for mode in (1, 2):
        ctx.dmap_config(mode)
        ctx.dmap_embed(None, 0.2834, k=10, m=2000)
        ctx.reset_profiling()
        t0 = time.perf_counter()
        for _ in range(5):
            val, vec, steps = ctx.dmap_embed(None, 0.2834, k=10, m=2000)
        wall = (time.perf_counter() - t0) / 5 * 1e3
        res[mode] = val
        print('mode %d: %d steps reported, %.3f ms of kernels, %.3f ms wall per solve' % (mode, steps, ctx.kernel_ms('lanczos')[0] / 5, wall), flush=True)
    good = float(np.max(np.abs(res[1] - res[2]) / res[1])) < 1e-10
    ok &= good
    print('eigenvalues of both schedules agree: %s' % ('ok' if good else 'FAILED'), flush=True)
    ctx.set_profiling(False)
    ctx.dmap_config(0)

print('EXPERIMENTAL KERNELS OK' if ok else 'EXPERIMENTAL KERNELS FAILED')
