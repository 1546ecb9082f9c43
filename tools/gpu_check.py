#!/usr/bin/env python
"""Development check run on the GPU box: every stage against the oracle + kernel timings.
Usage: python tools/gpu_check.py [sections...]   (smoke shapes full sweep eig rot)"""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import _capi, synth, Rotation_Histograms as RH  # noqa: E402
from oracle import esper_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def rel_err(D, Dr, floor=1e-3):
    return float(np.max(np.abs(D - Dr) / np.maximum(Dr, floor)))


def sec_smoke(ctx):
    import __graft_entry__ as g
    g.smoke()


def sec_shapes(ctx):
    rng = np.random.default_rng(0)
    for n in (1, 2, 3, 127, 129, 300):
        for P in (16, 1000, 4096):
            X = rng.standard_normal((n, P)).astype(np.float32)
            D = ctx.dist_rms(X)
            Dr = O.distances_gram_f64(O.standardize_stack(X.reshape(n, 1, P)))
            print('shape n=%d P=%d err=%.3e sym=%s diag0=%s' % (n, P, rel_err(D, Dr), np.array_equal(D, D.T), np.all(np.diag(D) == 0)), flush=True)
    for name in ('stage123_small', 'stage123_pristine', 'stage1_dupes'):
        g = np.load(os.path.join(GOLD, name + '.npz'))
        for flags in (_capi.DIST_DEFAULT, _capi.DIST_STANDARDIZE, _capi.DIST_STANDARDIZE | _capi.DIST_CENTER):
            D = ctx.dist_rms(g['stack'], flags=flags)
            print('golden %s flags=%d err=%.3e maxabs=%.3e' % (name, flags, rel_err(D, g['D'], 1e-2), np.abs(D - g['D']).max()), flush=True)


def kernel_report(ctx, names=('prep', 'gram', 'dist_finalize', 'refine', 'sweep', 'markov', 'lanczos', 'rot')):
    out = {}
    for k in names:
        ms, n = ctx.kernel_ms(k)
        if n:
            out[k] = (round(ms, 4), n)
    return out


def sec_full(ctx):
    t = time.time()
    stack = synth.synthetic_pd(pd=1, box=250, n_side=20, tau=5, snr=0.1)
    n, P = stack.shape[0], 250 * 250
    print('generated stack %s in %.1fs' % (stack.shape, time.time() - t), flush=True)
    t = time.time()
    Dr = O.distances_gram_f64(O.standardize_stack(stack))
    print('oracle f64 gram in %.1fs  D range [%.4f, %.4f]' % (time.time() - t, Dr[Dr > 0].min(), Dr.max()), flush=True)
    np.save(os.path.join(ROOT, 'gpurun_out', 'D_ref_full.npy'), Dr) if os.environ.get('SAVE_D') else None
    ctx.upload_images(stack.reshape(n, P)); ctx.sync()
    for sk in (0, 1, 8, 37, 74):
        ctx.dist_config(split_k=sk)
        D = ctx.dist_rms(None, n=n, P=P)
        ctx.set_profiling(True); ctx.reset_profiling()
        for _ in range(3):
            ctx.dist_rms(None, n=n, P=P, download=False)
        ctx.sync()
        rep = kernel_report(ctx)
        ctx.set_profiling(False)
        gram_ms = rep.get('gram', (0, 1))[0] / max(rep.get('gram', (0, 1))[1], 1)
        d2 = D * D; d2r = Dr * Dr
        off = ~np.eye(n, dtype=bool)
        bias = float(np.mean((d2 - d2r)[off] / d2r[off]))
        print('full split_k=%d err=%.3e d2 mean-rel-bias=%.3e rms=%.3e gram_ms=%.3f TFLOPs(alg)=%.1f rep=%s' % (
            sk, rel_err(D, Dr), bias, float(np.sqrt(np.mean(((d2 - d2r)[off] / d2r[off]) ** 2))), gram_ms,
            2.0 * n * n * P / (gram_ms * 1e-3) / 1e12 if gram_ms else 0, json.dumps(rep)), flush=True)
    ctx.dist_config(split_k=0)
    return stack, Dr


def sec_sweep(ctx, Dr=None):
    if Dr is None:
        g = np.load(os.path.join(GOLD, 'stage23_medium.npz')); Dr = g['D']
    m = Dr.shape[0]
    logEps = np.arange(-100, 100.2, 0.2)
    t = time.time(); Lr = O.eps_sweep_logsum(Dr, logEps, sequential=False); tor = time.time() - t
    ctx.upload_dist(Dr)
    coef = 1. / (2. * np.exp(logEps))
    ctx.set_profiling(True); ctx.reset_profiling()
    t = time.time()
    for _ in range(3):
        L = ctx.eps_sweep(None, coef, 10.0, m=m)
    tg = (time.time() - t) / 3
    print('sweep m=%d maxabs=%.3e oracle_s=%.2f gpu_wall_ms=%.3f rep=%s' % (m, np.abs(L - Lr).max(), tor, tg * 1e3, json.dumps(kernel_report(ctx))), flush=True)
    ctx.set_profiling(False)
    return Lr


def sec_eig(ctx, Dr=None, eps=None):
    if Dr is None:
        g = np.load(os.path.join(GOLD, 'stage23_medium.npz')); Dr = g['D']; eps = float(g['eps'])
    m = Dr.shape[0]
    t = time.time(); val_r, U_r = O.dmap_embed(Dr, eps); tor = time.time() - t
    A, d, Ms_r = O.markov_symmetric(Dr, eps)
    ctx.upload_dist(Dr)
    deg, Ms = ctx.markov(None, eps, m=m)
    print('markov m=%d deg rel=%.3e Ms maxrel=%.3e bit-equal=%s' % (m, np.abs(deg - d).max() / d.max(), np.abs(Ms - Ms_r).max() / np.abs(Ms_r).max(), np.array_equal(Ms, Ms_r)), flush=True)
    for k in (10, 16):
        ctx.set_profiling(True); ctx.reset_profiling()
        t = time.time()
        val, vec, iters = ctx.dmap_embed(None, eps, k=k, m=m)
        tg = time.time() - t
        rep = kernel_report(ctx); ctx.set_profiling(False)
        ev = np.abs(val - val_r[:k]) / val_r[:k]
        vs = [min(np.abs(vec[:, i] - U_r[:, i]).max(), np.abs(vec[:, i] + U_r[:, i]).max()) for i in range(k)]
        lam = np.sqrt(val)
        res = np.linalg.norm(Ms_r @ vec - vec * lam[None, :], axis=0)
        print('eig m=%d k=%d iters=%d wall_ms=%.2f oracle_s=%.2f val_rel_max=%.3e vec_err=%s resid_max=%.3e orth=%.3e rep=%s' % (
            m, k, iters, tg * 1e3, tor, ev.max(), np.array2string(np.array(vs), precision=1), res.max(),
            np.abs(vec.T @ vec - np.eye(k)).max(), json.dumps(rep)), flush=True)
        print("   lam", np.array2string(lam[:k], precision=6), flush=True)


def sec_rot(ctx):
    fx = np.load(os.path.join(GOLD, 'rot_inputs.npz'))
    cnt_ref = np.load(os.path.join(GOLD, 'rot_counts_dim8.npz'))
    for dim in (6, 8):
        ref = np.load(os.path.join(GOLD, 'rot_ref_dim%d.npz' % dim))
        preps, keys = [], []
        for key in fx:
            preps.append(RH.prepare(fx[key], dim=dim, PCA=True, CTF=True)); keys.append(key)
        groups = {}
        for p, k in zip(preps, keys):
            groups.setdefault(len(p['Rij_final']), []).append((p, k))
        for nr, items in groups.items():
            ctx.set_profiling(True); ctx.reset_profiling()
            t = time.time()
            th, counts = RH.sweep_many([p for p, _ in items], dim, ctx=ctx, want_counts=True)
            tg = time.time() - t
            rep = kernel_report(ctx); ctx.set_profiling(False)
            ok = [np.array_equal(th[i], ref['RotMatrices_%s' % k[2:]]) for i, (_, k) in enumerate(items)]
            print('rot dim=%d n_pd=%d all_equal=%s wall_ms=%.2f rep=%s' % (dim, len(items), all(ok), tg * 1e3, json.dumps(rep)), flush=True)
            if dim == 8:
                for i, (p, k) in enumerate(items):
                    if k in cnt_ref:
                        ns = len(p['comboList'])
                        c = counts[i, :ns].reshape(-1)
                        print('   counts %s equal=%s' % (k, np.array_equal(c, cnt_ref[k])), flush=True)


def sec_rot126(ctx):
    """BASELINE config 5 shape: 126 manifolds [2000, 15], dim 8, whole theta search in one call."""
    preps = []
    t = time.time()
    for pd in range(126):
        U0 = synth.random_manifold(m=2000, ncol=15, seed=pd)
        try:
            preps.append(RH.prepare(U0, dim=8, PCA=True, CTF=True))
        except Exception as e:
            print('prepare failed for', pd, e)
    groups = {}
    for p in preps:
        groups.setdefault(len(p['Rij_final']), []).append(p)
    print('prepared %d PDs (host conic fits) in %.2fs; groups %s' % (len(preps), time.time() - t, {k: len(v) for k, v in groups.items()}), flush=True)
    for nr, items in groups.items():
        RH.sweep_many(items, 8, ctx=ctx)
        ctx.set_profiling(True); ctx.reset_profiling()
        t = time.time()
        th = RH.sweep_many(items, 8, ctx=ctx)
        tg = time.time() - t
        rep = kernel_report(ctx); ctx.set_profiling(False)
        n_eval = sum(len(p['comboList']) for p in items) * nr * 160
        print('rot126 n_pd=%d n_rij=%d evals=%d wall_ms=%.2f kernel_ms=%.3f evals/s(kernel)=%.3e evals/s(wall)=%.3e' % (
            len(items), nr, n_eval, tg * 1e3, rep['rot'][0], n_eval / (rep['rot'][0] * 1e-3), n_eval / tg), flush=True)


def main():
    secs = sys.argv[1:] or ['smoke', 'shapes', 'sweep', 'eig', 'rot', 'full']
    ctx = _capi.Context(0)
    Dr = None
    for s in secs:
        print('==== section %s' % s, flush=True)
        t = time.time()
        try:
            if s == 'smoke': sec_smoke(ctx)
            elif s == 'shapes': sec_shapes(ctx)
            elif s == 'full':
                _, Dr = sec_full(ctx)
                sec_sweep(ctx, Dr)
                sec_eig(ctx, Dr, 0.2834)
            elif s == 'sweep': sec_sweep(ctx)
            elif s == 'eig': sec_eig(ctx)
            elif s == 'rot': sec_rot(ctx)
            elif s == 'rot126': sec_rot126(ctx)
        except Exception:
            traceback.print_exc()
        print('==== section %s done in %.1fs' % (s, time.time() - t), flush=True)


if __name__ == '__main__':
    main()
