#!/usr/bin/env python
"""numpy statement of the moment sweep in csrc/k2_sweep.cu (moment_stats/plan/accum/final kernels, opt-in with
ESPER_SWEEP_MOMENTS=1) and its error against the reference sweep (GaussianBandwidth.py:36-44) on the golden matrices.

    S_k = sum_{ij : c_k q < thr} w exp(-c_k q),  q = D^2

Non-zero q are grouped into geometric bins [e_b, e_b rho), rho = 1/(1 - 2/thr).  Per bin and k: nothing passes
(c_k e_b >= thr), everything passes (c_k e_{b+1} < thr: 17 scaled moments about the bin centre give the sum, Taylor
remainder <= e/17!), or the truncation boundary falls inside the bin (exact per-element evaluation; <= MS such k).
The functions below follow the kernels line by line (same plan, same bisection predicates, same Horner scheme).

Run: python tools/sweep_moments_proto.py [D.npy ...]   -> max |log S - reference| (tolerance of the tests: 1e-9)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MP, MS, MB = 16, 4, 32


def first_k(coef, edge, thr):
    """first k with coef[k]*edge < thr (bisection on non-increasing coefficients, as moment_plan_kernel)."""
    a, z = 0, len(coef)
    while a < z:
        mid = (a + z) >> 1
        if coef[mid] * edge < thr:
            z = mid
        else:
            a = mid + 1
    return a


def moment_plan(q_nonzero, coef, thr):
    mn, mx = q_nonzero.min(), q_nonzero.max()
    rho = 1.0 / (1.0 - 2.0 / thr) if thr > 4.0 else 2.0
    rho = min(rho, 2.0)
    bins, lo, status = [], mn, 0
    while len(bins) < MB:
        hi = lo * rho
        mu = 0.5 * (lo + hi)
        h = hi - mu
        ka = first_k(coef, lo, thr)
        kf = max(first_k(coef, hi, thr), ka)
        if kf - ka > MS:
            status = 2
        if not (h > 0.0) or not np.isfinite(hi):
            status = 3
        bins.append((lo, hi, mu, 1.0 / h if h > 0 else 0.0, ka, kf))
        if hi > mx:
            break
        lo = hi
        if len(bins) == MB:
            status = 1
    return status, bins


def moment_sweep(D, coef, thr, symmetric=True):
    D = np.asarray(D, dtype=np.float64)
    m = D.shape[0]
    if symmetric:
        iu = np.triu_indices(m)
        q = (D * D)[iu]
        w = np.where(iu[0] == iu[1], 1.0, 2.0)
    else:
        q = (D * D).ravel()
        w = np.ones_like(q)
    ok = q == q
    q, w = q[ok], w[ok]
    nz = w[q == 0].sum()
    qn, wn = q[q != 0], w[q != 0]
    out = np.empty(len(coef))
    if qn.size == 0:
        out[:] = math.log(nz) if nz > 0 else -np.inf
        return out, 0
    status, bins = moment_plan(qn, coef, thr)
    if status:
        return None, status
    S = np.full(len(coef), nz if 0 < thr else 0.0)
    for (lo, hi, mu, sc, ka, kf) in bins:
        sel = (qn >= lo) & (qn < hi)
        qb, wb = qn[sel], wn[sel]
        u = (qb - mu) * sc
        M = np.empty(MP + 1)
        pw = wb.copy()
        M[0] = pw.sum()
        for p in range(1, MP + 1):
            pw = pw * u
            M[p] = pw.sum()
        h = hi - mu
        for k in range(ka, len(coef)):
            c = coef[k]
            if k < kf:
                t = c * qb
                keep = t < thr
                S[k] += np.sum(wb[keep] * np.exp(-t[keep]))
            else:
                x = -(c * h)
                poly = M[MP]
                for p in range(MP - 1, -1, -1):
                    poly = poly * (x / (p + 1)) + M[p]
                S[k] += math.exp(-(c * mu)) * poly
    with np.errstate(divide='ignore'):
        return np.log(S), len(bins)


if __name__ == '__main__':
    from oracle import esper_oracle as O
    logEps = np.arange(-100, 100.2, 0.2)
    coef = 1. / (2. * np.exp(logEps))
    gold = os.path.join(ROOT, 'tests', 'golden')
    cases = [(n, np.load(os.path.join(gold, n + '.npz'))) for n in
             ['stage123_small', 'stage123_pristine', 'stage23_medium', 'stage1_dupes']]
    mats = [(n, g['D'], g['logSumWij'] if 'logSumWij' in g else None) for n, g in cases]
    for path in sys.argv[1:]:
        mats.append((os.path.basename(path), np.load(path), None))
    for name, D, ref in mats:
        if ref is None:
            ref = O.eps_sweep_logsum(D, logEps, sequential=False)
        L, nb = moment_sweep(D, coef, 10.0)
        if L is None:
            print('%-20s m=%5d  falls back to the general kernel (status %d)' % (name, D.shape[0], nb))
            continue
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(L), fin)
        print('%-20s m=%5d  bins %2d  max |dL| = %.2e' % (name, D.shape[0], nb, np.max(np.abs(L[fin] - ref[fin]))))
