#!/usr/bin/env python
"""numpy statement of the moment sweep in csrc/k2_sweep.cu (moment_stats/plan/accum/final kernels, opt-in with
ESPER_SWEEP_MOMENTS=1) and its error against the reference sweep (GaussianBandwidth.py:36-44) on the golden matrices.

    S_k = sum_{ij : c_k q < thr} w exp(-c_k q),  q = D^2

Non-zero q are grouped into geometric bins [e_b, e_b rho), rho = 1/(1 - 2/thr).  Per bin and k: nothing passes
(c_k e_b >= thr), everything passes (c_k e_{b+1} < thr: 17 scaled moments about the bin centre give the sum, Taylor
remainder <= e/17!), or the truncation boundary falls inside the bin (exact per-element evaluation; <= MS such k).
`accumulate_numpy` below follows moment_accum_kernel line by line; `plan` and `finalize` (moment_plan_kernel,
moment_final_kernel) are the host functions of manifoldem_esper_b200/moment_sweep.py that the row-block path uses.

Run: python tools/sweep_moments_proto.py [D.npy ...]   -> max |log S - reference| (tolerance of the tests: 1e-9)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from manifoldem_esper_b200 import moment_sweep as MSW  # noqa: E402  (the numpy statement lives with the host code that uses it)

MP, MS, MB, MV = MSW.MP, MSW.MS, MSW.MB, MSW.MV
PLAN_HDR, PLAN_BIN = MSW.PLAN_HDR, MSW.PLAN_BIN


def accumulate_numpy(q, w, plan_arr, coef, thr):
    """numpy statement of moment_accum_kernel: q, w flat arrays of D^2 values and weights -> table [MB, MV]."""
    coef = np.asarray(coef, dtype=np.float64)
    tot = np.zeros((MB, MV))
    ok = (q == q) & (q != 0)
    q, w = q[ok], w[ok]
    for b in range(int(plan_arr[1])):
        lo, hi, mu, sc, ka, kf = plan_arr[PLAN_HDR + PLAN_BIN * b: PLAN_HDR + PLAN_BIN * (b + 1)]
        sel = (q >= lo) & (q < hi)
        qb, wb = q[sel], w[sel]
        u = (qb - mu) * sc
        pw = wb.copy()
        tot[b, 0] = pw.sum()
        for p in range(1, MP + 1):
            pw = pw * u
            tot[b, p] = pw.sum()
        for s in range(int(kf) - int(ka)):
            t = coef[int(ka) + s] * qb
            keep = t < thr
            tot[b, MP + 1 + s] = np.sum(wb[keep] * np.exp(-t[keep]))
    return tot


def sweep_numpy(D, coef, thr, symmetric=True):
    """The whole moment sweep on the host (CPU tests): -> (logSum[k] or None, status or number of bins)."""
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
    nzm = q == 0
    mn = q[~nzm].min() if np.any(~nzm) else np.inf
    mx = q.max() if q.size else 0.0
    status, pl = MSW.plan(mn, mx, w[nzm].sum(), w.sum(), coef, thr)
    if status:
        return None, status
    S = MSW.finalize(pl, accumulate_numpy(q, w, pl, coef, thr), coef, thr)
    with np.errstate(divide='ignore'):
        return np.log(S), int(pl[1])


def moment_sweep(D, coef, thr, symmetric=True):
    """-> (logSum[k], bins) or (None, status): plan -> per-bin moments and boundary sums -> final evaluation."""
    return sweep_numpy(D, coef, thr, symmetric=symmetric)


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
