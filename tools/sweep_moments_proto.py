#!/usr/bin/env python
"""numpy statement of the moment sweep in csrc/k2_sweep.cu (moment_stats/plan/accum/final kernels, opt-in with
ESPER_SWEEP_MOMENTS=1) and its error against the reference sweep (GaussianBandwidth.py:36-44) on the golden matrices.

    S_k = sum_{ij : c_k q < thr} w exp(-c_k q),  q = D^2

Non-zero q are grouped into geometric bins [e_b, e_b rho), rho = 1/(1 - 2/thr).  Per bin and k: nothing passes
(c_k e_b >= thr), everything passes (c_k e_{b+1} < thr: 17 scaled moments about the bin centre give the sum, Taylor
remainder <= e/17!), or the truncation boundary falls inside the bin (exact per-element evaluation; <= MS such k).
The statement itself is manifoldem_esper_b200/moment_sweep.py (`plan`, `accumulate_numpy`, `finalize` follow
moment_plan_kernel, moment_accum_kernel and moment_final_kernel line by line; the row-block path uses `plan` and `finalize`
on the host).

Run: python tools/sweep_moments_proto.py [D.npy ...]   -> max |log S - reference| (tolerance of the tests: 1e-9)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from manifoldem_esper_b200 import moment_sweep as MSW  # noqa: E402  (the numpy statement lives with the host code that uses it)

MP, MS, MB = MSW.MP, MSW.MS, MSW.MB


def moment_sweep(D, coef, thr, symmetric=True):
    """-> (logSum[k], bins) or (None, status): plan -> per-bin moments and boundary sums -> final evaluation."""
    return MSW.sweep_numpy(D, coef, thr, symmetric=symmetric)


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
