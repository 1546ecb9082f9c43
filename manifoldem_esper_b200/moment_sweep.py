"""Host side of the moment sweep (csrc/k2_sweep.cu) for the row-block path.

    S_k = sum_{ij : c_k q_ij < thr} exp(-c_k q_ij),   q = D^2      (GaussianBandwidth.py:36-44)

On one GPU the plan and the final evaluation run on the device (moment_plan_kernel, moment_final_kernel).  When the
distance matrix is split by row blocks over several ranks, the bins must be the same everywhere, so the plan is made
here from the all-reduced statistics (min / max of the non-zero q, weight of the exact zeros), every rank accumulates
the moments and boundary sums of its rows against that plan (`esper_moment_accum_rows`), the [bins, values] table is
all-reduced and `finalize` evaluates the 1001 sums.  `plan` / `finalize` follow the two kernels statement by statement.
There is no CPU path for the per-element work here: the numpy statement of moment_accum_kernel used by the CPU tests lives
in tools/sweep_moments_proto.py (test infrastructure).
"""
from __future__ import annotations

import numpy as np

MP, MS, MB = 16, 4, 32                 # highest moment, boundary coefficients per bin, bins   (k2_sweep.cu)
MV = MP + 1 + MS                       # values per bin: 17 moments + MS boundary sums
PLAN_HDR, PLAN_BIN = 8, 6              # [status, B, zero weight, valid weight, mn, mx, rho, -] + B x [lo, hi, mu, 1/h, ka, kf]
PLAN_LEN = PLAN_HDR + PLAN_BIN * MB


def _first_k(coef, edge, thr):
    """first k with coef[k] * edge < thr (bisection; coef non-increasing)."""
    a, z = 0, len(coef)
    while a < z:
        mid = (a + z) >> 1
        if coef[mid] * edge < thr:
            z = mid
        else:
            a = mid + 1
    return a


def usable(coef, thr):
    """The moment sweep needs a non-increasing, finite, non-negative coefficient grid and 0 < thr < inf."""
    coef = np.asarray(coef, dtype=np.float64)
    return bool(np.all(np.isfinite(coef)) and np.all(coef >= 0) and np.all(np.diff(coef) <= 0) and 0.0 < thr < np.inf)


def plan(mn, mx, zero_weight, valid_weight, coef, thr):
    """-> (status, plan[PLAN_LEN]); status 0 = usable, 1 = more than MB bins, 2 = more than MS coefficients fall into one
    bin, 3 = degenerate edges, 4 = infinite values.  mn = +inf means "no non-zero element" (zero bins)."""
    coef = np.asarray(coef, dtype=np.float64)
    out = np.zeros(PLAN_LEN)
    rho = 1.0 / (1.0 - 2.0 / thr) if thr > 4.0 else 2.0
    if not (rho <= 2.0):
        rho = 2.0
    status, B = 0, 0
    if mn < np.inf:
        lo = float(mn)
        while B < MB:
            hi = lo * rho
            mu = 0.5 * (lo + hi)
            h = hi - mu
            ka = _first_k(coef, lo, thr)
            kf = max(_first_k(coef, hi, thr), ka)
            if kf - ka > MS:
                status = 2
            if not (h > 0.0) or not (hi < np.inf):
                status = 3
            out[PLAN_HDR + PLAN_BIN * B: PLAN_HDR + PLAN_BIN * (B + 1)] = (lo, hi, mu, 1.0 / h if h > 0.0 else 0.0, ka, kf)
            B += 1
            if hi > mx:
                break
            lo = hi
            if B == MB:
                status = 1
    if not (mx < np.inf):
        status = 4
    out[:PLAN_HDR] = (status, B, zero_weight, valid_weight, mn, mx, rho, 0.0)
    return status, out


def finalize(plan_arr, tot, coef, thr):
    """moment_final_kernel: plan + (all-reduced) table [MB, MV] -> S[k] for every coefficient (not the logarithm)."""
    coef = np.asarray(coef, dtype=np.float64)
    tot = np.asarray(tot, dtype=np.float64).reshape(MB, MV)
    S = np.full(len(coef), plan_arr[2] if 0.0 < thr else 0.0)
    for b in range(int(plan_arr[1])):
        lo, hi, mu, sc, ka, kf = plan_arr[PLAN_HDR + PLAN_BIN * b: PLAN_HDR + PLAN_BIN * (b + 1)]
        ka, kf = int(ka), int(kf)
        M = tot[b]
        if kf > ka:
            S[ka:kf] += M[MP + 1: MP + 1 + (kf - ka)]
        if kf < len(coef):
            c = coef[kf:]
            x = -(c * (hi - mu))
            poly = np.full(len(c), M[MP])
            for p in range(MP - 1, -1, -1):
                poly = poly * (x / (p + 1)) + M[p]
            S[kf:] += np.exp(-(c * mu)) * poly
    return S


def check_layout(lib):
    """The constants above against the library's (esper_moment_layout)."""
    import ctypes as C
    v = (C.c_int * 5)()
    lib.esper_moment_layout(v)
    if tuple(v) != (MP, MS, MB, PLAN_HDR, PLAN_BIN):
        raise RuntimeError('moment sweep layout mismatch between moment_sweep.py %s and libesper_b200 %s'
                           % ((MP, MS, MB, PLAN_HDR, PLAN_BIN), tuple(v)))

