"""Gaussian-kernel bandwidth estimation (interface of 1_Embedding/DM/GaussianBandwidth.py).

The 1001-point sweep log(sum_ij exp(-D_ij^2 / 2 eps)) runs on the GPU (K2, esper_eps_sweep); the
4-parameter tanh fit stays on the host with scipy's `curve_fit`, exactly as the reference does it.
"""
import warnings

import numpy as np
from scipy.optimize import OptimizeWarning, curve_fit

from . import _capi


def fun(xx, aa0, aa1, aa2, aa3):
    """tanh model, GaussianBandwidth.py:18-24."""
    return aa3 + aa2 * np.tanh(aa0 * xx + aa1)


def sweep(D, logEps, ctx=None, m=None):
    """logSumWij[k] of GaussianBandwidth.py:36-44.  D=None uses the ctx-resident distance matrix."""
    ctx = ctx or _capi.default_context()
    logEps = np.asarray(logEps, dtype=np.float64)
    eps = np.exp(logEps)
    if D is not None:
        m = np.shape(D)[0]
        ctx.upload_dist(D)
    # find_threshold (GaussianBandwidth.py:26-32): 1% of the mean row sum at the largest epsilon
    coef_max = np.array([1. / (2. * np.max(eps))])
    log_ss = ctx.eps_sweep(None, coef_max, np.inf, m=m)[0]
    thr = max(-np.log(0.01 * np.exp(log_ss) / m), 10)
    coef = 1. / (2. * eps)
    return ctx.eps_sweep(None, coef, thr, m=m)


def fit(logEps, logSumWij, a0, native=False):
    """The curve-fit retry loop of GaussianBandwidth.py:47-57 (host side).

    native=False: scipy's `curve_fit`, the reference's own call.  native=True: `esper_tanh_fit`, the same MINPACK lmdif
    algorithm with scipy's default tolerances restated in C (csrc/k9_fit.cu) -- identical to scipy's result bit for bit
    when scipy is given a libm-tanh model, 4e-9 relative from it with numpy's SIMD tanh (the forward-difference Jacobian
    amplifies one-ulp differences), and callable without the interpreter lock, so the fits of several PDs in flight
    run in parallel."""
    resnorm, tries = np.inf, 0
    logEps = np.asarray(logEps, dtype=np.float64)
    logSumWij = np.asarray(logSumWij, dtype=np.float64)
    with warnings.catch_warnings():
        warnings.simplefilter(action='ignore', category=OptimizeWarning)
        while resnorm > 100:
            if native:
                popt, resnorm, R_squared, info, _ = _capi.tanh_fit(logEps, logSumWij, a0)
                a0 = 1 * (np.random.rand(4, 1) - .5)
                tries += 1
                if tries >= 200 and resnorm > 100:        # curve_fit raises when MINPACK gives up; never spin forever
                    raise RuntimeError('Optimal parameters not found: %d starts without a usable tanh fit (last MINPACK code %d)' % (tries, info))
                continue
            popt, pcov = curve_fit(fun, logEps, logSumWij, p0=a0)
            resnorm = sum(np.sqrt(np.fabs(np.diag(pcov))))
            a0 = 1 * (np.random.rand(4, 1) - .5)
            residuals = logSumWij - fun(logEps, popt[0], popt[1], popt[2], popt[3])
            ss_res = np.sum(residuals ** 2)
            ss_tot = np.sum((logSumWij - np.mean(logSumWij)) ** 2)
            R_squared = 1 - (ss_res / ss_tot)
    return popt, resnorm, R_squared


def op(D, logEps, a0, ctx=None, m=None, native=False):
    """Same call and return tuple as the reference: (popt, logSumWij, resnorm, R_squared)."""
    logSumWij = sweep(D, logEps, ctx=ctx, m=m)
    popt, resnorm, R_squared = fit(logEps, logSumWij, a0, native=native)
    return (popt, logSumWij, resnorm, R_squared)
