"""Diffusion-map embedding per PD (interface of 1_Embedding/DM/DiffusionMaps.py).

`op(pyDir, PD)` loads `Data_Distances/PD_%s_dist.npy`, selects epsilon (GPU sweep + host tanh
fit), builds the symmetric Markov matrix and its leading eigenpairs on the GPU (K3 + K4) and
writes `Data_Manifolds/PD_%s_val.npy` (lambda^2, descending) and `PD_%s_vec.npy` ([m, k], column 0
the steady state).  The reference writes all m columns; its consumers read <= 16
(3_Manifold_Binning/Manifold_Binning.py:139, DM_Viewer.py:86-96), so k defaults to 16; `k='all'` writes the full
[m, m] matrix like the reference (Lanczos to exhaustion + QL rotations on the device: 0.4 s at m = 2000).
"""
import os
import sys

import numpy as np

from . import GaussianBandwidth, _capi

LOG_EPS = np.arange(-100, 100.2, 0.2)     # DiffusionMaps.py:73


def embed(Dist, k=16, eps=None, a0=None, tol=0.0, ctx=None, verbose=True):
    """Distance matrix -> dict(eps, popt, logSumWij, R2, slope, val, vec, iters).

    eps=None runs the automated bandwidth selection of DiffusionMaps.py:72-93; `a0` (4 values)
    is the tanh-fit start, drawn from numpy's global RNG like the reference when None.
    """
    ctx = ctx or _capi.default_context()
    Dist = np.ascontiguousarray(Dist, dtype=np.float64)
    m = np.shape(Dist)[0]
    ctx.upload_dist(Dist)
    out = {}
    if eps is None:
        if a0 is None:
            a0 = 1 * (np.random.rand(4, 1) - .5)
        popt, logSumWij, resnorm, R2 = GaussianBandwidth.op(None, LOG_EPS, a0, ctx=ctx, m=m)
        slope = popt[0] * popt[2]
        ln_eps = -(popt[1] / popt[0])
        eps = np.exp(ln_eps)
        if verbose:
            print('Coefficient of Determination: %s' % R2)
            print('Slope: %s' % slope)
            print('ln(epsilon): %s; epsilon: %s' % (ln_eps, eps))
        out.update(popt=popt, logSumWij=logSumWij, R2=R2, slope=slope, ln_eps=ln_eps)
    if k in ('all', None):
        k = m                                                # the reference's full U (DiffusionMaps.py:166-169): every column
    val, vec, iters = ctx.dmap_embed(None, eps, k=min(k, m), tol=tol, m=m)
    out.update(eps=eps, val=val, vec=vec, iters=iters)
    return out


def run(pyDir, PD, k=16, eps=None, ctx=None):
    """`op` that also returns a summary dict (eps, R2, slope, lanczos_steps) -- what the launcher records per PD."""
    outDir = os.path.join(pyDir, 'Data_Manifolds')
    if not os.path.exists(outDir):
        os.mkdir(outDir)
    dataDir = os.path.join(pyDir, 'Data_Distances')
    Dist = np.load(os.path.join(dataDir, 'PD_%s_dist.npy' % PD))
    res = embed(Dist, k=k, eps=eps, ctx=ctx)
    np.save(os.path.join(outDir, 'PD_%s_val.npy' % PD), res['val'])
    np.save(os.path.join(outDir, 'PD_%s_vec.npy' % PD), res['vec'])
    # what the reference prints (:94-96) plus the solver's step count
    return {'eps': float(res['eps']), 'R2': float(res['R2']) if 'R2' in res else None,
            'slope': float(res['slope']) if 'slope' in res else None, 'lanczos_steps': int(res['iters'])}


def op(pyDir, PD, k=16, eps=None, ctx=None):
    """Same call as the reference's DiffusionMaps.op(pyDir, PD) (:57): writes the two files, returns None."""
    run(pyDir, PD, k=k, eps=eps, ctx=ctx)


def embed_stack_to_files(pyDir, PD, stackPath=None, k=16, a0=None, ctx=None, n=None, P=None):
    """Distances.op (non-CTF branch) + DiffusionMaps.op of one PD in ONE library call (esper_pd_embed_fused): the stack goes from
    the MRC file (or, stackPath=None with n and P given, from the stack already resident on the device) to
    Data_Distances/PD_%s_dist.npy, Data_Manifolds/PD_%s_val.npy and PD_%s_vec.npy without the distance matrix making the
    round trip through a file.  `a0` is the tanh-fit start (drawn from numpy's global RNG like the reference when None)."""
    ctx = ctx or _capi.default_context()
    if stackPath is not None:
        n, N1, N2 = ctx.upload_mrc(stackPath)
        P = N1 * N2
    if a0 is None:
        a0 = 1 * (np.random.rand(4, 1) - .5)
    res = ctx.pd_embed_fused(None, LOG_EPS, a0, k=min(k, n), n=n, P=P, want_D=True)
    for sub_, name, arr in (('Data_Distances', 'PD_%s_dist.npy', res['D']), ('Data_Manifolds', 'PD_%s_val.npy', res['val']),
                            ('Data_Manifolds', 'PD_%s_vec.npy', res['vec'])):
        d = os.path.join(pyDir, sub_)
        if not os.path.exists(d):
            os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, name % PD), arr)
    return {'eps': float(res['eps']), 'R2': float(res['R2']), 'slope': float(res['slope']), 'lanczos_steps': int(res['iters'])}


if __name__ == '__main__':
    path1 = os.path.splitext(sys.argv[0])[0]
    path2, tail = os.path.split(path1)
    op(path2, sys.argv[1])
