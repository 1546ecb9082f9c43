"""Pairwise image distances per PD (interface of 1_Embedding/DM/Distances.py, non-CTF branch).

`op(pyDir, PD)` reads the PD's MRC stack, computes the RMS Euclidean distance matrix of the
per-image standardised images on the GPU (K0 + K1) and writes `Data_Distances/PD_%s_dist.npy`
(float64 [nS, nS]) exactly where the reference writes it.
"""
import os
import sys

import numpy as np

from . import _capi, mrc


def distances(stack, ctx=None, flags=_capi.DIST_DEFAULT):
    """float32 stack [nS, N, N] (or [nS, P]) -> float64 [nS, nS]; Distances.py:87-104."""
    ctx = ctx or _capi.default_context()
    return ctx.dist_rms(stack, flags=flags)


def op(pyDir, PD, CTF=False, dataDir=None, stackPath=None, outDir=None, ctx=None):
    parDir1 = os.path.abspath(os.path.join(pyDir, os.pardir))
    parDir2 = os.path.abspath(os.path.join(parDir1, os.pardir))
    if CTF:
        raise NotImplementedError('the CTF (defocus-tolerant kernel) branch of Distances.py:107-145 is outside '
                                  'the accelerated path; call op(..., CTF=False)')
    if dataDir is None:
        dataDir = os.path.join(parDir2, '0_Data_Inputs/CTF5k15k_SNRpt1_ELS_2D')
    if outDir is None:
        outDir = os.path.join(pyDir, 'Data_Distances')
    if not os.path.exists(outDir):
        os.mkdir(outDir)
    if stackPath is None:
        stackPath = os.path.join(dataDir, 'PD%s_SNR_tau5_stack.mrcs' % PD)
        if not os.path.exists(stackPath):               # the reference's noiseless naming (:59)
            alt = os.path.join(dataDir, 'PD_%s.mrcs' % PD)
            stackPath = alt if os.path.exists(alt) else stackPath
    print('CTF mode disabled.')
    PD_stack = mrc.mmap(stackPath, 'r')
    nS, N, N2 = PD_stack.data.shape
    print('Computing distances...')
    D = distances(np.asarray(PD_stack.data).reshape(nS, N * N2), ctx=ctx)
    np.save(os.path.join(outDir, 'PD_%s_dist.npy' % PD), D)
    print('Distances complete.')
    PD_stack.close()


if __name__ == '__main__':
    path1 = os.path.splitext(sys.argv[0])[0]
    path2, tail = os.path.split(path1)
    op(path2, sys.argv[1])
