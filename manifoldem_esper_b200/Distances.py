"""Pairwise image distances per PD (interface of 1_Embedding/DM/Distances.py).

`op(pyDir, PD)` reads the PD's MRC stack and writes `Data_Distances/PD_%s_dist.npy` (float64 [nS, nS]) exactly
where the reference writes it.  `CTF` is the reference's user parameter (Distances.py:48; shipped as True):
  CTF=False  RMS Euclidean distances of the per-image standardised images (K0 + K1, `esper_dist_rms`)
  CTF=True   defocus-tolerant kernel (`esper_ctf_dist`): reads `Hsp2D_5k15k_PD_%s.mrcs` and the microscope parameters
             from `Hsp2D_5k15k_PD_%s.star`, and also writes the Wiener-filtered stack `Hsp2D_5k15k_PD_%s_filtered.mrcs`
             that 3_Manifold_Binning reads.
"""
import os
import sys

import numpy as np

from . import _capi, mrc


def distances(stack, ctx=None, flags=_capi.DIST_DEFAULT):
    """float32 stack [nS, N, N] (or [nS, P]) -> float64 [nS, nS]; Distances.py:87-104."""
    ctx = ctx or _capi.default_context()
    return ctx.dist_rms(stack, flags=flags)


def star_params(star):
    """[nS, 5] float64 = (pixel size, Cs, defocus U, voltage, amplitude contrast) from a parsed star file
    (the columns read at Distances.py:68-72)."""
    cols = ['rlnPixelSize', 'rlnSphericalAberration', 'rlnDefocusU', 'rlnVoltage', 'rlnAmplitudeContrast']
    return np.ascontiguousarray(np.stack([np.asarray(star[c].values, dtype=np.float64) for c in cols], axis=1))


def distances_ctf(stack, params, ctx=None, flags=_capi.DIST_DEFAULT, wiener=True):
    """float32 stack [nS, N, N], params [nS, 5] -> (D float64 [nS, nS], Wiener-filtered stack float32 [nS, N, N] or None);
    Distances.py:107-145."""
    ctx = ctx or _capi.default_context()
    return ctx.ctf_dist(stack, params, flags=flags, wiener=wiener)


def op(pyDir, PD, CTF=True, dataDir=None, stackPath=None, outDir=None, ctx=None):
    parDir1 = os.path.abspath(os.path.join(pyDir, os.pardir))
    parDir2 = os.path.abspath(os.path.join(parDir1, os.pardir))
    if CTF:
        return _op_ctf(PD, dataDir or os.path.join(parDir2, '0_Data_Inputs/CTF5k15k_SNRpt1_ELS_2D'),
                       outDir or os.path.join(pyDir, 'Data_Distances'), ctx)
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
    ctx = ctx or _capi.default_context()
    nS, N, N2 = ctx.upload_mrc(stackPath)        # disk -> page-locked double buffer -> HBM (the reference: mrcfile.mmap, :75)
    print('Computing distances...')
    D = ctx.dist_rms(None, n=nS, P=N * N2)
    np.save(os.path.join(outDir, 'PD_%s_dist.npy' % PD), D)
    print('Distances complete.')


def _op_ctf(PD, dataDir, outDir, ctx):
    from . import Read_Alignment
    if not os.path.exists(outDir):
        os.mkdir(outDir)
    print('CTF mode enabled.')
    stackPath = os.path.join(dataDir, 'Hsp2D_5k15k_PD_%s.mrcs' % PD)
    alignPath = os.path.join(dataDir, 'Hsp2D_5k15k_PD_%s.star' % PD)
    wienerOut = os.path.join(dataDir, 'Hsp2D_5k15k_PD_%s_filtered.mrcs' % PD)
    params = star_params(Read_Alignment.parse_star(alignPath))
    PD_stack = mrc.mmap(stackPath, 'r')
    nS, N, N2 = PD_stack.data.shape
    if params.shape[0] != nS:
        raise ValueError('Distances.op: %s lists %d images, the stack holds %d' % (alignPath, params.shape[0], nS))
    print('Generating filters...')
    print('Performing CTF correction...')
    print('Computing distances...')
    D, W = distances_ctf(np.asarray(PD_stack.data), params, ctx=ctx)
    wiener_stack = mrc.new_mmap(wienerOut, shape=(nS, N, N), mrc_mode=2, overwrite=True)
    wiener_stack.data[:] = W
    wiener_stack.close()
    np.save(os.path.join(outDir, 'PD_%s_dist.npy' % PD), D)
    print('Distances complete.')
    PD_stack.close()


if __name__ == '__main__':
    path1 = os.path.splitext(sys.argv[0])[0]
    path2, tail = os.path.split(path1)
    op(path2, sys.argv[1])
