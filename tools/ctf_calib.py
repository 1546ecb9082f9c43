"""Calibration of the pair-by-pair refinement threshold of the CTF branch: error of the Gram expansion
D^2 = T1 + T1^T - 2 T2 (tensor-core products) relative to the summed terms, on noise-free and noisy stacks.
Run on a B200: python tools/ctf_calib.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manifoldem_esper_b200 import _capi, synth          # noqa: E402
from oracle import esper_oracle as O                    # noqa: E402


def terms(st, pa):
    nS, N, _ = st.shape
    F = np.zeros((nS, N * N), dtype=complex)
    C = np.zeros((nS, N * N))
    for ii in range(nS):
        tmp = st[ii]
        tmp = (tmp - tmp.mean()) / tmp.std()
        c, G = O.ctf_filters(N, *pa[ii])
        C[ii] = c.ravel()
        F[ii] = (np.fft.fft2(tmp.astype(np.float64)) * G).ravel()
    T1 = (C ** 2) @ (np.abs(F) ** 2).T
    return T1 + T1.T


ctx = _capi.Context(0)
for name, kw in [('pristine box48', dict(pd=2, box=48, n_side=4, tau=3, snr=None)),
                 ('pristine box96', dict(pd=2, box=96, n_side=6, tau=2, snr=None)),
                 ('snr1 box64', dict(pd=3, box=64, n_side=6, tau=2, snr=1.0)),
                 ('snr0.1 box64', dict(pd=3, box=64, n_side=6, tau=2, snr=0.1))]:
    st, pa = synth.ctf_stack(**kw)
    Dr = O.distances_ctf_direct(st, pa)
    S = terms(st, pa)
    D, _ = ctx.ctf_dist(st, pa, flags=0, wiener=False)
    off = ~np.eye(len(D), dtype=bool)
    e2 = np.abs(D ** 2 - Dr ** 2)[off] / S[off]
    ratio = (Dr ** 2)[off] / S[off]
    rel = np.abs(D - Dr)[off] / Dr[off]
    print('%-16s n=%d  D^2/S in [%.3g, %.3g]  err(D^2)/S max %.2e mean %.2e   rel err D max %.2e' %
          (name, len(D), ratio.min(), ratio.max(), e2.max(), e2.mean(), rel.max()))
    for lo, hi in [(0, .01), (.01, .05), (.05, .2), (.2, 1), (1, 9)]:
        m = (ratio >= lo) & (ratio < hi)
        if m.any():
            print('    D^2/S in [%.2f, %.2f): %6d pairs, rel err D max %.2e, err(D^2)/S max %.2e' % (lo, hi, m.sum(), rel[m].max(), e2[m].max()))
