#!/usr/bin/env python
"""numpy statement of the two operand splits of the stage-1 Gram kernel and their errors against the float64 Gram.

  3xTF32 (default)          c = hi + lo, hi = tf32(c), lo = tf32(c - hi);            G ~ hi hi^T + hi lo^T + lo hi^T
  3xFP16 (esper_gram_config 1)  c 2^s = hi + lo, hi = fp16(c 2^s), lo = fp16(c 2^s - hi);  same three products, x 2^-2s

Both significands have 11 bits, so both pairs carry 22 bits; the scale 2^s (k1_gram.cu: gram_f16_scale_log2, from P alone:
|c| <= 2 sqrt(P) for centred z-scores) keeps hi below 2^15 and lo out of the FP16 subnormal range.

Run: python tools/f16_split_proto.py      (benchmark-like noisy stack and a noise-free stack, 250x250)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tf32_rna(x):
    """cvt.rna.tf32.f32: round to nearest, ties away from zero, 10 explicit mantissa bits."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((b + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def split_tf32(c):
    hi = tf32_rna(c)
    lo = tf32_rna((c - hi).astype(np.float32))
    return hi.astype(np.float64), lo.astype(np.float64)


def f16_scale_log2(P):
    s = int(math.floor(math.log2(32768.0 / (2.0 * math.sqrt(P)))))
    return -1 if s < 0 else min(s, 8)


def split_f16(c, P=None):
    s = f16_scale_log2(c.shape[1] if P is None else P)
    sc = np.float32(2.0 ** s)
    x = c * sc
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    assert np.all(np.isfinite(hi.astype(np.float32)))
    return hi.astype(np.float64) / float(sc), lo.astype(np.float64) / float(sc)


def gram3(hi, lo):
    return hi @ hi.T + hi @ lo.T + lo @ hi.T


def split_errors(c):
    """c float32 [n, P] centred z-scores -> {name: (max |dG| / (|c_i||c_j|), max relative error of D)}."""
    c64 = c.astype(np.float64)
    G = c64 @ c64.T
    P = c.shape[1]
    nrm = np.diag(G)
    D2 = (nrm[:, None] + nrm[None, :] - 2 * G) / P
    off = ~np.eye(len(G), dtype=bool)
    out = {}
    for name, (hi, lo) in (('tf32x3', split_tf32(c)), ('f16x3', split_f16(c))):
        Gs = gram3(hi, lo)
        D2s = (nrm[:, None] + nrm[None, :] - 2 * Gs) / P
        relG = float(np.max(np.abs(Gs - G) / np.sqrt(np.outer(nrm, nrm))))
        relD = float(np.max(np.abs(np.sqrt(np.maximum(D2s, 0)) - np.sqrt(np.maximum(D2, 0)))[off] / np.maximum(np.sqrt(D2[off]), 1e-2)))
        out[name] = (relG, relD)
    return out


if __name__ == '__main__':
    from manifoldem_esper_b200 import synth
    from oracle import esper_oracle as O
    for label, stack in [('noisy', synth.synthetic_pd(pd=1, box=250, n_side=8, tau=5, snr=0.1)),
                         ('pristine', synth.synthetic_pd(pd=1, box=250, n_side=16, tau=1, snr=None))]:
        X = O.standardize_stack(stack).reshape(stack.shape[0], -1)
        c = (X - X[:64].mean(axis=0)).astype(np.float32)
        for name, (rg, rd) in split_errors(c).items():
            print('%-9s %-7s max|dG|/(|ci||cj|) = %.2e   max rel dD = %.2e' % (label, name, rg, rd))
