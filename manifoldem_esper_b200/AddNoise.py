"""Additive Gaussian noise at a target SNR + background standardisation of one image (interface of
0_Data_Inputs/Pristine_AddTauSNR/AddNoise.py), computed on the GPU by `esper_tau_snr_stack` (k8_dataprep.cu).

`find_SNR(image, SNR) -> (sig_mean, noise_std)` and `op(img_orig, SNR) -> standardised noisy image` keep the
reference's names and argument meaning.  The reference draws its noise from numpy's unseeded global generator; here
the stream is a counter-based Philox keyed by `seed` (default: a fresh seed per call, like the reference's
irreproducible draw).  Whole stacks go through `Generate_TauSNR.op`, one launch for all images.
"""
import os

import numpy as np

from . import _capi


def _fresh_seed():
    return int.from_bytes(os.urandom(8), 'little')


def find_SNR(image, SNR, ctx=None):
    """AddNoise.py:17-39 -> (signal mean, noise std)."""
    ctx = ctx or _capi.default_context()
    image = np.asarray(image, dtype=np.float32)
    _, params = ctx.tau_snr_stack(image[None], 1, SNR, seed=0, download=False)
    return params[0, 0], params[0, 1]


def op(img_orig, SNR, seed=None, ctx=None):
    """AddNoise.py:77-86: noisy, background-standardised copy of one [box, box] image (float32)."""
    ctx = ctx or _capi.default_context()
    image = np.asarray(img_orig, dtype=np.float32)
    stack, _ = ctx.tau_snr_stack(image[None], 1, SNR, seed=_fresh_seed() if seed is None else seed)
    return stack[0]
