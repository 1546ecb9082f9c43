"""Contrast-transfer function and Butterworth filter on the FFT grid (interface of
1_Embedding/DM/Generate_Filters.py).  Host-side numpy helpers: the GPU path (`esper_ctf_dist`) evaluates the
same formulas per pixel inside its FFT kernels; these functions serve callers that want the filters themselves
(e.g. `synth.ctf_stack`) and the tests.
"""
import math

import numpy as np


def create_grid(N, N2):
    """Radial frequency grid Q in units of Nyquist, zero frequency at the centre (Generate_Filters.py:25-32)."""
    if N % 2 == 1:
        a = np.arange(-(N - 1) / 2, (N - 1) / 2 + 1)
    else:
        a = np.arange(-N2, N / 2)
    X, Y = np.meshgrid(a, a)
    return (1. / (N / 2.)) * np.sqrt(X ** 2 + Y ** 2)


def create_filter(filter_type, NN, Qc, Q):
    """'Gauss' or 'Butter' low-pass of order NN and cut-off Qc (Generate_Filters.py:34-39)."""
    if filter_type == 'Gauss':
        return np.exp(-(np.log(2) / 2.) * (Q / Qc) ** 2)
    if filter_type == 'Butter':
        return np.sqrt(1. / (1 + (Q / Qc) ** (2 * NN)))
    raise ValueError('create_filter: unknown filter type %r' % (filter_type,))


def op(k, params):
    """CTF value at spatial frequency k [1/A]; params = [Cs mm, df A (underfocus > 0), keV, B, ampc]
    (Generate_Filters.py:41-58)."""
    Cs = params[0] * 1.0e7
    df, kev, B, ampc = params[1], params[2], params[3], params[4]
    wav = 12.3986 / np.sqrt(((2 * 511.0) + kev) * kev)            # electron wavelength [A]
    w1 = np.pi * Cs * wav * wav * wav
    w2 = np.pi * wav * df
    k2 = k * k
    sigm = B / math.sqrt(2 * math.log(2))
    wi = np.exp(-k2 / (2 * sigm ** 2))
    wr = (0.5 * w1 * k2 - w2) * k2
    return (np.sin(wr) - ampc * np.cos(wr)) * wi


def gen_ctf(N, px, Cs, df, volt, B, ampc):
    """[CTF, G]: CTF image and order-8 Butterworth filter (cut-off 0.5 Nyquist), both in FFT order
    (Generate_Filters.py:60-84)."""
    from numpy.fft import ifftshift
    Q = create_grid(N, N / 2.)
    CTF = ifftshift(op(Q / (2 * px), [Cs, df, volt, B, ampc]))
    G = ifftshift(create_filter('Butter', 8, 0.5, Q))
    return np.array([CTF, G], dtype='object')
