"""Seeded synthetic PD image stacks of the shape the reference pipeline consumes.

The reference's own inputs (0_Data_Inputs/2_PDs_2D/PD_00x.mrcs: 400 states = 20 x 20, two
degrees of freedom, 0_Data_Inputs/README.md:6) are not shipped, so benchmarks and parity
tests use this generator: a toy molecule with a central blob and two
satellites whose polar angles are the two conformational coordinates, then tau noisy copies
per state at a target SNR following 0_Data_Inputs/Pristine_AddTauSNR/AddNoise.py:17-86
(signal = pixels outside +-0.25 std; noise_std = sqrt(var(signal)/SNR); noise mean = signal
mean; standardise with the mean/std of the pixels outside the inscribed circle).
"""
from __future__ import annotations

import numpy as np


def pristine_states(box: int = 250, n_side: int = 20, pd: int = 1, n_pd: int = 126) -> np.ndarray:
    """float32 [n_side**2, box, box]: state (i, j) has satellite angles (a_i, b_j)."""
    c = (box - 1) / 2.0
    y, x = np.arange(box, dtype=np.float64), np.arange(box, dtype=np.float64)
    rad = 0.22 * box
    phase = pd * np.pi / n_pd  # in-plane rotation so PDs differ

    def blob(cx, cy, sig):
        ex = np.exp(-((x - cx) ** 2) / (2 * sig * sig))
        ey = np.exp(-((y - cy) ** 2) / (2 * sig * sig))
        return np.outer(ey, ex)

    core = blob(c, c, 0.10 * box)
    ang = (np.pi / 3.0) * np.arange(n_side) / max(n_side - 1, 1)
    satA = [0.8 * blob(c + rad * np.cos(phase + a), c + rad * np.sin(phase + a), 0.05 * box) for a in ang]
    satB = [0.5 * blob(c + rad * np.cos(phase + np.pi + b), c + rad * np.sin(phase + np.pi + b), 0.05 * box)
            for b in ang]
    out = np.empty((n_side * n_side, box, box), dtype=np.float32)
    for i in range(n_side):
        for j in range(n_side):
            out[i * n_side + j] = core + satA[i] + satB[j]
    return out


def _background_mask(box: int) -> np.ndarray:
    Y, X = np.ogrid[:box, :box]
    return np.sqrt((X - int(box / 2)) ** 2 + (Y - int(box / 2)) ** 2) > int(box / 2)


def noisy_stack(pristine: np.ndarray, tau: int = 5, snr: float = 0.1, seed: int = 0) -> np.ndarray:
    """float32 [n_states*tau, box, box]; image index = state*tau + copy (Generate_TauSNR.py:56-64)."""
    ns, box, _ = pristine.shape
    rng = np.random.default_rng(seed)
    bg = _background_mask(box)
    out = np.empty((ns * tau, box, box), dtype=np.float32)
    for s in range(ns):
        img = pristine[s].astype(np.float64)
        std = img.std()
        sig = img[(img <= -0.25 * std) | (img >= 0.25 * std)]
        sig_mean, noise_std = sig.mean(), np.sqrt(sig.var() / snr)
        for t in range(tau):
            g = rng.standard_normal((box, box), dtype=np.float32)
            noisy = img + (sig_mean + noise_std * g)
            b = noisy[bg]
            out[s * tau + t] = (noisy - b.mean()) / b.std()
    return out


def synthetic_pd(pd: int = 1, box: int = 250, n_side: int = 20, tau: int = 5, snr: float | None = 0.1,
                 seed: int | None = None) -> np.ndarray:
    """One PD stack, float32 [n_side**2 * tau, box, box].  `snr=None` -> tau exact duplicates."""
    pr = pristine_states(box, n_side, pd)
    if snr is None:
        return np.repeat(pr, tau, axis=0)
    return noisy_stack(pr, tau, snr, 1000 * pd if seed is None else seed)


def ctf_stack(pd: int = 1, box: int = 250, n_side: int = 20, tau: int = 5, snr: float = 0.1, seed: int | None = None,
              df_range=(5000, 15000), volt: float = 300., Cs: float = 2.7, ampc: float = 0.1, pix: float = 1.0):
    """CTF-modulated noisy PD stack and its microscope parameters, following
    0_Data_Inputs/Pristine_AddCtfSNR/GenStackAlign_CtfSnr.py:186-207: per image a random integer defocus in
    [5000, 15000) A, image <- -ifft2(fft2(image) * CTF).real, Gaussian noise from the "signal" pixels
    (|pixel| >= 1 std), background normalisation.  Returns (stack float32 [N, box, box],
    params float64 [N, 5] = pixel size, Cs, defocus, voltage, amplitude contrast)."""
    from . import Generate_Filters
    pr = np.repeat(pristine_states(box, n_side, pd), tau, axis=0)
    n = pr.shape[0]
    rng = np.random.default_rng(1000 * pd + 7 if seed is None else seed)
    bg = _background_mask(box)
    out = np.empty((n, box, box), dtype=np.float32)
    params = np.empty((n, 5), dtype=np.float64)
    for i in range(n):
        df = float(rng.integers(df_range[0], df_range[1]))
        ctf = Generate_Filters.gen_ctf(box, pix, Cs, df, volt, np.inf, ampc)[0].astype(float)
        img = -np.fft.ifft2(np.fft.fft2(pr[i].astype(np.float64)) * ctf).real
        if snr is not None:
            std = img.std()
            sig = img[(img <= -std) | (img >= std)]
            img = img + rng.normal(sig.mean(), np.sqrt(sig.var() / snr), (box, box))
        b = img[bg]
        out[i] = (img - b.mean()) / b.std()
        params[i] = (pix, Cs, df, volt, ampc)
    return out, params


def random_manifold(m: int = 2000, ncol: int = 15, seed: int = 0) -> np.ndarray:
    """float64 [m, ncol] stand-in for a PD_xxx_vec.npy manifold (config 5 throughput shape):
    two parabolic pairs (psi_1, psi_2), (psi_3, psi_4) like a 2-CM diffusion map, mixed by a
    small random rotation, plus noise columns."""
    rng = np.random.default_rng(seed)
    a, b = rng.uniform(-1, 1, m), rng.uniform(-1, 1, m)
    cols = [a, b, 2 * a * a - 1, 2 * b * b - 1, a * b, 4 * a ** 3 - 3 * a, 4 * b ** 3 - 3 * b]
    U = np.zeros((m, ncol))
    for c in range(ncol):
        U[:, c] = cols[c] if c < len(cols) else rng.standard_normal(m) * 0.3
    U[:, :7] += 0.05 * rng.standard_normal((m, 7))
    k = min(ncol, 8)
    Q, _ = np.linalg.qr(np.eye(k) + 0.15 * rng.standard_normal((k, k)))
    U[:, :k] = U[:, :k] @ Q
    return np.ascontiguousarray(U * (1.0 / np.sqrt(m)))


# ---- the device noise generator of esper_tau_snr_stack, restated in numpy (tests compare the two) -------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11) on uint32 arrays -> four uint32 arrays."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox_normal(seed: int, image: int, n_pixels: int) -> np.ndarray:
    """float64 [n_pixels] N(0,1) field of image `image` exactly as k8_dataprep.cu draws it: counter = (pixel pair,
    image, 0, 0), key = seed; two 53-bit uniforms per counter; Box-Muller (pixel 2k <- r cos, pixel 2k+1 <- r sin)."""
    pairs = (n_pixels + 1) // 2
    x0, x1, x2, x3 = philox4x32_10(np.arange(pairs, dtype=np.uint64), np.uint64(image), np.uint64(0), np.uint64(0),
                                   seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u0 = ((x0 >> np.uint64(5)).astype(np.float64) * 67108864.0 + (x1 >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0
    u1 = ((x2 >> np.uint64(5)).astype(np.float64) * 67108864.0 + (x3 >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0
    r = np.sqrt(-2.0 * np.log(1.0 - u0))
    z = np.empty(2 * pairs, dtype=np.float64)
    z[0::2] = r * np.cos(2.0 * np.pi * u1)
    z[1::2] = r * np.sin(2.0 * np.pi * u1)
    return z[:n_pixels]
