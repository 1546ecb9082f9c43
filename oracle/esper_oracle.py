"""CPU restatement (numpy/scipy) of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Every function cites the reference lines it follows (paths relative to the
upstream tree, evanseitz/ManifoldEM_ESPER).  The restatement is pinned in two ways
(see tests/golden/README.md and tools/gen_goldens.py):
  * golden outputs produced by executing the reference's own code in the authoring
    container (Distances.op non-CTF branch, GaussianBandwidth.op, DiffusionMaps.op,
    Generate_Nd_Rot.genNdRotations, Rotation_Histograms.op) are committed under
    tests/golden/ and compared with this module by `pytest -m "not gpu"`;
  * the reference ships no golden outputs or tests of its own for this path
    (SURVEY.md §4), so those generated vectors are the only pin.

Nothing here is imported by the product package.
"""
from __future__ import annotations

import math
import numpy as np

# --------------------------------------------------------------------------------------
# Stage 1 -- 1_Embedding/DM/Distances.py (non-CTF branch)
# --------------------------------------------------------------------------------------

def standardize_stack(stack_f32: np.ndarray) -> np.ndarray:
    """Per-image z-score, Distances.py:88-92.

    `PD_stack.data[i]/1.` stays float32 (numpy weak-scalar promotion), so mean, std and
    the two in-place updates are float32 arithmetic; the result is widened to float64
    when stored into `imgAll`.
    """
    stack_f32 = np.asarray(stack_f32)
    out = np.zeros(stack_f32.shape, dtype=np.float64)
    for i in range(stack_f32.shape[0]):
        img = stack_f32[i] / 1.
        img -= img.mean()
        img /= img.std()
        out[i] = img
    return out


def distances_as_written(img_all: np.ndarray, quiet: bool = True) -> np.ndarray:
    """The double Python loop of Distances.py:96-104 (one pair per iteration).

    Used for small cases and as the `as written` CPU baseline sample.  The reference
    also prints every (i, j); `quiet=False` reproduces that cost.
    """
    nS = img_all.shape[0]
    D = np.zeros((nS, nS))
    p = 2.
    for i in range(nS):
        for j in range(nS):
            if i > j:
                if not quiet:
                    print(i, j)
                D[i, j] = (np.sum(np.abs((img_all[i] - img_all[j])) ** p) / img_all[i].size) ** (1. / p)
            elif i == j:
                D[i, j] = 0
    return D + D.T - np.diag(np.diag(D))


def distances_direct(img_all: np.ndarray) -> np.ndarray:
    """Same quantity as Distances.py:96-104, one row of differences at a time (float64).

    Differs from `distances_as_written` only in the order of the float64 summation
    (<= a few ulp).
    """
    nS = img_all.shape[0]
    X = img_all.reshape(nS, -1)
    P = X.shape[1]
    D = np.zeros((nS, nS))
    for i in range(1, nS):
        diff = X[:i] - X[i]
        D[i, :i] = np.sqrt(np.einsum('ij,ij->i', diff, diff) / P)
    return D + D.T


def distances_gram_f64(img_all: np.ndarray) -> np.ndarray:
    """float64 Gram form of Distances.py:96-104 (dgemm); the full-size (N=2000) oracle.

    ||x-y||^2 = ||x||^2 + ||y||^2 - 2 x.y in float64 loses ~1e-12 relative on the
    SNR-0.1 data and ~4e-13 on pristine data (hard part 2); it is
    checked against `distances_direct` at small sizes in tests/test_oracle.py.
    """
    nS = img_all.shape[0]
    X = img_all.reshape(nS, -1)
    P = X.shape[1]
    n = np.einsum('ij,ij->i', X, X)
    G = X @ X.T
    D2 = (n[:, None] + n[None, :] - 2. * G) / P
    D2 = 0.5 * (D2 + D2.T)
    np.fill_diagonal(D2, 0.)
    return np.sqrt(np.maximum(D2, 0.))


# --------------------------------------------------------------------------------------
# Stage 2 -- 1_Embedding/DM/GaussianBandwidth.py
# --------------------------------------------------------------------------------------

def tanh_model(xx, aa0, aa1, aa2, aa3):
    """GaussianBandwidth.py:18-24."""
    return aa3 + aa2 * np.tanh(aa0 * xx + aa1)


def find_threshold(logEps: np.ndarray, D2: np.ndarray) -> float:
    """GaussianBandwidth.py:26-32 (the row sort there does not change the sum)."""
    eps = np.exp(logEps)
    d = 1. / (2. * np.max(eps)) * D2
    ss = np.sum(np.exp(-np.sort(d)))
    return max(-np.log(0.01 * ss / len(D2)), 10)


def eps_sweep_logsum(D: np.ndarray, logEps: np.ndarray, sequential: bool = True) -> np.ndarray:
    """GaussianBandwidth.py:36-44: L_k = log sum_{ij: t<thr} exp(-t), t = (1/(2 eps_k)) D_ij^2.

    `sequential=True` reproduces the reference's builtin `sum` (left-to-right float64)
    with `np.cumsum`; `False` uses numpy's pairwise `np.sum` (differs by <~1e-13 rel.).
    """
    logEps = np.asarray(logEps, dtype=np.float64)
    out = np.zeros(len(logEps))
    D2 = D * D
    thr = find_threshold(logEps, D2)
    for k in range(len(logEps)):
        eps = np.exp(logEps[k])
        d = 1. / (2. * eps) * D2
        d = -d[d < thr]
        W = np.exp(d)
        out[k] = np.log(np.cumsum(W)[-1] if sequential else np.sum(W))
    return out


def gaussian_bandwidth_op(D, logEps, a0, logSumWij=None, sequential=True):
    """GaussianBandwidth.op, GaussianBandwidth.py:34-59.

    Returns (popt, logSumWij, resnorm, R_squared).  The retry loop redraws `a0` from the
    global numpy RNG exactly where the reference does (after every fit), so seeding
    `np.random.seed(s)` before the call reproduces the reference run with the same seed.
    """
    from scipy.optimize import curve_fit, OptimizeWarning
    import warnings
    if logSumWij is None:
        logSumWij = eps_sweep_logsum(D, logEps, sequential=sequential)
    resnorm = np.inf
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', category=OptimizeWarning)
        while resnorm > 100:
            popt, pcov = curve_fit(tanh_model, logEps, logSumWij, p0=a0)
            resnorm = sum(np.sqrt(np.fabs(np.diag(pcov))))
            a0 = 1 * (np.random.rand(4, 1) - .5)
            res = logSumWij - tanh_model(logEps, popt[0], popt[1], popt[2], popt[3])
            ss_res = np.sum(res ** 2)
            ss_tot = np.sum((logSumWij - np.mean(logSumWij)) ** 2)
            R2 = 1 - (ss_res / ss_tot)
    return popt, logSumWij, resnorm, R2


def eps_from_popt(popt):
    """DiffusionMaps.py:91-93 -> (slope, ln_eps, eps)."""
    slope = popt[0] * popt[2]
    ln_eps = -(popt[1] / popt[0])
    return slope, ln_eps, np.exp(ln_eps)


# --------------------------------------------------------------------------------------
# Stage 3 -- 1_Embedding/DM/DiffusionMaps.py
# --------------------------------------------------------------------------------------

def markov_symmetric(Dist: np.ndarray, eps: float):
    """DiffusionMaps.py:109,131-152 -> (A, d, Ms).

    The reference forms dense diagonal matrices and multiplies them; every product with a
    diagonal matrix has exactly one non-zero term, so the element-wise form below performs
    the same roundings in the same order:
        M  = A @ inv(D)                  -> A_ij * (1/d_j)
        Ms = sqrtm(inv(D)) @ M @ sqrtm(D) -> (sqrt(1/d_i) * M_ij) * sqrt(d_j)
    (checked against the dense form by tools/gen_goldens.py and tests/test_oracle.py).
    """
    A = np.exp(-1. * ((Dist ** 2.) / (2. * eps)))
    d = np.array([np.sum(A[i], axis=0) for i in range(A.shape[0])])
    dinv = 1. / d
    M = A * dinv[None, :]
    Ms = (np.sqrt(dinv)[:, None] * M) * np.sqrt(d)[None, :]
    return A, d, Ms


def markov_symmetric_dense(Dist: np.ndarray, eps: float) -> np.ndarray:
    """The dense-matrix form exactly as DiffusionMaps.py:131-152 writes it (small m only)."""
    import scipy.linalg
    m = Dist.shape[0]
    A = np.exp(-1. * ((Dist ** 2.) / (2. * eps)))
    Dg = np.zeros((m, m))
    for i in range(m):
        Dg[i, i] = np.sum(A[i], axis=0)
    Dinv = np.linalg.inv(Dg)
    M = np.matmul(A, Dinv)
    Dhalf = scipy.linalg.sqrtm(Dg)
    Dinv_half = scipy.linalg.sqrtm(Dinv)
    return Dinv_half.dot(M).dot(Dhalf)


def dmap_embed(Dist: np.ndarray, eps: float):
    """DiffusionMaps.py:109-169 at a given eps -> (sdiag**2, U) as saved to PD_%s_val/vec.npy."""
    _, _, Ms = markov_symmetric(Dist, eps)
    U, s, _ = np.linalg.svd(Ms)
    return s ** 2., U


def diffusion_maps_op(Dist: np.ndarray, seed: int | None = None):
    """DiffusionMaps.op end to end (DiffusionMaps.py:72-93,109-169) -> dict.

    `seed` seeds the global numpy RNG before `a0` is drawn (the reference leaves it unseeded,
    DiffusionMaps.py:74).
    """
    if seed is not None:
        np.random.seed(seed)
    logEps = np.arange(-100, 100.2, 0.2)
    a0 = 1 * (np.random.rand(4, 1) - .5)
    popt, logSumWij, resnorm, R2 = gaussian_bandwidth_op(Dist, logEps, a0)
    slope, ln_eps, eps = eps_from_popt(popt)
    val, U = dmap_embed(Dist, eps)
    return dict(logEps=logEps, logSumWij=logSumWij, popt=popt, R2=R2, slope=slope,
                ln_eps=ln_eps, eps=eps, val=val, vec=U)


# ---------------------------------------------------------------------------------------------
# Stage 1, CTF branch -- 1_Embedding/DM/Distances.py:107-145, Generate_Filters.py -- SURVEY 8f-1
# ---------------------------------------------------------------------------------------------
def ctf_grid(N: int) -> np.ndarray:
    """Generate_Filters.py:25-32 (create_grid): radial frequency in units of Nyquist, centred."""
    if N % 2 == 1:
        a = np.arange(-(N - 1) / 2, (N - 1) / 2 + 1)
    else:
        a = np.arange(-(N / 2.), N / 2)
    X, Y = np.meshgrid(a, a)
    return (1. / (N / 2.)) * np.sqrt(X ** 2 + Y ** 2)


def ctf_value(k, Cs_mm, df, kev, ampc):
    """Generate_Filters.py:41-58 (op) with B = inf as at the call site Distances.py:121 (envelope = 1)."""
    Cs = Cs_mm * 1.0e7
    wav = (2 * 511.0) + kev
    wav = 12.3986 / np.sqrt(wav * kev)
    w1 = np.pi * Cs * wav * wav * wav
    w2 = np.pi * wav * df
    k2 = k * k
    wr = (0.5 * w1 * k2 - w2) * k2
    return (np.sin(wr) - ampc * np.cos(wr)) * np.exp(-k2 / np.inf)


def ctf_filters(N, px, Cs_mm, df, kev, ampc):
    """Generate_Filters.py:60-84 (gen_ctf): (CTF, Butterworth G), both ifftshift-ed to FFT order."""
    Q = ctf_grid(N)
    CTF = np.fft.ifftshift(ctf_value(Q / (2 * px), Cs_mm, df, kev, ampc))
    G = np.fft.ifftshift(np.sqrt(1. / (1 + (Q / 0.5) ** (2 * 8))))
    return CTF, G


def distances_ctf_op(stack_f32: np.ndarray, params: np.ndarray):
    """Distances.py:107-145: defocus-tolerant distances and the Wiener-filtered stack.

    params [nS, 5] = (pixel size, Cs [mm], defocus [A], voltage [keV], amplitude contrast) per image
    (the star-file columns read at Distances.py:68-72).  Returns (D [nS, nS] float64, wiener [nS, N, N] float32).
    """
    from scipy.fftpack import fft2, ifft2
    nS, N, _ = stack_f32.shape
    fy = complex(0) * np.ones((nS, N, N))
    CTF = np.zeros((nS, N, N))
    imgAll = np.zeros((nS, N, N))
    for ii in range(nS):
        tmp = stack_f32[ii]
        tmp = (tmp - tmp.mean()) / tmp.std()                              # :114-115 (float32 arithmetic)
        px, Cs, df, volt, ampc = params[ii]
        CTF[ii], G = ctf_filters(N, px, Cs, df, volt, ampc)               # :118-119
        image = ifft2(fft2(tmp.astype(np.float64)) * G).real               # :121-122 Butterworth low-pass
        fy[ii] = fft2(image.T)                                             # :124-125 (column-major flatten = transpose)
        imgAll[ii] = image                                                 # :126
    wienerD = 0.
    for i in range(nS):                                                    # gen_wiener :155-161, SNR = 5
        wienerD = wienerD + CTF[i] ** 2
    wienerD = -(wienerD + 1. / 5)                                          # :129
    wiener = np.zeros((nS, N, N), dtype=np.float32)
    for ii in range(nS):
        wiener[ii] = ifft2(fft2(imgAll[ii]) * (CTF[ii] / wienerD)).real   # :130-135
    fy = fy.reshape(nS, N ** 2)
    C = CTF.reshape(nS, N ** 2)
    CTFfy = C.conj() * fy                                                  # :141
    D = np.dot((abs(C) ** 2), (abs(fy) ** 2).T)                           # :142
    D = D + D.T - 2 * np.real(np.dot(CTFfy, CTFfy.conj().transpose()))     # :143
    D -= np.diag(D)                                                        # :144 (row-broadcast of the diagonal)
    with np.errstate(invalid='ignore'):
        D = np.sqrt(D)                                                     # :145
    return D, wiener


def distances_ctf_direct(stack_f32: np.ndarray, params: np.ndarray) -> np.ndarray:
    """The same quantity pair by pair without the Gram expansion: D_ij = |CTF_j fy_i - CTF_i fy_j| (cancellation-free
    check of the expansion at Distances.py:142-143)."""
    nS, N, _ = stack_f32.shape
    F = np.zeros((nS, N, N), dtype=complex)
    C = np.zeros((nS, N, N))
    for ii in range(nS):
        tmp = stack_f32[ii]
        tmp = (tmp - tmp.mean()) / tmp.std()
        C[ii], G = ctf_filters(N, *params[ii])
        F[ii] = np.fft.fft2(tmp.astype(np.float64)) * G
    D = np.zeros((nS, nS))
    for i in range(nS):
        for j in range(i + 1, nS):
            D[i, j] = D[j, i] = np.sqrt(np.sum(np.abs(C[j] * F[i] - C[i] * F[j]) ** 2))
    return D


# ---------------------------------------------------------------------------------------------
# PCA embedding (1_Embedding/PCA/PCA.py:33-71) -- SURVEY 8f-2
# ---------------------------------------------------------------------------------------------
def pca_op(stack_f32: np.ndarray, dim: int = 20):
    """PCA.py:46-68: Y [P, N] raw pixels as columns, minus the mean image; SVD; returns (eig_vals [N] = s^2,
    U [N, dim] = Y^T u[:, :dim])."""
    N = stack_f32.shape[0]
    P = int(np.prod(stack_f32.shape[1:]))
    Y = np.ndarray(shape=(P, N), dtype=float)                     # PCA.py:49
    for i in range(N):
        Y[:, i] = stack_f32[i].flatten()                          # PCA.py:50-51
    mean_all = np.mean(Y, axis=1)                                 # PCA.py:53
    for i in range(N):
        Y[:, i] -= mean_all                                       # PCA.py:54-55
    u, s, v = np.linalg.svd(Y, full_matrices=False)               # PCA.py:58
    eig_vals = s ** 2                                             # PCA.py:60
    W = np.hstack([u[:, i].reshape(P, 1) for i in range(dim)])    # PCA.py:66
    U = Y.T.dot(W)                                                # PCA.py:67
    return eig_vals, U


# --------------------------------------------------------------------------------------
# Stage 4 -- 2_Manifold_Rotations/Generate_Nd_Rot.py, Rotation_Histograms.py
# --------------------------------------------------------------------------------------

def plane_list(n: int):
    """Lexicographic plane order (0,1),(0,2),...,(n-2,n-1) walked by Generate_Nd_Rot.py:16-37."""
    return [(k, j) for k in range(n - 1) for j in range(k + 1, n)]


def gen_nd_rotations(n: int, theta) -> np.ndarray:
    """Generate_Nd_Rot.genNdRotations, Generate_Nd_Rot.py:4-45.

    Composite R = G_0 . G_1 ... G_{T-1} multiplied left to right with `ndarray.dot`, as the
    reference does (so the BLAS path, and hence the bits, are the same).
    """
    total = int(n * (n - 1.) / 2.)
    if isinstance(theta, list) and len(theta) != total:
        raise ValueError('%s angles required' % (total))
    if isinstance(theta, np.ndarray) and theta.shape[0] != total:
        raise ValueError('%sx1 array of angles required' % (total))
    R = None
    for r, (k, j) in enumerate(plane_list(n)):
        t = float(np.asarray(theta[r]).reshape(-1)[0])
        c, s = math.cos(t), math.sin(t)
        G = np.eye(n)
        G[k, k] = c
        G[k, j] = -1. * s
        G[j, k] = s
        G[j, j] = c
        R = G if R is None else R.dot(G)
    return R


def _fma(a: float, b: float, c: float) -> float:
    """Correctly rounded a*b + c (exact rational arithmetic; slow, small cases only)."""
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def gen_nd_rotations_givens(n: int, theta) -> np.ndarray:
    """Same product as `gen_nd_rotations`, applied as T column-pair updates.

    Multiplying on the right by G_r only changes columns (k, j); every other column is
    multiplied by 1 and summed with exact zeros.  A dgemm micro-kernel accumulates each output
    over the inner index in ascending order with fused multiply-adds, so the two-term sums are
        R[:,k] <- fma(R[:,j], s, fl(R[:,k]*c)) ;  R[:,j] <- fma(R[:,j], c, fl(R[:,k]*(-s)))
    which is the arithmetic of the C/CUDA code (esper_nd_rotation, K5) and reproduces the
    reference's `ndarray.dot` chain bit for bit on the golden vectors (tests/test_oracle.py).
    """
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    R = np.eye(n)
    for r, (k, j) in enumerate(plane_list(n)):
        c, s = math.cos(theta[r]), math.sin(theta[r])
        for i in range(n):
            a, b = float(R[i, k]), float(R[i, j])
            R[i, k] = _fma(b, s, a * c)
            R[i, j] = _fma(b, c, a * (-1. * s))
    return R


def normalize_manifold(U0: np.ndarray) -> np.ndarray:
    """Rotation_Histograms.py:100-106 (`normalize`, to_sum=False, copy=True)."""
    d = np.copy(U0)
    d -= np.min(d, axis=0)
    d /= np.ptp(d, axis=0)
    d = 2. * (d - np.min(d)) / np.ptp(d) - 1
    return d


def initial_subspace(U0: np.ndarray, dim: int, PCA: bool) -> np.ndarray:
    """Rotation_Histograms.py:108-112."""
    U = normalize_manifold(U0)
    return U[:, 0:dim] if PCA else U[:, 1:dim + 1]


def rot_hist_eval(U_init: np.ndarray, R: np.ndarray, v1: int, v2: int, bins: int = 51,
                  lo: float = -1.5, hi: float = 1.5):
    """One evaluation of the sweep body, Rotation_Histograms.py:304-312 -> (H, NumberOfZeros)."""
    U_rot = np.matmul(R, U_init.T).T
    proj = np.vstack((U_rot[:, v1], U_rot[:, v2])).T
    H, _ = np.histogramdd(proj, bins=bins, range=((lo, hi), (lo, hi)))
    zeros = 100 - (np.count_nonzero(H) / (float(bins ** 2))) * 100
    return H, zeros


def rot_sweep(U_init: np.ndarray, dim: int, comboList, Rij_final, itIdx: int = 1,
              start=-40, stop=40, step=.5, bins=51, collect=None):
    """Sweep hot loop, Rotation_Histograms.py:283-365 -> thetas_all [len(comboList), T].

    `collect`, if a list, receives (subspace index, Rij, step index, nonzero bin count).
    """
    T = int(dim * (dim - 1) / 2)
    thetas_all = np.zeros((T, len(comboList)))
    steps = np.abs(int((start - stop) / step))
    for ci, (v1, v2) in enumerate(comboList):
        thetas = np.zeros((T, 1))
        for Rij in Rij_final:
            meas = np.zeros(steps)
            theta_list = np.arange(start, stop, step)
            for si, t in enumerate(theta_list):
                thetas[Rij] = t * np.pi / 180
                R = gen_nd_rotations(dim, thetas)
                H, meas[si] = rot_hist_eval(U_init, R, v1, v2, bins)
                if collect is not None:
                    collect.append((ci, Rij, si, int(np.count_nonzero(H))))
            thetas[Rij] = theta_list[np.argmax(meas)] * np.pi / 180
        if itIdx > 1:
            thetas[0] += 45 * np.pi / 180
        thetas_all[:, ci] = thetas.T
    return thetas_all.T


def conic_fit2(U, v1, v2):
    """3_Manifold_Binning/ConicFit.py:205-264 -> (cF1, R2) (general conic, sympy rref solve)."""
    import sympy
    x, y = U[:, v1], U[:, v2]
    A = np.column_stack((x * y, y * y, x, y, np.ones(x.shape[0])))
    b = -1. * (x * x)
    AtA = np.matmul(A.T, A)
    Atb = np.matmul(A.T, b)
    rref = sympy.Matrix(np.column_stack((AtA, Atb))).rref(simplify=True)
    cF = np.array(rref[0].tolist(), dtype=float)[:, -1]
    tss = np.sum((b - np.mean(b)) ** 2)
    ssr = np.sum((b - np.matmul(A, cF)) ** 2)
    return np.array([1., cF[0], cF[1], cF[2], cF[3], cF[4]]), 1 - (ssr / tss)


def conic_fit3(U, v1, v2):
    """3_Manifold_Binning/ConicFit.py:267-283 -> (cF, R2) (parabola y = a x^2 + b x + c)."""
    import sympy
    x, b = U[:, v1], U[:, v2]
    A = np.column_stack((x ** 2, x, np.ones(x.shape[0])))
    AtA = np.matmul(A.T, A)
    Atb = np.matmul(A.T, b)
    rref = sympy.Matrix(np.column_stack((AtA, Atb))).rref()
    cF = np.array(rref[0].tolist(), dtype=float)[:, -1]
    tss = np.sum((b - np.mean(b)) ** 2)
    ssr = np.sum((b - np.matmul(A, cF)) ** 2)
    return cF, 1 - (ssr / tss)


def find_parabolas(dim, U_init, itIdx, CTF=True):
    """Rotation_Histograms.py:153-212 without the figures -> (R2_best_psi, U_init, itIdx)."""
    T = int(dim * (dim - 1) / 2)
    R2_best_vals = np.zeros(dim - 1)
    R2_best_psi = np.zeros(dim - 1, dtype=int)
    for v1 in range(dim - 1):
        R2_best = 0
        for v2 in range(v1 + 1, dim):
            R2 = conic_fit2(U_init, v1, v2)[1] if CTF else conic_fit3(U_init, v1, v2)[1]
            if R2 > R2_best and v2 != R2_best_psi[0]:
                R2_best = R2
                R2_best_vals[v1] = R2
                R2_best_psi[v1] = v2
    if (R2_best_vals[0] < 0.2) and (itIdx == 1):
        itIdx += 1
        thetas = np.zeros((T, 1))
        thetas[0] = 45 * np.pi / 180
        R = gen_nd_rotations(dim, thetas)
        U_init = np.matmul(R, U_init.T).T
        return find_parabolas(dim, U_init, itIdx, CTF)
    return R2_best_psi, U_init, itIdx


def rotation_plan(dim, R2_best_psi):
    """Rotation_Histograms.py:217-281 -> (comboList, saveEigs, comboFinal, Rij_final)."""
    comboList, saveEigs, catch = [], np.zeros(dim - 1), []
    for i in range(dim - 1):
        if i not in catch:
            if R2_best_psi[i] < i:
                comboList.append((i, i + 1))
                saveEigs[i] = i + 1
            else:
                comboList.append((i, R2_best_psi[i]))
                saveEigs[i] = R2_best_psi[i]
            catch.append(i)
            catch.append(R2_best_psi[i])
    comboFinal = []
    for c in range(1, dim - 1):
        c1, c2 = comboList[0]
        if c1 in comboList[c] or c2 in comboList[c]:
            continue
        comboFinal += [(c1, comboList[c][0]), (c1, comboList[c][1]),
                       (c2, comboList[c][0]), (c2, comboList[c][1])]
        break
    # Rij index r <-> plane (a, b) in lexicographic order; the reference walks (Rij_1, Rij_2)
    # one-based and tests the four comboFinal pairs in order at every r (lines 239-281).
    Rij_final = []
    for r, (a, b) in enumerate(plane_list(dim)):
        for q in range(4):
            pa, pb = comboFinal[q][0], comboFinal[q][1]
            if (a == pa and b == pb) or (b == pa and a == pb):
                Rij_final.append(r)
    return comboList, saveEigs, comboFinal, Rij_final


def rotation_histograms_op(U0, dim=6, PCA=False, CTF=True, collect=None):
    """Rotation_Histograms.op without I/O and figures -> dict of the three saved arrays."""
    U_init = initial_subspace(U0, dim, PCA)
    R2_best_psi, U_init, itIdx = find_parabolas(dim, U_init, 1, CTF)
    comboList, saveEigs, comboFinal, Rij_final = rotation_plan(dim, R2_best_psi)
    thetas_all = rot_sweep(U_init, dim, comboList, Rij_final, itIdx, collect=collect)
    return dict(RotMatrices=thetas_all, RotEigenfunctions=saveEigs,
                RotParameters=np.array([dim]), comboList=comboList, Rij_final=Rij_final,
                U_init=U_init, itIdx=itIdx)


# --------------------------------------------------------------------------------------
# Data preparation either side of the path
#   0_Data_Inputs/Pristine_AddTauSNR/AddNoise.py, Generate_TauSNR.py; Manifold_Binning.py:940-974
# --------------------------------------------------------------------------------------
def find_snr(image: np.ndarray, SNR: float):
    """AddNoise.py:17-39 -> (sig_mean, noise_std).  The reference flattens the float32 image into a Python list of
    numpy float32 scalars; `np.mean` / `np.var` of that list are float32 pairwise reductions over the same
    contiguous values, which is what the array calls below compute (bit-identical, see the golden test)."""
    img_1d = np.asarray(image).reshape(-1)
    img_std = np.sqrt(np.var(img_1d))
    keep = ~((-img_std * .25 < img_1d) & (img_1d < img_std * .25))
    sig = img_1d[keep]
    sig_mean = np.mean(sig)
    noise_std = np.sqrt(np.var(sig) / SNR)
    return sig_mean, noise_std


def background_mask(h: int, w: int) -> np.ndarray:
    """AddNoise.py:54-59: True inside the circle of radius int(h/2) around (int(w/2), int(h/2))."""
    center = [int(w / 2), int(h / 2)]
    radius = int(h / 2)
    Y, X = np.ogrid[:h, :w]
    return np.sqrt((X - center[0]) ** 2 + (Y - center[1]) ** 2) <= radius


def standardize_background(image: np.ndarray):
    """AddNoise.py:52-75: (image - mean(bg)) / std(bg), bg = pixels outside the inscribed circle; 0 if std(bg) == 0."""
    mask = background_mask(*image.shape[:2])
    bg = image[~mask]
    bg_mean, bg_std = np.mean(bg), np.std(bg)
    if bg_std == 0:
        return 0
    return (image - bg_mean) / bg_std


def add_noise_op(img_orig: np.ndarray, SNR: float, gauss: np.ndarray | None = None) -> np.ndarray:
    """AddNoise.op (:77-86).  gauss=None draws `np.random.normal(mean, std, shape)` from numpy's global generator like
    the reference (:45); otherwise `gauss` is a field of N(0,1) values and the noise is mean + std * gauss."""
    sig_mean, noise_std = find_snr(img_orig, SNR)
    row, col = img_orig.shape
    if gauss is None:
        noise = np.random.normal(sig_mean, noise_std, (row, col))
    else:
        noise = sig_mean + noise_std * np.asarray(gauss, dtype=np.float64)
    return standardize_background(img_orig + noise)


def tau_snr_stack(pristine: np.ndarray, tau: int, SNR: float | None, gauss: np.ndarray | None = None) -> np.ndarray:
    """Generate_TauSNR.py:56-64: float32 [ss*tau, box, box], image s*tau + t = copy t of state s (mode-2 MRC cast)."""
    ss, box, _ = pristine.shape
    out = np.empty((ss * tau, box, box), dtype=np.float32)
    k = 0
    for i in range(ss):
        for j in range(tau):
            if SNR:
                out[k] = add_noise_op(pristine[i], SNR, None if gauss is None else gauss[k])
            else:
                out[k] = pristine[i]
            k += 1
    return out


def bin_accumulate(stack: np.ndarray, order, offsets) -> np.ndarray:
    """3_Manifold_Binning/Manifold_Binning.py:338,940-974: `finalStackBin[b] += init_stack.data[img]` (float64
    accumulator, float32 images) for the images order[offsets[b]:offsets[b+1]] of every bin, in that order."""
    n_bins = len(offsets) - 1
    out = np.zeros((n_bins,) + stack.shape[1:])
    for b in range(n_bins):
        for k in range(offsets[b], offsets[b + 1]):
            out[b] += stack[order[k]]
    return out
