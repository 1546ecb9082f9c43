"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle and the
golden vectors generated from the reference.  Tolerances are the ones BASELINE.json's north_star
states: distances 1e-5 relative, eigenvalues 1e-6 relative at identical epsilon, eigenvectors 1e-4
up to sign, histograms / rotation angles bit-exact."""
import os

import numpy as np
import pytest

from conftest import golden
from manifoldem_esper_b200 import (PCA, AddNoise, DiffusionMaps, Distances, GaussianBandwidth, Generate_TauSNR,
                                   Rotation_Histograms as RH, _capi, mrc, synth)
from oracle import esper_oracle as O

pytestmark = pytest.mark.gpu

D_TOL = 1e-5          # relative, on D
D_FLOOR = 1e-2        # |dD| <= D_TOL * max(D, D_FLOOR)  (exact duplicates have D = 0)
VAL_TOL = 1e-6
VEC_TOL = 1e-4
LOG_EPS = np.arange(-100, 100.2, 0.2)


def d_err(D, Dr):
    return float(np.max(np.abs(D - Dr) / np.maximum(Dr, D_FLOOR)))


def vec_err(u, v):
    return min(np.abs(u - v).max(), np.abs(u + v).max())


# ---- stage 1 ----------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['stage123_small', 'stage123_pristine', 'stage1_dupes'])
def test_distances_vs_reference_golden(ctx, name):
    g = golden(name + '.npz')
    D = ctx.dist_rms(g['stack'])
    assert d_err(D, g['D']) < D_TOL
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0) and np.all(np.isfinite(D))
    if name == 'stage1_dupes':
        n = D.shape[0]
        assert np.all(D[np.arange(0, n, 2), np.arange(1, n, 2)] == 0)       # exact duplicates -> exactly 0


@pytest.mark.parametrize('refine', [0, 1])
@pytest.mark.parametrize('n', [1, 2, 3, 127, 129, 257, 513])
@pytest.mark.parametrize('P', [16, 1000, 128 * 128])
def test_distances_tile_and_k_tails(ctx, n, P, refine):
    """refine=0: the tensor-core contraction alone (no pair is recomputed); refine=1: the default flags, which compute stacks of
    at most 512 images pair by pair in the reference's arithmetic."""
    X = np.random.default_rng(n * 7 + P).standard_normal((n, P)).astype(np.float32)
    flags = _capi.DIST_DEFAULT if refine else (_capi.DIST_STANDARDIZE | _capi.DIST_CENTER)
    D = ctx.dist_rms(X, flags=flags)
    Dr = O.distances_gram_f64(O.standardize_stack(X.reshape(n, 1, P)))
    assert D.shape == (n, n) and d_err(D, Dr) < D_TOL
    if refine and 1 < n <= 512:
        assert d_err(D, O.distances_direct(O.standardize_stack(X.reshape(n, 1, P)))) < 1e-11


def test_distances_unstandardized_and_unaligned_P(ctx):
    X = np.random.default_rng(5).standard_normal((70, 999)).astype(np.float32) * 3 + 1
    D = ctx.dist_rms(X, flags=_capi.DIST_CENTER)                             # no z-score: caller's values as is
    Dr = O.distances_gram_f64(X.astype(np.float64).reshape(70, 1, 999))
    assert d_err(D, Dr) < D_TOL


def test_distances_linearity_property(ctx):
    """Size-independent property: scaling all images by s scales D by s (no standardisation)."""
    X = np.random.default_rng(9).standard_normal((300, 2048)).astype(np.float32)
    D1 = ctx.dist_rms(X, flags=_capi.DIST_CENTER)
    D2 = ctx.dist_rms(X * np.float32(4.0), flags=_capi.DIST_CENTER)          # exact power-of-two scaling
    assert np.array_equal(D2, 4.0 * D1)


def test_distances_full_size_vs_f64_gram(ctx):
    """BASELINE config 2 shape: N=2000, 250x250, tau=5, SNR 0.1 against the float64 Gram oracle."""
    stack = synth.synthetic_pd(pd=1, box=250, n_side=20, tau=5, snr=0.1)
    Dr = O.distances_gram_f64(O.standardize_stack(stack))
    D = ctx.dist_rms(stack)
    assert d_err(D, Dr) < D_TOL
    d2, d2r = D * D, Dr * Dr
    off = ~np.eye(D.shape[0], dtype=bool)
    assert abs(np.mean((d2 - d2r)[off] / d2r[off])) < 1e-7                   # no systematic accumulation bias
    # resident-stack chaining gives the same bits as the host-pointer call
    assert np.array_equal(ctx.dist_rms(None, n=2000, P=62500), D)


def test_distances_pristine_full_box(ctx):
    """BASELINE config 1 stand-in: noise-free N=400 stack at the full 250x250 box (small distances, the Gram
    cancellation regime) against direct float64 differences."""
    stack = synth.synthetic_pd(pd=4, box=250, n_side=20, tau=1, snr=None)
    Dr = O.distances_direct(O.standardize_stack(stack))
    D = ctx.dist_rms(stack)
    assert Dr[Dr > 0].min() < 0.1 and d_err(D, Dr) < D_TOL
    assert np.array_equal(D, D.T)


def test_distances_row_block_equals_full(ctx):
    """Row-block path (oversized PD): every block holds the same bits as the rows of the full matrix."""
    X = np.random.default_rng(11).standard_normal((700, 3000)).astype(np.float32)
    ctx.dist_config(split_k=4)
    D = ctx.dist_rms(X)
    for row0, nrows in [(0, 256), (256, 384), (640, 60)]:
        Dr = ctx.dist_rms_rows(X, 700, 3000, row0, nrows, download=True)
        assert np.array_equal(Dr, D[row0:row0 + nrows])
    with pytest.raises(ValueError):
        ctx.dist_rms_rows(X, 700, 3000, 100, 200)                             # row0 must be a multiple of 128
    ctx.dist_config(split_k=0)


def test_distances_errors(ctx):
    with pytest.raises(ValueError):
        ctx.dist_rms(None, n=7, P=9)                                          # nothing resident of that shape
    with pytest.raises(ValueError):
        ctx.dist_rms(np.zeros((0, 4), np.float32))


def test_distances_op_files(ctx, tmp_path):
    g = golden('stage123_small.npz')
    root = tmp_path / 'tree'
    dm = root / '1_Embedding' / 'DM'
    data = root / '0_Data_Inputs' / 'CTF5k15k_SNRpt1_ELS_2D'
    os.makedirs(dm); os.makedirs(data)
    mrc.write_stack(str(data / 'PD003_SNR_tau5_stack.mrcs'), g['stack'])
    Distances.op(str(dm), '003', CTF=False, ctx=ctx)
    D = np.load(str(dm / 'Data_Distances' / 'PD_003_dist.npy'))
    assert D.dtype == np.float64 and D.flags.c_contiguous and d_err(D, g['D']) < D_TOL
    assert DiffusionMaps.op(str(dm), '003', k=8, ctx=ctx) is None                 # the reference's op returns nothing (:57)
    val = np.load(str(dm / 'Data_Manifolds' / 'PD_003_val.npy'))
    vec = np.load(str(dm / 'Data_Manifolds' / 'PD_003_vec.npy'))
    assert val.shape == (8,) and vec.shape == (D.shape[0], 8) and abs(val[0] - 1) < 1e-12


# ---- stage 2 ----------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['stage123_small', 'stage123_pristine', 'stage23_medium', 'stage1_dupes'])
def test_eps_sweep_vs_reference_golden(ctx, name):
    g = golden(name + '.npz')
    D = g['D']
    L = GaussianBandwidth.sweep(D, LOG_EPS, ctx=ctx)
    Lr = g['logSumWij'] if 'logSumWij' in g else O.eps_sweep_logsum(D, LOG_EPS)
    assert np.max(np.abs(L - Lr)) < 1e-9
    m = D.shape[0]
    assert abs(L[0] - np.log((D == 0).sum())) < 1e-12
    assert abs(L[-1] - np.log(m * m)) < 1e-12                                  # plateaus: log(#zeros), log(m^2)


def test_bandwidth_op_same_seed_same_epsilon(ctx):
    g = golden('stage23_medium.npz')
    np.random.seed(int(g['seed']))
    a0 = 1 * (np.random.rand(4, 1) - .5)
    popt, L, resnorm, R2 = GaussianBandwidth.op(g['D'], LOG_EPS, a0, ctx=ctx)
    eps = np.exp(-(popt[1] / popt[0]))
    assert abs(eps - float(g['eps'])) / float(g['eps']) < 1e-7
    assert abs(R2 - float(g['R2'])) < 1e-9


def test_eps_sweep_general_coefficients(ctx):
    """Unsorted coefficients, no truncation, zero coefficient: the closed-form shortcuts stay exact."""
    D = golden('stage123_small.npz')['D']
    coef = np.array([0.0, 1e-30, 3.0, 0.5, 1e6, 2.0, 1e-3])
    for thr in (10.0, np.inf, 0.5):
        L = ctx.eps_sweep(D, coef, thr)
        ref = []
        for c in coef:
            d = c * (D * D)
            ref.append(np.log(np.sum(np.exp(-d[d < thr]))))
        assert np.allclose(L, ref, rtol=0, atol=1e-12)


def test_eps_sweep_nonsymmetric_input(ctx):
    """A caller-supplied D need not be symmetric for the sweep (the reference sums all m^2 entries)."""
    D = np.abs(np.random.default_rng(2).standard_normal((150, 150))) + 0.1
    np.fill_diagonal(D, 0.0)
    coef = 1. / (2. * np.exp(np.arange(-12, 12, 0.2)))
    L = ctx.eps_sweep(D, coef, 10.0)
    ref = [np.log(np.sum(np.exp(-(c * D * D)[(c * D * D) < 10.0]))) for c in coef]
    assert np.allclose(L, ref, rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        ctx.dmap_embed(D, 0.3, k=4)                                           # the eigensolver needs symmetry


@pytest.fixture
def f16_operands(ctx):
    """3xFP16 operand planes are the default; the fixture pins the mode so these tests also hold under ESPER_GRAM_F16=0."""
    ctx.gram_config(1)
    yield ctx
    ctx.gram_config(1)


@pytest.mark.parametrize('n,P', [(129, 1000), (257, 16), (300, 4096), (640, 16384), (700, 3000)])
def test_f16_operands_tile_and_k_tails(f16_operands, n, P):
    """esper_gram_config(1): 3xFP16 operand planes (kind::f16) against the float64 oracle and the 3xTF32 path."""
    ctx = f16_operands
    X = np.random.default_rng(n * 7 + P).standard_normal((n, P)).astype(np.float32)
    D = ctx.dist_rms(X)
    assert ctx.gram_info() == 1
    Dr = O.distances_gram_f64(O.standardize_stack(X.reshape(n, 1, P)))
    assert D.shape == (n, n) and d_err(D, Dr) < D_TOL
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    ctx.gram_config(0)
    D0 = ctx.dist_rms(X)
    assert ctx.gram_info() == 0 and d_err(D, D0) < 2e-6


def test_f16_operands_row_block_equals_full(f16_operands):
    """Row-block path with FP16 planes (gram_kernel<true>): every block holds the same bits as the rows of the full matrix."""
    ctx = f16_operands
    X = np.random.default_rng(11).standard_normal((700, 3000)).astype(np.float32)
    ctx.dist_config(split_k=4)
    try:
        D = ctx.dist_rms(X)
        assert ctx.gram_info() == 1
        for row0, nrows in [(0, 256), (256, 384), (640, 60)]:
            Dr = ctx.dist_rms_rows(X, 700, 3000, row0, nrows, download=True)
            assert ctx.gram_info() == 1 and np.array_equal(Dr, D[row0:row0 + nrows])
    finally:
        ctx.dist_config(split_k=0)
    assert d_err(D, O.distances_gram_f64(O.standardize_stack(X.reshape(700, 1, 3000)))) < D_TOL


def test_f16_operands_fall_back_where_the_range_is_not_bounded(f16_operands):
    ctx = f16_operands
    X = np.random.default_rng(5).standard_normal((300, 999)).astype(np.float32) * 3e6 + 1
    D = ctx.dist_rms(X, flags=_capi.DIST_CENTER)                             # no z-score: 3xTF32 planes
    assert ctx.gram_info() == 0
    Dr = O.distances_gram_f64(X.astype(np.float64).reshape(300, 1, 999))
    assert d_err(D, Dr) < D_TOL
    ctx.dist_rms(X[:100])                                                    # a single 128-row tile
    assert ctx.gram_info() == 0
    with pytest.raises(ValueError):
        ctx.gram_config(3)


def test_f16_operands_outlier_pixels(f16_operands):
    """A hot pixel makes one z-score ~ sqrt(P): the largest value the FP16 planes must hold."""
    ctx = f16_operands
    X = np.random.default_rng(21).standard_normal((260, 4096)).astype(np.float32) * 1e-3
    X[7, 100] = 5e4
    X[9, 4000] = -7e4
    D = ctx.dist_rms(X)
    assert ctx.gram_info() == 1 and np.all(np.isfinite(D))
    assert d_err(D, O.distances_gram_f64(O.standardize_stack(X.reshape(260, 1, 4096)))) < D_TOL


def test_f16_operands_pristine_and_duplicates(f16_operands):
    """Noise-free stack (Gram cancellation regime) with every image duplicated: exact zeros, refinement, symmetry."""
    ctx = f16_operands
    stack = synth.synthetic_pd(pd=4, box=250, n_side=14, tau=2, snr=None)
    Dr = O.distances_direct(O.standardize_stack(stack))
    D = ctx.dist_rms(stack)
    assert ctx.gram_info() == 1
    assert Dr[Dr > 0].min() < 0.1 and d_err(D, Dr) < D_TOL
    assert np.array_equal(D, D.T)
    n = D.shape[0]
    assert np.all(D[np.arange(0, n, 2), np.arange(1, n, 2)] == 0)


def test_f16_operands_full_size_pipeline(f16_operands):
    """BASELINE config 2 with FP16 operand planes: distances, no accumulation bias, eigenvalues at fixed epsilon."""
    ctx = f16_operands
    stack = synth.synthetic_pd(pd=1, box=250, n_side=20, tau=5, snr=0.1)
    Dr = O.distances_gram_f64(O.standardize_stack(stack))
    D = ctx.dist_rms(stack)
    assert ctx.gram_info() == 1 and d_err(D, Dr) < D_TOL
    d2, d2r = D * D, Dr * Dr
    off = ~np.eye(D.shape[0], dtype=bool)
    assert abs(np.mean((d2 - d2r)[off] / d2r[off])) < 1e-7
    eps = 0.2834
    val_r, _ = O.dmap_embed(Dr, eps)
    val, vec, _ = ctx.dmap_embed(None, eps, k=10, m=2000)
    assert np.max(np.abs(val - val_r[:10]) / val_r[:10]) < VAL_TOL


@pytest.mark.parametrize('name', ['stage123_small', 'stage123_pristine', 'stage23_medium', 'stage1_dupes'])
def test_moment_sweep_vs_reference_golden(ctx, name):
    """esper_sweep_config(1): the moment sweep against the reference run's logSumWij and against the general kernel."""
    g = golden(name + '.npz')
    D = g['D']
    Lr = g['logSumWij'] if 'logSumWij' in g else O.eps_sweep_logsum(D, LOG_EPS)
    ctx.sweep_config(0)
    try:
        L0 = GaussianBandwidth.sweep(D, LOG_EPS, ctx=ctx)
        assert ctx.sweep_info()[0] == 0
    finally:
        ctx.sweep_config(1)
    L1 = GaussianBandwidth.sweep(D, LOG_EPS, ctx=ctx)
    path, bins = ctx.sweep_info()
    assert path == 1 and 1 <= bins <= 32
    assert np.array_equal(np.isfinite(L1), np.isfinite(Lr))
    fin = np.isfinite(Lr)
    assert np.max(np.abs(L1[fin] - Lr[fin])) < 1e-9
    assert np.max(np.abs(L1[fin] - L0[fin])) < 1e-12


def test_moment_sweep_nonsymmetric_thresholds_and_fallbacks(ctx):
    rng = np.random.default_rng(5)
    D = np.sqrt(1.9 + 0.1 * rng.random((300, 300)))                           # concentrated like a high-noise stack
    np.fill_diagonal(D, 0.0)
    D[3, 7] = 0.0                                                             # an off-diagonal zero, non-symmetric matrix
    coef = 1. / (2. * np.exp(LOG_EPS))
    ctx.sweep_config(1)
    try:
        for thr in (10.0, 3.0, 25.0):
            L = ctx.eps_sweep(D, coef, thr)
            assert ctx.sweep_info() == (1, 1)
            ref = []
            for c in coef:
                t = c * (D * D)
                ref.append(np.log(np.sum(np.exp(-t[t < thr]))))
            assert np.allclose(L, ref, rtol=0, atol=1e-12), thr
        # returns to the general kernel: spread of more than 32 bins, unsorted coefficients, no truncation
        W = np.exp(rng.uniform(-12, 12, (120, 120)))
        L = ctx.eps_sweep(W, coef, 10.0)
        assert ctx.sweep_info()[0] == 0
        ref = [np.log(np.sum(np.exp(-(c * W * W)[(c * W * W) < 10.0]))) for c in coef]
        assert np.allclose(L, ref, rtol=0, atol=1e-12)
        ctx.eps_sweep(D, coef[::-1].copy(), 10.0)
        assert ctx.sweep_info()[0] == 0
        ctx.eps_sweep(D, coef, np.inf)
        assert ctx.sweep_info()[0] == 0
        with pytest.raises(ValueError):
            ctx.sweep_config(5)
    finally:
        ctx.sweep_config(1)


def test_moment_sweep_row_block_entry_points(ctx):
    """esper_moment_stats_rows / esper_moment_accum_rows on a resident row block against the host plan and the numpy
    statement of the accumulation (tools/sweep_moments_proto.py); finalize reproduces the general row-block sweep."""
    import importlib.util
    from manifoldem_esper_b200 import moment_sweep as MSW
    spec = importlib.util.spec_from_file_location('sweep_moments_proto', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools', 'sweep_moments_proto.py'))
    proto = importlib.util.module_from_spec(spec); spec.loader.exec_module(proto)
    X = np.random.default_rng(31).standard_normal((600, 4096)).astype(np.float32)
    row0, nrows = 256, 344                                                    # a short block must be the last one
    Dr = ctx.dist_rms_rows(X, 600, 4096, row0, nrows, download=True)
    coef = 1. / (2. * np.exp(LOG_EPS))
    st = ctx.moment_stats_rows(nrows, 600)
    q = (Dr * Dr).ravel()
    assert st[0] == q[q != 0].min() and st[1] == q.max() and st[2] == (q == 0).sum() and st[3] == q.size
    status, pl = MSW.plan(st[0], st[1], st[2], st[3], coef, 10.0)
    assert status == 0
    tab = ctx.moment_accum_rows(nrows, 600, coef, 10.0, pl)
    assert np.allclose(tab, proto.accumulate_numpy(q, np.ones_like(q), pl, coef, 10.0), rtol=1e-12, atol=1e-12)
    sums = MSW.finalize(pl, tab, coef, 10.0)
    ref = ctx.eps_sweep_rows(nrows, 600, coef, 10.0)
    assert np.allclose(sums, ref, rtol=1e-12, atol=0)
    bad = pl.copy(); bad[0] = 1.0
    with pytest.raises(ValueError):
        ctx.moment_accum_rows(nrows, 600, coef, 10.0, bad)


def test_moment_sweep_full_size(ctx):
    """BASELINE config 2 shape: D^2 spans 5 %, one bin; both kernels on the same resident matrix."""
    rng = np.random.default_rng(11)
    X = rng.standard_normal((2000, 8192))
    G = X @ X.T
    nrm = np.diag(G)
    D = np.sqrt(np.maximum(nrm[:, None] + nrm[None, :] - 2 * G, 0) / 8192) * (1.38 / np.sqrt(2))
    D = 0.5 * (D + D.T)
    np.fill_diagonal(D, 0.0)
    coef = 1. / (2. * np.exp(LOG_EPS))
    ctx.sweep_config(0)
    try:
        L0 = ctx.eps_sweep(D, coef, 10.0)
    finally:
        ctx.sweep_config(1)
    L1 = ctx.eps_sweep(None, coef, 10.0, m=2000)
    path, bins = ctx.sweep_info()
    assert path == 1 and bins <= 2
    assert np.max(np.abs(L1 - L0)) < 1e-12
    assert np.max(np.abs(L1 - O.eps_sweep_logsum(D, LOG_EPS, sequential=False))) < 1e-10


# ---- stage 3 ----------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['stage123_pristine', 'stage23_medium'])
def test_markov_matrix_vs_oracle(ctx, name):
    g = golden(name + '.npz')
    A, d, Ms_r = O.markov_symmetric(g['D'], float(g['eps']))
    deg, Ms = ctx.markov(g['D'], float(g['eps']))
    assert np.max(np.abs(deg - d) / d) < 1e-14
    assert np.max(np.abs(Ms - Ms_r)) / np.max(np.abs(Ms_r)) < 1e-14
    assert np.allclose(Ms, Ms.T, rtol=1e-8, atol=1e-8)                         # the reference's is_symmetric check


@pytest.mark.parametrize('name,k', [('stage123_pristine', 12), ('stage23_medium', 16), ('stage123_small', 16)])
def test_embedding_vs_reference_golden_at_fixed_eps(ctx, name, k):
    g = golden(name + '.npz')
    val, vec, iters = ctx.dmap_embed(g['D'], float(g['eps']), k=k)
    assert np.max(np.abs(val - g['val'][:k]) / g['val'][:k]) < VAL_TOL
    assert vec_err(vec[:, 0], g['vec'][:, 0]) < VEC_TOL and np.all(vec[:, 0] > 0)
    lam = np.sqrt(g['val'][:k + 1])
    gaps = np.minimum(np.abs(np.diff(lam))[:-1], np.abs(np.diff(lam))[1:]) if k > 2 else []
    for i in range(1, k - 1):                                                 # well separated -> per vector
        if gaps[i - 1] > 1e-3 * lam[1]:
            assert vec_err(vec[:, i], g['vec'][:, i]) < VEC_TOL, i
    # all k as a subspace (principal angles), orthonormality, residuals
    s = np.linalg.svd(vec.T @ g['vec'][:, :k], compute_uv=False)
    assert s.min() > 1 - 1e-6 or lam[k - 1] - lam[k] < 1e-3 * lam[1]
    assert np.max(np.abs(vec.T @ vec - np.eye(k))) < 1e-10
    Ms = O.markov_symmetric(g['D'], float(g['eps']))[2]
    assert np.max(np.linalg.norm(Ms @ vec - vec * np.sqrt(val)[None, :], axis=0)) < 1e-9


def test_embed_end_to_end_matches_reference_run(ctx):
    """Distances -> sweep -> seeded fit -> embedding on the GPU vs the reference run with the same seed."""
    g = golden('stage123_pristine.npz')
    D = ctx.dist_rms(g['stack'])
    np.random.seed(int(g['seed']))
    res = DiffusionMaps.embed(D, k=8, ctx=ctx, verbose=False)
    assert d_err(D, g['D']) < 1e-11                                           # small stack: every pair in the reference's arithmetic
    assert abs(res['eps'] - float(g['eps'])) / float(g['eps']) < 1e-7
    val_r, U_r = O.dmap_embed(g['D'], res['eps'])                             # parity at identical epsilon
    assert np.max(np.abs(res['val'] - val_r[:8]) / val_r[:8]) < VAL_TOL
    for i in range(4):
        assert vec_err(res['vec'][:, i], U_r[:, i]) < VEC_TOL


def test_full_size_pipeline_vs_oracle(ctx):
    """BASELINE config 2 end to end on the GPU (stack -> distances -> sweep -> embedding) against the oracle run on
    the same stack, at the north-star tolerances: D 1e-5, sweep 1e-9, eigenvalues 1e-6 and eigenvectors 1e-4 at
    identical epsilon (the separated ones per vector, the bulk-edge cluster as a subspace)."""
    stack = synth.synthetic_pd(pd=7, box=250, n_side=20, tau=5, snr=0.1)
    D_ref = O.distances_gram_f64(O.standardize_stack(stack))
    D = ctx.dist_rms(stack)
    assert d_err(D, D_ref) < D_TOL
    L = GaussianBandwidth.sweep(None, LOG_EPS, ctx=ctx, m=2000)               # on the resident GPU matrix
    L_ref = O.eps_sweep_logsum(D_ref, LOG_EPS, sequential=False)
    assert np.max(np.abs(L - L_ref)) < 1e-6                                   # carries the 2e-7 distance error
    a0 = np.array([[-0.2], [0.3], [-0.4], [0.1]])
    popt, _, _ = GaussianBandwidth.fit(LOG_EPS, L_ref, a0)
    eps = float(np.exp(-(popt[1] / popt[0])))
    val_r, U_r = O.dmap_embed(D_ref, eps)
    k = 10
    val, vec, iters = ctx.dmap_embed(None, eps, k=k, m=2000)                  # GPU distances -> GPU embedding
    assert np.max(np.abs(val - val_r[:k]) / val_r[:k]) < VAL_TOL
    lam = np.sqrt(val_r[:k + 1])
    for i in range(k):
        gap = min(abs(lam[i] - lam[i - 1]) if i else 1.0, abs(lam[i] - lam[i + 1]))
        if gap > 2e-3 * lam[1]:
            assert vec_err(vec[:, i], U_r[:, i]) < VEC_TOL, i
    sep = [i for i in range(k) if min(abs(lam[i] - lam[i - 1]) if i else 1.0, abs(lam[i] - lam[i + 1])) > 2e-3 * lam[1]]
    assert len(sep) >= 5                                                      # steady state + the signal eigenvectors
    # the whole leading block as a subspace (principal angles) unless the cut falls inside a cluster
    s_ = np.linalg.svd(vec.T @ U_r[:, :k], compute_uv=False)
    assert s_.min() > 1 - 1e-6 or lam[k - 1] - lam[k] < 2e-3 * lam[1]


# ---- stage 1, CTF branch -----------------------------------------------------------
def offdiag_rel(D, Dr):
    off = ~np.eye(D.shape[0], dtype=bool)
    return float(np.max(np.abs(D - Dr)[off] / Dr[off]))


@pytest.mark.parametrize('shape', [(5, 250, 250), (9, 32, 32), (7, 25, 25), (4, 7), (3, 129), (6, 1000), (2, 128 * 128)])
def test_image_stats_bit_identical_to_numpy(ctx, shape):
    """The float32 statistics of the CTF branch reproduce numpy's pairwise summation bit for bit."""
    rng = np.random.default_rng(sum(shape))
    x = (rng.standard_normal(shape) * 3 + 1.5).astype(np.float32)
    mean, std = ctx.image_stats(x.reshape(shape[0], -1))
    for i in range(shape[0]):
        assert mean[i] == x[i].mean() and std[i] == x[i].std(), i


@pytest.mark.parametrize('name', ['ctf_small', 'ctf_odd', 'ctf_pristine'])
def test_ctf_distances_vs_reference_golden(ctx, name):
    """Defocus-tolerant distances and Wiener-filtered stack against the reference's Distances.op (CTF = True)."""
    g = golden(name + '.npz')
    D, W = ctx.ctf_dist(g['stack'], g['params'])
    assert offdiag_rel(D, g['D']) < D_TOL
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0) and np.all(np.isfinite(D))
    assert W.dtype == np.float32 and np.max(np.abs(W - g['wiener'])) < 2e-6 * np.abs(g['wiener']).max()


def test_ctf_refinement_of_near_duplicates(ctx):
    """Noise-free images that differ only in defocus, plus exact duplicates: the Gram expansion cancels, the
    pair-by-pair pass restores 1e-5 and exact zeros."""
    st, pa = synth.ctf_stack(pd=2, box=48, n_side=4, tau=3, snr=None)
    st[1], pa[1] = st[0], pa[0]                                               # an exact duplicate
    Dr = O.distances_ctf_direct(st, pa)
    D, _ = ctx.ctf_dist(st, pa, wiener=False)
    off = ~np.eye(len(D), dtype=bool)
    assert D[0, 1] == 0.0 and D[1, 0] == 0.0
    m = off & (Dr > 0)
    assert np.max(np.abs(D - Dr)[m] / Dr[m]) < D_TOL
    D_raw, _ = ctx.ctf_dist(st, pa, flags=0, wiener=False)                    # without the pass: worse on close pairs
    assert np.max(np.abs(D_raw - Dr)[m] / Dr[m]) >= np.max(np.abs(D - Dr)[m] / Dr[m])


def test_ctf_medium_vs_oracle_and_chain(ctx):
    """N=400 images of 64x64 at SNR 0.1 with per-image defocus; then the resident matrix feeds stages 2-3."""
    st, pa = synth.ctf_stack(pd=6, box=64, n_side=10, tau=4, snr=0.1)
    Dr, Wr = O.distances_ctf_op(st, pa)
    D, W = ctx.ctf_dist(st, pa)
    assert offdiag_rel(D, Dr) < D_TOL
    assert np.max(np.abs(W - Wr)) < 2e-6 * np.abs(Wr).max()
    L = GaussianBandwidth.sweep(None, LOG_EPS, ctx=ctx, m=400)               # sweep on the resident CTF distances
    assert np.max(np.abs(L - O.eps_sweep_logsum(Dr, LOG_EPS, sequential=False))) < 1e-4
    popt, _, _ = GaussianBandwidth.fit(LOG_EPS, L, np.array([[-0.2], [0.3], [-0.4], [0.1]]))
    eps = float(np.exp(-(popt[1] / popt[0])))
    val, vec, _ = ctx.dmap_embed(None, eps, k=6, m=400)
    val_r, U_r = O.dmap_embed(D, eps)
    assert np.max(np.abs(val - val_r[:6]) / val_r[:6]) < VAL_TOL


def test_ctf_op_files_and_errors(ctx, tmp_path):
    """Distances.op(pyDir, PD) with the reference's default CTF = True: star + stack in, dist.npy + filtered stack out."""
    from manifoldem_esper_b200 import Read_Alignment
    g = golden('ctf_odd.npz')
    dm = tmp_path / '1_Embedding' / 'DM'
    dd = tmp_path / '0_Data_Inputs' / 'CTF5k15k_SNRpt1_ELS_2D'
    dm.mkdir(parents=True); dd.mkdir(parents=True)
    name = 'Hsp2D_5k15k_PD_005'
    mrc.write_stack(str(dd / (name + '.mrcs')), g['stack'])
    pa = g['params']
    Read_Alignment.write_star(str(dd / (name + '.star')), pa[:, 2], volt=pa[:, 3], Cs=pa[:, 1], ampc=pa[:, 4], pix=pa[:, 0], name=name)
    Distances.op(str(dm), '005', ctx=ctx)
    D = np.load(str(dm / 'Data_Distances' / 'PD_005_dist.npy'))
    assert D.dtype == np.float64 and offdiag_rel(D, g['D']) < D_TOL
    W = np.asarray(mrc.mmap(str(dd / (name + '_filtered.mrcs'))).data)
    assert W.shape == g['wiener'].shape and np.max(np.abs(W - g['wiener'])) < 2e-6 * np.abs(g['wiener']).max()
    with pytest.raises(ValueError):
        ctx.ctf_dist(g['stack'], pa[:-1])                                     # one parameter row per image
    with pytest.raises(ValueError):
        ctx.ctf_dist(g['stack'][:, :, :-1], pa)                               # square images only
    bad = pa.copy(); bad[3, 0] = 0.0
    with pytest.raises(ValueError):
        ctx.ctf_dist(g['stack'], bad)                                         # pixel size must be positive


# ---- PCA embedding ----------------------------------------------------------------
PCA_FLOOR = 2e-8      # FP32 pixels and 3xTF32 products: absolute accuracy of an eigenvalue ~1e-8 of the largest one


def pca_check(val, U, val_r, U_r, k, min_sep=3):
    """Eigenvalues 1e-6 relative (plus the absolute floor, which only matters for the rank-deficient noise-free
    stack whose trailing reference eigenvalues are rounding noise); separated eigenvectors 1e-4 up to sign."""
    assert np.all(np.abs(val - val_r[:k]) <= VAL_TOL * val_r[:k] + PCA_FLOOR * val_r[0])
    sep = 0
    for i in range(k):
        gap = min(val_r[i - 1] - val_r[i] if i else np.inf, val_r[i] - val_r[i + 1])
        if gap > 1e-2 * val_r[i] and gap > 1e-3 * val_r[0]:
            sep += 1
            assert vec_err(U[:, i], U_r[:, i]) < VEC_TOL * np.abs(U_r[:, i]).max(), i
    assert sep >= min_sep
    # scores are v s: column norms reproduce the eigenvalues, columns sum to zero (centred data)
    assert np.allclose(np.linalg.norm(U, axis=0) ** 2, val, rtol=1e-9)
    assert np.abs(U.sum(axis=0)).max() < 1e-8 * np.abs(U).max()
    # the significant leading block as a subspace
    ks = int(np.sum(val_r[:k] > 1e-4 * val_r[0]))
    while ks > 1 and val_r[ks - 1] - val_r[ks] < 1e-2 * val_r[ks - 1]:
        ks -= 1
    s_ = np.linalg.svd((U[:, :ks] / np.linalg.norm(U[:, :ks], axis=0)).T @ (U_r[:, :ks] / np.linalg.norm(U_r[:, :ks], axis=0)),
                       compute_uv=False)
    assert s_.min() > 1 - 1e-6


@pytest.mark.parametrize('name', ['pca_pristine', 'pca_small'])
def test_pca_vs_reference_golden(ctx, name):
    g = golden(name + '.npz')
    val, U, iters = PCA.embed(g['stack'], dim=20, ctx=ctx)
    pca_check(val, U, g['val'], g['vec'], 20)


def test_pca_medium_vs_oracle(ctx):
    """N=800 images of 96x96 at SNR 0.1: signal components above a dense noise bulk."""
    stack = synth.synthetic_pd(pd=4, box=96, n_side=20, tau=2, snr=0.1)
    val_r, U_r = O.pca_op(stack, dim=21)
    val, U, iters = PCA.embed(stack, dim=20, ctx=ctx)
    pca_check(val, U, val_r, U_r, 20, min_sep=2)


def test_pca_op_files_and_errors(ctx, tmp_path):
    g = golden('pca_pristine.npz')
    pdir = tmp_path / '1_Embedding' / 'PCA'
    ddir = tmp_path / '0_Data_Inputs' / '_Pristine_2D'
    pdir.mkdir(parents=True); ddir.mkdir(parents=True)
    mrc.write_stack(str(ddir / 'PD_005.mrcs'), g['stack'])
    PCA.op(str(pdir), '005')
    val = np.load(str(pdir / 'Data_Manifolds' / 'PD_005_val.npy'))
    vec = np.load(str(pdir / 'Data_Manifolds' / 'PD_005_vec.npy'))
    assert val.shape == (20,) and vec.shape == (64, 20) and vec.dtype == np.float64
    assert np.all(np.abs(val - g['val'][:20]) <= VAL_TOL * g['val'][:20] + PCA_FLOOR * g['val'][0])
    with pytest.raises(ValueError):
        PCA.embed(g['stack'][:10], dim=20, ctx=ctx)                           # rank of 10 centred images is 9
    with pytest.raises(ValueError):
        ctx.pca_embed(None, k=4, n=64, P=77)                                  # no resident stack of that shape


def test_eigensolver_modes_agree(ctx):
    """The barrier-free chain kernel with its on-device convergence test (mode 0, default), the resident barrier kernel
    (mode 3) and the streamed barrier kernel (mode 1) return the same eigenpairs; the barrier kernels run the same schedule."""
    g = golden('stage23_medium.npz')
    res = {}
    for mode in (3, 1, 0):
        ctx.dmap_config(mode)
        try:
            res[mode] = ctx.dmap_embed(g['D'], float(g['eps']), k=10, tol=1e-11)      # tight: the modes stop at different steps
        finally:
            ctx.dmap_config(0)
    assert ctx.stat('lanczos_fallbacks') == 0                                 # the chain kernel converged by itself
    v3, U3, it3 = res[3]
    for mode in (1, 0):
        v, U, it = res[mode]
        assert np.max(np.abs(v - v3) / v3) < 1e-12, mode
        for i in range(10):
            assert vec_err(U[:, i], U3[:, i]) < 1e-9 or i >= 7, (mode, i)       # bulk-edge vectors: subspace only
    assert abs(res[0][2] - it3) <= 40                                         # Krylov dimension the chain kernel's checker accepted
    with pytest.raises(ValueError):
        ctx.dmap_config(7)


@pytest.mark.parametrize('m', [33, 64, 129, 257, 400])
def test_merged_chain_kernel_over_sizes(ctx, m):
    """The default eigensolver schedule (one persistent cluster kernel, one exchange per step) on leading blocks of the medium golden:
    partial clusters, few rows per CTA, rows in registers only -- same pairs as the resident barrier kernel, the step count of the
    result identical from call to call (fixed schedule of convergence tests), no fallback to the barrier kernels."""
    g = golden('stage23_medium.npz')
    D = np.ascontiguousarray(g['D'][:m, :m])
    k = 6
    ctx.dmap_config(3)
    try:
        v3, U3, it3 = ctx.dmap_embed(D, float(g['eps']), k=k, tol=1e-10)
    finally:
        ctx.dmap_config(0)
    f0, m0, c0 = ctx.stat('lanczos_fallbacks'), ctx.stat('merged_launches'), ctx.stat('cluster_launches')
    runs = [ctx.dmap_embed(D, float(g['eps']), k=k, tol=1e-10) for _ in range(3)]
    assert ctx.stat('lanczos_fallbacks') == f0
    if ctx.stat('cluster_launches') > c0:                                      # clusters available on this device: the merged kernel ran
        assert ctx.stat('merged_launches') - m0 == 3
    v, U, it = runs[0]
    for v2, U2, it2 in runs[1:]:
        assert it2 == it and np.array_equal(v2, v) and np.array_equal(U2, U)    # bit-reproducible
    assert np.max(np.abs(v - v3) / v3) < 1e-11
    lam = np.sqrt(v3)
    for i in range(k):
        gap = min(abs(lam[i] - lam[i - 1]) if i else 1.0, abs(lam[i] - lam[i + 1]) if i + 1 < k else 1.0)
        if gap > 1e-3 * lam[1]:
            assert vec_err(U[:, i], U3[:, i]) < 1e-7, (m, i)
    assert abs(it - it3) <= 48


def test_eigensolver_on_duplicated_images(ctx):
    """Every image twice (stage1_dupes golden): Ms has rank n / 2, the Krylov space of the default kernel ends early (breakdown path:
    the workers stop by themselves, checker 0 accepts T_n as exact, vectors by inverse iteration) -- leading pairs as the oracle's SVD."""
    g = golden('stage1_dupes.npz')
    D = g['D']
    eps = float(np.median(D[D > 0]) ** 2 / 2.0)
    val_ref, U_ref = O.dmap_embed(D, eps)
    k = 6
    f0 = ctx.stat('lanczos_fallbacks')
    val, vec, iters = ctx.dmap_embed(D, eps, k=k)
    assert iters <= D.shape[0] // 2 + 8                                        # the rank (plus the few steps rounding noise survives)
    assert np.max(np.abs(val - val_ref[:k]) / val_ref[:k]) < VAL_TOL
    lam = np.sqrt(val_ref)
    for i in range(k):
        gap = min(abs(lam[i] - lam[i - 1]) if i else 1.0, abs(lam[i] - lam[i + 1]))
        if gap > 2e-3 * lam[1]:
            assert vec_err(vec[:, i], U_ref[:, i]) < VEC_TOL, i
    ctx.dmap_config(3)
    try:
        v3, U3, it3 = ctx.dmap_embed(D, eps, k=k)
    finally:
        ctx.dmap_config(0)
    assert np.max(np.abs(val - v3) / v3) < 1e-10


@pytest.mark.parametrize('case', ['eps/10', 'eps*10', 'eps*1e4', 'eps/1e3', 'duplicated', 'near-duplicate'])
def test_default_eigensolver_over_regimes(ctx, case):
    """The default schedule against the resident barrier kernel where the merged recurrence could suffer: nearly rank-one Ms (large
    bandwidth), nearly diagonal Ms (tiny bandwidth: every eigenvalue ~1, the Krylov space ends at once), every image twice (rank
    n / 2) and one near-duplicate pair -- same pairs (the default may fall back to the barrier kernels; it must not return others)."""
    g = golden('stage23_medium.npz')
    D, eps = g['D'], float(g['eps'])
    if case == 'duplicated':
        D = np.repeat(np.repeat(D[:150, :150], 2, axis=0), 2, axis=1)
    elif case == 'near-duplicate':
        D = D[:200, :200].copy(); D[1, :] = D[0, :] * (1 + 1e-9); D[:, 1] = D[1, :]; D[1, 1] = 0; D[0, 1] = D[1, 0] = 1e-9
    else:
        eps *= {'eps/10': 0.1, 'eps*10': 10.0, 'eps*1e4': 1e4, 'eps/1e3': 1e-3}[case]
    D = np.ascontiguousarray(D)
    ctx.dmap_config(3)
    try:
        v3, U3, it3 = ctx.dmap_embed(D, eps, k=8)
    finally:
        ctx.dmap_config(0)
    v0, U0, it0 = ctx.dmap_embed(D, eps, k=8)
    assert np.max(np.abs(v0 - v3) / np.maximum(v3, 1e-300)) < 1e-9, case
    lam = np.sqrt(v3)
    for i in range(4):
        gap = min(abs(lam[i] - lam[i - 1]) if i else 1.0, abs(lam[i] - lam[i + 1]))
        if gap > 1e-3 * lam[1]:
            assert vec_err(U0[:, i], U3[:, i]) < 1e-8, (case, i)


def test_eigensolver_without_speculative_batch(ctx):
    """esper_dmap_config(2): batches of 8 steps issued after each failed convergence test; same pairs, no more steps."""
    for name, k in (('stage23_medium', 10), ('stage123_pristine', 12)):
        g = golden(name + '.npz')
        ctx.dmap_config(1)
        v1, U1, it1 = ctx.dmap_embed(g['D'], float(g['eps']), k=k, tol=1e-11)
        ctx.dmap_config(2)
        try:
            v2, U2, it2 = ctx.dmap_embed(g['D'], float(g['eps']), k=k, tol=1e-11)
        finally:
            ctx.dmap_config(0)
        assert it2 <= it1 and np.max(np.abs(v1 - v2) / v1) < 1e-10
        for i in range(k):
            assert vec_err(U1[:, i], U2[:, i]) < 1e-7 or i >= 7       # bulk-edge vectors: subspace only


def test_embed_argument_errors(ctx):
    D = golden('stage123_small.npz')['D']
    with pytest.raises(ValueError):
        ctx.dmap_embed(D, -1.0, k=4)
    with pytest.raises(ValueError):
        ctx.dmap_embed(D, 0.3, k=D.shape[0] + 1)
    val, vec, _ = ctx.dmap_embed(D, 0.3, k=1)
    assert val[0] == 1.0 and vec.shape == (D.shape[0], 1)


# ---- stage 4 ----------------------------------------------------------------------------------
def test_histograms_bit_exact_vs_numpy(ctx):
    fx = golden('rot_inputs.npz')
    rng = np.random.default_rng(3)
    for key, dim in [('PD001', 6), ('PD064', 8)]:
        Ui = O.initial_subspace(fx[key], dim, PCA=True)
        T = dim * (dim - 1) // 2
        Rs, v1s, v2s, Href = [], [], [], []
        for e in range(24):
            th = np.zeros((T, 1)); th[rng.integers(0, T, 3)] = rng.uniform(-0.7, 0.7, (3, 1))
            R = O.gen_nd_rotations(dim, th)
            v1, v2 = sorted(rng.choice(dim, 2, replace=False))
            H, _ = O.rot_hist_eval(Ui, R, v1, v2)
            Rs.append(R); v1s.append(v1); v2s.append(v2); Href.append(H.astype(np.int32))
        nz, H = ctx.rot_hist_eval(Ui, np.stack(Rs), [0] * 24, v1s, v2s, want_hist=True)
        assert np.array_equal(H, np.stack(Href))
        assert np.array_equal(nz, [np.count_nonzero(h) for h in Href])
        assert np.all(H.sum(axis=(1, 2)) <= Ui.shape[0])


def test_histogram_edge_rule(ctx):
    U = np.array([[-1.5, 1.5], [1.5, -1.5], [0.0, 0.0], [1.6, 0.0], [np.nextafter(1.5, 2), 0.0], [np.nan, 0.0]])
    nz, H = ctx.rot_hist_eval(U, np.eye(2)[None], [0], [0], [1], want_hist=True)
    Hr, _ = O.rot_hist_eval(U[:5], np.eye(2), 0, 1)
    assert np.array_equal(H[0], Hr.astype(np.int32)) and nz[0] == 3


@pytest.mark.parametrize('dim', [6, 8])
def test_rotation_sweep_vs_reference_run(ctx, dim):
    """RotMatrices for the 12 committed fixtures equal the reference's output bit for bit; per-evaluation
    non-zero-bin counts equal the ones spied from the reference run."""
    fx, ref, cnt = golden('rot_inputs.npz'), golden('rot_ref_dim%d.npz' % dim), golden('rot_counts_dim8.npz')
    keys = list(fx.keys())
    preps = [RH.prepare(fx[k], dim=dim, PCA=True, CTF=True) for k in keys]
    thetas, counts = RH.sweep_many(preps, dim, ctx=ctx, want_counts=True)
    for i, k in enumerate(keys):
        assert np.array_equal(thetas[i], ref['RotMatrices_' + k[2:]]), k
        assert np.array_equal(preps[i]['saveEigs'], ref['RotEigenfunctions_' + k[2:]])
        if dim == 8 and k in cnt:
            assert np.array_equal(counts[i, :len(preps[i]['comboList'])].reshape(-1), cnt[k])


@pytest.mark.parametrize('dim', [6, 8])
def test_device_prepare_and_sweep_vs_reference_run(ctx, dim):
    """normalisation + all conic fits on the device (esper_manifold_prepare), partner selection on the
    host, sweep on the manifolds left resident: outputs equal the reference run's files bit for bit."""
    fx, ref = golden('rot_inputs.npz'), golden('rot_ref_dim%d.npz' % dim)
    keys = list(fx.keys())
    preps, thetas = RH.run_many([fx[k] for k in keys], dim=dim, PCA=True, CTF=True, ctx=ctx)
    assert preps[0]['resident'] is not None                 # no retry: the sweep used the device-resident manifolds
    for i, k in enumerate(keys):
        host = RH.prepare(fx[k], dim=dim, PCA=True, CTF=True)
        assert np.array_equal(preps[i]['U_init'], host['U_init']), k          # normalisation is bit-identical
        assert preps[i]['comboList'] == host['comboList'] and preps[i]['Rij_final'] == host['Rij_final']
        assert np.array_equal(thetas[i], ref['RotMatrices_' + k[2:]]), k
        assert np.array_equal(preps[i]['saveEigs'], ref['RotEigenfunctions_' + k[2:]])


@pytest.mark.parametrize('kind', [2, 3])
def test_conic_fits_vs_host_rref(ctx, kind):
    """R^2 and coefficients of every pair equal the oracle's (the reference's sympy `rref` solve) to rounding."""
    fx = golden('rot_inputs.npz')
    keys = list(fx.keys())[:3]
    dim = 7
    Ui, R2, coef = ctx.manifold_prepare(np.stack([fx[k][:, :8] for k in keys]), dim, first=1, kind=kind, want_coef=True)
    for i, k in enumerate(keys):
        U = O.normalize_manifold(fx[k][:, :8])[:, 1:dim + 1]
        assert np.array_equal(Ui[i], U)
        q = 0
        for v1 in range(dim - 1):
            for v2 in range(v1 + 1, dim):
                if kind == 2:
                    cF1, r2 = O.conic_fit2(U, v1, v2)
                    cF = cF1[1:]
                else:
                    cF, r2 = O.conic_fit3(U, v1, v2)
                assert abs(R2[i, q] - r2) < 1e-11, (k, v1, v2)
                assert np.allclose(coef[i, q, :len(cF)], cF, rtol=1e-8, atol=1e-10), (k, v1, v2)
                q += 1
    # the fits alone on caller-supplied manifolds, and on the resident copy
    R2b = ctx.conic_fit_batch(Ui, kind=kind)
    assert np.array_equal(R2b, R2)
    assert np.array_equal(ctx.conic_fit_batch(None, kind=kind, shape=Ui.shape), R2)


def test_device_prepare_retry_path(ctx):
    """A manifold without any parabolic pair (R^2 < 0.2 in row 0) takes the 45-degree pre-rotation (:202-209)."""
    rng = np.random.default_rng(5)
    blob = rng.standard_normal((600, 8))
    good = golden('rot_inputs.npz')['PD001'][:600, :8]
    preps = RH.prepare_many([blob, good], dim=6, PCA=False, CTF=True, ctx=ctx)
    for p, U0 in zip(preps, (blob, good)):
        host = RH.prepare(U0, dim=6, PCA=False, CTF=True)
        assert p['itIdx'] == host['itIdx'] and p['comboList'] == host['comboList'] and p['Rij_final'] == host['Rij_final']
        assert np.array_equal(p['U_init'], host['U_init'])
    assert preps[0]['itIdx'] == 2 and preps[0]['resident'] is None
    th = RH.sweep_many(preps, 6, ctx=ctx)
    th_host = RH.sweep_many([RH.prepare(U0, dim=6, PCA=False, CTF=True) for U0 in (blob, good)], 6, ctx=ctx)
    assert all(np.array_equal(a, b) for a, b in zip(th, th_host))


def test_manifold_prepare_argument_errors(ctx):
    U = np.zeros((1, 10, 6))
    with pytest.raises(ValueError):
        ctx.manifold_prepare(U, 6, first=1)                      # first + dim > ncol
    with pytest.raises(ValueError):
        ctx.manifold_prepare(U, 4, first=0, kind=7)
    with pytest.raises(ValueError):
        ctx.conic_fit_batch(None, kind=2, shape=(3, 10, 4))      # nothing resident of that shape
    with pytest.raises(ValueError):
        ctx.rot_sweep(None, np.array([[[0, 1]]]), [1], [[0]], RH.theta_grid(), shape=(2, 10, 4))


def test_rotation_op_files(ctx, tmp_path):
    fx, ref = golden('rot_inputs.npz'), golden('rot_ref_dim6.npz')
    root = tmp_path / 'tree'
    os.makedirs(root / '2_Manifold_Rotations'); os.makedirs(root / '1_Embedding' / 'PCA' / 'Data_Manifold')
    np.save(str(root / '1_Embedding' / 'PCA' / 'Data_Manifold' / 'PD_017_vec.npy'), fx['PD017'])
    RH.op(str(root / '2_Manifold_Rotations'), '017', dim=6, PCA=True, ctx=ctx)
    out = root / '2_Manifold_Rotations' / 'Data_Rotations' / 'PD017'
    assert np.array_equal(np.load(str(out / 'PD017_RotMatrices.npy')), ref['RotMatrices_017'])
    assert np.array_equal(np.load(str(out / 'PD017_RotEigenfunctions.npy')), ref['RotEigenfunctions_017'])
    assert np.array_equal(np.load(str(out / 'PD017_RotParameters.npy')), np.array([6]))


def test_rotation_sweep_argument_errors(ctx):
    U = np.zeros((1, 10, 6))
    with pytest.raises(ValueError):
        ctx.rot_sweep(U, np.array([[[0, 9]]]), [1], [[0]], RH.theta_grid())   # v2 out of range
    with pytest.raises(ValueError):
        ctx.rot_sweep(U, np.array([[[0, 1]]]), [1], [[15]], RH.theta_grid())  # Rij out of range


# ---- data preparation and formats --------------------------------------------------
def test_snr_params_vs_reference_golden(ctx):
    """AddNoise.find_SNR per state: the reference computes it in float32, the device in float64."""
    g = golden('tau_snr.npz')
    _, params = ctx.tau_snr_stack(g['pristine'], 2, float(g['snr']), seed=1, download=False)
    assert np.max(np.abs(params - g['params']) / np.abs(g['params'])) < 1e-5
    sm, ns = AddNoise.find_SNR(g['pristine'][3], float(g['snr']), ctx=ctx)
    assert sm == params[3, 0] and ns == params[3, 1]


@pytest.mark.parametrize('box,ns,tau', [(24, 6, 10), (25, 4, 3), (250, 3, 2)])
def test_tau_snr_stack_vs_oracle_on_the_same_noise_field(ctx, box, ns, tau):
    """The device stack equals the oracle's AddNoise.op applied with the SAME N(0,1) field (the numpy restatement of the
    device's Philox + Box-Muller stream); every image is standardised on its background ring."""
    pr = golden('tau_snr.npz')['pristine'] if box == 24 else synth.pristine_states(box=box, n_side=2, pd=4)[:ns]
    seed = 0x1234567890ABCDEF + box
    stack, params = ctx.tau_snr_stack(pr, tau, 0.1, seed=seed)
    gauss = np.stack([synth.philox_normal(seed, k, box * box).reshape(box, box) for k in range(pr.shape[0] * tau)])
    ref = O.tau_snr_stack(pr, tau, 0.1, gauss=gauss)
    assert stack.shape == ref.shape and stack.dtype == np.float32
    assert np.max(np.abs(stack - ref)) < 2e-5
    bg = ~O.background_mask(box, box)
    for img in stack:
        assert abs(img[bg].astype(np.float64).mean()) < 1e-5 and abs(img[bg].astype(np.float64).std() - 1) < 1e-5
    again, _ = ctx.tau_snr_stack(pr, tau, 0.1, seed=seed)
    other, _ = ctx.tau_snr_stack(pr, tau, 0.1, seed=seed + 1)
    assert np.array_equal(again, stack) and not np.array_equal(other, stack)
    # copies of one state carry independent noise
    a, b = stack[0].astype(np.float64).ravel(), stack[1].astype(np.float64).ravel()
    assert abs(np.corrcoef(a - a.mean(), b - b.mean())[0, 1]) < 0.5


def test_tau_snr_resident_chain_and_duplicates(ctx):
    pr = synth.pristine_states(box=32, n_side=4, pd=9)
    stack, _ = ctx.tau_snr_stack(pr, 3, 0.5, seed=77)
    D_res = ctx.dist_rms(None, n=48, P=1024)                      # embeds the stack where it was generated
    assert np.array_equal(ctx.download_images(48, 1024), stack.reshape(48, 1024))
    assert np.array_equal(D_res, ctx.dist_rms(stack))
    dup, params = ctx.tau_snr_stack(pr, 3, None)
    assert np.array_equal(dup, np.repeat(pr, 3, axis=0)) and not params.any()
    odd = synth.pristine_states(box=25, n_side=2, pd=1)            # P not a multiple of 4
    assert np.array_equal(ctx.tau_snr_stack(odd, 2, 0)[0], np.repeat(odd, 2, axis=0))
    with pytest.raises(ValueError):
        ctx.tau_snr_stack(pr[:, :, :16], 2, 0.1)
    with pytest.raises(ValueError):
        ctx.tau_snr_stack(pr, 0, 0.1)


def test_generate_tau_snr_op_files(ctx, tmp_path):
    g = golden('tau_snr.npz')
    root = tmp_path / '0_Data_Inputs'
    os.makedirs(root / 'Pristine_2D'); os.makedirs(root / 'Pristine_AddTauSNR')
    mrc.write_stack(str(root / 'Pristine_2D' / 'PD_001.mrcs'), g['pristine'])
    Generate_TauSNR.op(str(root / 'Pristine_AddTauSNR'), '001', seed=5, ctx=ctx)
    out = mrc.mmap(str(root / 'Pristine_AddTauSNR' / 'Modified_Stacks' / 'PD001_SNR_tau10_stack.mrcs'))
    assert out.data.shape == g['stack'].shape
    ref, _ = ctx.tau_snr_stack(g['pristine'], 10, 0.1, seed=5)
    assert np.array_equal(np.asarray(out.data), ref)
    # same statistics as the reference run's stack (different noise stream): per-image spread of the whole image
    assert abs(np.asarray(out.data).std() / g['stack'].std() - 1) < 0.05
    out.close()
    Generate_TauSNR.op(str(root / 'Pristine_AddTauSNR'), '001', apply_noise=False, tau=2, ctx=ctx)
    dup = mrc.mmap(str(root / 'Pristine_AddTauSNR' / 'Modified_Stacks' / 'PD001_tau2_stack.mrcs'))
    assert np.array_equal(np.asarray(dup.data), np.repeat(g['pristine'], 2, axis=0))
    dup.close()


def test_upload_mrc_streams_the_stack(ctx, tmp_path):
    rng = np.random.default_rng(11)
    stack = rng.standard_normal((300, 250, 250)).astype(np.float32)          # 75 MB: three staging chunks
    path = str(tmp_path / 'PD_009.mrcs')
    mrc.write_stack(path, stack)
    assert ctx.upload_mrc(path) == (300, 250, 250)
    assert np.array_equal(ctx.download_images(300, 62500), stack.reshape(300, -1))
    assert ctx.upload_mrc(path, first=290, count=50) == (10, 250, 250)        # clipped at the end of the file
    assert np.array_equal(ctx.download_images(10, 62500), stack[290:].reshape(10, -1))
    D = ctx.dist_rms(None, n=10, P=62500)
    assert np.array_equal(D, ctx.dist_rms(stack[290:]))
    # extended header (nsymbt) and the error paths
    raw = bytearray(open(path, 'rb').read(1024))
    small = stack[:4, :8, :8].copy()
    hdr = np.frombuffer(bytes(raw), dtype='<i4').copy()
    hdr[0:3] = (8, 8, 4); hdr[23] = 80
    p2 = str(tmp_path / 'ext.mrcs')
    with open(p2, 'wb') as f:
        f.write(hdr.tobytes()); f.write(b'x' * 80); f.write(small.tobytes())
    assert ctx.upload_mrc(p2) == (4, 8, 8)
    assert np.array_equal(ctx.download_images(4, 64), small.reshape(4, 64))
    hdr[3] = 1
    p3 = str(tmp_path / 'mode1.mrcs')
    with open(p3, 'wb') as f:
        f.write(hdr.tobytes()); f.write(b'x' * 80); f.write(small.tobytes())
    with pytest.raises(ValueError):
        ctx.upload_mrc(p3)
    with pytest.raises(ValueError):
        ctx.upload_mrc(str(tmp_path / 'missing.mrcs'))
    with pytest.raises(ValueError):
        ctx.upload_mrc(path, first=300)


def test_bin_accumulate_bit_exact(ctx):
    """Manifold_Binning.py:940-974: float64 bin sums in the caller's order equal the reference loop bit for bit."""
    rng = np.random.default_rng(2)
    stack = (rng.standard_normal((400, 33, 33)) * np.exp(rng.uniform(-8, 8, (400, 1, 1)))).astype(np.float32)
    order = rng.permutation(400)[:371].astype(np.int32)
    cuts = np.sort(rng.choice(372, 19, replace=False))
    offsets = np.concatenate([[0], cuts[:7], [cuts[7]], cuts[7:], [371]]).astype(np.int32)     # 21 bins, one of them empty
    ref = O.bin_accumulate(stack, order, offsets)
    sums = ctx.bin_accumulate(stack, order, offsets)
    assert np.array_equal(sums, ref.reshape(ref.shape[0], -1))
    assert np.array_equal(ctx.bin_accumulate(None, order, offsets, n=400, P=33 * 33), sums)    # resident stack
    with pytest.raises(ValueError):
        ctx.bin_accumulate(stack, [400], [0, 1])
    with pytest.raises(ValueError):
        ctx.bin_accumulate(stack, [1, 2], [0, 2, 1])


def test_whole_chain_through_files_with_the_launcher(ctx, tmp_path):
    """Generate_TauSNR.op -> launcher (Distances.op non-CTF, DiffusionMaps.op, Rotation_Histograms.op) on a tree laid out
    like the reference's, every stage reading the previous stage's files; each output checked against the oracle applied to
    the same input file."""
    from manifoldem_esper_b200 import launcher
    root = tmp_path / 'tree'
    dm = root / '1_Embedding' / 'DM'
    rot = root / '2_Manifold_Rotations'
    prep = root / '0_Data_Inputs' / 'Pristine_AddTauSNR'
    data = root / '0_Data_Inputs' / 'CTF5k15k_SNRpt1_ELS_2D'          # the data directory Distances.op reads (:50)
    for d in (dm, rot, prep, root / '0_Data_Inputs' / 'Pristine_2D', data):
        os.makedirs(d)
    pds = ['001', '002']
    for pd in pds:
        mrc.write_stack(str(root / '0_Data_Inputs' / 'Pristine_2D' / ('PD_%s.mrcs' % pd)),
                        synth.pristine_states(box=32, n_side=8, pd=int(pd)))
        Generate_TauSNR.op(str(prep), pd, SNR=2.0, tau=5, seed=int(pd), ctx=ctx)
        os.replace(str(prep / 'Modified_Stacks' / ('PD%s_SNR_tau5_stack.mrcs' % pd)), str(data / ('PD%s_SNR_tau5_stack.mrcs' % pd)))
    work = launcher._stage_fns(str(root), {'dist', 'dm', 'rot'}, 4, False, 8, False, ctx.device, ctf=False)
    recs = launcher.run_sharded(pds, work, 0, 1)
    for pd, rec in zip(pds, recs):
        stack = np.asarray(mrc.mmap(str(data / ('PD%s_SNR_tau5_stack.mrcs' % pd))).data)
        assert stack.shape == (320, 32, 32)
        D = np.load(str(dm / 'Data_Distances' / ('PD_%s_dist.npy' % pd)))
        assert d_err(D, O.distances_gram_f64(O.standardize_stack(stack))) < D_TOL
        val = np.load(str(dm / 'Data_Manifolds' / ('PD_%s_val.npy' % pd)))
        vec = np.load(str(dm / 'Data_Manifolds' / ('PD_%s_vec.npy' % pd)))
        assert val.shape == (8,) and vec.shape == (320, 8) and val[0] == 1.0 and np.all(np.diff(val) <= 0)
        try:
            ref = O.rotation_histograms_op(vec, dim=4, PCA=False, CTF=True)
        except IndexError:                      # no mixed pair of parabolic subspaces: the reference fails too (:266)
            assert not rec['ok'] and 'IndexError' in rec['error']
            continue
        assert rec['ok'], rec
        out = rot / 'Data_Rotations' / ('PD%s' % pd)
        assert np.array_equal(np.load(str(out / ('PD%s_RotMatrices.npy' % pd))), ref['RotMatrices'])
        assert np.array_equal(np.load(str(out / ('PD%s_RotEigenfunctions.npy' % pd))), ref['RotEigenfunctions'])
        # the consumer's side of the contract: the files load the way 3_Manifold_Binning/Manifold_Binning.py:88-139 loads them
        from consumer_contract import load_like_manifold_binning
        c = load_like_manifold_binning(str(root), pd)
        assert c['dim'] == 4 and c['rotMatrix'].shape[1] == 6
    # the same tree through the fused stage (Distances.op + DiffusionMaps.op in one library call): same files
    import shutil
    D1 = np.load(str(dm / 'Data_Distances' / 'PD_001_dist.npy')); vec1 = np.load(str(dm / 'Data_Manifolds' / 'PD_001_vec.npy'))
    shutil.rmtree(str(dm / 'Data_Distances')); shutil.rmtree(str(dm / 'Data_Manifolds'))
    work = launcher._stage_fns(str(root), {'embed'}, 4, False, 8, False, ctx.device, ctf=False)
    recs = launcher.run_sharded(pds, work, 0, 1)
    assert all(r['ok'] for r in recs), recs
    assert np.array_equal(np.load(str(dm / 'Data_Distances' / 'PD_001_dist.npy')), D1)
    vec2 = np.load(str(dm / 'Data_Manifolds' / 'PD_001_vec.npy'))
    assert vec2.shape == vec1.shape and all(vec_err(vec1[:, i], vec2[:, i]) < 1e-3 for i in range(3))   # own random fit starts: eps differs ~1e-6
