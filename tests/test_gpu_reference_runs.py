"""GPU parity against reference runs at the BASELINE sizes and on the whole shipped fixture set (all through the C ABI).

The goldens come from tools/gen_goldens.py --only-big (the reference's own code executed at N=400 / 250x250 and
N=2000 / 250x250) and from the reference's Rotation_Histograms.op on all 126 shipped Data_Manifolds_126 fixtures."""
import os

import numpy as np
import pytest

from conftest import golden
from manifoldem_esper_b200 import DiffusionMaps, GaussianBandwidth, Rotation_Histograms as RH, _capi, synth
from oracle import esper_oracle as O

pytestmark = pytest.mark.gpu

D_TOL, D_FLOOR, VAL_TOL, VEC_TOL = 1e-5, 1e-2, 1e-6, 1e-4
LOG_EPS = np.arange(-100, 100.2, 0.2)


def d_err(D, Dr):
    return float(np.max(np.abs(D - Dr) / np.maximum(Dr, D_FLOOR)))


def vec_err(u, v):
    return min(np.abs(u - v).max(), np.abs(u + v).max())


def synth_stack(g):
    pd, box, n_side, tau, snr = g['synth']
    return synth.synthetic_pd(pd=int(pd), box=int(box), n_side=int(n_side), tau=int(tau), snr=None if snr < 0 else float(snr))


def check_vectors(val_r, U_r, val, vec, k):
    """Separated eigenvectors one by one (1e-4 up to sign), the rest as a subspace (SURVEY hard part 4)."""
    lam = np.sqrt(val_r[:k + 1])
    sep = []
    for i in range(k):
        gap = min(abs(lam[i] - lam[i - 1]) if i else 1.0, abs(lam[i] - lam[i + 1]))
        if gap > 2e-3 * lam[1]:
            assert vec_err(vec[:, i], U_r[:, i]) < VEC_TOL, i
            sep.append(i)
    s_ = np.linalg.svd(vec.T @ U_r[:, :k], compute_uv=False)
    assert s_.min() > 1 - 1e-6 or lam[k - 1] - lam[k] < 2e-3 * lam[1]
    return sep


def test_config1_standin_whole_chain_vs_reference_run(ctx):
    """Pristine N=400 at the full 250x250 box (the stand-in for the reference's own CPU-runnable inputs): Distances.op,
    GaussianBandwidth.op and DiffusionMaps.op of the reference against the chained GPU pipeline at the north-star tolerances."""
    g = golden('stage1_p2.npz')
    stack = synth_stack(g)
    D = ctx.dist_rms(stack)
    assert d_err(D, g['D']) < 1e-11                          # <= 512 images: every pair in the reference's own arithmetic
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    L = GaussianBandwidth.sweep(None, LOG_EPS, ctx=ctx, m=D.shape[0])
    assert np.max(np.abs(L - g['logSumWij'])) < 1e-9
    np.random.seed(int(g['seed']))
    res = DiffusionMaps.embed(D, k=16, ctx=ctx, verbose=False)
    assert abs(res['eps'] - float(g['eps'])) / float(g['eps']) < 1e-7
    k = 16
    val_g, vec_g = g['val'], g['vec']
    assert np.max(np.abs(res['val'] - val_g[:k]) / val_g[:k]) < VAL_TOL       # eps agrees to 1e-7: well inside the tolerance
    sep = check_vectors(val_g, vec_g, res['val'], res['vec'], k)
    assert len(sep) >= 3


def test_config2_distances_vs_reference_run_rows(ctx):
    """BASELINE config 2 (N=2000, 250x250, SNR 0.1): every 25th row of the matrix the reference's pair loop produced."""
    g = golden('stage1_c2_rows.npz')
    stack = synth_stack(g)
    D = ctx.dist_rms(stack)
    step = int(g['step'])
    assert d_err(D[::step], g['D_rows']) < D_TOL
    assert ctx.gram_info() == 1                              # the FP16-plane tensor-core contraction produced it
    off = D[::step] > 0
    d2, d2r = (D[::step] ** 2)[off], (g['D_rows'] ** 2)[off]
    assert abs(np.mean((d2 - d2r) / d2r)) < 1e-7             # no systematic accumulation bias against the reference's float64 sums


def test_config2_stage23_vs_reference_run(ctx):
    """GaussianBandwidth.op + DiffusionMaps.op of the reference at N=2000 on the float64-Gram distance matrix of the config-2
    stack: sweep 1e-9, same-seed epsilon 1e-7, eigenvalues 1e-6 and eigenvectors 1e-4 at the reference's epsilon."""
    g = golden('stage23_c2.npz')
    stack = synth_stack(g)
    D = O.distances_gram_f64(O.standardize_stack(stack))
    assert np.max(np.abs(D[::100] - g['D_rows'])) < 1e-12     # the matrix the reference run was given
    L = GaussianBandwidth.sweep(D, LOG_EPS, ctx=ctx)
    assert ctx.sweep_info()[0] == 1                          # moment sweep
    assert np.max(np.abs(L - g['logSumWij'])) < 1e-9
    np.random.seed(int(g['seed']))
    a0 = 1 * (np.random.rand(4, 1) - .5)
    popt, _, _, R2 = GaussianBandwidth.op(None, LOG_EPS, a0, ctx=ctx, m=2000)
    eps = np.exp(-(popt[1] / popt[0]))
    assert abs(eps - float(g['eps'])) / float(g['eps']) < 1e-7 and abs(R2 - float(g['R2'])) < 1e-9
    k = 16
    for mode in (0, 1):                                      # barrier-free chain kernel and the streamed barrier kernel
        ctx.dmap_config(mode)
        try:
            val, vec, iters = ctx.dmap_embed(None, float(g['eps']), k=k, m=2000)
        finally:
            ctx.dmap_config(0)
        assert np.max(np.abs(val - g['val'][:k]) / g['val'][:k]) < VAL_TOL, mode
        lam = np.sqrt(g['val'])
        for i in range(7):
            gap = min(abs(lam[i] - lam[i - 1]) if i else 1.0, abs(lam[i] - lam[i + 1]))
            if gap > 2e-3 * lam[1]:
                assert vec_err(vec[:, i], g['vec'][:, i]) < VEC_TOL, (mode, i)
        s_ = np.linalg.svd(vec[:, :7].T @ g['vec'][:, :7], compute_uv=False)
        assert s_.min() > 1 - 1e-6, mode                      # the signal eigenvectors as a subspace


def test_pd_embed_fused_equals_the_staged_calls(ctx):
    """esper_pd_embed_fused = esper_dist_rms -> esper_eps_sweep -> esper_tanh_fit -> esper_dmap_embed, bit for bit."""
    g = golden('stage123_small.npz')
    stack = g['stack']
    n = stack.shape[0]
    a0 = np.array([-0.2, 0.3, -0.4, 0.1])
    res = ctx.pd_embed_fused(stack, LOG_EPS, a0, k=8, want_D=True)
    D = ctx.dist_rms(stack)
    L = ctx.eps_sweep(None, 1. / (2. * np.exp(LOG_EPS)), 10.0, m=n)
    popt, resnorm, R2, info, _ = _capi.tanh_fit(LOG_EPS, L, a0)
    eps = float(np.exp(-(popt[1] / popt[0])))
    val, vec, iters = ctx.dmap_embed(None, eps, k=8, m=n)
    assert np.array_equal(res['D'], D) and np.array_equal(res['logSumWij'], L) and np.array_equal(res['popt'], popt)
    assert res['eps'] == eps and res['iters'] == iters and np.array_equal(res['val'], val) and np.array_equal(res['vec'], vec)
    assert abs(res['R2'] - R2) < 1e-15 and resnorm <= 100
    # against the reference run of the same stack (its own random start: epsilon agrees to the fit's reproducibility)
    assert d_err(res['D'], g['D']) < D_TOL and np.max(np.abs(res['logSumWij'] - g['logSumWij'])) < 1e-9
    assert abs(res['eps'] - float(g['eps'])) / float(g['eps']) < 1e-4
    # the resident stack, no distance matrix requested
    res2 = ctx.pd_embed_fused(None, LOG_EPS, a0, k=8, n=n, P=stack.shape[1] * stack.shape[2])
    assert 'D' not in res2 and np.array_equal(res2['val'], val) and np.array_equal(res2['vec'], vec)
    with pytest.raises(ValueError):
        ctx.pd_embed_fused(None, LOG_EPS, a0, k=8, n=7, P=9)
    with pytest.raises(ValueError):
        ctx.pd_embed_fused(stack, LOG_EPS, a0[:3], k=8)


def test_pd_embed_fused_retry_when_the_first_fit_is_unusable(ctx):
    """A start vector whose fit fails the reference's `resnorm > 100` test: the wrapper continues the retry loop of
    GaussianBandwidth.py:47-57 on the resident matrix."""
    g = golden('stage123_small.npz')
    a_bad = np.array([0.0, 0.0, 0.0, 0.0])                  # zero slope: singular Jacobian -> no covariance -> resnorm = inf
    np.random.seed(3)
    res = ctx.pd_embed_fused(g['stack'], LOG_EPS, a_bad, k=6)
    assert res['resnorm'] <= 100 and abs(res['eps'] - float(g['eps'])) / float(g['eps']) < 1e-4
    assert abs(res['val'][0] - 1.0) < 1e-12 and res['vec'].shape == (g['stack'].shape[0], 6)


def test_tf32_operand_planes_stay_available(ctx):
    """esper_gram_config(0): the 3xTF32 planes (round-1 default) against the reference goldens and the full-size oracle."""
    ctx.gram_config(0)
    try:
        X = np.random.default_rng(77).standard_normal((700, 3000)).astype(np.float32)
        D = ctx.dist_rms(X)
        assert ctx.gram_info() == 0
        assert d_err(D, O.distances_gram_f64(O.standardize_stack(X.reshape(700, 1, 3000)))) < D_TOL
    finally:
        ctx.gram_config(1)
    D1 = ctx.dist_rms(X)
    assert ctx.gram_info() == 1 and d_err(D1, D) < 2e-6


@pytest.mark.parametrize('dim', [6, 8])
def test_all_126_fixtures_vs_reference_run(ctx, dim):
    """P5: RotMatrices / RotEigenfunctions of ALL 126 shipped Data_Manifolds_126 fixtures equal the files the reference's
    Rotation_Histograms.op wrote, bit for bit (device normalisation + conic fits + sweep; partner selection on the host)."""
    fx, ref = golden('rot_inputs_all.npz'), golden('rot_ref_dim%d.npz' % dim)
    keys = sorted(fx.keys())
    assert len(keys) == 126
    preps, thetas = RH.run_many([fx[k] for k in keys], dim=dim, PCA=True, CTF=True, ctx=ctx)
    bad = [k for i, k in enumerate(keys)
           if not (np.array_equal(thetas[i], ref['RotMatrices_' + k[2:]]) and np.array_equal(preps[i]['saveEigs'], ref['RotEigenfunctions_' + k[2:]]))]
    assert not bad, bad


def test_rowblock_cuda_backend_single_rank(ctx):
    """The row-block (oversized PD) entry points on one GPU, world = 1: esper_dist_rms_rows, esper_eps_sweep_rows,
    esper_degree_rows, esper_markov_rows, esper_symv_rows_d, esper_lanczos_start_d / ortho_d, esper_ritz_vectors_d with the NCCL
    all-gather, then the fused SYMV + packet gather (esper_comm_*, esper_symv_allgather_d, esper_comm_wait_d), twice on one
    communicator: the same eigenpairs as esper_dmap_embed on the full matrix."""
    import torch
    import torch.distributed as dist
    from manifoldem_esper_b200 import rowblock as RB
    if not dist.is_initialized():
        dist.init_process_group('nccl', init_method='tcp://127.0.0.1:29617', rank=0, world_size=1, device_id=torch.device('cuda', 0))
    try:
        stack = synth.synthetic_pd(pd=3, box=64, n_side=12, tau=5, snr=0.5)          # N=720, P=4096
        n, P = stack.shape[0], stack.shape[1] * stack.shape[2]
        X = np.ascontiguousarray(stack.reshape(n, P))
        ctx.dist_config(split_k=4)                           # the same K chunks in the full matrix and in the row blocks: same bits
        D = ctx.dist_rms(X)
        eps = 0.35
        val_r, vec_r, _ = ctx.dmap_embed(None, eps, k=8, m=n, tol=1e-11)
        dev = torch.device('cuda', 0)
        outs = []
        for fused in (False, True):
            be = RB.CudaBackend(ctx, X, n, P, dev, fused=fused)
            comm = RB.TorchComm(dev, stream=be.stream)
            for rep in range(2 if fused else 1):                                     # a second embedding on the same communicator
                out = RB.embed_oversized(be, comm, n, k=8, eps=eps, tol=1e-11)
                outs.append(out)
                assert np.max(np.abs(out['val'] - val_r) / val_r) < 1e-10
                for i in range(5):
                    assert vec_err(out['vec'][:, i], vec_r[:, i]) < 1e-8, (fused, i)
            be.comm_close()
        assert np.array_equal(outs[0]['val'], outs[1]['val']) and np.array_equal(outs[0]['vec'], outs[1]['vec'])   # fused == NCCL, bit for bit
        assert np.array_equal(outs[1]['vec'], outs[2]['vec'])
        # the automatic bandwidth selection over the block sums (moment tables and the general sweep agree)
        be = RB.CudaBackend(ctx, X, n, P, dev, fused=False)
        comm = RB.TorchComm(dev, stream=be.stream)
        o1 = RB.embed_oversized(be, comm, n, k=4, moments=True)
        o2 = RB.embed_oversized(be, comm, n, k=4, moments=False)
        assert o1['sweep_path'] == 'moments' and o2['sweep_path'] == 'general'
        assert np.max(np.abs(o1['logSumWij'] - o2['logSumWij'])) < 1e-12 and abs(o1['eps'] - o2['eps']) / o2['eps'] < 1e-7
        np.random.seed(12345)
        L = GaussianBandwidth.sweep(D, LOG_EPS, ctx=ctx)
        assert np.max(np.abs(o2['logSumWij'] - L)) < 1e-12
    finally:
        ctx.dist_config(split_k=0)
        dist.destroy_process_group()


def test_short_row_block_must_be_the_last_one(ctx):
    """a block of nrows % 128 != 0 that does not end at row n would be written past its buffer: rejected."""
    X = np.random.default_rng(1).standard_normal((700, 512)).astype(np.float32)
    fresh = _capi.Context(0)
    try:
        with pytest.raises(ValueError):
            fresh.dist_rms_rows(X, 700, 512, 0, 100)
        Dr = fresh.dist_rms_rows(X, 700, 512, 640, 60, download=True)
        assert Dr.shape == (60, 700)
    finally:
        fresh.close()


def test_rowblock_two_ranks_vs_single_gpu():
    """P6 on two GPUs (skipped when fewer are visible): tools/rowblock_check.py under torchrun -- BASELINE config 2 split over two
    ranks by row blocks: every rank's distance rows are bit-equal to the single-GPU matrix, the eigenpairs agree with the
    single-GPU solve, and the fused SYMV + packet all-gather over NVLink peer memory equals the NCCL all-gather bit for bit."""
    import re
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29643', os.path.join(root, 'tools', 'rowblock_check.py'), 'small']
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, (res.stdout[-2000:], res.stderr[-2000:])
    out = res.stdout
    assert len(re.findall(r'bit-equal to single-GPU D: True', out)) == 2, out[-3000:]
    assert len(re.findall(r'fused == NCCL bit for bit at fixed eps: True', out)) == 2, out[-3000:]
    diffs = [float(x) for x in re.findall(r'val rel diff ([0-9.e+-]+)', out)]
    assert len(diffs) == 2 and max(diffs) < 1e-9, out[-3000:]


def test_full_spectrum_on_request(ctx):
    """k = m (the reference saves the full U of its SVD, DiffusionMaps.py:166-169): Lanczos to exhaustion + QL rotations applied on
    the device.  All m eigenpairs of the N=400 golden matrix: eigenvalues against the oracle's SVD, U orthonormal, residuals, the
    leading columns against the reference run."""
    g = golden('stage23_medium.npz')
    D, eps = g['D'], float(g['eps'])
    m = D.shape[0]
    val_r, U_r = O.dmap_embed(D, eps)
    val, vec, iters = ctx.dmap_embed(D, eps, k=m)
    assert val.shape == (m,) and vec.shape == (m, m) and iters == m - 1
    lam, lam_r = np.sqrt(val), np.sqrt(val_r)
    assert np.max(np.abs(lam - lam_r)) < 1e-12                                 # every eigenvalue, absolute (lambda_0 = 1)
    assert np.max(np.abs(val[:32] - g['val'][:32]) / g['val'][:32]) < VAL_TOL
    assert np.max(np.abs(vec.T @ vec - np.eye(m))) < 1e-10
    _, _, Ms = O.markov_symmetric(D, eps)
    assert np.max(np.abs(Ms @ vec - vec * lam[None, :])) < 1e-10
    for i in range(5):
        assert vec_err(vec[:, i], g['vec'][:, i]) < VEC_TOL, i
    # an intermediate request (64 < k < m) takes the same path
    val2, vec2, _ = ctx.dmap_embed(None, eps, k=100, m=m)
    assert np.array_equal(val2, val[:100]) and np.array_equal(vec2, vec[:, :100])
    # duplicated images: the Krylov space ends early, fewer than m pairs exist -> the call says so
    idx = np.repeat(np.arange(80), 2)
    Dd = np.ascontiguousarray(D[np.ix_(idx, idx)])                             # 160 images, every one twice: rank(Ms) <= 80
    with pytest.raises(_capi.EsperError):
        ctx.dmap_embed(Dd, eps, k=160)
    val3, vec3, it3 = ctx.dmap_embed(Dd, eps, k=70)                            # ... but the pairs that exist are available
    val3r, _ = O.dmap_embed(Dd, eps)
    assert it3 <= 100 and np.max(np.abs(np.sqrt(val3) - np.sqrt(val3r[:70]))) < 1e-12
