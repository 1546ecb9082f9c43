"""CPU execution of the moment-sweep kernels of csrc/k2_sweep.cu through a minimal CUDA execution-model stand-in
(tests/emu/cuda_emu.h: one OS thread per CUDA thread, barriers for __syncthreads and warp shuffles).  The kernel text is
cut from the .cu source at test time, so what runs here is the code that is compiled for the GPU -- control flow, indexing,
plan, reductions -- with host arithmetic.  Test infrastructure only; the product never runs these kernels on the CPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from manifoldem_esper_b200 import moment_sweep as MSW


def _proto():
    import importlib.util
    spec = importlib.util.spec_from_file_location('sweep_moments_proto', os.path.join(ROOT, 'tools', 'sweep_moments_proto.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


PROTO = _proto()
from oracle import esper_oracle as O

SRC = os.path.join(ROOT, 'manifoldem_esper_b200', 'csrc', 'k2_sweep.cu')
EMU = os.path.join(ROOT, 'tests', 'emu')
LOG_EPS = np.arange(-100, 100.2, 0.2)


def check_run(res):
    """The stand-in needs up to ~2000 OS threads; an environment that refuses them skips instead of failing."""
    if res.returncode != 0 and 'Resource temporarily unavailable' in (res.stderr or ''):
        pytest.skip('cannot create the threads of the execution-model stand-in here')
    assert res.returncode == 0, (res.stderr or '')[-2000:]


def cut(text, start, stop):
    a = text.index(start)
    return text[a:text.index(stop, a)]


@pytest.fixture(scope='module')
def emu_bin(tmp_path_factory):
    d = tmp_path_factory.mktemp('emu')
    src = open(SRC).read()
    helpers = cut(src, '__device__ __forceinline__ double warp_min(double v) {', 'constexpr int TR = 8;')
    kernels = cut(src, 'constexpr int MP = 16;', '}  // namespace sweep')
    assert 'moment_accum_kernel' in kernels and 'moment_table_kernel' in kernels and 'exp_neg_fast' in helpers
    (d / 'moment_kernels.inc').write_text(helpers + '\n' + kernels)
    exe = str(d / 'moment_emu')
    cmd = ['g++', '-O1', '-std=c++20', '-pthread', '-Wno-unknown-pragmas', '-I', EMU, '-I', str(d),
           os.path.join(EMU, 'moment_emu.cpp'), '-o', exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    return exe, d


def run_emu(emu_bin, D, coef, thr, sym, sms=3):
    exe, d = emu_bin
    D = np.ascontiguousarray(D, dtype=np.float64)
    rows, m = D.shape
    D.tofile(str(d / 'D.bin'))
    np.ascontiguousarray(coef, dtype=np.float64).tofile(str(d / 'coef.bin'))
    res = subprocess.run([exe, str(d / 'D.bin'), str(m), str(rows), str(int(sym)), str(d / 'coef.bin'), str(len(coef)), repr(float(thr)),
                          str(d / 'out.bin'), str(sms)], capture_output=True, text=True, timeout=600)
    check_run(res)
    out = np.fromfile(str(d / 'out.bin'))
    n = len(coef)
    return dict(status=int(out[0]), bins=int(out[1]), logsum=out[2:2 + n], table=out[2 + n:2 + n + MSW.MB * MSW.MV].reshape(MSW.MB, MSW.MV),
                plan=out[2 + n + MSW.MB * MSW.MV:])


@pytest.mark.parametrize('name', ['stage123_small', 'stage123_pristine', 'stage1_dupes'])
def test_moment_kernels_on_the_cpu_vs_reference_golden(emu_bin, name):
    g = golden(name + '.npz')
    D = g['D']
    coef = 1. / (2. * np.exp(LOG_EPS))
    ref = g['logSumWij'] if 'logSumWij' in g else O.eps_sweep_logsum(D, LOG_EPS)
    r = run_emu(emu_bin, D, coef, 10.0, sym=True)
    L_np, bins_np = PROTO.sweep_numpy(D, coef, 10.0)
    assert r['status'] == 0 and r['bins'] == bins_np
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(r['logsum']), fin)
    assert np.max(np.abs(r['logsum'][fin] - ref[fin])) < 1e-9               # the GPU tests' tolerance
    assert np.max(np.abs(r['logsum'][fin] - L_np[fin])) < 1e-12              # the numpy statement of the same plan


def test_moment_kernels_on_the_cpu_full_matrix_row_block_and_limits(emu_bin):
    rng = np.random.default_rng(5)
    D = np.sqrt(1.9 + 0.1 * rng.random((90, 90)))
    np.fill_diagonal(D, 0.0)
    D[3, 7] = 0.0
    coef = 1. / (2. * np.exp(LOG_EPS))
    for thr in (10.0, 3.0, 25.0):
        r = run_emu(emu_bin, D, coef, thr, sym=False)
        ref = []
        with np.errstate(divide='ignore'):
            for c in coef:
                t = c * (D * D)
                ref.append(np.log(np.sum(np.exp(-t[t < thr]))))
        assert r['status'] == 0 and r['bins'] == 1 and np.allclose(r['logsum'], ref, rtol=0, atol=1e-12), thr
    # a row block (the oversized-PD path): plan and table against the host statement
    blk = D[16:53]
    r = run_emu(emu_bin, blk, coef, 10.0, sym=False, sms=2)
    q = (blk * blk).ravel()
    st, pl = MSW.plan(q[q != 0].min(), q.max(), float((q == 0).sum()), float(q.size), coef, 10.0)
    assert st == 0 and np.array_equal(r['plan'][:8 + 6], pl[:8 + 6])
    tab = PROTO.accumulate_numpy(q, np.ones_like(q), pl, coef, 10.0)
    assert np.allclose(r['table'], tab, rtol=1e-13, atol=1e-13)
    # more than 32 bins / too many coefficients per bin: status tells the host to run the general kernel
    W = np.exp(rng.uniform(-12, 12, (40, 40)))
    assert run_emu(emu_bin, W, coef, 10.0, sym=False)['status'] == 1
    fine = 1. / (2. * np.exp(np.arange(-3, 3, 0.01)))
    assert run_emu(emu_bin, D, fine, 10.0, sym=False)['status'] == 2
    # odd and tiny orders, symmetric enumeration of the upper triangle
    for m in (1, 2, 5, 33):
        Dm = np.sqrt(1.0 + rng.random((m, m)))
        Dm = 0.5 * (Dm + Dm.T)
        np.fill_diagonal(Dm, 0.0)
        r = run_emu(emu_bin, Dm, coef, 10.0, sym=True)
        L_np, _ = PROTO.sweep_numpy(Dm, coef, 10.0)
        assert r['status'] == 0 and np.allclose(r['logsum'], L_np, rtol=0, atol=1e-12), m


# ---- K0: stats_split_kernel<HALF> and the cluster variant ---------------------------------------------------------------
K0 = os.path.join(ROOT, 'manifoldem_esper_b200', 'csrc', 'k0_prep.cu')
DEVUTIL = os.path.join(ROOT, 'manifoldem_esper_b200', 'csrc', 'devutil.cuh')


def arena_shared(text):
    import re
    text = re.sub(r'extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];', r'\1* \2 = emu_dyn<\1>();', text)
    return re.sub(r'__shared__\s+(\w+)\s+(\w+)\[([^\]]+)\];', r'\1* \2 = emu_smem<\1>(\3);', text)


@pytest.fixture(scope='module')
def prep_bin(tmp_path_factory):
    d = tmp_path_factory.mktemp('emu_prep')
    dev = open(DEVUTIL).read()
    (d / 'devutil_body.inc').write_text(cut(dev, 'struct RowStat', '}  // namespace esper'))
    src = open(K0).read()
    parts = [cut(src, '__device__ __forceinline__ void row_stats(', '// Statistics of the images i = blockIdx.x * stride'),
             cut(src, 'template <bool HALF>\n__global__ void __launch_bounds__(512) stats_split_kernel', '// ------------------------------------------------------------------------------------------------\n// numpy-exact float32 statistics.')]
    text = arena_shared('\n'.join(parts))
    assert '__shared__' not in text and 'stats_split_kernel' in text
    (d / 'prep_kernels.inc').write_text(text)
    exe = str(d / 'prep_emu')
    cmd = ['g++', '-O1', '-std=c++20', '-pthread', '-ffp-contract=off', '-Wno-unknown-pragmas', '-I', EMU, '-I', str(d),
           os.path.join(EMU, 'prep_emu.cpp'), '-o', exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    return exe, d


def tf32_rna(x):
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return ((b + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def run_prep(prep_bin, raw, ld, standardize, xbar, half, scale_log2, cs):
    exe, d = prep_bin
    n, P = raw.shape
    raw.astype(np.float32).tofile(str(d / 'raw.bin'))
    xb = '-'
    if xbar is not None:
        xbar.astype(np.float32).tofile(str(d / 'xbar.bin'))
        xb = str(d / 'xbar.bin')
    res = subprocess.run([exe, str(d / 'raw.bin'), str(n), str(P), str(ld), str(int(standardize)), xb, str(int(half)),
                          repr(float(2.0 ** scale_log2)), str(cs), str(d / 'hi.bin'), str(d / 'lo.bin'), str(d / 'st.bin')],
                         capture_output=True, text=True, timeout=900)
    check_run(res)
    dt = np.float16 if half else np.float32
    hi = np.fromfile(str(d / 'hi.bin'), dtype=dt).reshape(n, ld)
    lo = np.fromfile(str(d / 'lo.bin'), dtype=dt).reshape(n, ld)
    st = np.fromfile(str(d / 'st.bin')).reshape(n, 3)
    return hi, lo, st


def expected_planes(raw, st, xbar, standardize, half, scale_log2, ld):
    """The kernel's arithmetic from ITS mean32 / std32: z = fl32(fl32(x - mean) / std), c = fl32(z - xbar), split."""
    n, P = raw.shape
    mean, std = st[:, 0].astype(np.float32)[:, None], st[:, 1].astype(np.float32)[:, None]
    z = ((raw - mean).astype(np.float32) / std).astype(np.float32) if standardize else raw.astype(np.float32)
    c = (z - xbar[None, :]).astype(np.float32) if xbar is not None else z
    if half:
        cs_ = (c * np.float32(2.0 ** scale_log2)).astype(np.float32)
        h = cs_.astype(np.float16)
        l = (cs_ - h.astype(np.float32)).astype(np.float32).astype(np.float16)
    else:
        h = tf32_rna(c)
        l = tf32_rna((c - h).astype(np.float32))
    H = np.zeros((n, ld), dtype=h.dtype); L = np.zeros((n, ld), dtype=h.dtype)
    H[:, :P] = h; L[:, :P] = l
    return H, L, np.sum(c.astype(np.float64) ** 2, axis=1)


@pytest.mark.parametrize('half', [0, 1])
@pytest.mark.parametrize('P,cs', [(1000, 0), (999, 0), (250, 0), (7, 0)])
def test_split_kernels_on_the_cpu(prep_bin, P, cs, half):
    rng = np.random.default_rng(P + 10 * cs + half)
    n = 5
    raw = (rng.standard_normal((n, P)) * 2.5 + 0.7).astype(np.float32)
    raw[1, 3] = 400.0                                                         # an outlier pixel: a large z-score
    xbar = (rng.standard_normal(P) * 0.1).astype(np.float32)
    ld = -(-P // (64 if half else 32)) * (64 if half else 32)
    s = 6
    for standardize, xb in ((1, xbar), (1, None), (0, xbar)):
        hi, lo, st = run_prep(prep_bin, raw, ld, standardize, xb, half, s, cs)
        if standardize:
            m64, s64 = raw.astype(np.float64).mean(axis=1), raw.astype(np.float64).std(axis=1)
            assert np.all(np.abs(st[:, 0] - m64) <= 2e-7 * np.abs(m64) + 1e-9) and np.all(np.abs(st[:, 1] - s64) <= 2e-7 * s64)
        else:
            assert np.all(st[:, 0] == 0) and np.all(st[:, 1] == 1)
        H, L, nrm = expected_planes(raw, st, xb, standardize, half, s, ld)
        assert np.array_equal(hi.view(np.uint16 if half else np.uint32), H.view(np.uint16 if half else np.uint32))
        assert np.array_equal(lo.view(np.uint16 if half else np.uint32), L.view(np.uint16 if half else np.uint32))
        assert np.allclose(st[:, 2], nrm, rtol=1e-13, atol=0)
        # hi + lo carries 22 bits of the centred z-score
        c = (H.astype(np.float64) + L.astype(np.float64))[:, :P] / (2.0 ** s if half else 1.0)
        z = ((raw - st[:, 0].astype(np.float32)[:, None]).astype(np.float32) / st[:, 1].astype(np.float32)[:, None]).astype(np.float32) if standardize else raw
        cref = (z - xb[None, :]).astype(np.float32) if xb is not None else z
        floor = 2.0 ** -25 / 2.0 ** s if half else 0.0                        # half the FP16 subnormal spacing, unscaled
        assert np.all(np.abs(c - cref) <= 2.0 ** -22 * np.abs(cref) + floor)


def test_gram_tile_maps_exhaustively(tmp_path):
    """tile_coords / tile_map / wide_coords / wide_tile_count of k1_gram.cu on the host for every tile count up to 80
    (N = 10 240 images): the wide enumeration covers the upper triangle exactly once, the row-block map its rectangle."""
    src = open(os.path.join(ROOT, 'manifoldem_esper_b200', 'csrc', 'k1_gram.cu')).read()
    text = cut(src, '// Upper-triangular tile index t -> (ta, tb)'.replace('(ta, tb)', '(ti, tj)'), '// ---- wide (128 x 256) tiles')
    text += cut(src, 'constexpr int WIDE_TRI = -3;', '__device__ __forceinline__ void umma_tf32_w(')
    assert 'tile_map' in text and 'wide_coords' in text and 'wide_tile_count' in text
    (tmp_path / 'tilemap.inc').write_text(text)
    exe = str(tmp_path / 'tilemap_emu')
    res = subprocess.run(['g++', '-O1', '-std=c++17', '-I', str(tmp_path), os.path.join(EMU, 'tilemap_emu.cpp'), '-o', exe],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    res = subprocess.run([exe, '80'], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and res.stdout.strip() == 'ok 80', res.stdout[-500:]


# ---- K1 epilogues: split-K partial tiles -> distances / Gram matrix, plain and wide tile enumeration ------------------
@pytest.fixture(scope='module')
def finalize_bin(tmp_path_factory):
    d = tmp_path_factory.mktemp('emu_fin')
    src = open(os.path.join(ROOT, 'manifoldem_esper_b200', 'csrc', 'k1_gram.cu')).read()
    text = cut(src, '// Upper-triangular tile index t -> (ti, tj)', '// ---- wide (128 x 256) tiles')
    text += cut(src, '// Distance epilogue: D_ij = sqrt(', '// General product epilogue (CTF branch)')
    assert 'dist_finalize_kernel' in text and 'gram_finalize_kernel' in text and 'gscale' in text
    (d / 'finalize.inc').write_text(text)
    exe = str(d / 'finalize_emu')
    res = subprocess.run(['g++', '-O1', '-std=c++20', '-pthread', '-ffp-contract=off', '-Wno-unknown-pragmas', '-I', EMU, '-I', str(d),
                          os.path.join(EMU, 'finalize_emu.cpp'), '-o', exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    return exe, d


def tile_native(T):
    """[128, 128] tile -> the layout the Gram kernels store: element (r, c) at ((c >> 2) * 128 + r) * 4 + (c & 3)."""
    return T.reshape(128, 32, 4).transpose(1, 0, 2).reshape(-1)


@pytest.mark.parametrize('n,wide', [(300, False), (300, True), (640, True), (129, True)])
@pytest.mark.parametrize('gram', [0, 1])
def test_gram_epilogues_on_the_cpu(finalize_bin, n, wide, gram):
    exe, d = finalize_bin
    rng = np.random.default_rng(n + wide)
    P, split_k, s_log2 = 777, 3, 5
    C = rng.standard_normal((n, P)).astype(np.float32)
    G = C.astype(np.float64) @ C.astype(np.float64).T
    nrm = np.diag(G).copy()
    nt = -(-n // 128)
    Gp = np.zeros((nt * 128, nt * 128))
    Gp[:n, :n] = G
    gscale = 2.0 ** (-2 * s_log2) if (wide and not gram) else 1.0              # the FP16 path hands scaled sums to the epilogue
    if wide:
        nt2 = (nt + 1) // 2
        tiles = [(i, 2 * tc + h) for i in range(nt) for tc in range(i // 2, nt2) for h in (0, 1)]   # virtual tiles 2w + h
    else:
        tiles = [(i, j) for i in range(nt) for j in range(i, nt)]
    part = np.zeros((split_k, len(tiles), 128 * 128), dtype=np.float32)
    wts = rng.dirichlet(np.ones(split_k))
    for t, (ti, tj) in enumerate(tiles):
        if tj >= nt or tj < ti:
            part[:, t] = np.nan                                                    # redundant / out-of-range virtual tile: never read
            continue
        T = Gp[ti * 128:(ti + 1) * 128, tj * 128:(tj + 1) * 128] / gscale
        for ch in range(split_k):
            part[ch, t] = tile_native((T * wts[ch]).astype(np.float32))
    part.tofile(str(d / 'part.bin')); nrm.tofile(str(d / 'nrm.bin'))
    res = subprocess.run([exe, str(d / 'part.bin'), str(d / 'nrm.bin'), str(n), str(P), str(nt), str(len(tiles)), str(-3 if wide else -1),
                          str(split_k), repr(gscale), str(gram), str(d / 'out.bin')], capture_output=True, text=True, timeout=900)
    check_run(res)
    out = np.fromfile(str(d / 'out.bin')).reshape(n, n)
    assert np.array_equal(out, out.T) and not np.any(out == -7.0)
    Gs = np.zeros((nt * 128, nt * 128))                                            # what the kernel sums: float32 partials, FP64 adds
    for t, (ti, tj) in enumerate(tiles):
        if tj < nt and tj >= ti:
            tot = np.zeros(128 * 128)
            for ch in range(split_k):
                tot += part[ch, t].astype(np.float64)
            Gs[ti * 128:(ti + 1) * 128, tj * 128:(tj + 1) * 128] = tot.reshape(32, 128, 4).transpose(1, 0, 2).reshape(128, 128)
    Gs = np.triu(Gs) + np.triu(Gs, 1).T
    Gs = Gs[:n, :n]
    if gram:
        ref = Gs.copy()
        np.fill_diagonal(ref, nrm)
        assert np.array_equal(out, ref)
    else:
        d2 = (nrm[:, None] + nrm[None, :] - 2.0 * (Gs * gscale)) * (1.0 / P)
        ref = np.sqrt(np.maximum(d2, 0.0))
        np.fill_diagonal(ref, 0.0)
        assert np.allclose(out, ref, rtol=1e-15, atol=0) and np.all(np.diag(out) == 0)
        true = np.sqrt(np.maximum(nrm[:, None] + nrm[None, :] - 2 * G, 0) / P)
        off = ~np.eye(n, dtype=bool)
        assert np.max(np.abs(out - true)[off] / true[off]) < 1e-5


# ---- K4: the persistent cooperative Lanczos kernels (grid barriers between concurrently running blocks) -----------------
def scalar_shared(text):
    import re
    return re.sub(r'__shared__\s+(\w+)\s+(\w+);', r'\1& \2 = *emu_smem<\1>(1);', text)


@pytest.fixture(scope='module')
def lanczos_bin(tmp_path_factory):
    d = tmp_path_factory.mktemp('emu_lan')
    src = open(os.path.join(ROOT, 'manifoldem_esper_b200', 'csrc', 'k4_eigen.cu')).read()
    text = cut(src, '__device__ __forceinline__ double ld_cg(const double* p)', '// ------------------------------------------------------------------------------------------------\n// Small-m variant (the whole Lanczos vector fits in shared memory): ONE grid barrier per step.')
    text = text.replace('asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");', 'emu_red_release_add(bar);')
    text = text.replace('asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");', 'seen = emu_ld_acquire(bar);')
    text = text.replace('asm volatile("fence.acq_rel.gpu;" ::: "memory");', '__atomic_thread_fence(__ATOMIC_SEQ_CST);')
    assert 'asm' not in text and 'lanczos_resident_kernel' in text and 'lanczos_steps_kernel' in text
    text = scalar_shared(arena_shared(text))
    assert '__shared__' not in text
    (d / 'lanczos.inc').write_text(text)
    exe = str(d / 'lanczos_emu')
    res = subprocess.run(['g++', '-O1', '-std=c++20', '-pthread', '-Wno-unknown-pragmas', '-I', EMU, '-I', str(d),
                          os.path.join(EMU, 'lanczos_emu.cpp'), '-o', exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    return exe, d


@pytest.mark.parametrize('variant,m,grid', [(0, 45, 3), (1, 45, 4), (2, 45, 4), (1, 29, 3)])
def test_lanczos_kernels_on_the_cpu(lanczos_bin, variant, m, grid):
    """Full-reorthogonalisation Lanczos on a small Markov matrix: alpha / beta against a numpy Lanczos with the same start,
    orthonormal basis, Ritz values equal to the leading eigenvalues; batches as the host issues them (state carried over)."""
    exe, d = lanczos_bin
    g = golden('stage123_pristine.npz')
    D = g['D'][:m, :m]
    Ms = O.markov_symmetric_dense(D, float(g['eps']) if 'eps' in g else 0.0128)
    lam, U = np.linalg.eigh(Ms)
    v0 = U[:, -1] * np.sign(U[:, -1].sum())
    q = np.cos(np.arange(m) * 0.7 + 0.3)
    for _ in range(2):
        q -= v0 * (v0 @ q)
    q /= np.linalg.norm(q)
    steps = 14
    Ms.tofile(str(d / 'Ms.bin')); np.stack([v0, q]).tofile(str(d / 'V0.bin'))
    res = subprocess.run([exe, str(d / 'Ms.bin'), str(m), str(d / 'V0.bin'), str(steps), str(variant), str(grid), '5', str(d / 'out.bin')],
                         capture_output=True, text=True, timeout=900)
    check_run(res)
    out = np.fromfile(str(d / 'out.bin'))
    alpha, beta, V = out[:steps], out[steps:2 * steps], out[2 * steps:].reshape(steps + 2, m)
    # numpy Lanczos with full (two-pass) reorthogonalisation against [v0, q1, ...]
    Vn = [v0, q]; an, bn = [], []
    for j in range(steps):
        w = Ms @ Vn[-1]
        B = np.array(Vn)
        a = 0.0
        for _ in range(2):
            hh = B @ w
            w = w - B.T @ hh
            a += hh[-1]
        b = np.linalg.norm(w)
        an.append(a); bn.append(b); Vn.append(w / b)
    assert np.allclose(alpha, an, rtol=1e-9, atol=1e-13) and np.allclose(beta, bn, rtol=1e-9, atol=1e-13)
    assert np.max(np.abs(V[:steps + 1] @ V[:steps + 1].T - np.eye(steps + 1))) < 1e-12
    T = np.diag(alpha) + np.diag(beta[:-1], 1) + np.diag(beta[:-1], -1)
    th = np.linalg.eigvalsh(T)[::-1]
    assert abs(th[0] - lam[-2]) / lam[-2] < 1e-8                          # leading non-trivial eigenvalue


@pytest.fixture(scope='module')
def chain_bin(tmp_path_factory):
    """The barrier-free Lanczos kernel of csrc/k4_chain.cu (workers + on-device convergence checkers) for the CPU stand-in."""
    import re
    d = tmp_path_factory.mktemp('emu_chain')
    src = open(os.path.join(ROOT, 'manifoldem_esper_b200', 'csrc', 'k4_chain.cu')).read()
    (d / 'chain_structs.inc').write_text(cut(src, 'struct Ctl {', '// ---- packet and flag primitives'))
    text = cut(src, '// ---- begin chain kernel text', '// ---- end chain kernel text')
    text = re.sub(r'extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];', r'\1* \2 = emu_dyn<\1>();', text)
    text = re.sub(r'__shared__\s+(\w+(?:\s+\w+)?)\s+(\w+)\[([^\]]+)\];', r'\1* \2 = emu_smem<\1>(\3);', text)
    text = re.sub(r'__shared__\s+(\w+(?:\s+\w+)?)\s+(\w+);', r'\1& \2 = *emu_smem<\1>(1);', text)
    assert '__shared__' not in text and 'asm' not in text and 'lanczos_chain_kernel' in text and 'checker_main' in text and 'lanczos_cluster_kernel' in text and 'lanczos_merged_kernel' in text
    (d / 'chain.inc').write_text(text)
    exe = str(d / 'chain_emu')
    res = subprocess.run(['g++', '-O1', '-std=c++20', '-pthread', '-Wno-unknown-pragmas', '-I', EMU, '-I', str(d),
                          os.path.join(EMU, 'chain_emu.cpp'), '-o', exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    return exe, d


@pytest.mark.parametrize('m,grid,kw,max_steps,expect', [(45, 5, 4, 40, 'converged'), (20, 10, 4, 19, 'exhausted'), (45, 5, 4, 10, 'noconv'),
                                                         (100, 9, 6, 90, 'converged'), (40, 16, 4, 36, 'converged')])
def test_chain_lanczos_kernel_on_the_cpu(chain_bin, m, grid, kw, max_steps, expect):
    """Workers (rows of Ms and of the basis on chip, packet hand-overs, no grid barrier) and checker CTAs (multisection
    Sturm counts, inverse iteration, stop word) together: alpha / beta against a numpy Lanczos with the same start,
    orthonormal basis, Ritz pairs of the winning checker against numpy's eigh; exhausted Krylov space and max_steps."""
    run_chain_case(chain_bin, m, grid, kw, max_steps, expect, 0)


@pytest.mark.parametrize('m,grid,S,kw,max_steps,expect', [(45, 6, 2, 4, 40, 'converged'),      # 3 clusters of 2, the last one: 1 worker + 1 checker
                                                           (100, 12, 4, 6, 90, 'converged'),     # 10 rows per worker, 10 workers: 4 + 4 + 2, two checkers
                                                           (20, 8, 4, 4, 19, 'exhausted'),       # 3 rows per worker, 7 workers, Krylov space exhausted
                                                           (45, 8, 8, 4, 10, 'noconv'),          # ONE cluster (no packet level), max_steps reached
                                                           (61, 9, 3, 4, 50, 'converged'),       # odd cluster size: gather groups of 4 lanes for 3 clusters
                                                           (50, 4, 2, 4, 45, 'converged')])      # 17 rows per worker (10 in shared memory, 7 in registers)
def test_cluster_lanczos_kernel_on_the_cpu(chain_bin, m, grid, S, kw, max_steps, expect):
    """The cluster schedule (lanczos_cluster_kernel): partial coefficients and the vector exchanged through distributed shared
    memory (st.async on the consumer's mbarrier) inside clusters of S CTAs and through packets between the clusters; same checks
    as for the flat kernel.  Covers clusters that are not full, one cluster, rows in registers only and in shared memory."""
    run_chain_case(chain_bin, m, grid, kw, max_steps, expect, S)


@pytest.mark.parametrize('m,grid,S,kw,max_steps,expect', [(45, 6, 2, 4, 40, 'converged'), (100, 12, 4, 6, 90, 'converged'), (20, 8, 4, 4, 19, 'exhausted'),
                                                           (45, 8, 8, 4, 10, 'noconv'), (61, 9, 3, 4, 50, 'converged'), (50, 4, 2, 4, 45, 'converged')])
def test_merged_lanczos_kernel_on_the_cpu(chain_bin, m, grid, S, kw, max_steps, expect):
    """The merged schedule (lanczos_merged_kernel): the unorthogonalised candidate is gathered together with the reduction of its
    coefficients (ONE exchange per step), Ms q_{j+1} follows from Ms w_j by linearity and the three-term relation of the basis.
    Same cases and checks as the two-exchange cluster kernel: alpha / beta against a numpy Lanczos with direct products,
    orthonormal basis, Ritz pairs; the single vector buffer is guarded by the count-based mbarrier with remote arrives."""
    run_chain_case(chain_bin, m, grid, kw, max_steps, expect, S, merged=1)


def run_chain_case(chain_bin, m, grid, kw, max_steps, expect, S, merged=0):
    exe, d = chain_bin
    g = golden('stage23_medium.npz' if m > 64 else 'stage123_pristine.npz')
    Ms = O.markov_symmetric_dense(g['D'][:m, :m], float(g['eps']))
    lam, U = np.linalg.eigh(Ms)
    v0 = U[:, -1] * np.sign(U[:, -1].sum())
    q = np.cos(np.arange(m) * 0.7 + 0.3)
    for _ in range(2):
        q -= v0 * (v0 @ q)
    q /= np.linalg.norm(q)
    Ms.tofile(str(d / 'Ms.bin')); np.stack([v0, q]).tofile(str(d / 'V0.bin'))
    res = subprocess.run([exe, str(d / 'Ms.bin'), str(m), str(d / 'V0.bin'), str(kw), '2', '1e-10', str(grid), str(max_steps), str(d / 'out.bin'), str(S), str(merged)],
                         capture_output=True, text=True, timeout=900, env=dict(os.environ, EMU_TIMEOUT_S='240'))
    check_run(res)
    out = np.fromfile(str(d / 'out.bin'))
    status, n_final, winner, steps_run, abort = (int(x) for x in out[:5])
    o = 5
    alpha = out[o:o + max_steps]; o += max_steps
    beta = out[o:o + max_steps]; o += max_steps
    theta = out[o:o + 24]; o += 24
    S = out[o:o + max_steps * kw].reshape(max_steps, kw); o += max_steps * kw
    V = out[o:].reshape(max_steps + 2, m)
    assert abort == 0 and 1 <= steps_run <= max_steps
    Vn = [v0, q]; an, bn = [], []
    for j in range(steps_run):
        w = Ms @ Vn[-1]
        B = np.array(Vn)
        a = 0.0
        for _ in range(2):
            hh = B @ w
            w = w - B.T @ hh
            a += hh[-1]
        b = np.linalg.norm(w)
        an.append(a); bn.append(b); Vn.append(w / b)
    assert np.allclose(alpha[:steps_run], an, rtol=1e-9, atol=1e-13) and np.allclose(beta[:steps_run], bn, rtol=1e-9, atol=1e-13)
    assert np.max(np.abs(V[:steps_run + 1] @ V[:steps_run + 1].T - np.eye(steps_run + 1))) < 1e-12
    if expect == 'noconv':
        assert status == 2 and winner == -1 and steps_run == max_steps
        return
    assert status == 1 and winner >= 0 and n_final <= steps_run
    if expect == 'exhausted':
        assert n_final == m - 1
    else:
        assert n_final < max_steps                                             # stopped by a passing check, not by the cap
        # the sizes tested are a fixed schedule and the oldest passing test wins: the step count of the result is reproducible
        chunk = -(-m // (grid - 1)); n_checkers = min(grid - (-(-m // chunk)), 4)
        first = min(max_steps, max(3 * (kw + 2) + 8, 12))
        assert (n_final - first) % (2 if n_checkers >= 2 else 3) == 0
    true = lam[::-1][1:kw + 1]
    assert np.max(np.abs(theta[:kw] - true) / true) < 1e-12
    Uv = V[1:n_final + 1].T @ S[:n_final]
    for i in range(kw):
        u = Uv[:, i] / np.linalg.norm(Uv[:, i])
        assert np.linalg.norm(Ms @ u - theta[i] * u) < 1e-9


def test_moment_kernels_on_the_cpu_random_cases(emu_bin):
    """The kernel text against its numpy statement on random spreads, zero patterns, grids and thresholds (plan status,
    bin count and every sum)."""
    rng = np.random.default_rng(1)
    agree = 0
    for trial in range(14):
        m = int(rng.integers(2, 70))
        spread = 10 ** rng.uniform(-3, 1.0)
        base = 10 ** rng.uniform(-4, 4)
        D = np.sqrt(base * np.exp(rng.uniform(0, spread * np.log(10), (m, m))))
        D = 0.5 * (D + D.T)
        if trial % 3:
            np.fill_diagonal(D, 0)
        step = [0.2, 0.1, 0.5][trial % 3]
        coef = 1. / (2. * np.exp(np.arange(-30, 30 + step / 2, step) + np.log(base)))
        thr = [10.0, 3.0, 25.0, 6.5][trial % 4]
        r = run_emu(emu_bin, D, coef, thr, sym=True, sms=1 + trial % 4)
        L, info = PROTO.sweep_numpy(D, coef, thr, symmetric=True)
        if L is None:
            assert r['status'] == info
            continue
        assert r['status'] == 0 and r['bins'] == info
        fin = np.isfinite(L)
        assert np.array_equal(np.isfinite(r['logsum']), fin) and (not fin.any() or np.max(np.abs(r['logsum'][fin] - L[fin])) < 1e-12)
        agree += 1
    assert agree >= 6


def test_umma_instruction_descriptors_match_the_cutlass_encoder(tmp_path):
    """The four tcgen05 instruction descriptors of k1_gram.cu (TF32 / FP16 operands, N = 128 / 256) against
    cute::UMMA::make_instr_desc of the CUTLASS headers vendored in the image (library code used as a checker only)."""
    import glob
    import re
    inc = [p for p in glob.glob('/opt/prime-rl/.venv/lib/python3*/site-packages/*/data/cutlass/include') +
           glob.glob('/opt/prime-rl/.venv/lib/python3*/site-packages/*/3rdparty/cutlass/include')
           if os.path.exists(os.path.join(p, 'cute', 'arch', 'mma_sm100_desc.hpp'))]
    if not inc:
        pytest.skip('no CUTLASS headers with mma_sm100_desc.hpp in this image')
    src = open(os.path.join(ROOT, 'manifoldem_esper_b200', 'csrc', 'k1_gram.cu')).read()
    defs = re.findall(r'^constexpr (?:int|uint32_t) (?:BM|BN|BN_W|IDESC|IDESC_W|IDESC_WH|IDESC_H) = [^;]+;', src, flags=re.M)
    assert len(defs) == 7, defs
    prog = '#include <cstdint>\n#include <cute/arch/mma_sm100_desc.hpp>\n#include <cutlass/half.h>\n#include <cutlass/tfloat32.h>\n' + '\n'.join(defs) + '''
template <class T, int N> constexpr uint32_t ref() {
    return uint32_t(cute::UMMA::make_instr_desc<T, T, float, 128, N, cute::UMMA::Major::K, cute::UMMA::Major::K>());
}
int main() {
    return (ref<cutlass::tfloat32_t, 128>() == IDESC ? 0 : 1) + (ref<cutlass::tfloat32_t, 256>() == IDESC_W ? 0 : 2)
         + (ref<cutlass::half_t, 256>() == IDESC_WH ? 0 : 4) + (ref<cutlass::half_t, 128>() == IDESC_H ? 0 : 8);
}
'''
    (tmp_path / 'idesc.cu').write_text(prog)
    exe = str(tmp_path / 'idesc')
    res = subprocess.run(['nvcc', '-std=c++17', '-Wno-deprecated-gpu-targets', '-I', inc[0], '-o', exe, str(tmp_path / 'idesc.cu')],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    assert subprocess.run([exe]).returncode == 0
