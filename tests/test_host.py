"""CPU: host-side logic, the C-ABI surface and the no-fallback rule (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden
from manifoldem_esper_b200 import (ConicFit, Distances, Generate_Nd_Rot, Rotation_Histograms as RH, _capi,
                                   launcher, mrc, synth)
from oracle import esper_oracle as O


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'esper_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(esper_[a-z_0-9]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), 'libesper_b200.so does not export %s' % s
    assert sorted(_capi.SIGNATURES) == syms                      # the ctypes table covers the whole header
    assert _capi.load_library().esper_abi_version() == 1


def test_library_is_sm100a_tcgen05_code(lib_built):
    import subprocess
    sass = subprocess.run(['cuobjdump', '-sass', lib_built], capture_output=True, text=True).stdout
    assert 'sm_100a' in sass
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'LDTM'):               # tcgen05.mma, TMA, tcgen05.ld
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(_capi.EsperError):
        _capi.Context(0)
    with pytest.raises(_capi.EsperError):
        Distances.distances(np.zeros((4, 16), np.float32))


def test_nd_rotation_bits_and_errors(lib_built):
    g = golden('nd_rot.npz')
    for n in (2, 3, 6, 8):
        th, R = g['theta_%d' % n], g['R_%d' % n]
        T = n * (n - 1) // 2
        for i in range(th.shape[0]):
            assert np.array_equal(Generate_Nd_Rot.genNdRotations(n, th[i].reshape(T, 1)), R[i])
            assert np.array_equal(Generate_Nd_Rot.genNdRotations(n, th[i]), R[i])
            assert np.array_equal(Generate_Nd_Rot.genNdRotations(n, list(th[i])), R[i])
        with pytest.raises(ValueError):
            Generate_Nd_Rot.genNdRotations(n, [0.0] * (T + 1))
        with pytest.raises(ValueError):
            Generate_Nd_Rot.genNdRotations(n, np.zeros((T + 1, 1)))
    assert np.array_equal(Generate_Nd_Rot.genNdRotations(1, []), np.eye(1))


def test_hist_edges_equal_numpy_linspace(lib_built):
    for bins, lo, hi in [(51, -1.5, 1.5), (7, 0.1, 0.9), (64, -3.0, 2.0)]:
        assert np.array_equal(_capi.hist_edges(bins, lo, hi), np.linspace(lo, hi, bins + 1))


def test_conic_fits_equal_sympy_rref():
    U0 = golden('rot_inputs.npz')['PD017']
    Ui = O.initial_subspace(U0, 6, True)
    for v1, v2 in [(0, 1), (0, 2), (1, 4), (3, 5)]:
        a = ConicFit.fit2(Ui, v1, v2)
        cF1, R2 = O.conic_fit2(Ui, v1, v2)
        assert np.array_equal(a[0], cF1) and a[3] == R2
        b = ConicFit.fit3(Ui, v1, v2)
        cF, R2b = O.conic_fit3(Ui, v1, v2)
        assert np.array_equal(b[0], cF) and b[1] == R2b


@pytest.mark.parametrize('dim', [6, 8])
def test_select_partners_equals_find_parabolas(dim):
    """The partner selection used with the device R^2 table takes the decisions of findParabolas (:159-177)."""
    fx = golden('rot_inputs.npz')
    for key in ['PD002', 'PD064']:
        Ui = O.initial_subspace(fx[key], dim, True)
        R2 = [ConicFit.fit2(Ui, v1, v2)[3] for v1 in range(dim - 1) for v2 in range(v1 + 1, dim)]
        psi, vals = RH.select_partners(dim, R2)
        psi_ref, _, it = O.find_parabolas(dim, Ui, 1, True)
        assert it == 1 and np.array_equal(psi, psi_ref) and vals[0] >= 0.2


@pytest.mark.parametrize('dim', [6, 8])
def test_rotation_plan_equals_oracle(dim):
    fx = golden('rot_inputs.npz')
    for key in ['PD001', 'PD048', 'PD126']:
        p = RH.prepare(fx[key], dim=dim, PCA=True, CTF=True)
        Ui = O.initial_subspace(fx[key], dim, True)
        psi, Ui2, it = O.find_parabolas(dim, Ui, 1, True)
        cl, se, cf, rij = O.rotation_plan(dim, psi)
        assert p['comboList'] == [(a, int(b)) for a, b in cl]
        assert p['Rij_final'] == rij and p['itIdx'] == it
        assert np.array_equal(p['saveEigs'], se) and np.array_equal(p['U_init'], Ui2)
    assert np.array_equal(RH.normalize(fx['PD001']), O.normalize_manifold(fx['PD001']))
    assert np.array_equal(RH.theta_grid(), np.arange(-40, 40, .5) * np.pi / 180)


def test_mrc_roundtrip(tmp_path):
    st = np.random.default_rng(0).standard_normal((5, 8, 8)).astype(np.float32)
    p = str(tmp_path / 'PD_001.mrcs')
    mrc.write_stack(p, st)
    assert os.path.getsize(p) == 1024 + st.nbytes
    m = mrc.mmap(p, 'r')
    assert m.data.shape == (5, 8, 8) and np.array_equal(np.asarray(m.data), st)
    m.close()
    with open(p, 'r+b') as f:
        f.truncate(1024 + 10)
    with pytest.raises(ValueError):
        mrc.mmap(p)


def test_star_params_columns(tmp_path):
    """Distances.star_params picks the five columns the reference reads (Distances.py:68-72), per image."""
    from manifoldem_esper_b200 import Read_Alignment
    df = np.array([5000., 9000.])
    Read_Alignment.write_star(str(tmp_path / 's.star'), df, volt=200, Cs=2.0, ampc=0.07, pix=1.5)
    par = Distances.star_params(Read_Alignment.parse_star(str(tmp_path / 's.star')))
    assert par.shape == (2, 5) and par.dtype == np.float64
    assert np.array_equal(par, [[1.5, 2.0, 5000., 200., 0.07], [1.5, 2.0, 9000., 200., 0.07]])


def test_argument_errors_raise_valueerror():
    with pytest.raises(ValueError):
        _capi._check_stack(np.zeros((0, 4), np.float32))
    with pytest.raises(ValueError):
        _capi._check_square(np.zeros((3, 4)))


def test_synthetic_generator_is_seeded():
    a = synth.synthetic_pd(pd=2, box=16, n_side=3, tau=2, snr=0.1)
    b = synth.synthetic_pd(pd=2, box=16, n_side=3, tau=2, snr=0.1)
    assert a.dtype == np.float32 and a.shape == (18, 16, 16) and np.array_equal(a, b)
    d = synth.synthetic_pd(pd=2, box=16, n_side=3, tau=2, snr=None)
    assert np.array_equal(d[0], d[1]) and not np.array_equal(d[0], d[2])


def test_pd_parsing_and_sharding():
    pds = launcher.parse_pds('1-5,9,12-13')
    assert pds == ['001', '002', '003', '004', '005', '009', '012', '013']
    parts = [launcher.shard(pds, r, 3) for r in range(3)]
    assert sorted(sum(parts, [])) == pds and max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        launcher.shard(pds, 3, 3)

    def work(pd):
        if pd == '003':
            raise RuntimeError('boom')
        return {'x': int(pd)}
    recs = launcher.run_sharded(pds, work, 0, 1)
    assert [r['ok'] for r in recs] == [True, True, False, True, True, True, True, True]
    assert 'boom' in recs[2]['error']
    # several PDs in flight: one work function per host thread, records still in PD order, failures isolated
    import threading
    import time as _time
    seen = {}

    def make(tag):
        def w(pd):
            seen.setdefault(tag, []).append((pd, threading.get_ident()))
            _time.sleep(0.01)
            return work(pd)
        return w
    recs3 = launcher.run_sharded(pds, [make(t) for t in range(3)], 0, 1, inflight=3)
    assert [r['PD'] for r in recs3] == pds and [r['ok'] for r in recs3] == [r['ok'] for r in recs]
    assert sorted(pd for v in seen.values() for pd, _ in v) == pds and len(seen) == 3
    assert all(len({tid for _, tid in v}) == 1 for v in seen.values())          # a work function stays on its thread
    with pytest.raises(ValueError):
        launcher.run_sharded(pds, work, 0, 1, inflight=2)


def test_philox_restatement_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (the generator of esper_tau_snr_stack)."""
    def kat(c, k):
        r = synth.philox4x32_10(*[np.array([x], dtype=np.uint64) for x in c], k[0], k[1])
        return [int(x[0]) for x in r]
    assert kat((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert kat((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert kat((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    z = synth.philox_normal(99, 3, 200001)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01 and z.shape == (200001,)


def test_mrc_info_host_entry(lib_built, tmp_path):
    st = np.arange(3 * 5 * 7, dtype=np.float32).reshape(3, 5, 7)
    p = str(tmp_path / 'a.mrcs')
    mrc.write_stack(p, st)
    assert _capi.mrc_info(p) == ((7, 5, 3), 2, 1024)
    with pytest.raises(ValueError):
        _capi.mrc_info(str(tmp_path / 'missing.mrcs'))
    with open(str(tmp_path / 'short.mrcs'), 'wb') as f:
        f.write(open(p, 'rb').read()[:1100])
    with pytest.raises(ValueError):
        _capi.mrc_info(str(tmp_path / 'short.mrcs'))


def test_hostbind_is_best_effort_without_a_gpu():
    from manifoldem_esper_b200 import hostbind
    assert hostbind._parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    before = os.sched_getaffinity(0)
    info = hostbind.bind_to_gpu_numa_node(0)               # no CUDA device here: nothing is changed, nothing raises
    assert info['bound'] is False and os.sched_getaffinity(0) == before


# ---- numpy statements of the opt-in kernels: the algorithms the CUDA code follows, pinned on the CPU ----
def _tool(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'tools', name + '.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize('name', ['stage123_small', 'stage123_pristine', 'stage23_medium', 'stage1_dupes'])
def test_moment_sweep_statement_vs_reference_golden(name):
    """The plan of moment_plan_kernel / moment_accum_kernel / moment_final_kernel (k2_sweep.cu), restated in numpy,
    reproduces the reference run's logSumWij (GaussianBandwidth.py:36-44) on every golden matrix."""
    proto = _tool('sweep_moments_proto')
    g = golden(name + '.npz')
    log_eps = np.arange(-100, 100.2, 0.2)
    coef = 1. / (2. * np.exp(log_eps))
    ref = g['logSumWij'] if 'logSumWij' in g else O.eps_sweep_logsum(g['D'], log_eps)
    L, bins = proto.moment_sweep(g['D'], coef, 10.0)
    assert L is not None and 1 <= bins <= proto.MB
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(L), fin)
    assert np.max(np.abs(L[fin] - ref[fin])) < 1e-9                           # the GPU tests' tolerance


def test_moment_sweep_statement_plan_limits():
    proto = _tool('sweep_moments_proto')
    coef = 1. / (2. * np.exp(np.arange(-100, 100.2, 0.2)))
    rng = np.random.default_rng(3)
    W = np.exp(rng.uniform(-12, 12, (40, 40)))                                # D^2 spans e^48: more than 32 bins
    assert proto.moment_sweep(W, coef, 10.0, symmetric=False) == (None, 1)
    fine = 1. / (2. * np.exp(np.arange(-3, 3, 0.01)))                         # 22 coefficients per bin boundary
    D = np.sqrt(1.9 + 0.1 * rng.random((30, 30)))
    np.fill_diagonal(D, 0.0)
    assert proto.moment_sweep(D, fine, 10.0, symmetric=False) == (None, 2)
    for thr in (3.0, 10.0, 25.0):                                             # rho = 2, 1.25, 1.087
        L, bins = proto.moment_sweep(D, coef, thr, symmetric=False)
        ref = []
        for c in coef:
            t = c * (D * D)
            ref.append(np.log(np.sum(np.exp(-t[t < thr]))))
        assert bins == 1 and np.max(np.abs(L - np.array(ref))) < 1e-12


def test_f16_split_statement_matches_the_tf32_split():
    """3xFP16 operand planes (stats_split_kernel<true>) carry the same 22 bits as the 3xTF32 planes: equal Gram and
    distance errors on a noisy and on a noise-free (cancellation regime) stack; the scale keeps FP16 in range."""
    proto = _tool('f16_split_proto')
    assert [proto.f16_scale_log2(P) for P in (16, 4096, 16384, 62500, 250000, 10 ** 9)] == [8, 8, 7, 6, 5, -1]
    for snr, n_side, tau in ((0.1, 6, 3), (None, 9, 1)):
        stack = synth.synthetic_pd(pd=2, box=64, n_side=n_side, tau=tau, snr=snr)
        X = O.standardize_stack(stack).reshape(stack.shape[0], -1)
        c = (X - X[:16].mean(axis=0)).astype(np.float32)
        err = proto.split_errors(c)
        assert err['f16x3'][0] < 1.5e-7 and err['tf32x3'][0] < 1.5e-7         # 2^-22-level Gram error for both
        assert err['f16x3'][1] < 1.5 * err['tf32x3'][1] + 1e-9
        assert err['f16x3'][1] < 1e-5
    # a hot pixel: the largest z-score a row of P pixels can hold stays finite after scaling
    P = 4096
    x = np.full((1, P), -1.0 / np.sqrt(P - 1), dtype=np.float32)
    x[0, 7] = np.sqrt(P - 1)
    hi, lo = proto.split_f16(np.repeat(x, 2, axis=0))
    assert np.all(np.isfinite(hi)) and abs(hi[0, 7] + lo[0, 7] - float(x[0, 7])) < 1e-5


# ---- stage 2b in C: esper_tanh_fit against scipy's curve_fit (the reference's call, GaussianBandwidth.py:47-57) ----
@pytest.mark.parametrize('name', ['stage123_small', 'stage123_pristine', 'stage23_medium'])
def test_tanh_fit_vs_scipy_curve_fit(lib_built, name):
    from scipy.optimize import curve_fit
    from manifoldem_esper_b200 import GaussianBandwidth as GB
    g = golden(name + '.npz')
    log_eps = np.arange(-100, 100.2, 0.2)
    L = g['logSumWij']
    rng = np.random.RandomState(int(g['seed']) if 'seed' in g else 0)
    for trial in range(4):
        a0 = rng.rand(4, 1) - .5                                   # the reference's start (DiffusionMaps.py:74)
        try:
            popt, pcov = curve_fit(GB.fun, log_eps, L, p0=a0)
        except RuntimeError:
            continue
        resnorm = sum(np.sqrt(np.fabs(np.diag(pcov))))
        p, rn, r2, info, nfev = _capi.tanh_fit(log_eps, L, a0)
        if not np.isfinite(resnorm) or resnorm > 100:
            assert rn > 100                                           # both ask for another start
            continue
        assert 1 <= info <= 4 and nfev > 4
        assert np.max(np.abs(p - popt) / np.maximum(np.abs(popt), 1e-3)) < 1e-7
        assert abs(np.exp(-(p[1] / p[0])) - np.exp(-(popt[1] / popt[0]))) / np.exp(-(popt[1] / popt[0])) < 1e-7   # epsilon
        assert abs(rn - resnorm) / resnorm < 1e-5
        res = L - GB.fun(log_eps, *popt)
        assert abs(r2 - (1 - np.sum(res ** 2) / np.sum((L - np.mean(L)) ** 2))) < 1e-10
    if 'eps' in g and 'seed' in g:                                     # the epsilon of the reference run itself
        np.random.seed(int(g['seed']))
        a0 = 1 * (np.random.rand(4, 1) - .5)
        p, _, _ = GB.fit(log_eps, L, a0, native=True)
        assert abs(np.exp(-(p[1] / p[0])) - float(g['eps'])) / float(g['eps']) < 1e-7


def test_tanh_fit_is_minpack_lmdif_bit_for_bit(lib_built):
    """With the same libm tanh inside the model, scipy's MINPACK and the C restatement walk the same trajectory."""
    import math
    from scipy.optimize import leastsq
    g = golden('stage23_medium.npz')
    log_eps = np.arange(-100, 100.2, 0.2)
    L = g['logSumWij']

    def resid(a):
        return np.array([a[3] + a[2] * math.tanh(a[0] * x + a[1]) - y for x, y in zip(log_eps, L)])
    for a0 in ([-0.2, 0.3, -0.4, 0.1], [0.4, -0.1, 0.3, 0.45]):
        ref = leastsq(resid, np.array(a0), full_output=1)
        p, rn, r2, info, nfev = _capi.tanh_fit(log_eps, L, a0)
        assert np.array_equal(p, ref[0]) and nfev == ref[2]['nfev'] and info == ref[4]


def test_tanh_fit_argument_errors(lib_built):
    x = np.arange(10.0)
    with pytest.raises(ValueError):
        _capi.tanh_fit(x, np.full(10, -np.inf), [0.1, 0.1, 0.1, 0.1])   # curve_fit: "array must not contain infs or NaNs"
    with pytest.raises(ValueError):
        _capi.tanh_fit(x, x[:5], [0.1, 0.1, 0.1, 0.1])
    with pytest.raises(ValueError):
        _capi.tanh_fit(x[:3], x[:3], [0.1, 0.1, 0.1, 0.1])


def test_moment_sweep_layout_matches_the_library(lib_built):
    """moment_sweep.py builds the plan the device kernels read: its constants must be the library's."""
    from manifoldem_esper_b200 import moment_sweep
    moment_sweep.check_layout(_capi.load_library())
    coef = 1. / (2. * np.exp(np.arange(-100, 100.2, 0.2)))
    assert moment_sweep.usable(coef, 10.0) and not moment_sweep.usable(coef[::-1], 10.0) and not moment_sweep.usable(coef, np.inf)
    status, pl = moment_sweep.plan(1.85, 1.95, 2000.0, 4e6, coef, 10.0)
    assert status == 0 and pl.shape == (moment_sweep.PLAN_LEN,) and pl[1] == 1 and 0 <= pl[13] - pl[12] <= moment_sweep.MS


def test_product_package_never_touches_the_oracle_or_the_test_tools():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may use oracle/; tools/ holds test statements only."""
    pkg = os.path.join(ROOT, 'manifoldem_esper_b200')
    for fn in sorted(os.listdir(pkg)):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(from|import)\s+(oracle|tools)\b', src, flags=re.M), fn
            assert not re.search(r'importlib[^\n]*(oracle|tools)', src), fn


def test_moment_sweep_statement_random_spreads_and_thresholds():
    """Random orders, spreads (0.1 % to 16 decades), zeros, grids and thresholds: wherever the plan is usable the sums agree
    with the brute-force sweep to 1e-13, everything else is sent to the general kernel."""
    proto = _tool('sweep_moments_proto')
    rng = np.random.default_rng(0)
    used = 0
    for trial in range(80):
        m = int(rng.integers(2, 30))
        spread = 10 ** rng.uniform(-3, 1.2)
        base = 10 ** rng.uniform(-6, 6)
        D = np.sqrt(base * np.exp(rng.uniform(0, spread * np.log(10), (m, m))))
        D = 0.5 * (D + D.T)
        if rng.random() < 0.7:
            np.fill_diagonal(D, 0)
        step = rng.choice([0.2, 0.1, 0.5, 1.0])
        log_eps = np.arange(-40, 40 + step / 2, step) + np.log(base)
        coef = 1. / (2. * np.exp(log_eps))
        thr = float(rng.choice([10, 10, 3, 25, 6.5, 100]))
        L, info = proto.sweep_numpy(D, coef, thr, symmetric=True)
        if L is None:
            assert info in (1, 2)
            continue
        ref = []
        with np.errstate(divide='ignore'):
            for c in coef:
                t = c * (D * D)
                ref.append(np.log(np.sum(np.exp(-t[t < thr]))))
        ref = np.array(ref)
        fin = np.isfinite(ref)
        assert np.array_equal(np.isfinite(L), fin) and (not fin.any() or np.max(np.abs(L[fin] - ref[fin])) < 1e-13)
        used += 1
    assert used >= 40


def test_tridiag_eigen_all_matches_numpy(lib_built):
    """esper_tridiag_eigen_all (host half of the full-spectrum path: implicit QL with recorded rotations) against numpy's eigh."""
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 17, 160):
        a, b = rng.uniform(0.0, 0.05, n), np.abs(rng.normal(0.006, 0.002, n))
        T = np.diag(a) + np.diag(b[:n - 1], 1) + np.diag(b[:n - 1], -1)
        w, v = np.linalg.eigh(T)
        th, Z = _capi.tridiag_eigen_all(a, b)
        assert np.max(np.abs(th - w[::-1])) < 1e-15 and np.all(np.diff(th) <= 0)
        assert np.max(np.abs(Z.T @ Z - np.eye(n))) < 1e-13
        assert np.max(np.abs(T @ Z - Z * th[None, :])) < 1e-15
    # a matrix that splits (a zero off-diagonal element) and one with repeated eigenvalues
    a, b = np.array([3.0, 1.0, 2.0, 2.0, 5.0]), np.array([0.5, 0.0, 0.0, 0.25, 0.0])
    th, Z = _capi.tridiag_eigen_all(a, b)
    T = np.diag(a) + np.diag(b[:4], 1) + np.diag(b[:4], -1)
    assert np.allclose(th, np.linalg.eigvalsh(T)[::-1], atol=1e-14) and np.max(np.abs(T @ Z - Z * th[None, :])) < 1e-14
