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
from oracle import esper_oracle as O

SRC = os.path.join(ROOT, 'manifoldem_esper_b200', 'csrc', 'k2_sweep.cu')
EMU = os.path.join(ROOT, 'tests', 'emu')
LOG_EPS = np.arange(-100, 100.2, 0.2)


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
    assert res.returncode == 0, res.stderr[-2000:]
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
    L_np, bins_np = MSW.sweep_numpy(D, coef, 10.0)
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
    tab = MSW.accumulate_numpy(q, np.ones_like(q), pl, coef, 10.0)
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
        L_np, _ = MSW.sweep_numpy(Dm, coef, 10.0)
        assert r['status'] == 0 and np.allclose(r['logsum'], L_np, rtol=0, atol=1e-12), m
