"""CPU dry run of bench.py's GPU arm: every Python line of run_ours() executes against a fake device layer (a stand-in
for _capi.Context that answers with oracle results on a small stack, and no-op CUDA streams / events), and the JSON line
is checked against the driver's contract.  This is a test of the HARNESS (flags, bookkeeping, JSON keys), not of any
kernel: nothing here is a product path, and the fake never leaves this file."""
import io
import json
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest

from conftest import ROOT  # noqa: F401
from oracle import esper_oracle as O


class FakeCtx:
    """Answers the calls bench.py makes on a _capi.Context; counts them like the library counts launches."""

    def __init__(self, device, stream=None):
        self.n_launch = 0
        self.prof = 0
        self.gram_calls = 0
        self.cfg = {}
        self.stack = None

    def upload_images(self, imgs):
        self.stack = np.asarray(imgs)

    def sync(self):
        pass

    def dmap_config(self, mode=0):
        self.cfg['dmap'] = mode

    def gram_config(self, mode=0):
        self.cfg['gram'] = mode

    def close(self):
        pass

    def pd_embed_fused(self, imgs, logEps, a0, k=16, n=None, P=None, out=None, **kw):
        from manifoldem_esper_b200 import _capi
        D = self.dist_rms(imgs, n=n, P=P, download=True, out=out)
        L = self.eps_sweep(None, 1. / (2. * np.exp(np.asarray(logEps))), 10.0, m=n)
        popt, resnorm, R2, info, _ = _capi.tanh_fit(logEps, L, a0)
        eps = float(np.exp(-(popt[1] / popt[0])))
        val, vec, it = self.dmap_embed(None, eps, k=k, m=n)
        r = dict(logSumWij=L, popt=popt, resnorm=resnorm, R2=R2, eps=eps, slope=popt[0] * popt[2], val=val, vec=vec, iters=it)
        if out is not None:
            r['D'] = D
        return r

    def sweep_config(self, mode=0):
        self.cfg['sweep'] = mode

    def set_profiling(self, on=True):
        self.prof = int(on)

    def reset_profiling(self):
        self.gram_calls = 0

    def kernel_ms(self, name):
        return (1.1 * self.gram_calls, self.gram_calls) if name == 'gram' else (0.2 * self.gram_calls, self.gram_calls)

    def launch_count(self):
        return self.n_launch

    def dist_rms(self, imgs=None, n=None, P=None, flags=7, download=True, out=None):
        st = self.stack if imgs is None else np.asarray(imgs)
        side = int(round(np.sqrt(st.shape[1])))
        self.D = O.distances_gram_f64(O.standardize_stack(st.reshape(st.shape[0], side, side)))
        self.n_launch += 6
        if self.prof:
            self.gram_calls += 1
        if out is not None:
            out[...] = self.D
        return self.D if download else None

    def eps_sweep(self, D, coef, thr, m=None):
        self.n_launch += 2
        return O.eps_sweep_logsum(self.D, -np.log(2.0 * np.asarray(coef)), sequential=False)

    def dmap_embed(self, D, eps, k=16, tol=0.0, max_iter=0, m=None):
        self.n_launch += 5
        val, U = O.dmap_embed(self.D, eps)
        return val[:k], U[:, :k], 57


class _Evt:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 12.5


class _Stream:
    cuda_stream = 0

    def wait_event(self, e):
        pass


@pytest.mark.parametrize('flags', [[], ['--gram-tf32', '--sweep-general', '--eig-mode', '2', '--scipy-fit']])
def test_bench_gpu_arm_dry_run(monkeypatch, flags):
    import torch
    import bench
    from manifoldem_esper_b200 import _capi, synth

    n_side, box, tau = 6, 16, 2
    monkeypatch.setattr(bench, 'N_IMG', n_side * n_side * tau)
    monkeypatch.setattr(bench, 'BOX', box)
    monkeypatch.setattr(bench, 'P', box * box)
    monkeypatch.setattr(bench, 'TAU', tau)
    real_pd = synth.synthetic_pd
    monkeypatch.setattr(synth, 'synthetic_pd', lambda pd=1, box=box, n_side=20, tau=tau, snr=0.1: real_pd(pd=pd, box=box, n_side=6, tau=tau, snr=snr))
    monkeypatch.setattr(_capi, 'Context', FakeCtx)
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self)
    real_empty = torch.empty
    monkeypatch.setattr(torch, 'empty', lambda *a, **k: real_empty(*a, **{kk: vv for kk, vv in k.items() if kk != 'device'}))
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, 'tensor', lambda *a, **k: real_tensor(*a, **{kk: vv for kk, vv in k.items() if kk != 'device'}))
    fake_cuda = types.SimpleNamespace(set_device=lambda d: None, Stream=_Stream, Event=_Evt, synchronize=lambda: None)
    monkeypatch.setattr(torch, 'cuda', fake_cuda)
    monkeypatch.setattr(bench, 'ClockSampler', lambda idx: types.SimpleNamespace(
        start=lambda: None, stop=lambda *a: {'sm_mhz': 1965.0, 'sm_max_mhz': 1965.0, 'reasons': [], 'samples': 3}))
    monkeypatch.setattr(bench, 'cpu_sample', lambda stack, **kw: (200.0, {'dist_loop_s': 176.0, 'sweep_s': 23.0, 'kernel_svd_s': 1.0,
                                                                         'vectorised_dist_s': 0.8, 'vectorised_total_s': 25.0}))
    for var in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        monkeypatch.delenv(var, raising=False)
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--steps', '4', '--warmup', '3', '--inflight', '2', '--no-extras'] + flags)
    buf = io.StringIO()
    with redirect_stdout(buf):
        assert bench.main() == 0
    lines = [ln for ln in buf.getvalue().splitlines() if ln.startswith('{')]
    assert len(lines) == 1                                           # ONE JSON line on rank 0
    d = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'):
        assert key in d, key
    assert d['unit'] == 'PD/s' and d['n_gpus'] == 1 and d['steps'] == 4 and d['warmup'] == 3 and d['vs_baseline'] is None
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['data'] == 'synthetic' and 'workload' in d['config']
    assert abs(d['value'] - 4 / 12.5e-3) < 1e-9 and abs(d['ms_per_step'] - 12.5 / 4) < 1e-12
    for key in ('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'):
        assert key in d['e2e'], key
    assert d['e2e']['h2d_bytes_per_step'] == bench.N_IMG * bench.P * 4
    for key in ('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'):
        assert key in d['roofline'], key
    assert d['roofline']['bound'] == 'tensor' and d['roofline']['launches_timed'] == 4 and abs(d['roofline']['avg_launch_ms'] - 1.1) < 1e-9
    assert d['roofline']['launches_pipelined'] == 4 and abs(d['roofline']['avg_launch_ms_pipelined'] - 1.1) < 1e-9
    assert abs(d['roofline']['frac'] - d['roofline']['achieved'] / d['roofline']['peak']) < 1e-12
    for key in ('value', 'unit', 'cores', 'kind', 'sample'):
        assert key in d['cpu_baseline'], key
    assert d['gpu_launches'] == 4 * 13                               # counted over the timed region only
    assert d['config']['kernels'] == {'gram_f16': not flags, 'sweep_moments': not flags}
    assert d['dtype'] == ('tf32x3+f64' if flags else 'f16x3+f64')
    assert ('curve_fit' in d['config']['tanh_fit']) == bool(flags) and ('esper_pd_embed_fused' in d['config']['api']) == (not flags)
    for key in ('frac_issued', 'frac_in_step_vs_sustained', 'peak_sustained', 'traffic_note'):
        assert key in d['roofline'], key
    assert 'frac_of_h2d_floor' in d['e2e'] and 'c5_rot' not in d and 'experimental_probe' not in d
