#!/usr/bin/env python
"""One stage of the opt-in kernels against the default kernels on this GPU, at the benchmark shape, as ONE
JSON line.  bench.py runs the stages `sweep`, `prep`, `eig` and -- last, because a tensor-core / TMA kernel that has never
run is the one that can fault or wait on an mbarrier -- `gram` in separate subprocesses AFTER its measurements, each under a
time-out: a fault in a probe cannot touch the numbers already taken, and nothing measured here enters `value`.  It records
whether kernels that have never run on hardware work on this box and what they would gain.

    python tools/experimental_probe.py sweep|prep|eig|gram
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import _capi, synth  # noqa: E402

stage = sys.argv[1] if len(sys.argv) > 1 else 'sweep'
out = {'stage': stage, 'ok': False}
try:
    ctx = _capi.Context(0)
    X = synth.synthetic_pd(pd=5, box=250, n_side=20, tau=5, snr=0.1).reshape(2000, 62500)
    n, P = X.shape
    coef = 1. / (2. * np.exp(np.arange(-100, 100.2, 0.2)))
    reps = 5

    def timed(name, fn):
        fn()
        ctx.set_profiling(True); ctx.reset_profiling()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        ctx.sync()
        wall = (time.perf_counter() - t0) / reps * 1e3
        ms = ctx.kernel_ms(name)[0] / reps
        ctx.set_profiling(False)
        return r, ms, wall

    if stage == 'sweep':
        ctx.dist_rms(X, download=False)
        ctx.sweep_config(0)
        L0, ms0, _ = timed('sweep', lambda: ctx.eps_sweep(None, coef, 10.0, m=n))
        ctx.sweep_config(1)
        L1, ms1, _ = timed('sweep', lambda: ctx.eps_sweep(None, coef, 10.0, m=n))
        path, bins = ctx.sweep_info()
        diff = float(np.max(np.abs(L1 - L0)))
        out.update(ok=bool(path == 1 and diff < 1e-12), path=path, bins=bins, max_abs_diff_logsum=diff,
                   kernel_ms_general=round(ms0, 4), kernel_ms_moments=round(ms1, 4))
    elif stage == 'prep':
        ctx.prep_config(0)
        D0, ms0, _ = timed('prep', lambda: ctx.dist_rms(X if ctx.launch_count() == 0 else None, n=n, P=P))
        ctx.prep_config(1)
        D1, ms1, _ = timed('prep', lambda: ctx.dist_rms(None, n=n, P=P))
        cs = ctx.prep_info()
        off = ~np.eye(n, dtype=bool)
        rel = float(np.max(np.abs(D1 - D0)[off] / D0[off]))
        out.update(ok=bool(cs == 4 and rel < 1e-6 and np.array_equal(D1, D1.T)), cluster_size=cs, max_rel_diff_D=rel,
                   kernel_ms_one_cta=round(ms0, 4), kernel_ms_cluster=round(ms1, 4))
    elif stage == 'eig':
        ctx.dist_rms(X, download=False)
        ctx.dmap_config(1)
        (v1, _, it1), ms1, w1 = timed('lanczos', lambda: ctx.dmap_embed(None, 0.2834, k=10, m=n))
        ctx.dmap_config(2)
        (v2, _, it2), ms2, w2 = timed('lanczos', lambda: ctx.dmap_embed(None, 0.2834, k=10, m=n))
        rel = float(np.max(np.abs(v1 - v2) / v1))
        out.update(ok=bool(rel < 1e-9 and it2 <= it1), max_rel_diff_val=rel, steps_speculative=it1, steps_mode2=it2,
                   kernel_ms_speculative=round(ms1, 4), kernel_ms_mode2=round(ms2, 4), wall_ms_speculative=round(w1, 3), wall_ms_mode2=round(w2, 3))
    elif stage == 'gram':
        ctx.gram_config(0)
        D0, ms0, _ = timed('gram', lambda: ctx.dist_rms(X if ctx.launch_count() == 0 else None, n=n, P=P))
        ctx.gram_config(1)
        D1, ms1, _ = timed('gram', lambda: ctx.dist_rms(None, n=n, P=P))
        off = ~np.eye(n, dtype=bool)
        rel = float(np.max(np.abs(D1 - D0)[off] / D0[off]))
        out.update(ok=bool(ctx.gram_info() == 1 and rel < 2e-6), max_rel_diff_D=rel, kernel_ms_tf32=round(ms0, 4), kernel_ms_f16=round(ms1, 4))
    else:
        out['error'] = 'unknown stage'
except Exception as e:                                   # noqa: BLE001 -- the probe reports, it never raises
    out['error'] = '%s: %s' % (type(e).__name__, str(e)[:300])
print(json.dumps(out), flush=True)
