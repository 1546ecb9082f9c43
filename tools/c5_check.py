#!/usr/bin/env python
"""BASELINE config 5 shape (126 manifolds [2000, 15], 8-D subspace): stage-4 preparation on the device
(esper_manifold_prepare: normalisation + 28 conic fits per PD) next to the host (numpy) preparation, then the sweep."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import Rotation_Histograms as RH, _capi, synth  # noqa: E402

n_pd, dim = 126, 8
U0s = [synth.random_manifold(m=2000, ncol=15, seed=s) for s in range(n_pd)]
ctx = _capi.Context(0)
RH.run_many(U0s[:4], dim=dim, PCA=True, CTF=True, ctx=ctx)                     # warm-up
t0 = time.perf_counter()
host = [RH.prepare(u, dim=dim, PCA=True, CTF=True) for u in U0s[:8]]
t_host = (time.perf_counter() - t0) / 8 * n_pd
ctx.set_profiling(True); ctx.reset_profiling()
t0 = time.perf_counter()
preps = RH.prepare_many(U0s, dim=dim, PCA=True, CTF=True, ctx=ctx)
t_prep = time.perf_counter() - t0
t0 = time.perf_counter()
thetas = RH.sweep_many(preps, dim, ctx=ctx)
t_sweep = time.perf_counter() - t0
same = all(h['comboList'] == p['comboList'] and h['Rij_final'] == p['Rij_final'] and np.array_equal(h['U_init'], p['U_init'])
           for h, p in zip(host, preps))
th_host = RH.sweep_many(host, dim, ctx=ctx)
same_theta = all(np.array_equal(a, b) for a, b in zip(th_host, thetas[:8]))
print('C5 prepare: host %.2f s (extrapolated from 8 PDs) | device %.4f s wall, kernels %.3f ms in %d launches | sweep %.4f s wall, kernels %.3f ms'
      % (t_host, t_prep, ctx.kernel_ms('conic')[0], ctx.kernel_ms('conic')[1], t_sweep, ctx.kernel_ms('rot')[0]))
print('decisions identical to host: %s, thetas identical: %s, resident sweep: %s' % (same, same_theta, preps[0]['resident'] is not None))
