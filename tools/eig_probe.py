"""Eigensolver schedules on the benchmark matrix (N=2000, 250x250, SNR 0.1): steps, kernel and wall time per mode, agreement.
    python tools/eig_probe.py [n_side=20] [k=10]
mode 0 = barrier-free chain kernel (default), 3 = resident barrier kernel, 1 = streamed barrier kernel."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manifoldem_esper_b200 import _capi, synth

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 20
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
stack = synth.synthetic_pd(pd=1, box=250, n_side=n_side, tau=5, snr=0.1)
n = stack.shape[0]
ctx = _capi.Context(0)
ctx.dist_rms(stack, download=False)
eps = 0.2834
ref = None
for mode in (3, 0, 1, 0):
    ctx.dmap_config(mode)
    ctx.dmap_embed(None, eps, k=k, m=n)                       # warm-up (attributes, allocations)
    ctx.set_profiling(True); ctx.reset_profiling()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        val, vec, it = ctx.dmap_embed(None, eps, k=k, m=n)
    wall = (time.perf_counter() - t0) / reps * 1e3
    ms, cnt = ctx.kernel_ms('lanczos')
    mk, _ = ctx.kernel_ms('markov')
    ctx.set_profiling(False)
    if ref is None:
        ref = (val, vec)
    dv = float(np.max(np.abs(val - ref[0]) / ref[0]))
    du = max(min(np.abs(vec[:, i] - ref[1][:, i]).max(), np.abs(vec[:, i] + ref[1][:, i]).max()) for i in range(min(k, 7)))
    print('mode %d: steps %d  lanczos kernels %.3f ms (%d launches)  markov %.3f ms  wall %.3f ms  |dval| %.2e  |dvec(0..6)| %.2e  fallbacks %d'
          % (mode, it, ms / reps, cnt // reps, mk / reps, wall, dv, du, ctx.stat('lanczos_fallbacks')), flush=True)
os.environ['ESPER_DEBUG_LANCZOS'] = '1'
ctx.dmap_config(0)
ctx.dmap_embed(None, eps, k=k, m=n)
