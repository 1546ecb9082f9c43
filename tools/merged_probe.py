"""Merged (one exchange per step) against the two-exchange cluster schedule of the chain eigensolver on the benchmark matrix
(N=2000, 250x250, SNR 0.1): steps, kernel and wall time, agreement with the resident barrier kernel, reproducibility of the
step count, timeline of one step.
    python tools/merged_probe.py [n_side=20] [k=10] [trace_step=100]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manifoldem_esper_b200 import _capi, synth

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 20
k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
os.environ['ESPER_TRACE_STEP'] = sys.argv[3] if len(sys.argv) > 3 else '100'
stack = synth.synthetic_pd(pd=1, box=250, n_side=n_side, tau=5, snr=0.1)
n = stack.shape[0]
ctx = _capi.Context(0)
ctx.dist_rms(stack, download=False)
eps = 0.2834
ctx.dmap_config(3)
ref = ctx.dmap_embed(None, eps, k=k, m=n)
print('reference (resident barrier kernel): steps %d' % ref[2], flush=True)
ctx.dmap_config(0)
for merged in (1, 0, 1):
    os.environ['ESPER_CHAIN_MERGED'] = str(merged)
    os.environ.pop('ESPER_DEBUG_LANCZOS', None)
    c0 = ctx.stat('cluster_launches'); m0 = ctx.stat('merged_launches'); f0 = ctx.stat('lanczos_fallbacks')
    ctx.dmap_embed(None, eps, k=k, m=n)                       # warm-up (attributes, allocations)
    ctx.set_profiling(True); ctx.reset_profiling()
    reps = 8
    its = []
    t0 = time.perf_counter()
    for _ in range(reps):
        val, vec, it = ctx.dmap_embed(None, eps, k=k, m=n)
        its.append(it)
    wall = (time.perf_counter() - t0) / reps * 1e3
    ms, cnt = ctx.kernel_ms('lanczos')
    ctx.set_profiling(False)
    dv = float(np.max(np.abs(val - ref[0]) / ref[0]))
    du = max(min(np.abs(vec[:, i] - ref[1][:, i]).max(), np.abs(vec[:, i] + ref[1][:, i]).max()) for i in range(min(k, 7)))
    print('merged %d: steps %s  lanczos kernels %.3f ms (%d launches)  wall %.3f ms  |dval| %.2e  |dvec(0..6)| %.2e  cluster launches %d  merged launches %d  fallbacks %d'
          % (merged, sorted(set(its)), ms / reps, cnt // reps, wall, dv, du, ctx.stat('cluster_launches') - c0, ctx.stat('merged_launches') - m0, ctx.stat('lanczos_fallbacks') - f0), flush=True)
    os.environ['ESPER_DEBUG_LANCZOS'] = '1'
    ctx.dmap_embed(None, eps, k=k, m=n)
    sys.stderr.flush()
