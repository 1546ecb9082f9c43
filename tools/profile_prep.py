#!/usr/bin/env python
"""The kernels added for SURVEY 8f-3 / 8f-4 at the benchmark shapes -- the command profiled with ncu.
  tau/SNR stack generation (400 states x tau 5, 250x250), stage-4 preparation (126 manifolds [2000, 15], dim 8),
  bin accumulation (2000 images into 20 bins).  Prints CUDA-event kernel times."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import Rotation_Histograms as RH, _capi, synth  # noqa: E402

ctx = _capi.Context(0)
pr = synth.pristine_states(box=250, n_side=20, pd=1)
U0s = [synth.random_manifold(m=2000, ncol=15, seed=s) for s in range(126)]
rng = np.random.default_rng(0)
order = rng.permutation(2000).astype(np.int32)
offsets = np.arange(0, 2001, 100, dtype=np.int32)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx.set_profiling(True)
for r in range(reps):
    ctx.reset_profiling()
    ctx.tau_snr_stack(pr, 5, 0.1, seed=7, download=False)
    t_gen = ctx.kernel_ms('dataprep')[0]
    D = ctx.dist_rms(None, n=2000, P=62500, download=False)
    ctx.reset_profiling()
    sums = ctx.bin_accumulate(None, order, offsets, n=2000, P=62500)
    t_bin = ctx.kernel_ms('dataprep')[0]
    preps = RH.prepare_many(U0s, dim=8, PCA=True, CTF=True, ctx=ctx)
    t_conic = ctx.kernel_ms('conic')[0]
print('tau_snr_stack (2000 x 250^2): %.3f ms | bin_accumulate: %.3f ms (%.0f GB/s of image reads) | manifold_prepare (126 PDs): %.3f ms'
      % (t_gen, t_bin, 2000 * 62500 * 4 / (t_bin * 1e-3) / 1e9, t_conic))
