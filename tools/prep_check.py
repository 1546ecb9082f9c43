"""K0 at the benchmark shape: kernel time of the prep stage per PD (CUDA events) for both operand formats, and the distances
against the float64 oracle.  Usage: python tools/prep_check.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manifoldem_esper_b200 import _capi, synth
from oracle import esper_oracle as O
stack = synth.synthetic_pd(pd=1, box=250, n_side=20, tau=5, snr=0.1)
Dr = O.distances_gram_f64(O.standardize_stack(stack))
ctx = _capi.Context(0)
ctx.upload_images(stack.reshape(2000, -1)); ctx.sync()
for mode in (1, 0):
    ctx.gram_config(mode)
    D = ctx.dist_rms(None, n=2000, P=62500)
    ctx.set_profiling(True); ctx.reset_profiling()
    for _ in range(10):
        ctx.dist_rms(None, n=2000, P=62500, download=False)
    ctx.sync()
    ms, cnt = ctx.kernel_ms('prep')
    ctx.set_profiling(False)
    print('planes=%s: prep %.4f ms per PD (%d launches); max rel err of D vs the float64 oracle %.2e'
          % ('f16' if mode else 'tf32', ms / 10, cnt // 10, float(np.max(np.abs(D - Dr) / np.maximum(Dr, 1e-2)))), flush=True)
