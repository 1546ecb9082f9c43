#!/usr/bin/env python
"""One PD (BASELINE config 2 shape) through stages 1-3 in one library call, a few times -- the command profiled with ncu.
Usage: python tools/profile_step.py [steps] [eig_mode]     eig_mode: esper_dmap_config (0 chain kernel, 1 streamed, 3 resident)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import _capi, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
N, BOX = 2000, 250
stack = synth.synthetic_pd(pd=1, box=BOX, n_side=20, tau=5, snr=0.1).reshape(N, BOX * BOX)
ctx = _capi.Context(0)
ctx.upload_images(stack); ctx.sync()
ctx.dmap_config(mode)
logEps = np.arange(-100, 100.2, 0.2)
a0 = np.array([-0.2, 0.3, -0.4, 0.1])
for s in range(steps):
    r = ctx.pd_embed_fused(None, logEps, a0, k=10, n=N, P=BOX * BOX)
print('ok eps=%.6f iters=%d lam1=%.6f' % (r['eps'], r['iters'], np.sqrt(r['val'][1])))
