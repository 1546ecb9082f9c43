#!/usr/bin/env python
"""One PD (BASELINE config 2 shape) through stages 1-3, a few times -- the command profiled with ncu.
Usage: python tools/profile_step.py [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import GaussianBandwidth as GB, _capi, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0        # esper_dmap_config: 0 latency (resident), 1 throughput (streamed)
N, BOX = 2000, 250
stack = synth.synthetic_pd(pd=1, box=BOX, n_side=20, tau=5, snr=0.1).reshape(N, BOX * BOX)
ctx = _capi.Context(0)
ctx.upload_images(stack); ctx.sync()
ctx.dmap_config(mode)
logEps = np.arange(-100, 100.2, 0.2)
coef = 1. / (2. * np.exp(logEps))
a0 = np.array([[-0.2], [0.3], [-0.4], [0.1]])
for s in range(steps):
    ctx.dist_rms(None, n=N, P=BOX * BOX, download=False)
    L = ctx.eps_sweep(None, coef, 10.0, m=N)
    popt, resnorm, R2 = GB.fit(logEps, L, a0)
    eps = float(np.exp(-(popt[1] / popt[0])))
    val, vec, iters = ctx.dmap_embed(None, eps, k=10, m=N)
print('ok eps=%.6f iters=%d lam1=%.6f' % (eps, iters, np.sqrt(val[1])))
