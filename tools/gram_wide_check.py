#!/usr/bin/env python
"""First hardware check of the EXPERIMENTAL 128x256-tile Gram kernel (`gram_wide_kernel`, ESPER_GRAM_WIDE=1).

It has only been compiled and desk-checked so far (no GPU minutes were left when it was written), so run it under a
short timeout -- a barrier mismatch in a tcgen05/TMA pipeline shows up as a kernel that never ends:

    gpurun --timeout 300 -- 'timeout 120 python tools/gram_wide_check.py'

Compares distances of the default and the wide kernel with the float64 oracle on small, odd and full-size shapes, then
times both at the benchmark shape (N = 2000, 250x250).  Expected if the model in profiles/r1_gram_chain_cut_cost.md
holds: ~0.95 ms against ~1.23 ms.  If it is correct and faster, make it the default for the full symmetric path
(run_gram, k1_gram.cu) and re-run `pytest -m gpu`.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import _capi, synth  # noqa: E402
from oracle import esper_oracle as O  # noqa: E402


def run(ctx, X, wide):
    if wide:
        os.environ['ESPER_GRAM_WIDE'] = '1'
    else:
        os.environ.pop('ESPER_GRAM_WIDE', None)
    return ctx.dist_rms(X)


ctx = _capi.Context(0)
rng = np.random.default_rng(0)
ok = True
for n, P in [(129, 64), (256, 1000), (300, 4096), (640, 128 * 128), (2000, 62500)]:
    X = synth.synthetic_pd(pd=3, box=250, n_side=20, tau=5, snr=0.1).reshape(2000, 62500) if n == 2000 else \
        rng.standard_normal((n, P)).astype(np.float32)
    Dr = O.distances_gram_f64(O.standardize_stack(X.reshape(n, 1, P)))
    errs = []
    for wide in (False, True):
        D = run(ctx, X, wide)
        errs.append(float(np.max(np.abs(D - Dr) / np.maximum(Dr, 1e-2))))
        ok &= errs[-1] < 1e-5 and np.array_equal(D, D.T) and not np.diag(D).any()
    print('n=%d P=%d  rel err default %.2e  wide %.2e' % (n, P, errs[0], errs[1]), flush=True)
ctx.set_profiling(True)
for wide in (False, True):
    run(ctx, X, wide)
    ctx.reset_profiling()
    for _ in range(5):
        if wide:
            os.environ['ESPER_GRAM_WIDE'] = '1'
        ctx.dist_rms(None, n=2000, P=62500, download=False)
    ctx.sync()
    print('%s gram kernel: %.3f ms' % ('wide 128x256' if wide else 'default 128x128', ctx.kernel_ms('gram')[0] / 5), flush=True)
print('PARITY OK' if ok else 'PARITY FAILED')
