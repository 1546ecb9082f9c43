#!/usr/bin/env python
"""The 128x256-tile Gram kernel (`gram_wide_kernel`, default) against the 128x128 kernel (ESPER_GRAM_WIDE=0).

    gpurun --timeout 300 -- 'timeout 120 python tools/gram_wide_check.py [quick]'

Compares the distances of both kernels with the float64 oracle on small, odd and full-size shapes (quick: the full-size
case is compared with the 128x128 kernel instead), then times both at the benchmark shape (N = 2000, 250x250).
First run (round 1): parity 1.3e-7 on all shapes, bit-identical to the 128x128 kernel at full size, 1.107 vs 1.259 ms.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import _capi, synth  # noqa: E402
from oracle import esper_oracle as O  # noqa: E402


def run(ctx, X, wide):
    os.environ['ESPER_GRAM_WIDE'] = '1' if wide else '0'
    return ctx.dist_rms(X)


ctx = _capi.Context(0)
rng = np.random.default_rng(0)
ok = True
quick = len(sys.argv) > 1 and sys.argv[1] == 'quick'          # random full-size stack, full size checked against the default kernel
shapes = [(300, 4096), (640, 128 * 128), (2000, 62500)] if quick else [(129, 64), (256, 1000), (300, 4096), (640, 128 * 128), (2000, 62500)]
for n, P in shapes:
    if n == 2000 and not quick:
        X = synth.synthetic_pd(pd=3, box=250, n_side=20, tau=5, snr=0.1).reshape(2000, 62500)
    else:
        X = rng.standard_normal((n, P), dtype=np.float32)
    if n == 2000 and quick:
        Dr = run(ctx, X, False)
    else:
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
        ctx.dist_rms(None, n=2000, P=62500, download=False)
    ctx.sync()
    print('%s gram kernel: %.3f ms' % ('wide 128x256' if wide else 'default 128x128', ctx.kernel_ms('gram')[0] / 5), flush=True)
print('PARITY OK' if ok else 'PARITY FAILED')
