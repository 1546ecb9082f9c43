"""Default eigensolver schedule against the resident barrier kernel over bandwidth regimes (nearly rank-one Ms ... nearly diagonal Ms) and
duplicated images on the medium golden: eigenvalues, leading eigenvectors, step counts, fallbacks.
    python tools/eps_regimes_probe.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manifoldem_esper_b200 import _capi
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'stage23_medium.npz'))
ctx = _capi.Context(0)
D0 = g['D']; eps0 = float(g['eps'])
cases = [('medium', D0, eps0 * f) for f in (1e-3, 1e-2, 0.1, 1.0, 10.0, 1e2, 1e4)]
Dd = np.repeat(np.repeat(D0[:150, :150], 2, axis=0), 2, axis=1)            # every image twice
cases += [('duplicated', Dd, eps0 * f) for f in (0.1, 1.0, 10.0)]
Dn = D0[:200, :200].copy(); Dn[1, :] = Dn[0, :] * (1 + 1e-9); Dn[:, 1] = Dn[1, :]; Dn[1, 1] = 0; Dn[0, 1] = Dn[1, 0] = 1e-9   # a near-duplicate pair
cases += [('near-duplicate', Dn, eps0)]
for name, D, eps in cases:
    out = {}
    for mode in (3, 0):
        ctx.dmap_config(mode)
        f0 = ctx.stat('lanczos_fallbacks')
        try:
            v, U, it = ctx.dmap_embed(np.ascontiguousarray(D), eps, k=8)
            out[mode] = (v, U, it, ctx.stat('lanczos_fallbacks') - f0, None)
        except Exception as e:                                            # noqa: BLE001
            out[mode] = (None, None, -1, ctx.stat('lanczos_fallbacks') - f0, str(e)[:80])
    ctx.dmap_config(0)
    (v3, U3, it3, f3, e3), (v0, U0, it0, fb0, e0) = out[3], out[0]
    if v3 is None or v0 is None:
        print('%-15s eps %.3e: mode 3 %s | mode 0 %s' % (name, eps, e3 or 'ok', e0 or 'ok'), flush=True)
        continue
    dv = float(np.max(np.abs(v0 - v3) / np.maximum(v3, 1e-300)))
    du = max(min(np.abs(U0[:, i] - U3[:, i]).max(), np.abs(U0[:, i] + U3[:, i]).max()) for i in range(4))
    print('%-15s eps %.3e: steps %3d / %3d  fallbacks %d  |dval| %.2e  |dvec(0..3)| %.2e  lambda_1..3 %s' % (name, eps, it3, it0, fb0, dv, du, np.sqrt(v3[1:4])), flush=True)
