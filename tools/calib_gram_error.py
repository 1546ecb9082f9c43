import sys, numpy as np
sys.path.insert(0, '/root/repo')
from manifoldem_esper_b200 import _capi, synth
from oracle import esper_oracle as O
ctx = _capi.Context(0)
for name, stack in [('pristine', synth.synthetic_pd(pd=4, box=250, n_side=20, tau=1, snr=None)),
                    ('snr1', synth.synthetic_pd(pd=4, box=250, n_side=20, tau=1, snr=1.0)),
                    ('snr10', synth.synthetic_pd(pd=4, box=250, n_side=20, tau=1, snr=10.0))]:
    X = O.standardize_stack(stack)
    Dr = O.distances_direct(X)
    n = Dr.shape[0]; off = ~np.eye(n, dtype=bool)
    for sk in (0, 148, 400):
        ctx.dist_config(split_k=sk, refine_tau=0.0)
        D = ctx.dist_rms(stack, flags=_capi.DIST_STANDARDIZE | _capi.DIST_CENTER)
        rel = np.abs(D - Dr)[off] / Dr[off]
        d2rel = ((D * D - Dr * Dr)[off] / (Dr * Dr)[off])
        rho = 1 - (Dr * Dr)[off] / 2
        # implied relative Gram error: dD2/D2 = epsG * 2G/(D2 P), G/P ~ rho (uncentred) -> epsG ~ d2rel * D2 / (2 rho)
        print('%s split_k=%d: D range [%.3f, %.3f] max rel err D %.2e at D=%.3f; mean d2 bias %.2e; err for D>1: %.2e' % (
            name, sk, Dr[off].min(), Dr[off].max(), rel.max(), Dr[off][rel.argmax()], d2rel.mean(), rel[Dr[off] > 1.0].max() if (Dr[off] > 1).any() else -1), flush=True)
    # binned error vs D^2
    ctx.dist_config(split_k=0, refine_tau=0.0)
    D = ctx.dist_rms(stack, flags=_capi.DIST_STANDARDIZE | _capi.DIST_CENTER)
    d2 = (Dr * Dr)[off]; rel = np.abs(D - Dr)[off] / Dr[off]
    for lo, hi in [(0, .05), (.05, .2), (.2, .5), (.5, 1), (1, 1.6), (1.6, 4)]:
        m = (d2 >= lo) & (d2 < hi)
        if m.any(): print('   D^2 in [%.2f,%.2f): n=%d max rel err %.2e' % (lo, hi, m.sum(), rel[m].max()), flush=True)
