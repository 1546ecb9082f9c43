import sys, os, time, numpy as np
sys.path.insert(0, '/root/repo')
from manifoldem_esper_b200 import _capi, synth, GaussianBandwidth
from oracle import esper_oracle as O
stack = synth.synthetic_pd(pd=7, box=250, n_side=20, tau=5, snr=0.1)
D_ref = O.distances_gram_f64(O.standardize_stack(stack))
LOG_EPS = np.arange(-100, 100.2, 0.2)
L_ref = O.eps_sweep_logsum(D_ref, LOG_EPS, sequential=False)
popt, _, _ = GaussianBandwidth.fit(LOG_EPS, L_ref, np.array([[-0.2], [0.3], [-0.4], [0.1]]))
eps = float(np.exp(-(popt[1] / popt[0])))
val_r, U_r = O.dmap_embed(D_ref, eps)
ctx = _capi.Context(0)
D = ctx.dist_rms(stack)
val, vec, it = ctx.dmap_embed(None, eps, k=10, m=2000)
off = ~np.eye(2000, dtype=bool)
print('ESPER_DRAIN_KB=%s D err %.2e d2 bias %.2e val rel err %s' % (os.environ.get('ESPER_DRAIN_KB'), np.max(np.abs(D-D_ref)[off]/D_ref[off]), np.mean(((D*D-D_ref*D_ref)[off])/(D_ref*D_ref)[off]), ' '.join('%.1e' % x for x in np.abs(val-val_r[:10])/val_r[:10])))
ctx.set_profiling(True); ctx.reset_profiling()
for _ in range(5): ctx.dist_rms(None, n=2000, P=62500, download=False)
ctx.sync(); print('   gram ms', ctx.kernel_ms('gram')[0]/5)
