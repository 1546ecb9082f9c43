"""CTF branch at BASELINE config-2 size (N=2000 images of 250x250, per-image defocus): kernel times and a parity check of a
random sub-block of pairs and of a few Wiener images against the oracle.  Run on a B200: python tools/ctf_full.py [n] [box]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manifoldem_esper_b200 import _capi, synth          # noqa: E402
from oracle import esper_oracle as O                    # noqa: E402

n_side = 20 if len(sys.argv) < 2 else int(sys.argv[1])
box = 250 if len(sys.argv) < 3 else int(sys.argv[2])
t = time.time()
st, pa = synth.ctf_stack(pd=4, box=box, n_side=n_side, tau=5, snr=0.1)
n = st.shape[0]
print('stack %r generated in %.1f s' % (st.shape, time.time() - t))
ctx = _capi.Context(0)
D, W = ctx.ctf_dist(st, pa)                              # warm-up (allocations)
ctx.set_profiling(True)
reps = 3
for _ in range(reps):
    t = time.time(); ctx.ctf_dist(st, pa, download=False, wiener=False); ctx.sync(); t_nw = time.time() - t
NAMES = ['ctf_stats', 'ctf_gen', 'ctf_rows', 'ctf_cols', 'ctf_irows', 'gram', 'dist_finalize', 'refine']
print('kernel ms per call (no wiener):', {k: round(ctx.kernel_ms(k)[0] / reps, 3) for k in NAMES}, 'wall %.1f ms' % (t_nw * 1e3))
ctx.reset_profiling()
for _ in range(reps):
    ctx.ctf_dist(st, pa, download=False, wiener=True)
print('kernel ms per call (with wiener):', {k: round(ctx.kernel_ms(k)[0] / reps, 3) for k in NAMES})

rng = np.random.default_rng(1)
sub = np.sort(rng.choice(n, size=48, replace=False))
Dr = O.distances_ctf_direct(st[sub], pa[sub])
off = ~np.eye(len(sub), dtype=bool)
print('sub-block max rel err of D: %.2e' % np.max(np.abs(D[np.ix_(sub, sub)] - Dr)[off] / Dr[off]))
# Wiener images: S = sum over ALL images of CTF^2
S = np.zeros((box, box))
for i in range(n):
    S += O.ctf_filters(box, *pa[i])[0] ** 2
err = 0.0
for i in sub[:6]:
    tmp = st[i]; tmp = (tmp - tmp.mean()) / tmp.std()
    C, G = O.ctf_filters(box, *pa[i])
    img = np.fft.ifft2(np.fft.fft2(tmp.astype(np.float64)) * G).real
    w = np.fft.ifft2(np.fft.fft2(img) * (C / -(S + 0.2))).real.astype(np.float32)
    err = max(err, np.max(np.abs(W[i] - w)) / np.abs(w).max())
print('wiener max err relative to image max: %.2e' % err)
print('D range [%.1f, %.1f], symmetric %s' % (D[~np.eye(n, dtype=bool)].min(), D.max(), np.array_equal(D, D.T)))
