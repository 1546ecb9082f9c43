"""Step counts of the eigensolver schedules on the medium golden (N=400) at a tight tolerance, with the chain kernel's own report."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manifoldem_esper_b200 import _capi
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'stage23_medium.npz'))
ctx = _capi.Context(0)
tol = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-11
for mode in (3, 0):
    ctx.dmap_config(mode)
    if mode == 0: os.environ['ESPER_DEBUG_LANCZOS'] = '1'
    v, U, it = ctx.dmap_embed(g['D'], float(g['eps']), k=10, tol=tol)
    print('mode', mode, 'steps', it, flush=True)
