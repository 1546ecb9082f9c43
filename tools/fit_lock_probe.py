#!/usr/bin/env python
"""Stage 2b under several host threads: scipy's curve_fit (the reference's call, holds the interpreter lock) against
esper_tanh_fit (the same MINPACK algorithm in C, called through ctypes, lock released).  No GPU needed.

    python tools/fit_lock_probe.py [sweep.npy]

The sweep is a logSumWij curve of 1001 points (default: the golden of the N = 400 reference run).  Prints wall time per fit
for 1, 2, 4 and 8 threads; bench.py runs 4 pipeline threads per GPU.
"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import GaussianBandwidth as GB  # noqa: E402

log_eps = np.arange(-100, 100.2, 0.2)
L = np.load(sys.argv[1]) if len(sys.argv) > 1 else np.load(os.path.join(ROOT, 'tests', 'golden', 'stage23_medium.npz'))['logSumWij']
a0 = np.array([[-0.2], [0.3], [-0.4], [0.1]])
reps = 40
print('threads  curve_fit [ms/fit]  esper_tanh_fit [ms/fit]')
for nt in (1, 2, 4, 8):
    row = []
    for native in (False, True):
        def work():
            for _ in range(reps):
                GB.fit(log_eps, L, a0, native=native)
        th = [threading.Thread(target=work) for _ in range(nt)]
        t = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        row.append((time.perf_counter() - t) / (reps * nt) * 1e3)
    print('%7d  %18.2f  %23.2f' % (nt, row[0], row[1]))
