#!/usr/bin/env python
"""BASELINE config 4 on ONE GPU (N=50000 images of 128x128 px as a single row block): the command profiled with ncu for the
HBM-bound kernels of the row-block path (symv_rows_kernel streams the 20 GB Ms block once per Lanczos step).
Usage: python tools/profile_c4.py [n=50000] [max_steps=40]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manifoldem_esper_b200 import _capi, rowblock, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
max_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tau = n // 400
n = 400 * tau
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
dist.init_process_group('nccl', init_method='tcp://127.0.0.1:29631', rank=0, world_size=1, device_id=dev)
ctx = _capi.Context(0)
ctx.dist_config(split_k=8)
ctx.tau_snr_stack(synth.pristine_states(box=128, n_side=20, pd=1), tau, 0.1, seed=1000, download=False)
ctx.sync()
be = rowblock.CudaBackend(ctx, None, n, 128 * 128, dev, fused=False)
comm = rowblock.TorchComm(dev, stream=be.stream)
t0 = time.perf_counter()
try:
    r = rowblock.embed_oversized(be, comm, n, k=10, eps=0.2834, max_steps=max_steps, check_every=max_steps)
    print('converged in %d steps' % r['steps'])
except RuntimeError as e:                                     # max_steps is a cap for the profile, not a convergence budget
    print('stopped:', e)
torch.cuda.synchronize()
print('n=%d: %.2f s' % (n, time.perf_counter() - t0))
ctx.set_profiling(True); ctx.reset_profiling()
q = torch.randn(n, dtype=torch.float64, device=dev)
w = torch.zeros(n, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for _ in range(10):
    ctx.symv_rows_d(n, n, q.data_ptr(), w.data_ptr())
ctx.sync()
ms, cnt = ctx.kernel_ms('lanczos')
print('symv_rows_kernel: %.3f ms per launch, %.1f GB/s of Ms' % (ms / cnt, n * n * 8 / (ms / cnt * 1e-3) / 1e9))
dist.destroy_process_group()
