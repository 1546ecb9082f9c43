#!/usr/bin/env python
"""The north-star target once: the full 126-PD tau=5, SNR 0.1 manifold set (N=2000 images of 250x250 px per PD) through
stages 1-4 with the output files, sharded over the GPUs of one node (PD p -> rank p mod W, no data-path collective).

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 tools/northstar_run.py [--pds 1-126] [--inflight 2] [--root /tmp/esper_tree]

Every PD: 400 pristine states (synth.pristine_states, PD index = in-plane rotation of the toy molecule) -> esper_tau_snr_stack on
the device (tau = 5 noisy copies at SNR 0.1; the 500 MB stack never exists on the host) -> esper_pd_embed_fused (distances,
epsilon sweep, tanh fit, Markov matrix, 16 leading eigenpairs) -> PD_%s_dist.npy / _val.npy / _vec.npy -> Rotation_Histograms.op
(dim 8) -> the three rotation files.  Rank 0 prints one JSON summary line; afterwards the consumer's loader
(3_Manifold_Binning/Manifold_Binning.py:88-139, restated in tests/consumer_contract.py) reads every PD's files."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from manifoldem_esper_b200 import DiffusionMaps, Rotation_Histograms, _capi, launcher, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pds', default='1-126')
    ap.add_argument('--inflight', type=int, default=2)
    ap.add_argument('--root', default='/tmp/esper_northstar')
    ap.add_argument('--dim', type=int, default=8)
    ap.add_argument('--box', type=int, default=250)
    args = ap.parse_args()
    rank, local_rank, world = launcher.dist_env()
    gather = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('gloo')

        def gather(local):
            out = [None] * world
            dist.all_gather_object(out, local)
            return out
    dm_dir = os.path.join(args.root, '1_Embedding', 'DM')
    rot_dir = os.path.join(args.root, '2_Manifold_Rotations')
    for d in (dm_dir, rot_dir):
        os.makedirs(d, exist_ok=True)
    box, P = args.box, args.box * args.box

    def make_work():
        ctx = _capi.Context(local_rank)

        def work(pd):
            info = {}
            t = time.time()
            prist = synth.pristine_states(box=box, n_side=20, pd=int(pd))
            info['states_s'] = time.time() - t; t = time.time()
            ctx.tau_snr_stack(prist, 5, 0.1, seed=1000 * int(pd), download=False)
            ctx.sync()
            info['stack_s'] = time.time() - t; t = time.time()
            np.random.seed(int(pd))
            info.update(DiffusionMaps.embed_stack_to_files(dm_dir, pd, stackPath=None, k=16, ctx=ctx, n=400 * 5, P=P))
            info['embed_s'] = time.time() - t; t = time.time()
            Rotation_Histograms.op(rot_dir, pd, dim=args.dim, PCA=False, ctx=ctx)
            info['rot_s'] = time.time() - t
            return info
        return work

    works = [make_work() for _ in range(max(1, args.inflight))]
    pds = launcher.parse_pds(args.pds)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t0 = time.time()
    recs = launcher.run_sharded(pds, works if len(works) > 1 else works[0], rank, world, gather, inflight=len(works))
    wall = time.time() - t0
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    wall_all = time.time() - t0
    if rank == 0:
        from consumer_contract import load_like_manifold_binning
        failed = [r for r in recs if not r.get('ok')]
        loaded, load_err = 0, []
        for r in recs:
            if r.get('ok'):
                try:
                    load_like_manifold_binning(args.root, r['PD'])
                    loaded += 1
                except Exception as e:                 # noqa: BLE001
                    load_err.append('%s: %s' % (r['PD'], str(e)[:80]))
        ok = [r for r in recs if r.get('ok')]
        agg = {k: float(np.sum([r.get(k, 0.0) for r in ok])) for k in ('states_s', 'stack_s', 'embed_s', 'rot_s')}
        print(json.dumps({'n_pd': len(recs), 'world': world, 'inflight': args.inflight, 'wall_s': wall_all, 'rank0_wall_s': wall,
                          'pd_per_s_with_files': len(ok) / wall_all if wall_all > 0 else None,
                          'failed': [(r['PD'], r.get('error', '')[:100]) for r in failed],
                          'stage_seconds_summed_over_pds': agg,
                          'lanczos_steps_mean': float(np.mean([r['lanczos_steps'] for r in ok])) if ok else None,
                          'eps_range': [float(np.min([r['eps'] for r in ok])), float(np.max([r['eps'] for r in ok]))] if ok else None,
                          'consumer_loader_ok': loaded, 'consumer_loader_errors': load_err[:5],
                          'bytes_written_GB': sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(args.root) for f in fs) / 1e9}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
