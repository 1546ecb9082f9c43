"""Shard the per-PD stages over the GPUs of one node (replaces the serial shell loops
1_Embedding/DM/1_Dist_Batch.sh:6-11, 2_DM_Batch.sh:6-11, 2_Manifold_Rotations/1_RotHist_Batch.sh:6-11).

PDs are independent, so the data path has no collective: rank r takes PDs r, r+W, r+2W, ...
`torch.distributed` is used only to gather the per-PD status records on rank 0.

    torchrun --nproc-per-node 8 -m manifoldem_esper_b200.launcher --root <tree> --stages dist,dm,rot --pds 1-126

`--root` is a directory laid out like the reference tree (1_Embedding/DM, 2_Manifold_Rotations, ...).
A failing PD is recorded and skipped (the reference's shell loop behaves the same way); the exit
status is non-zero if any PD failed.  `--skip-existing` resumes an interrupted run.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import traceback


def parse_pds(spec: str):
    """'1-5,9,12-13' -> ['001','002',...] zero-padded like the reference's batch scripts."""
    out = []
    for part in spec.split(','):
        part = part.strip()
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            out += list(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return ['%03d' % p for p in out]


def shard(items, rank: int, world: int):
    """Static round-robin shard: the items of `rank` (no communication needed)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('bad rank/world %s/%s' % (rank, world))
    return list(items[rank::world])


def dist_env():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def _run_one(pd, work_fn, rank):
    t0 = time.time()
    rec = {'PD': pd, 'rank': rank}
    try:
        extra = work_fn(pd)
        rec['ok'] = True
        if isinstance(extra, dict):
            rec.update(extra)
    except Exception as e:                      # noqa: BLE001 -- isolate one PD's failure
        rec.update(ok=False, error='%s: %s' % (type(e).__name__, e), trace=traceback.format_exc(limit=4))
    rec['seconds'] = time.time() - t0
    return rec


def run_sharded(pds, work_fn, rank=0, world=1, gather=None, inflight=1):
    """Run work_fn(pd) for this rank's PDs with per-PD error isolation.

    `work_fn` is a callable, or -- with `inflight` > 1 -- a list of `inflight` callables (one per host thread, each
    with its own GPU context and stream): the threads take this rank's PDs from a shared queue, so the file I/O,
    host-to-device copy and host-side fit of one PD overlap the kernels of the others (the library calls release the
    interpreter lock).  Records come back in PD order either way.

    Returns this rank's records, or on rank 0 all ranks' records when `gather` (a callable
    taking the local list and returning the list of all ranks' lists, e.g. built on
    torch.distributed.all_gather_object) is given.
    """
    mine = shard(pds, rank, world)
    fns = list(work_fn) if isinstance(work_fn, (list, tuple)) else [work_fn]
    if inflight < 1 or (inflight > 1 and len(fns) != inflight):
        raise ValueError('run_sharded: inflight=%s needs that many work functions (got %d)' % (inflight, len(fns)))
    if inflight == 1:
        records = [_run_one(pd, fns[0], rank) for pd in mine]
    else:
        import queue
        import threading
        todo = queue.Queue()
        for i, pd in enumerate(mine):
            todo.put((i, pd))
        slots = [None] * len(mine)

        def worker(fn):
            while True:
                try:
                    i, pd = todo.get_nowait()
                except queue.Empty:
                    return
                slots[i] = _run_one(pd, fn, rank)
        threads = [threading.Thread(target=worker, args=(fn,)) for fn in fns]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        records = slots
    if gather is not None:
        all_lists = gather(records)
        if rank == 0:
            merged = [r for lst in all_lists for r in lst]
            merged.sort(key=lambda r: r['PD'])
            return merged
    return records


def _stage_fns(root, stages, dim, pca, k, skip_existing, device, ctf=True):
    from . import PCA, Distances, DiffusionMaps, Rotation_Histograms, _capi
    ctx = _capi.Context(device)
    dm_dir = os.path.join(root, '1_Embedding', 'DM')
    rot_dir = os.path.join(root, '2_Manifold_Rotations')
    pca_dir = os.path.join(root, '1_Embedding', 'PCA')

    def work(pd):
        info = {}
        if 'dist' in stages:
            out = os.path.join(dm_dir, 'Data_Distances', 'PD_%s_dist.npy' % pd)
            if not (skip_existing and os.path.exists(out)):
                t = time.time(); Distances.op(dm_dir, pd, CTF=ctf, ctx=ctx); info['dist_s'] = time.time() - t
        if 'embed' in stages:                        # dist + dm of the non-CTF branch in one library call (esper_pd_embed_fused)
            out = os.path.join(dm_dir, 'Data_Manifolds', 'PD_%s_vec.npy' % pd)
            if not (skip_existing and os.path.exists(out)):
                data = os.path.join(root, '0_Data_Inputs', 'CTF5k15k_SNRpt1_ELS_2D')
                path = os.path.join(data, 'PD%s_SNR_tau5_stack.mrcs' % pd)
                if not os.path.exists(path):
                    path = os.path.join(data, 'PD_%s.mrcs' % pd)
                t = time.time(); info.update(DiffusionMaps.embed_stack_to_files(dm_dir, pd, stackPath=path, k=k, ctx=ctx)); info['embed_s'] = time.time() - t
        if 'dm' in stages:
            out = os.path.join(dm_dir, 'Data_Manifolds', 'PD_%s_vec.npy' % pd)
            if not (skip_existing and os.path.exists(out)):
                t = time.time(); info.update(DiffusionMaps.run(dm_dir, pd, k=k, ctx=ctx) or {}); info['dm_s'] = time.time() - t
        if 'pca' in stages:
            out = os.path.join(pca_dir, 'Data_Manifolds', 'PD_%s_vec.npy' % pd)
            if not (skip_existing and os.path.exists(out)):
                t = time.time(); PCA.op(pca_dir, pd, ctx=ctx); info['pca_s'] = time.time() - t
        if 'rot' in stages:
            out = os.path.join(rot_dir, 'Data_Rotations', 'PD%s' % pd, 'PD%s_RotMatrices.npy' % (pd,))
            if not (skip_existing and os.path.exists(out)):
                t = time.time(); Rotation_Histograms.op(rot_dir, pd, dim=dim, PCA=pca, ctx=ctx); info['rot_s'] = time.time() - t
        return info
    return work


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--root', required=True)
    ap.add_argument('--pds', default='1-126')
    ap.add_argument('--stages', default='dist,dm,rot', help='comma list of dist, dm, embed (= dist + dm of the non-CTF branch in one library call), pca, rot')
    ap.add_argument('--dim', type=int, default=6)
    ap.add_argument('--pca', action='store_true')
    ap.add_argument('--no-ctf', action='store_true', help='Distances: CTF = False branch (the reference ships CTF = True)')
    ap.add_argument('-k', type=int, default=16)
    ap.add_argument('--skip-existing', action='store_true')
    ap.add_argument('--inflight', type=int, default=1, help='PDs in flight per GPU (host threads, each with its own context and stream)')
    args = ap.parse_args(argv)
    rank, local_rank, world = dist_env()
    gather = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('gloo')              # status records only; the data path has no collective

        def gather(local):
            out = [None] * world
            dist.all_gather_object(out, local)
            return out
    if world > 1:
        from . import hostbind
        hostbind.bind_to_gpu_numa_node(local_rank)      # host buffers of this rank next to its GPU (best effort)
    works = [_stage_fns(args.root, set(args.stages.split(',')), args.dim, args.pca, args.k, args.skip_existing, local_rank, ctf=not args.no_ctf)
             for _ in range(max(1, args.inflight))]
    recs = run_sharded(parse_pds(args.pds), works if len(works) > 1 else works[0], rank, world, gather, inflight=len(works))
    failed = [r for r in recs if not r.get('ok')]
    if rank == 0:
        for r in recs:
            print(json.dumps({k: v for k, v in r.items() if k != 'trace'}))
        print(json.dumps({'n_pd': len(recs), 'failed': [r['PD'] for r in failed], 'world': world}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 1 if failed else 0


if __name__ == '__main__':
    sys.exit(main())
