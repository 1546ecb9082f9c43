"""CPU, world_size 2 over gloo: the PD-sharding launcher partitions work with no data-path collective
and gathers the per-PD records on rank 0."""
import json
import os
import socket
import subprocess
import sys

from conftest import ROOT

WORKER = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from manifoldem_esper_b200 import launcher
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
pds = launcher.parse_pds('1-11')
def gather(local):
    out = [None] * world
    dist.all_gather_object(out, local)
    return out
def work(pd):
    if pd == '004':
        raise ValueError('bad stack')
    return {'val0': float(int(pd)) * 2}
recs = launcher.run_sharded(pds, work, rank, world, gather)
if rank == 0:
    print('RESULT ' + json.dumps([{k: r[k] for k in ('PD', 'rank', 'ok')} for r in recs]))
dist.barrier()
dist.destroy_process_group()
'''


def free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS='1')
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    line = [l for l in res.stdout.splitlines() if l.startswith('RESULT ')][0]
    recs = json.loads(line[len('RESULT '):])
    assert [r['PD'] for r in recs] == ['%03d' % i for i in range(1, 12)]
    assert all(r['rank'] == (int(r['PD']) - 1) % 2 for r in recs)           # round-robin ownership
    assert [r['ok'] for r in recs] == [pd != 4 for pd in range(1, 12)]      # one PD failed, the rest went on
