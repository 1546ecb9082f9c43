#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/northstar_run.py --pds 1-126 --inflight 2 > gpurun_out/r2s3_northstar_g8.log 2> gpurun_out/r2s3_northstar_g8.err; echo "northstar rc=$?"
tail -2 gpurun_out/r2s3_northstar_g8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 > gpurun_out/r2s3_bench_g8.json 2> gpurun_out/r2s3_bench_g8.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s3_bench_g8.json').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), d['e2e'].get('h2d_probe_gbs'), d['e2e'].get('frac_of_h2d_floor'))
print(json.dumps(d.get('c4_rowblock'))[:1500])
PY
