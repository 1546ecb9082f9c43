#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/s3_bench_n2.json 2> gpurun_out/s3_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python -m pytest tests -m gpu -q -k "rowblock or two_rank or rank" > gpurun_out/s3_n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s3_n2_pytest.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s3_bench_n2.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), d['kernel_ms_per_step'], d['config']['lanczos_steps'], d['clocks']); print(json.dumps(d.get('c4_rowblock'))[:1500])
PY
