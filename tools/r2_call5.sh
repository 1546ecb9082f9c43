#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_c5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c5_pytest.log
tail -5 gpurun_out/r2_c5_pytest.log
timeout 600 python tools/northstar_run.py --pds 1-6 --inflight 2 > gpurun_out/r2_c5_ns1.log 2>&1; echo "northstar rc=$?" >> gpurun_out/r2_c5_ns1.log
tail -3 gpurun_out/r2_c5_ns1.log
timeout 300 python tools/eig_probe.py > gpurun_out/r2_c5_eig.log 2>&1; tail -8 gpurun_out/r2_c5_eig.log
timeout 600 python bench.py --steps 24 --warmup 3 > gpurun_out/r2_c5_bench.json 2> gpurun_out/r2_c5_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_c5_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f serial %.3f'%(d['value'], d['e2e']['value'], d['config']['serial_ms_per_step']), d['kernel_ms_per_step'], d['config']['lanczos_steps'], d.get('cpu_baseline',{}).get('value'))
PY
