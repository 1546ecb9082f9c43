#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/merged_probe.py 20 10 100 > gpurun_out/s3c_probe.log 2>&1; grep "^merged\|status=" gpurun_out/s3c_probe.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s3c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s3c_pytest.log
timeout 300 python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/s3c_bench.json 2> gpurun_out/s3c_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s3c_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f serial %.3f'%(d['value'], d['e2e']['value'], d['config']['serial_ms_per_step']), d['kernel_ms_per_step'], d['config']['lanczos_steps'])
PY
