#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s3c_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s3c_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3c_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s3c_smoke.log
timeout 900 python bench.py > gpurun_out/s3c_bench.json 2> gpurun_out/s3c_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s3c_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f serial %.3f'%(d['value'], d['e2e']['value'], d['config']['serial_ms_per_step']), d['kernel_ms_per_step'], d['config']['lanczos_steps'], d['clocks'], d['roofline']['frac'], d['e2e'].get('frac_of_h2d_floor'), d['gpu_launches'])
PY
