#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s3b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s3b_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3b_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s3b_smoke.log
timeout 900 python bench.py > gpurun_out/s3b_bench.json 2> gpurun_out/s3b_bench.err; echo "bench rc=$?"
for v in "--inflight 2" "--inflight 4"; do
  tag=$(echo "d $v" | tr -d ' -')
  timeout 300 python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-extras $v > gpurun_out/s3b_bench_$tag.json 2> gpurun_out/s3b_bench_$tag.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s3b_bench*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value %.1f e2e %.1f serial %.3f'%(d['value'], d['e2e']['value'], d['config']['serial_ms_per_step']), d['kernel_ms_per_step'], d['config']['lanczos_steps'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
