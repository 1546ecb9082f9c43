#!/bin/bash
# full GPU suite + bench variants (eigensolver schedule x PDs in flight)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s3f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3f_pytest.log
tail -6 gpurun_out/s3f_pytest.log
for v in "" "--eig-mode 0" "--eig-mode 0 --inflight 2" "--eig-mode 0 --inflight 3" "--eig-mode 0 --inflight 4" "--eig-mode 0 --inflight 1"; do
  tag=$(echo "d $v" | tr -d ' -')
  timeout 300 python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-extras $v > gpurun_out/s3f_bench_$tag.json 2> gpurun_out/s3f_bench_$tag.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/s3f_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value %.1f e2e %.1f serial %.3f'%(d['value'], d['e2e']['value'], d['config']['serial_ms_per_step']), d['kernel_ms_per_step'], d['config']['lanczos_steps'])
    except Exception as e: print(f, 'ERR', e)
PY
