#!/bin/bash
mkdir -p gpurun_out
for v in "--inflight 4" "--inflight 6" "--inflight 8" "--inflight 4 --eig-mode 4" "--inflight 6 --eig-mode 4" "--inflight 8 --eig-mode 4"; do
  tag=$(echo "$v" | tr -d ' -')
  timeout 300 python bench.py --steps 48 --warmup 3 --no-cpu-baseline --no-extras $v > gpurun_out/r2_c6_bench_$tag.json 2> gpurun_out/r2_c6_bench_$tag.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_c6_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('bench_')[1], 'value %.1f e2e %.1f serial %.3f'%(d['value'], d['e2e']['value'], d['config']['serial_ms_per_step']), d['kernel_ms_per_step']['lanczos'], d['config']['lanczos_steps'], d['clocks'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-300:])
PY
