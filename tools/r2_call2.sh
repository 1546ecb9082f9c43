#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/eig_probe.py > gpurun_out/r2_c2_eig.log 2>&1; echo "eig_probe rc=$?" >> gpurun_out/r2_c2_eig.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --eig-mode 0 > gpurun_out/r2_c2_bench_m0.json 2> gpurun_out/r2_c2_bench_m0.err; echo "bench rc=$?" >> gpurun_out/r2_c2_bench_m0.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --eig-mode 0 --inflight 2 > gpurun_out/r2_c2_bench_m0_nf2.json 2> gpurun_out/r2_c2_bench_m0_nf2.err
tail -8 gpurun_out/r2_c2_eig.log; tail -8 gpurun_out/r2_c2_pytest.log; python - <<'PY'
import json
for f in ('gpurun_out/r2_c2_bench_m0.json','gpurun_out/r2_c2_bench_m0_nf2.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value %.1f e2e %.1f serial %.3f'%(d['value'], d['e2e']['value'], d['config']['serial_ms_per_step']), d['kernel_ms_per_step'], d['config']['lanczos_steps'])
    except Exception as e: print(f, 'ERR', e)
PY
