#!/bin/bash
# full GPU suite + bench variants + ncu launch list and full captures (one call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_c4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c4_pytest.log
tail -4 gpurun_out/r2_c4_pytest.log
for v in "" "--inflight 2" "--eig-mode 0" "--inflight 2 --eig-mode 0" "--inflight 3"; do
  tag=$(echo "nf4 $v" | tr -d ' -')
  timeout 300 python bench.py --steps 24 --warmup 3 --no-cpu-baseline --no-extras $v > gpurun_out/r2_c4_bench_$tag.json 2> gpurun_out/r2_c4_bench_$tag.err
done
timeout 600 python bench.py --steps 24 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c4_bench_full.json 2> gpurun_out/r2_c4_bench_full.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_c4_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value %.1f e2e %.1f serial %.3f'%(d['value'], d['e2e']['value'], d['config']['serial_ms_per_step']), d['kernel_ms_per_step'], d['config']['lanczos_steps'], {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk in ('evaluations_per_s','evaluations_per_s_kernels','error')}) for k, v in d.items() if k in ('c5_rot',)})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 python tools/profile_step.py 3 0 > gpurun_out/r2_c4_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_profile_step.csv python tools/profile_step.py 3 0 > gpurun_out/r2_c4_ncu1.log 2>&1
timeout 300 python tools/profile_step.py 2 0 > gpurun_out/r2_c4_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gram_wide|stats_split|lanczos_chain|dist_finalize|moment_accum' -c 10 -o gpurun_out/r2_step_full python tools/profile_step.py 2 0 > gpurun_out/r2_c4_ncu2.log 2>&1
timeout 300 python tools/profile_c4.py 50000 24 > gpurun_out/r2_c4_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'symv_rows|dots_split|update_split' -s 20 -c 6 -o gpurun_out/r2_c4_full python tools/profile_c4.py 50000 24 > gpurun_out/r2_c4_ncu3.log 2>&1
tail -3 gpurun_out/r2_c4_plain3.log; ls -la gpurun_out/*.ncu-rep
