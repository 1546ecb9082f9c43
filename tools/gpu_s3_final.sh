#!/bin/bash
# end of session 3: full GPU suite, smoke, default bench + reference arm, ncu launch list and full capture of the default kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s3z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s3z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s3z_smoke.log
timeout 900 python bench.py > gpurun_out/s3z_bench.json 2> gpurun_out/s3z_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s3z_ref.json 2> gpurun_out/s3z_ref.err; echo "ref rc=$?"
export ESPER_CHAIN_NOCOOP=1
timeout 300 python tools/profile_step.py 3 0 > gpurun_out/s3z_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s3_launches_profile_step.csv python tools/profile_step.py 3 0 > gpurun_out/s3z_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/profile_step.py 2 0 > gpurun_out/s3z_plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'lanczos_merged|stats_split|gram_wide|ritz_rows|ritz_scale' -c 10 -o gpurun_out/r2s3_step_full python tools/profile_step.py 2 0 > gpurun_out/s3z_ncu2.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/r2s3_step_full.ncu-rep
unset ESPER_CHAIN_NOCOOP
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s3z_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), d['kernel_ms_per_step'], d['config']['lanczos_steps'], d['clocks'], d['roofline']['frac'], d['e2e'].get('frac_of_h2d_floor'))
r=json.loads(open('gpurun_out/s3z_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['ms_per_step'])
PY
