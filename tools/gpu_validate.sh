#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/val_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/val_pytest.log
tail -4 gpurun_out/val_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/val_smoke.log
timeout 900 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/val_ref.json 2> gpurun_out/val_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/val_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), d['kernel_ms_per_step'], d['config']['lanczos_steps'], d['clocks'], d['roofline']['frac'], d['roofline']['frac_issued'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['c5_rot'].get('evaluations_per_s'))
r=json.loads(open('gpurun_out/val_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['ms_per_step'], r['config'].get('measured_s_per_step'))
PY
timeout 300 python tools/profile_step.py 2 1 > gpurun_out/val_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'stats_split|lanczos_resident' -c 6 -o gpurun_out/step_full_b python tools/profile_step.py 2 1 > gpurun_out/val_ncu.log 2>&1
ls -la gpurun_out/step_full_b.ncu-rep
