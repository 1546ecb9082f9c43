mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s3_last_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s3_last_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s3_last_bench.json 2> gpurun_out/s3_last_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s3_last_bench.json').read().strip().splitlines()[-1]); print('value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), d['kernel_ms_per_step'], d['clocks'], d['gpu_launches'])
PY
