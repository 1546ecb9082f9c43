#!/bin/bash
# session 3 baseline: cluster-schedule timeline, full GPU suite, smoke, default bench
mkdir -p gpurun_out
for stp in 20 150; do
ESPER_TRACE_STEP=$stp timeout 240 python tools/cluster_probe.py 20 10 8 > gpurun_out/s3_probe8_s$stp.log 2>&1; echo "trace step $stp rc=$?"; grep -v "checker [123]" gpurun_out/s3_probe8_s$stp.log | tail -22
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_pytest.log
tail -4 gpurun_out/s3_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s3_smoke.log
timeout 900 python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/s3_bench.json
