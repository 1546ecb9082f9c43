#!/bin/bash
mkdir -p gpurun_out
for stp in 20 165; do
ESPER_TRACE_STEP=$stp timeout 240 python tools/cluster_probe.py 20 10 8 > gpurun_out/cl_probe8_s$stp.log 2>&1; echo "trace step $stp rc=$?"; grep -v "checker [123]" gpurun_out/cl_probe8_s$stp.log | tail -19
done
timeout 900 python -m pytest tests -m gpu -x -q -k "embed or eig or dmap or pca or full_size or stage23 or chain or rowblock" > gpurun_out/cl_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/cl_pytest.log
