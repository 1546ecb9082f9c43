#!/bin/bash
mkdir -p gpurun_out
ESPER_TRACE_CLUSTERS=8 timeout 300 python tools/merged_probe.py 20 10 100 > gpurun_out/s3c_probe.log 2>&1; grep -v "checker [123]" gpurun_out/s3c_probe.log | grep -A19 "^merged 1" | head -21; grep "^merged" gpurun_out/s3c_probe.log
timeout 900 python -m pytest tests -m gpu -x -q -k "embed or eig or dmap or full_size or stage23 or chain or fused or pca" > gpurun_out/s3c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s3c_pytest.log
