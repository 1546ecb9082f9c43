#!/bin/bash
mkdir -p gpurun_out
for stp in 20 150; do
timeout 300 python tools/merged_probe.py 20 10 $stp > gpurun_out/s3_merged_s$stp.log 2>&1; echo "trace step $stp rc=$?"; grep -v "checker [123]" gpurun_out/s3_merged_s$stp.log | tail -60
done
timeout 900 python -m pytest tests -m gpu -x -q -k "embed or eig or dmap or full_size or stage23 or chain or fused" > gpurun_out/s3_merged_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s3_merged_pytest.log
