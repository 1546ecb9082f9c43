#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/merged_probe.py 20 10 100 > gpurun_out/s3c_probe.log 2>&1; grep "^merged\|status=\|checker [01]\|kernel timeline\|Ritz vectors" gpurun_out/s3c_probe.log | head -14
timeout 300 python tools/merged_probe.py 20 16 100 > gpurun_out/s3c_probe16.log 2>&1; grep "^merged\|status=\|Ritz vectors" gpurun_out/s3c_probe16.log | head -6
python tools/modes_probe.py 1e-11 2>&1 | grep "mode\|status\|Ritz vectors" | tail -4
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s3c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s3c_pytest.log
