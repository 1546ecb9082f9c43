#!/bin/bash
mkdir -p gpurun_out
python tools/modes_probe.py 1e-11 2>&1 | grep "mode\|status" | tail; python tools/modes_probe.py 1e-8 2>&1 | grep "mode\|status" | tail
timeout 300 python tools/merged_probe.py 20 10 100 > gpurun_out/s3c_probe.log 2>&1; grep "^merged\|status=\|checker 0" gpurun_out/s3c_probe.log
timeout 300 python tools/merged_probe.py 20 16 100 > gpurun_out/s3c_probe16.log 2>&1; grep "^merged\|status=\|checker 0" gpurun_out/s3c_probe16.log
