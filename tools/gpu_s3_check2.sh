#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/merged_probe.py 20 10 100 > gpurun_out/s3c_probe.log 2>&1; grep "^merged\|status=\|checker [01]\|kernel timeline\|Ritz vectors\|recurrence" gpurun_out/s3c_probe.log | head -8
python tools/modes_probe.py 1e-11 2>&1 | grep "mode\|status\|Ritz vectors\|recurrence" | tail -5
