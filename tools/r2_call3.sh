#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/eig_probe.py > gpurun_out/r2_c3_eig.log 2>&1; echo "eig_probe rc=$?" >> gpurun_out/r2_c3_eig.log
timeout 900 python -m pytest tests -m gpu -q -x -k "embed or eigen or pca or smoke or fused" > gpurun_out/r2_c3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c3_pytest.log
tail -12 gpurun_out/r2_c3_eig.log; tail -6 gpurun_out/r2_c3_pytest.log
