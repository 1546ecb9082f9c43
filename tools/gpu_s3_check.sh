#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "merged_chain or modes_agree or fused_equals" > gpurun_out/s3c_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/s3c_pytest.log
