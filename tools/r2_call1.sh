#!/bin/bash
# first GPU call of round 2: the chain eigensolver on hardware, then the full suite with the new defaults, then a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_c1_smi.txt 2>&1
timeout 300 python tools/eig_probe.py > gpurun_out/r2_c1_eig.log 2>&1; echo "eig_probe rc=$?" >> gpurun_out/r2_c1_eig.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_c1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err; echo "bench rc=$?" >> gpurun_out/r2_c1_bench.err
tail -5 gpurun_out/r2_c1_eig.log; tail -5 gpurun_out/r2_c1_pytest.log; tail -c 600 gpurun_out/r2_c1_bench.json
