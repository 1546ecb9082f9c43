#!/bin/bash
# session 3: ncu launch list of one PD and full capture of the default kernels.  ESPER_CHAIN_NOCOOP=1: Nsight Compute cannot launch a
# kernel that is both cooperative and clustered; the eigensolver kernel is launched with the cluster attribute only (same code, same grid).
mkdir -p gpurun_out
export ESPER_CHAIN_NOCOOP=1
timeout 300 python tools/profile_step.py 3 0 > gpurun_out/s3p_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s3_launches_profile_step.csv python tools/profile_step.py 3 0 > gpurun_out/s3p_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 300 python tools/profile_step.py 2 0 > gpurun_out/s3p_plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'lanczos_merged|stats_split|gram_wide|ritz_fix' -c 8 -o gpurun_out/r2s3_step_full python tools/profile_step.py 2 0 > gpurun_out/s3p_ncu2.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/s3p_ncu2.log
