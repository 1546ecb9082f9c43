#!/bin/bash
mkdir -p gpurun_out
echo "--- variant: no cooperative attribute"
ESPER_CHAIN_NOCOOP=1 timeout 300 python tools/profile_step.py 2 0 > gpurun_out/s3n_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/s3n_plain.log
ESPER_CHAIN_NOCOOP=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s3n_launches_nocoop.csv python tools/profile_step.py 3 0 > gpurun_out/s3n_ncu_nocoop.log 2>&1; echo "nocoop list rc=$?"; grep -c "merged" gpurun_out/s3n_launches_nocoop.csv; grep merged gpurun_out/s3n_launches_nocoop.csv | tail -1 | cut -c1-300
echo "--- variant: application replay"
timeout 600 ncu --replay-mode application --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s3n_launches_app.csv python tools/profile_step.py 3 0 > gpurun_out/s3n_ncu_app.log 2>&1; echo "app replay rc=$?"; grep merged gpurun_out/s3n_launches_app.csv | tail -1 | cut -c1-300
