#!/bin/bash
mkdir -p gpurun_out
for cs in 4 16 2; do
ESPER_CLUSTER_SIZE=$cs timeout 300 python tools/merged_probe.py 20 10 100 > gpurun_out/s3_merged_cs$cs.log 2>&1; echo "cluster size $cs rc=$?"; grep "^merged 1\|clusters of\|status" gpurun_out/s3_merged_cs$cs.log | head -3
done
