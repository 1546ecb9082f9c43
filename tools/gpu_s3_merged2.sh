#!/bin/bash
mkdir -p gpurun_out
for ord in 1 9 10; do
ESPER_MERGED_ORDER=$ord timeout 300 python tools/merged_probe.py 20 10 ${TRACE_STEP:-100} > gpurun_out/s3_merged_o$ord.log 2>&1; echo "order $ord rc=$?"; grep -v "checker [123]" gpurun_out/s3_merged_o$ord.log | grep -A19 "^merged 1" | head -21 | grep "^merged\|vector complete\|product done\|step end\|step start"
done
