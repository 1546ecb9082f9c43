#!/bin/bash
# repeated runs of the eigensolver tests and of the probe: rare races / hangs of the persistent kernels would show up here
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  timeout 600 python -m pytest tests -m gpu -x -q -k "eig or embed or chain or fused or regimes or merged or dmap or stage23 or full_size" > gpurun_out/soak_$i.log 2>&1; echo "round $i rc=$? $(tail -1 gpurun_out/soak_$i.log)"
done
for i in 1 2 3; do timeout 300 python tools/merged_probe.py 20 10 100 2>&1 | grep "^merged" | cut -c1-120; done
