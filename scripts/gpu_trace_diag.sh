#!/bin/bash
# event traces (first field launch, pair 0) of the product and of timing-diagnostic variants: gpurun_out/trace_<variant>.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
D=pixel-nerf-yolo_b200/csrc/build/diag
for v in ${VARIANTS:-product noepi nogather}; do
  rm -f gpurun_out/trace_$v.txt
  lib=$D/lib_$v.so; [ "$v" = "product" ] && lib=product
  PNR_TRACE=gpurun_out/trace_$v.txt PNR_TRACE_STAGES=1 timeout 200 python scripts/diag_field_time.py $lib 1 2>&1 | tail -1
  # keep the first two launches (coarse + fine of the first step)
  awk '/^# launch/{n++} n<=2' gpurun_out/trace_$v.txt > gpurun_out/trace_$v.tmp && mv gpurun_out/trace_$v.tmp gpurun_out/trace_$v.txt
done
