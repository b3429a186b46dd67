#!/bin/bash
# Upper bounds for the next kernel redesign: step time of the field kernel with one cost removed at a time (timing-diagnostic
# builds from scripts/build_diag_variants.sh; wrong results by construction, never used by the product).
cd "$(dirname "$0")/.."
D=pixel-nerf-yolo_b200/csrc/build/diag
mkdir -p gpurun_out
{
timeout 120 python scripts/diag_field_time.py product 5
for v in ${VARIANTS:-nogather noweights noepi noweights_noepi none3 noring noring_noepi noring_none3}; do
  timeout 120 python scripts/diag_field_time.py $D/lib_$v.so 5 2>&1 | tail -1
done
} | tee gpurun_out/diag_bounds.txt
