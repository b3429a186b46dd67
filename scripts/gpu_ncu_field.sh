#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-x}
PNR_CLUSTER=${PNR_CLUSTER:-1} python scripts/profile_field.py 4096 3 > gpurun_out/prof_plain_$TAG.log 2>&1 && \
PNR_CLUSTER=${PNR_CLUSTER:-1} ncu --set full --clock-control none --import-source on -k regex:field_umma -s 2 -c 1 -f -o gpurun_out/prof_field_$TAG \
    python scripts/profile_field.py 4096 3 > gpurun_out/ncu_field_$TAG.log 2>&1
tail -3 gpurun_out/ncu_field_$TAG.log
