#!/bin/bash
# ncu evidence: (1) launch list of the bench command, (2) full capture of the fused field kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-r01}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
python scripts/profile_field.py 4096 3 > gpurun_out/prof_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:field_ -s 2 -c 2 -f -o gpurun_out/prof_field_$TAG \
    python scripts/profile_field.py 4096 3 > gpurun_out/ncu_field_$TAG.log 2>&1
ls -la gpurun_out/ | grep $TAG
