#!/bin/bash
# Round-2 ncu captures (each only after the plain command exited 0):
#   1. `ncu --set full` of the two field-kernel launches of one bench-sized render (16 384 rays),
#   2. `ncu --set full` of the training GEMMs inside one config-3 step (a heavy rowgemm and a heavy wgrad launch),
#   3. launch list of the bench command.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-r02}
python scripts/profile_field.py 16384 2 > gpurun_out/prof_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:field_pair -s 2 -c 2 -f -o gpurun_out/prof_field_$TAG \
    python scripts/profile_field.py 16384 2 > gpurun_out/ncu_field_$TAG.log 2>&1
echo "field capture exit $?"
TP=bf16 python scripts/train_profile.py > gpurun_out/train_plain_$TAG.log 2>&1 && \
TP=bf16 ncu --set full --clock-control none --import-source on -k regex:rowgemm -s 60 -c 6 -f -o gpurun_out/prof_rowgemm_$TAG \
    python scripts/train_profile.py > gpurun_out/ncu_rowgemm_$TAG.log 2>&1
echo "rowgemm capture exit $?"
TP=bf16 ncu --set full --clock-control none --import-source on -k regex:wgrad -s 36 -c 4 -f -o gpurun_out/prof_wgrad_$TAG \
    python scripts/train_profile.py > gpurun_out/ncu_wgrad_$TAG.log 2>&1
echo "wgrad capture exit $?"
TP=bf16 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_train_launches.csv python scripts/train_profile.py > gpurun_out/ncu_trainlist_$TAG.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
ls -la gpurun_out/*.ncu-rep | tail -5
