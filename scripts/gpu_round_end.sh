#!/bin/bash
# Everything the judge looks at, in one GPU call: tests, smoke, both bench arms, ncu launch list + full captures.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-r01c}
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/test_full_$TAG.log 2>&1
echo "tests exit $? $(tail -1 gpurun_out/test_full_$TAG.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $? $(tail -1 gpurun_out/smoke_$TAG.log)"
timeout 600 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -1 gpurun_out/bench_$TAG.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-300
# ncu: launch list of the bench command, then full captures (each only after the plain command exited 0)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
python scripts/profile_field.py 16384 2 > gpurun_out/prof_plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:field_pair -s 2 -c 2 -f -o gpurun_out/prof_field_$TAG \
    python scripts/profile_field.py 16384 2 > gpurun_out/ncu_field_$TAG.log 2>&1
python scripts/gather_bench.py > gpurun_out/gather_bench_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gather_encode_kernel -s 9 -c 1 -f -o gpurun_out/prof_gather_$TAG \
    python scripts/gather_bench.py > gpurun_out/ncu_gather_$TAG.log 2>&1
cat gpurun_out/gather_bench_$TAG.log
ls -la gpurun_out/ | grep $TAG
