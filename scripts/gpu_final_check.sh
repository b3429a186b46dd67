#!/bin/bash
# HEAD validation without the full ncu captures (kernels unchanged since the last gpu_round_end.sh): tests, smoke, both bench
# arms, ncu launch list of the bench command.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-r01h}
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/test_full_$TAG.log 2>&1
echo "tests exit $? $(tail -1 gpurun_out/test_full_$TAG.log)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $? $(tail -1 gpurun_out/smoke_$TAG.log)"
timeout 600 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -1 gpurun_out/bench_$TAG.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-200
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list exit $?"
