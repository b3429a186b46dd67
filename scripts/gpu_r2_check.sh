#!/bin/bash
# Round-2 HEAD validation on one B200: GPU tests, smoke, both bench arms.  Usage: scripts/gpu_r2_check.sh TAG [pytest -k expr]
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-r02a}
KEXPR=${2:-}
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -k "$KEXPR" > gpurun_out/test_$TAG.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/test_$TAG.log 2>&1
fi
echo "tests exit $? $(tail -1 gpurun_out/test_$TAG.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/test_$TAG.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $? $(tail -1 gpurun_out/smoke_$TAG.log)"
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -1 gpurun_out/bench_$TAG.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2> gpurun_out/bench_ref_$TAG.err; echo "ref exit $?"; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-200
