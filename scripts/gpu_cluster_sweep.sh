#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
: > gpurun_out/sweep.txt
for cs in ${SWEEP:-1 2 4}; do
  PNR_CLUSTER=$cs timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "field and bf16 or render_bf16 or edge" > gpurun_out/test_cs$cs.log 2>&1
  echo "cs=$cs tests exit $? $(tail -1 gpurun_out/test_cs$cs.log)" >> gpurun_out/sweep.txt
  PNR_CLUSTER=$cs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cs$cs.log 2>gpurun_out/bench_cs$cs.err
  echo "cs=$cs bench exit $?" >> gpurun_out/sweep.txt
  python - <<PY >> gpurun_out/sweep.txt
import json
try:
    d=json.loads(open("gpurun_out/bench_cs$cs.log").read().strip().splitlines()[-1])
    print("   rays/s", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("   parse fail", e)
PY
done
cat gpurun_out/sweep.txt
