#!/bin/bash
# quick correctness + speed check of the fused kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "umma or field or render or edge or full_image or latent_widths" > gpurun_out/test_quick.log 2>&1
echo "tests exit $? $(tail -1 gpurun_out/test_quick.log)"
for cs in ${SWEEP:-1 2}; do
  PNR_CLUSTER=$cs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q$cs.log 2>gpurun_out/bench_q$cs.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_q$cs.log").read().strip().splitlines()[-1])
    print("cs=$cs rays/s", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("cs=$cs parse fail", e); print(open("gpurun_out/bench_q$cs.err").read()[-1500:])
PY
done
PNR_PROF=1 PNR_CLUSTER=1 timeout 300 python scripts/profile_field.py 16384 1 2>&1 | tail -18
