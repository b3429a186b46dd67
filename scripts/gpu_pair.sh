#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for pairflag in ${PAIRS:-1}; do
  PNR_PAIR=$pairflag timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "field and bf16 or render or edge or full_image or latent_widths" > gpurun_out/test_pair$pairflag.log 2>&1
  echo "PAIR=$pairflag tests exit $? $(tail -1 gpurun_out/test_pair$pairflag.log)"
  grep -E "pnr:|Error|error|FAILED|assert" gpurun_out/test_pair$pairflag.log | head -8
  PNR_PAIR=$pairflag timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair$pairflag.log 2>gpurun_out/bench_pair$pairflag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_pair$pairflag.log").read().strip().splitlines()[-1])
    print("PAIR=$pairflag rays/s", round(d["value"]), "ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3), d["clocks"])
except Exception as e:
    print("PAIR=$pairflag parse fail", e); print(open("gpurun_out/bench_pair$pairflag.err").read()[-800:])
PY
done
PNR_PROF=1 python scripts/profile_field.py 8192 1 2>&1 | tail -16
