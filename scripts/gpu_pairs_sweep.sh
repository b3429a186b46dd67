#!/bin/bash
# is the kernel bound by chip-wide L2->SM bandwidth?  cycles per super group vs number of resident CTA pairs
cd "$(dirname "$0")/.."
for p in 74 56 37 18 8; do
  echo "== PNR_MAX_PAIRS=$p"
  PNR_MAX_PAIRS=$p PNR_PROF=1 timeout 300 python scripts/profile_field.py 8192 1 2>&1 | grep -E "mma_total|mma_wait_weights|mma_wait_chunk|mma_issue|gather_total|gather_wait" | head -6
done
