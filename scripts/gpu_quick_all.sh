#!/bin/bash
# Like gpu_quick.sh but runs every selected test (no -x) and prints each assertion message.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-q}
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -k "$2" > gpurun_out/test_$TAG.log 2>&1
echo "tests exit $? $(tail -1 gpurun_out/test_$TAG.log)"
grep -E "^(FAILED|ERROR)|^E  .*(Error|err)" gpurun_out/test_$TAG.log | head -40
