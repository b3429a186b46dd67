#!/bin/bash
# Quick GPU check of a test subset.  Usage: scripts/gpu_quick.sh TAG "pytest -k expression"
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-q}
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x -k "$2" > gpurun_out/test_$TAG.log 2>&1
echo "tests exit $? $(tail -1 gpurun_out/test_$TAG.log)"
grep -E "^(FAILED|ERROR)|Error|error:" gpurun_out/test_$TAG.log | head -20
