#!/bin/bash
# event trace of CTA pair 0 of the field kernel (trace instantiation, PNR_TRACE): gpurun_out/trace_<tag>.txt
# usage: gpu_trace.sh <tag> [stages]   (stages: also log every MMA stage)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/trace_$1.txt
if [ -n "$2" ]; then export PNR_TRACE_STAGES=1; fi
PNR_TRACE=gpurun_out/trace_$1.txt python scripts/profile_field.py 16384 1 2>&1 | tail -3
wc -l gpurun_out/trace_$1.txt
