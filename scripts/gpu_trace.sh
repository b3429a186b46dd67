#!/bin/bash
# event trace of CTA pair 0 of the field kernel (PROF instantiation, PNR_TRACE): gpurun_out/trace_<tag>.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/trace_$1.txt
PNR_TRACE=gpurun_out/trace_$1.txt python scripts/profile_field.py 16384 1 2>&1 | tail -24
wc -l gpurun_out/trace_$1.txt
