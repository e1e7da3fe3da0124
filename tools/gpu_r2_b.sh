#!/bin/bash
mkdir -p gpurun_out
python tools/fa_trace.py > gpurun_out/fa_trace_fast.log 2>&1
CVIT_FA_EXACT=1 python tools/fa_trace.py > gpurun_out/fa_trace_exact.log 2>&1
cat gpurun_out/fa_trace_fast.log gpurun_out/fa_trace_exact.log
