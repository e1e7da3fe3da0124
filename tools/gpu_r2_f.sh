#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "wpackn or groupnorm" > gpurun_out/t_wpn.log 2>&1; echo "rc=$?" >> gpurun_out/t_wpn.log
tail -25 gpurun_out/t_wpn.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "head" > gpurun_out/t_head.log 2>&1; echo "rc=$?" >> gpurun_out/t_head.log
grep -E "parity|passed|failed|rc=|Error" gpurun_out/t_head.log | tail -20
timeout 300 python tools/head_probe.py > gpurun_out/head_probe.log 2>&1; tail -24 gpurun_out/head_probe.log
