#!/bin/bash
mkdir -p gpurun_out
for mode in tma2 tma1; do
  export CVIT_WPACKN_MODE=$mode
  timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "wpackn or folded" > gpurun_out/t_wpn_$mode.log 2>&1; echo "rc=$?" >> gpurun_out/t_wpn_$mode.log
  echo "== $mode"; tail -6 gpurun_out/t_wpn_$mode.log
  timeout 300 python tools/head_probe.py > gpurun_out/head_probe_$mode.log 2>&1; grep -E "wpackn|head total" gpurun_out/head_probe_$mode.log
done
