#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  CVIT_HEAD_FUSE_GN=$1 CVIT_HEAD_WPACKN=$2 python -m pytest tests/test_gpu_host.py -q -k "train_then_eval" > gpurun_out/t_te_$1$2.log 2>&1
  echo "fuse=$1 wpackn=$2: $(tail -1 gpurun_out/t_te_$1$2.log)"; grep -A3 "AssertionError:" gpurun_out/t_te_$1$2.log | head -5
done
