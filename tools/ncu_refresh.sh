#!/bin/bash
# ncu --set full of the hot kernels at BASELINE shapes (one launch each), extracted to small JSON under gpurun_out/.
# Each ncu command runs only after the same command exited 0 without ncu. usage (on the GPU box): bash tools/ncu_refresh.sh v14
set -u
tag=${1:-vN}
log=gpurun_out/ncu_full_${tag}.log
python tools/ncu_target.py 128 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|attention_tcgen05|layernorm" -s 6 -c 6 \
      -o /tmp/r01_full_${tag} python tools/ncu_target.py 128 > $log 2>&1
python tools/ncu_extract.py /tmp/r01_full_${tag}.ncu-rep > gpurun_out/r01_ncu_full_${tag}_vit.json 2>> $log
python tools/head_probe.py > /dev/null 2>&1 && \
  ncu --set full --clock-control none -s 54 -c 27 -o /tmp/r01_head_full_${tag} python tools/head_probe.py >> $log 2>&1
python tools/ncu_extract.py /tmp/r01_head_full_${tag}.ncu-rep > gpurun_out/r01_ncu_full_${tag}_head.json 2>> $log
ls -la /tmp/*.ncu-rep; tail -3 $log; wc -c gpurun_out/r01_ncu_full_${tag}_*.json
