#!/bin/bash
# ncu --set full of the hot kernels at BASELINE shapes (one launch each), extracted to small JSON under gpurun_out/, plus
# the launch list of the bench. Each ncu command runs only after the same command exited 0 without ncu.
# usage (on the GPU box): bash tools/ncu_refresh.sh r02 v1
set -u
rnd=${1:-r02}
tag=${2:-vN}
log=gpurun_out/${rnd}_ncu_full_${tag}.log
python tools/ncu_target.py 128 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05|attention_tcgen05|layernorm" -s 6 -c 6 \
      -o /tmp/${rnd}_full_${tag} python tools/ncu_target.py 128 > $log 2>&1
python tools/ncu_extract.py /tmp/${rnd}_full_${tag}.ncu-rep > gpurun_out/${rnd}_ncu_full_${tag}_hot_kernels.json 2>> $log
# head forward: the third volume of tools/head_probe.py (25 launches per volume, warm)
python tools/head_probe.py > /dev/null 2>&1 && \
  ncu --set full --clock-control none -k regex:"conv3d|gemm_tcgen05|gn_f|groupnorm" -s 50 -c 25 -o /tmp/${rnd}_head_full_${tag} python tools/head_probe.py >> $log 2>&1
python tools/ncu_extract.py /tmp/${rnd}_head_full_${tag}.ncu-rep > gpurun_out/${rnd}_ncu_full_${tag}_head_kernels.json 2>> $log
# training: the tcgen05 weight-gradient kernels of one step
python tools/train_probe.py > /dev/null 2>&1 && \
  ncu --set full --clock-control none -k regex:"wgrad_tc|wgrad_mn" -c 14 -o /tmp/${rnd}_train_full_${tag} python tools/train_probe.py >> $log 2>&1
python tools/ncu_extract.py /tmp/${rnd}_train_full_${tag}.ncu-rep > gpurun_out/${rnd}_ncu_full_${tag}_train_kernels.json 2>> $log
# launch list of the bench's ViT leg (one timed step after three warm-up steps)
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dataset > /dev/null 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 700 --csv --log-file gpurun_out/${rnd}_launches_${tag}.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dataset >> $log 2>&1
ls -la /tmp/*.ncu-rep; tail -3 $log; wc -c gpurun_out/${rnd}_ncu_full_${tag}_*.json gpurun_out/${rnd}_launches_${tag}.csv
