#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_host.py tests/test_gpu_train.py -q > gpurun_out/t_train.log 2>&1; echo "rc=$?" >> gpurun_out/t_train.log
tail -12 gpurun_out/t_train.log
python tools/train_probe.py > gpurun_out/train_probe.log 2>&1; tail -45 gpurun_out/train_probe.log
