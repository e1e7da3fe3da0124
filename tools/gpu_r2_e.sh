#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -x -q -k "groupnorm or convT or conv3d_dilated or conv3d_halo or linear_bias" > gpurun_out/t_gn.log 2>&1; echo "rc=$?" >> gpurun_out/t_gn.log
tail -25 gpurun_out/t_gn.log
python -m pytest tests/test_gpu_parity.py -x -q -s -k "head" > gpurun_out/t_head.log 2>&1; echo "rc=$?" >> gpurun_out/t_head.log
grep -E "parity|passed|failed|rc=|Error" gpurun_out/t_head.log | tail -20
python tools/head_probe.py > gpurun_out/head_probe.log 2>&1; tail -40 gpurun_out/head_probe.log
python tools/head_probe.py 128 --nofuse > gpurun_out/head_probe_nofuse.log 2>&1; tail -3 gpurun_out/head_probe_nofuse.log
