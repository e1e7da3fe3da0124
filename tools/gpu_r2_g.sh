#!/bin/bash
mkdir -p gpurun_out
python tools/head_probe.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:wpackn -s 8 -c 4 -o gpurun_out/prof_wpn python tools/head_probe.py > gpurun_out/ncu_wpn.log 2>&1
tail -5 gpurun_out/ncu_wpn.log; ls -la gpurun_out/*.ncu-rep
