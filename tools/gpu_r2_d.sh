#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/t_gpu_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_gpu_all.log
tail -5 gpurun_out/t_gpu_all.log
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | tail -3
tail -5 gpurun_out/bench_full.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
wc -c gpurun_out/bench_full.json gpurun_out/bench_ref.json
