#!/bin/bash
# Round-2 GPU call A: new attention fast pass (correctness + timing vs the exact pass and SDPA), parity sweep, bench per operand mode.
mkdir -p gpurun_out
python -c "import h5py; print('h5py', h5py.__version__)" > gpurun_out/h5py_probe.txt 2>&1
nvidia-smi -L > gpurun_out/gpus.txt
python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" > gpurun_out/t_attn.log 2>&1; echo "rc=$?" >> gpurun_out/t_attn.log
python tools/attn_ab.py > gpurun_out/attn_fast.log 2>&1
CVIT_FA_EXACT=1 python tools/attn_ab.py > gpurun_out/attn_exact.log 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -s -k "sweep or one_slice or width" > gpurun_out/t_parity.log 2>&1; echo "rc=$?" >> gpurun_out/t_parity.log
python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed.err
python bench.py --no-cpu-baseline --steps 5 --warmup 3 --operands bf16 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
CVIT_FA_EXACT=1 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/bench_mixed_exact.json 2> gpurun_out/bench_mixed_exact.err
tail -3 gpurun_out/t_attn.log gpurun_out/attn_fast.log gpurun_out/attn_exact.log
tail -25 gpurun_out/t_parity.log
