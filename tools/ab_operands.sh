for i in 1 2; do
for op in bf16 fp16; do
python bench.py --no-cpu-baseline --operands $op > gpurun_out/ab_${op}_$i.json 2> gpurun_out/ab_${op}_$i.err
python -c "
import json; d=json.loads(open('gpurun_out/ab_${op}_$i.json').read().strip().splitlines()[-1]); print('$op', d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], {k:round(v['ms'],3) for k,v in d['kernels'].items()})"
done; done
nvidia-smi --query-gpu=power.limit,power.draw,temperature.gpu --format=csv
