#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r2dp${N}_bench.json 2> gpurun_out/r2dp${N}_bench.err
echo "bench rc=$?"; grep -v "Backbone\|^$\|OMP_NUM\|\*\*\*\*" gpurun_out/r2dp${N}_bench.err | tail -8
python - <<PY
import json
d = json.load(open('gpurun_out/r2dp${N}_bench.json'))
print({k: d[k] for k in ('n_gpus','value','ms_per_step','dp_check')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'idx', d['e2e_indexed']['ms_per_step'], 'serial', d['serial_floor']['ms_per_step'])
for k,v in d.get('secondary',{}).items(): print(k, v.get('ms_per_step'), v.get('value'), v.get('per_gpu_batch'), v.get('error'))
PY
