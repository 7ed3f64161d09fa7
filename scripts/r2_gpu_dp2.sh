#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -p no:cacheprovider > gpurun_out/r2dp2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2dp2_pytest.log
tail -30 gpurun_out/r2dp2_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2dp2_bench.json 2> gpurun_out/r2dp2_bench.err
echo "bench rc=$?"; grep -v "Backbone\|^$" gpurun_out/r2dp2_bench.err | tail -15
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2dp2_bench.json'))
print({k: d[k] for k in ('n_gpus','value','ms_per_step','dp_check')})
print(d['e2e']['value'], d['e2e']['ms_per_step'])
for k,v in d.get('secondary',{}).items(): print(k, v.get('ms_per_step'), v.get('value'), v.get('per_gpu_batch'), v.get('error'))
PY
