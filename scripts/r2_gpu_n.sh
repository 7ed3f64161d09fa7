#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -12 gpurun_out/r2n_pytest.log
for multi in 8 1 16; do
ODPD_BENCH_STEPS_PER_REPLAY=$multi timeout 600 python bench.py --steps 320 --warmup 16 --no-cpu --no-secondary > gpurun_out/r2n_bench_m$multi.json 2> gpurun_out/r2n_bench_m$multi.err
tail -2 gpurun_out/r2n_bench_m$multi.err | grep -v Backbone
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_m$multi.json')); print('multi=$multi', d['ms_per_step'], d['value'], d['kernel_ms'], d['e2e']['ms_per_step'], d['e2e_indexed']['ms_per_step'], d['serial_floor']['ms_per_step'], d['run']['final_loss'])"
done
ODPD_BENCH_STEPS_PER_REPLAY=8 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > gpurun_out/r2n_bench_k20.json 2> gpurun_out/r2n_bench_k20.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_k20.json')); print('K20', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e_indexed']['ms_per_step'])"
