#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -12 gpurun_out/r2m_pytest.log
for fuse in 1 0; do
ODPD_FUSE_ADAMW=$fuse timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu --no-secondary > gpurun_out/r2m_bench_fuse$fuse.json 2> gpurun_out/r2m_bench_fuse$fuse.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_fuse$fuse.json')); print('fuse=$fuse', d['ms_per_step'], d['value'], d['kernel_ms'], d['e2e']['ms_per_step'], d['e2e_indexed']['ms_per_step'], d['gpu_launches'], d['run']['final_loss'])"
done
