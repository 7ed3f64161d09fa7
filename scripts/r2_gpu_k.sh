#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -12 gpurun_out/r2k_pytest.log
for pdl in 1 0; do
ODPD_PDL=$pdl timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu --no-secondary > gpurun_out/r2k_bench_pdl$pdl.json 2> gpurun_out/r2k_bench_pdl$pdl.err
python -c "
import json; d=json.load(open('gpurun_out/r2k_bench_pdl$pdl.json')); print('pdl=$pdl', d['ms_per_step'], d['value'], d['kernel_ms'], d['e2e']['ms_per_step'], d['e2e_indexed']['ms_per_step'], d['serial_floor']['ms_per_step'])"
done
