#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -30 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --workload c5q --steps 100 --warmup 10 --settle 300 --no-cpu --no-secondary > gpurun_out/r2j_c5q.json 2> gpurun_out/r2j_c5q.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_c5q.json')); print('c5q', d['ms_per_step'], d['value'], d['kernel_ms'])"
timeout 600 python bench.py --workload c3 --steps 50 --warmup 5 --settle 300 --no-cpu --no-secondary > gpurun_out/r2j_c3.json 2> gpurun_out/r2j_c3.err
python -c "
import json; d=json.load(open('gpurun_out/r2j_c3.json')); print('c3', d['ms_per_step'], d['value'], d['kernel_ms'])"
