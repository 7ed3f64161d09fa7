#!/bin/bash
# round-2 GPU session A: tests, first real-data bench line, sanitizer
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -30 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/r2a_bench.err; head -c 3000 gpurun_out/r2a_bench.json
timeout 600 compute-sanitizer --tool memcheck --launch-timeout 900 python scripts/sanitize_small.py > gpurun_out/r2a_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r2a_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --launch-timeout 900 python scripts/sanitize_small.py > gpurun_out/r2a_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -5 gpurun_out/r2a_racecheck.log
