#!/bin/bash
# one GPU session: tests, headline bench, reference arm, all workloads, ncu launch list + full capture of the dominant kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 400 python bench.py --steps 400 --warmup 20 2>gpurun_out/bench_c2a.err | tee gpurun_out/r1_chunked_bench_c2a_1gpu.json | cut -c1-400
STEPS=200 bash scripts/bench_all.sh 2>&1 | tail -12
cp gpurun_out/bench_all.jsonl gpurun_out/r1_chunked_bench_all_workloads.jsonl
# launch list: eager launches (no graph), steady-state warm-up, 7 train steps + the per-kernel section
ODPD_GRAPHS=0 ODPD_TWARM=64 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1_chunked_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
ODPD_GRAPHS=0 ODPD_TWARM=64 timeout 400 ncu --set full --clock-control none --import-source on -k regex:gru_ -s 12 -c 4 -o gpurun_out/r1_chunked_gru python scripts/prof_step.py dgru 13 64 2048 5 > gpurun_out/ncu_chunked.log 2>&1
tail -2 gpurun_out/ncu_chunked.log
