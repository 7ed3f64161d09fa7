#!/bin/bash
# round-2 profiling session: ncu --set full of the dominant kernels (C2a chunked GRU fwd / lean fused bwd, C3 delta fwd / bwd),
# and the launch list of a short bench run (per-launch durations: cold-cache, serialised — shares only)
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"gru_fwd_kernel|gru_bwdf_kernel" -c 4 -o gpurun_out/r2_gru -f python scripts/ktime.py dgru 13 64 2048 0,0,64 > gpurun_out/r2_ncu_gru.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"delta_fwd_kernel|delta_bwd_kernel" -c 2 -o gpurun_out/r2_delta -f python scripts/ktime.py deltagru_tcnskip 15 256 2048 1,1,0 > gpurun_out/r2_ncu_delta.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --settle 40 --no-cpu --no-secondary > gpurun_out/r2_launches_bench.log 2>&1
ls -la gpurun_out/r2_gru.ncu-rep gpurun_out/r2_delta.ncu-rep gpurun_out/r2_launches.csv
