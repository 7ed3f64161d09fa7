#!/bin/bash
mkdir -p gpurun_out
{
ODPD_ROLE_FLIP=2 python scripts/ktime.py dgru 13 64 2048 8,4,64
ODPD_ROLE_FLIP=3 python scripts/ktime.py dgru 13 64 2048 8,4,64
ODPD_ROLE_FLIP=2 python scripts/ktime.py dgru 23 256 2048 1,1,0
} > gpurun_out/r2c_ktime.jsonl 2> gpurun_out/r2c_ktime.err
cat gpurun_out/r2c_ktime.jsonl; tail -3 gpurun_out/r2c_ktime.err
# ncu baselines: C2a chunked kernels and the C3 kernels (full set + source), one step each
ncu --set full --import-source on --clock-control none -k regex:"gru_bwd_kernel|gru_fwd_kernel" -c 4 -o gpurun_out/r2c_gru -f python scripts/ktime.py dgru 13 64 2048 8,4,64 > gpurun_out/r2c_ncu_gru.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"delta_fwd_kernel|delta_bwd_kernel" -c 2 -o gpurun_out/r2c_delta -f python scripts/ktime.py deltagru_tcnskip 15 256 2048 1,1,0 > gpurun_out/r2c_ncu_delta.log 2>&1
ls -la gpurun_out/*.ncu-rep
