#!/bin/bash
mkdir -p gpurun_out
python scripts/ktime.py pgjanet 15 128 4096 1,1,0 0,0,0 > gpurun_out/r2_janet_ktime.jsonl 2>/dev/null
python scripts/ktime.py dvrjanet 15 128 4096 1,1,0 0,0,0 >> gpurun_out/r2_janet_ktime.jsonl 2>/dev/null
cat gpurun_out/r2_janet_ktime.jsonl
ncu --set full --import-source on --clock-control none -k regex:"pgjanet_fwd_kernel|pgjanet_bwd_kernel" -c 2 -o gpurun_out/r2_pgjanet -f python scripts/ktime.py pgjanet 15 128 1024 1,1,0 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:"dvrjanet_fwd_kernel|dvrjanet_bwd_kernel" -c 2 -o gpurun_out/r2_dvrjanet -f python scripts/ktime.py dvrjanet 15 128 1024 1,1,0 > /dev/null 2>&1
ls -la gpurun_out/r2_pgjanet.ncu-rep gpurun_out/r2_dvrjanet.ncu-rep
