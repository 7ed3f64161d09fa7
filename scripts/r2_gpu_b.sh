#!/bin/bash
mkdir -p gpurun_out
{
python scripts/ktime.py deltagru_tcnskip 15 128 2048 1,1,0
python scripts/ktime.py deltagru_tcnskip 15 256 2048 1,1,0
python scripts/ktime.py deltagru_tcnskip 15 296 2048 1,1,0
python scripts/ktime.py dgru 23 256 2048 1,1,0 0,0,0 2,2,128
python scripts/ktime.py dgru 13 64 2048 1,1,0 8,4,64 4,4,64 16,8,64 8,8,64
ODPD_ROLE_FLIP=1 python scripts/ktime.py dgru 13 64 2048 1,1,0 8,4,64
python scripts/ktime.py dgru 13 148 2048 1,1,0
} > gpurun_out/r2b_ktime.jsonl 2> gpurun_out/r2b_ktime.err
cat gpurun_out/r2b_ktime.jsonl; tail -3 gpurun_out/r2b_ktime.err
