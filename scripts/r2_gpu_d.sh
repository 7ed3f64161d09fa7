#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -25 gpurun_out/r2d_pytest.log
{
python scripts/ktime.py dgru 13 64 2048 1,1,0 8,8,64 8,4,64 8,6,64 8,16,64 8,10,64
ODPD_BWD_SPLIT=0 python scripts/ktime.py dgru 13 64 2048 8,4,64
python scripts/ktime.py dgru 23 256 2048 1,1,0 2,2,128 1,2,128 1,3,128
python scripts/ktime.py gru 32 8 1024 0,0,0
} > gpurun_out/r2d_ktime.jsonl 2> gpurun_out/r2d_ktime.err
cat gpurun_out/r2d_ktime.jsonl; grep -v Backbone gpurun_out/r2d_ktime.err | tail -5
