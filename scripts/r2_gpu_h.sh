#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -k "delta or tres or golden or full_size or train or iqview" > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -8 gpurun_out/r2h_pytest.log
{
python scripts/ktime.py deltagru_tcnskip 15 256 2048 1,1,0
python scripts/ktime.py deltagru_tcnskip 15 128 2048 1,1,0
} > gpurun_out/r2h_ktime.jsonl 2> gpurun_out/r2h_ktime.err
cat gpurun_out/r2h_ktime.jsonl; grep -v Backbone gpurun_out/r2h_ktime.err | tail -5
