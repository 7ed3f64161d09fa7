#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -k "gru or train or chunk or iqview or full_size or snippet" > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -25 gpurun_out/r2f_pytest.log
{
python scripts/ktime.py dgru 13 64 2048 1,1,0 8,4,64 8,6,64 8,5,64 8,8,64 0,0,64
python scripts/ktime.py dgru 23 256 2048 1,1,0
python scripts/ktime.py gru 32 8 1024 0,0,0
ODPD_BWD_FORM=0 python scripts/ktime.py dgru 13 64 2048 8,4,64
} > gpurun_out/r2f_ktime.jsonl 2> gpurun_out/r2f_ktime.err
cat gpurun_out/r2f_ktime.jsonl; grep -v Backbone gpurun_out/r2f_ktime.err | tail -5
ncu --metrics gpu__time_duration.sum,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"gru_bwd" -c 4 --csv --log-file gpurun_out/r2f_launches.csv python scripts/ktime.py dgru 13 64 2048 8,6,64 > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2f_launches.csv')))
hdr = next(i for i,r in enumerate(rows) if r and r[0]=='ID')
h = rows[hdr]; ci = {n:i for i,n in enumerate(h)}
for r in rows[hdr+1:]:
    if len(r) > ci['Metric Value']: print(r[ci['Kernel Name']][:40], r[ci['Grid Size']], r[ci['Block Size']], r[ci['Metric Name']], r[ci['Metric Value']])
PY
