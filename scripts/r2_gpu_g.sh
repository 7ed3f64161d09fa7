#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -15 gpurun_out/r2g_pytest.log
{
python scripts/ktime.py deltagru_tcnskip 15 256 2048 1,1,0
python scripts/ktime.py deltagru_tcnskip 15 128 2048 1,1,0
python scripts/ktime.py deltagru 15 256 2048 1,1,0
} > gpurun_out/r2g_ktime.jsonl 2> gpurun_out/r2g_ktime.err
cat gpurun_out/r2g_ktime.jsonl; grep -v Backbone gpurun_out/r2g_ktime.err | tail -5
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2g_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2g_bench.json'))
print({k: d[k] for k in ('value','ms_per_step','kernel_ms','value_sync_loss','serial_floor')})
print(d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_indexed']['ms_per_step'])
print(d['time_chunks']); print(d['time_chunk_events'])
print(d.get('gpu_reference'))
for k,v in d.get('secondary',{}).items(): print(k, v.get('ms_per_step'), v.get('value'), v.get('time_chunks'), v.get('error'))
PY
