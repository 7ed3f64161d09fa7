#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_achieved.jsonl
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider > gpurun_out/r2f1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f1_pytest.log
tail -6 gpurun_out/r2f1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f1_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2f1_smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f1_bench_ref.json 2> gpurun_out/r2f1_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f1_bench_k20.json 2> gpurun_out/r2f1_bench_k20.err
echo "bench rc=$?"; grep -v "Backbone\|^$" gpurun_out/r2f1_bench_k20.err | tail -5
timeout 900 python bench.py --steps 400 --warmup 16 --no-secondary > gpurun_out/r2f1_bench_k400.json 2> gpurun_out/r2f1_bench_k400.err
python - <<'PY'
import json
for f in ('r2f1_bench_k20', 'r2f1_bench_k400'):
    d = json.load(open(f'gpurun_out/{f}.json'))
    print(f, {k: d[k] for k in ('value','ms_per_step','kernel_ms','value_sync_loss','gpu_launches')})
    print('  e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_sequential'], 'idx', d['e2e_indexed']['ms_per_step'], 'serial', d['serial_floor']['ms_per_step'])
    print('  chunks', [(c['dir'], c['chunks'], c['warmup_steps'], c['serial_reruns'], round(c['worst_boundary_mismatch_over_tolerance'],3)) for c in d['time_chunks']], d['time_chunk_events'])
    print('  gpu_ref', d.get('gpu_reference',{}).get('ms_per_step'), d.get('gpu_reference',{}).get('native_over_gpu_reference'), 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_c_port',{}).get('value'))
    print('  roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['traffic'], d['clocks'])
    for k,v in d.get('secondary',{}).items(): print('   ', k, v.get('ms_per_step'), v.get('value'), v.get('time_chunks'), v.get('error'), (v.get('gpu_reference') or {}).get('ms_per_step'))
r = json.load(open('gpurun_out/r2f1_bench_ref.json')); print('ref', r['value'], r['ms_per_step'], r['cpu_baseline']['cores'], r['cpu_c_port']['value'])
PY
