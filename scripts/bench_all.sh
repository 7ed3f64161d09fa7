#!/bin/bash
# secondary workloads, 1 GPU, no CPU leg; one JSON line each -> gpurun_out/bench_all.jsonl
mkdir -p gpurun_out
: > gpurun_out/bench_all.jsonl
for w in c1 c2a c2b c3 c3b c3s c4p c4d c5g c5q lstm; do
  timeout 300 python bench.py --workload $w --steps ${STEPS:-100} --warmup 5 --no-cpu 2>>gpurun_out/bench_all.err | tail -1 >> gpurun_out/bench_all.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad line', l[:200]); continue
    print(f"{d['config']['workload'][:70]:70s} {d['ms_per_step']:8.3f} ms/step {d['value']:.3e} samp/s  e2e {d['e2e']['value']:.3e}  fwd {d['kernel_ms']['fwd']:.3f} bwd {d['kernel_ms']['bwd']:.3f} loss {d['config']['final_loss']:.4g}")
PY
