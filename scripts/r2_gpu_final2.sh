#!/bin/bash
# end-of-round validation: full GPU suite, smoke, racecheck over the edge cases, one bench line
cd /root/repo
rm -f gpurun_out/parity_achieved.jsonl
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 compute-sanitizer --tool racecheck --launch-timeout 600 --error-exitcode 0 python -m pytest tests/test_gpu_edges.py -q -m gpu --tb=line -p no:cacheprovider -k "not more_sequences" > gpurun_out/r2_edges_racecheck.log 2>&1
grep -v 'Initialized\|Host Frame' gpurun_out/r2_edges_racecheck.log | tail -4
timeout 150 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["kernel_ms"], {k:round(v["ms_per_step"],4) for k,v in d["secondary"].items() if "ms_per_step" in v})
PY
