#!/bin/bash
# round-2 final validation: full GPU suite, smoke, both bench arms, sanitizers
cd /root/repo
rm -f gpurun_out/parity_achieved.jsonl
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_reference_arm.json 2> gpurun_out/r2_final_reference_arm.err; tail -c 300 gpurun_out/r2_final_reference_arm.json; echo
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python - <<'PY'
import json
d=json.loads(open('/root/repo/gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["kernel_ms"], d["clocks"], {k:round(v["ms_per_step"],4) for k,v in d["secondary"].items() if "ms_per_step" in v}, {k:v.get("ms_per_step", v) for k,v in d["secondary"].get("other_backbones",{}).items() if isinstance(v,dict)})
PY
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 900 python scripts/sanitize_small.py > gpurun_out/r2c_memcheck.log 2>&1; tail -2 gpurun_out/r2c_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --launch-timeout 900 python scripts/sanitize_small.py > gpurun_out/r2c_racecheck.log 2>&1; tail -2 gpurun_out/r2c_racecheck.log
