#!/bin/bash
# round-2 session script: memcheck + racecheck over every cell incl. the layered path and the row f-4 cells
cd /root/repo
timeout 1500 compute-sanitizer --tool memcheck --launch-timeout 900 python scripts/sanitize_small.py > gpurun_out/r2b_memcheck.log 2>&1
tail -4 gpurun_out/r2b_memcheck.log
timeout 2400 compute-sanitizer --tool racecheck --launch-timeout 900 python scripts/sanitize_small.py > gpurun_out/r2b_racecheck.log 2>&1
tail -4 gpurun_out/r2b_racecheck.log
grep -c "Race reported\|hazard" gpurun_out/r2b_racecheck.log
