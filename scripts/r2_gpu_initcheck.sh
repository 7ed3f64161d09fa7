#!/bin/bash
# initcheck (reads of never-written global memory) over the small-case sweep of every cell
cd /root/repo
timeout 500 compute-sanitizer --tool initcheck --launch-timeout 600 --error-exitcode 0 python scripts/sanitize_small.py > gpurun_out/r2_initcheck.log 2>&1
grep -c 'Uninitialized' gpurun_out/r2_initcheck.log
grep -A3 'Uninitialized' gpurun_out/r2_initcheck.log | grep ' at \| in ' | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -30
tail -3 gpurun_out/r2_initcheck.log
