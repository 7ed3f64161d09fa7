#!/bin/bash
# racecheck over the frame-length edge cases (partial blocks, single-block pipelines) + the new quantised TRes sizes
cd /root/repo
timeout 240 compute-sanitizer --tool racecheck --launch-timeout 600 --error-exitcode 0 python -m pytest tests/test_gpu_edges.py -q -m gpu --tb=line -p no:cacheprovider -k "not more_sequences" > gpurun_out/r2_edges_racecheck.log 2>&1
grep -c 'hazard' gpurun_out/r2_edges_racecheck.log
grep -v Initialized gpurun_out/r2_edges_racecheck.log | grep -v 'Host Frame' | tail -25
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu --tb=short -k tres_qat 2>&1 | grep -v 'Initialized\|Replace\|quant the\|INT Quant\|No pretrained' | tail -15
