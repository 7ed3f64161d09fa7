#!/bin/bash
# memcheck over the frame-length / batch-size edge cases (short frames are where an unguarded lane reads past a buffer without faulting)
cd /root/repo
timeout 500 compute-sanitizer --tool memcheck --launch-timeout 600 --error-exitcode 0 python -m pytest tests/test_gpu_edges.py -q -m gpu --tb=line -p no:cacheprovider > gpurun_out/r2_edges_memcheck.log 2>&1
grep -c 'Invalid\|misaligned' gpurun_out/r2_edges_memcheck.log
grep -v Initialized gpurun_out/r2_edges_memcheck.log | tail -12
