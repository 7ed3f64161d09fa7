#!/bin/bash
# frame-length / batch-size edge cases of every backbone (tests/test_gpu_edges.py) + the eval-segment length
cd /root/repo
(timeout 600 python -m pytest tests/test_gpu_edges.py -q -m gpu --tb=short 2>&1 | grep -v Initialized | tail -30
 timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --tb=short -k eval_path 2>&1 | grep -v Initialized | tail -30) > gpurun_out/r2_edges.log 2>&1
tail -c 4000 gpurun_out/r2_edges.log
