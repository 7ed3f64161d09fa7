#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -k "vdlstm or lstm" > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -40 gpurun_out/r2l_pytest.log
