#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --launch-timeout 900 python scripts/sanitize_small.py > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r2_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --launch-timeout 900 python scripts/sanitize_small.py > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -v "^Backbone" gpurun_out/r2_racecheck.log | tail -25
for form in 0 1; do
ODPD_BWD_FORM=$form timeout 900 compute-sanitizer --tool racecheck --launch-timeout 900 python scripts/sanitize_small.py dgru gru qgru > gpurun_out/r2_racecheck_form$form.log 2>&1
echo "racecheck form $form rc=$?"; grep "RACECHECK SUMMARY\|Error" gpurun_out/r2_racecheck_form$form.log | head -5
done
