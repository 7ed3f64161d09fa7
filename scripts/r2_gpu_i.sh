#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"delta_fwd_kernel|delta_bwd_kernel" -c 2 -o gpurun_out/r2i_delta -f python scripts/ktime.py deltagru_tcnskip 15 256 2048 1,1,0 > gpurun_out/r2i_ncu_delta.log 2>&1
ls -la gpurun_out/r2i_delta.ncu-rep
