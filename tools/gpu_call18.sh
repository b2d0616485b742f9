#!/bin/bash
mkdir -p gpurun_out
FASTARC_VARIANTS=12 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fast_nms_tma -s 4 -c 1 -f -o gpurun_out/r2_fast_arc12 python tools/prof_kernels.py --what fastarc > gpurun_out/r2_ncu_fast_arc12.log 2>&1; tail -2 gpurun_out/r2_ncu_fast_arc12.log
ls -la gpurun_out/*.ncu-rep
