#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_kernels.py --what matchers > gpurun_out/r2_matchers.json 2>&1; tail -1 gpurun_out/r2_matchers.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sbp_device_kernel' -c 40 -o gpurun_out/r2_matchers_ncu \
    python tools/prof_kernels.py --what matchers > gpurun_out/ncu_matchers.log 2>&1 || true
ls -la gpurun_out/r2_matchers_ncu.ncu-rep
