#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_matchers.py tests/test_gpu_bench_config.py tests/test_gpu_facade_vs_ref.py -x -q > gpurun_out/r2_gputests_h.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_h.log ); tail -3 gpurun_out/r2_gputests_h.log
timeout 300 python tools/prof_kernels.py --what matchers > gpurun_out/r2_matchers.json 2>&1; tail -1 gpurun_out/r2_matchers.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbp_device -c 1 -f -o gpurun_out/r2_sbp0_new python tools/prof_kernels.py --what matchers > gpurun_out/r2_ncu_sbp0.log 2>&1; tail -2 gpurun_out/r2_ncu_sbp0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbp_device -s 76 -c 1 -f -o gpurun_out/r2_sbp2_new python tools/prof_kernels.py --what matchers > gpurun_out/r2_ncu_sbp2.log 2>&1; tail -2 gpurun_out/r2_ncu_sbp2.log
