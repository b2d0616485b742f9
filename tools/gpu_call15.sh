#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_matchers.py tests/test_gpu_bench_config.py tests/test_gpu_facade_vs_ref.py -x -q > gpurun_out/r2_gputests_i.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_i.log ); tail -3 gpurun_out/r2_gputests_i.log
timeout 300 python tools/prof_kernels.py --what matchers > gpurun_out/r2_matchers.json 2>&1; tail -1 gpurun_out/r2_matchers.json
ORBFE_SBP_SMEM_KB=100 timeout 300 python tools/prof_kernels.py --what matchers > gpurun_out/r2_matchers_100k.json 2>&1; tail -1 gpurun_out/r2_matchers_100k.json
timeout 600 python bench.py --no-parity > gpurun_out/bench_sm200.json 2> gpurun_out/bench_sm200.err; tail -c 1500 gpurun_out/bench_sm200.json
ORBFE_SBP_SMEM_KB=100 timeout 600 python bench.py --no-parity > gpurun_out/bench_sm100.json 2> gpurun_out/bench_sm100.err; tail -c 600 gpurun_out/bench_sm100.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbp_device -c 1 -f -o gpurun_out/r2_sbp0_new python tools/prof_kernels.py --what matchers > gpurun_out/r2_ncu_sbp0.log 2>&1; tail -1 gpurun_out/r2_ncu_sbp0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbp_device -s 76 -c 1 -f -o gpurun_out/r2_sbp2_new python tools/prof_kernels.py --what matchers > gpurun_out/r2_ncu_sbp2.log 2>&1; tail -1 gpurun_out/r2_ncu_sbp2.log
