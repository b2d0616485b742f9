#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_comm.py tests/test_gpu_matchers.py tests/test_gpu_configs.py -x -q > gpurun_out/r2_gputests_g.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_g.log ); tail -3 gpurun_out/r2_gputests_g.log
timeout 300 python tools/prof_kernels.py --what matchers > gpurun_out/r2_matchers.json 2>&1; tail -1 gpurun_out/r2_matchers.json
