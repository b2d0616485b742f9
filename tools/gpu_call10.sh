#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_f.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_f.log ); tail -4 gpurun_out/r2_gputests_f.log
timeout 300 python tools/prof_kernels.py --what small > gpurun_out/r2_small_f.json 2>&1; tail -1 gpurun_out/r2_small_f.json
