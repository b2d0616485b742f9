#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_facade_vs_ref.py -x -q > gpurun_out/r2_facade_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_facade_tests.log )
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gputests_b.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_b.log )
timeout 600 python tools/prof_kernels.py --what config5 > gpurun_out/r2_config5_csa.json 2> gpurun_out/r2_config5_csa.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2 -s 1 -c 1 -o gpurun_out/r2_knn2_csa \
    python tools/prof_kernels.py --what config5 --groups 1000 --iters 1 --warmup 1 > gpurun_out/ncu_knn2_csa.log 2>&1 || true
tail -4 gpurun_out/r2_facade_tests.log; tail -4 gpurun_out/r2_gputests_b.log; cat gpurun_out/r2_config5_csa.json
