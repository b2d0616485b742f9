#!/bin/bash
# Round-2 GPU call 1: tests, bench, launch list, ncu captures of every hot kernel, configs 3 and 5.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests.log )
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
timeout 600 python tools/prof_kernels.py --what config3,config5,small > gpurun_out/r2_prof_kernels.json 2> gpurun_out/r2_prof_kernels.err
# launch list: shares of GPU time (one pass over 64 frames per step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --frames 64 --repeat 1 --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1
# -s counts matching launches: launch 0 is the first warm-up batch of the device-resident path (64 frames); resize has 7 per batch
for k in fast_nms_tma describe_fused resize_level cell_select_kernel level_select_kernel sbp_device; do
    skip=1; [ $k = resize_level ] && skip=7
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/r2_$k \
        python bench.py --frames 64 --repeat 1 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_$k.log 2>&1 || true
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2 -s 1 -c 1 -o gpurun_out/r2_knn2 \
    python tools/prof_kernels.py --what config5 --groups 1000 --iters 1 --warmup 1 > gpurun_out/ncu_knn2.log 2>&1 || true
timeout 900 ncu --set full --clock-control none -k regex:'fast_nms_tma|resize_level|describe_fused' -s 13 -c 13 -o gpurun_out/r2_config3 \
    python tools/prof_kernels.py --what config3 --batch 2 --iters 1 --warmup 1 > gpurun_out/ncu_config3.log 2>&1 || true
timeout 600 ncu --set full --clock-control none -k regex:'bow_descend|distinctive|hamming_csr|undistort|bow_db' -c 6 -o gpurun_out/r2_small \
    python tools/prof_kernels.py --what small > gpurun_out/ncu_small.log 2>&1 || true
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r2_gputests.log
