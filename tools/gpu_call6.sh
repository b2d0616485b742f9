#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_c.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_c.log )
tail -4 gpurun_out/r2_gputests_c.log
timeout 300 python tools/prof_kernels.py --what small > gpurun_out/r2_small_c.json 2>&1
cat gpurun_out/r2_small_c.json | tail -2
NG=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu.py --what rig --steps 10 --warmup 3 > gpurun_out/r2_multi_gpu_$NG.json 2> gpurun_out/r2_multi_gpu_$NG.err
echo "multi rc=$?"; cat gpurun_out/r2_multi_gpu_$NG.json | cut -c1-1500
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -c 600 gpurun_out/r2_bench_c.err
