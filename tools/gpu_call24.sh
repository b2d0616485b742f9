#!/bin/bash
mkdir -p gpurun_out
NG=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu.py --what rig --steps 20 --warmup 5 > gpurun_out/r2_multi_gpu_${NG}_pipe.json 2> gpurun_out/r2_multi_gpu_${NG}_pipe.err
echo "multi rc=$?"; cut -c1-2200 gpurun_out/r2_multi_gpu_${NG}_pipe.json; tail -3 gpurun_out/r2_multi_gpu_${NG}_pipe.err | cut -c1-300
