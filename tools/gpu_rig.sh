#!/bin/bash
# usage: gpu_rig.sh N -- configs[3] (camera rig) on N GPUs: serial step (NCCL / fused exchange) and the pipelined throughput form
NG=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu.py --what rig --steps 20 --warmup 5 > gpurun_out/r2_multi_gpu_${NG}_pipelined.json 2> gpurun_out/r2_multi_gpu_${NG}_pipelined.err
echo "multi rc=$?"; cut -c1-2400 gpurun_out/r2_multi_gpu_${NG}_pipelined.json; tail -3 gpurun_out/r2_multi_gpu_${NG}_pipelined.err | cut -c1-300
