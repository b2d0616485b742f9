#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_comm.py tests/test_gpu_facade_vs_ref.py -q > gpurun_out/r2_comm_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_comm_tests.log )
tail -5 gpurun_out/r2_comm_tests.log
NG=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu.py --what rig,dbsweep --steps 10 --warmup 3 > gpurun_out/r2_multi_gpu_$NG.json 2> gpurun_out/r2_multi_gpu_$NG.err
echo "multi rc=$?"; cat gpurun_out/r2_multi_gpu_$NG.json; tail -5 gpurun_out/r2_multi_gpu_$NG.err
