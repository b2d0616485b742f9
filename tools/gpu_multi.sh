#!/bin/bash
# usage: gpu_multi.sh N  -- configs 4 and 5 on N GPUs through the C-ABI + a short weak-scaling bench line
NG=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus_$NG.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 tools/multi_gpu.py --what rig,dbsweep --steps 10 --warmup 3 > gpurun_out/r2_multi_gpu_$NG.json 2> gpurun_out/r2_multi_gpu_$NG.err
echo "multi rc=$?"; cut -c1-1700 gpurun_out/r2_multi_gpu_$NG.json; tail -3 gpurun_out/r2_multi_gpu_$NG.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $NG --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n$NG.json 2> gpurun_out/r2_bench_n$NG.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n$NG.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','n_gpus','ms_per_step','host_numa_pinning_rank0','parity_checked')}, d['e2e']['value'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2_bench_n$NG.err').read()[-1500:])
PY
