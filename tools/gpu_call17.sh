#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_kernels.py --what fastarc > gpurun_out/r2_fastarc.json 2> gpurun_out/r2_fastarc.err; tail -c 1800 gpurun_out/r2_fastarc.json; tail -3 gpurun_out/r2_fastarc.err
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_alu tools/ubench_alu.cu && /tmp/ubench_alu > gpurun_out/r2_ubench_alu.txt 2>&1; cat gpurun_out/r2_ubench_alu.txt
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_k.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_k.log ); tail -4 gpurun_out/r2_gputests_k.log
timeout 600 python bench.py > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; tail -c 300 gpurun_out/bench_u.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_u.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','parity_checked','host_ms_per_batch')}, d['e2e'], d['roofline']['stage_ms_per_batch_extractor_alone'], d['cpu_baseline'])
except Exception as e: print('parse failed',e)
PY
timeout 200 python tools/e2e_probe.py > gpurun_out/r2_e2e_probe.txt 2>&1; cat gpurun_out/r2_e2e_probe.txt
