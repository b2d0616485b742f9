#!/bin/bash
# 2-GPU box: the headline bench line at N=1 (corrected stage accounting) and N=2 (C++ e2e driver under torchrun), configs 3/4 at N=2
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
    r=d['roofline']; print({k:d[k] for k in ('value','ms_per_step','parity_checked')}, d['e2e']['value'], {k:r[k] for k in ('bound','achieved','frac','traffic','launch_ms')}, r['alu'], r['extract_all_kernels'])
except Exception as e: print('parse failed',e)
PY
bash tools/gpu_multi.sh 2
