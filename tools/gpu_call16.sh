#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_matchers.py tests/test_gpu_bench_config.py tests/test_gpu_facade_vs_ref.py tests/test_gpu_comm.py -x -q > gpurun_out/r2_gputests_j.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_j.log ); tail -3 gpurun_out/r2_gputests_j.log
timeout 300 python tools/prof_kernels.py --what matchers,h2d,exchange1 > gpurun_out/r2_matchers.json 2>&1; tail -3 gpurun_out/r2_matchers.json
timeout 600 python bench.py > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err; tail -c 300 gpurun_out/bench_t.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_t.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','parity_checked','host_ms_per_batch')}, d['e2e'])
except Exception as e: print('parse failed',e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbp_device -s 76 -c 1 -f -o gpurun_out/r2_sbp2_new python tools/prof_kernels.py --what matchers > gpurun_out/r2_ncu_sbp2.log 2>&1; tail -1 gpurun_out/r2_ncu_sbp2.log
