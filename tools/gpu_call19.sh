#!/bin/bash
mkdir -p gpurun_out
FASTARC_VARIANTS=12,12:3,16:3 timeout 300 python tools/prof_kernels.py --what fastarc > gpurun_out/r2_fastarc2.json 2> gpurun_out/r2_fastarc2.err; tail -c 900 gpurun_out/r2_fastarc2.json; tail -3 gpurun_out/r2_fastarc2.err
( timeout 600 python -m pytest tests/test_gpu_extract.py -x -q > gpurun_out/r2_gputests_l.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_l.log ); tail -3 gpurun_out/r2_gputests_l.log
for cfg in "2 4" "3 4" "2 8" "3 2"; do set -- $cfg
ORBFE_E2E_EXTRACTORS=$1 ORBFE_CHUNKS=$2 timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_v_$1_$2.json 2> gpurun_out/bench_v_$1_$2.err; tail -c 300 gpurun_out/bench_v_$1_$2.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_v_$1_$2.json').read().strip().splitlines()[-1])
    print("nex=$1 chunks=$2", {k:d[k] for k in ('value','parity_checked','host_ms_per_batch')}, d['e2e'])
except Exception as e: print('parse failed',e)
PY
done
ORBFE_E2E_PY=1 timeout 600 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_v_py.json 2> gpurun_out/bench_v_py.err; tail -c 300 gpurun_out/bench_v_py.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_v_py.json').read().strip().splitlines()[-1])
    print("python e2e", {k:d[k] for k in ('value','parity_checked','host_ms_per_batch')}, d['e2e'])
except Exception as e: print('parse failed',e)
PY
