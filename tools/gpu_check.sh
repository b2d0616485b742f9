#!/bin/bash
# quick validation on one B200: GPU tests, the bench line, the latency-bound kernels
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "rc=$?" >> gpurun_out/gputests.log ); tail -4 gpurun_out/gputests.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    r=d['roofline']; print({k:d[k] for k in ('value','ms_per_step','parity_checked','single_frame_latency_ms')}, d['e2e']['value'], r['stage_ms_per_batch_extractor_alone'], r['frac'], r['extract_all_kernels']['frac'])
except Exception as e: print('parse failed',e)
PY
timeout 300 python tools/prof_kernels.py --what small,matchers,latency > gpurun_out/prof_small.json 2>&1; tail -2 gpurun_out/prof_small.json | cut -c1-900
ORBFE_PDL=0 timeout 300 python tools/prof_kernels.py --what latency > gpurun_out/prof_latency_nopdl.json 2>&1; tail -1 gpurun_out/prof_latency_nopdl.json | cut -c1-700
ORBFE_PDL=0 timeout 600 python bench.py --no-cpu-baseline --no-parity --steps 10 > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_nopdl.json').read().strip().splitlines()[-1])
    print("PDL off:", {k:d[k] for k in ('value','single_frame_latency_ms')}, d['e2e']['value'])
except Exception as e: print('parse failed',e)
PY
