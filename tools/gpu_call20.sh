#!/bin/bash
mkdir -p gpurun_out
for cfg in "3 2 3" "3 2 4" "4 2 3" "4 1 4" "4 2 4" "3 1 3"; do set -- $cfg
ORBFE_E2E_EXTRACTORS=$1 ORBFE_CHUNKS=$2 ORBFE_E2E_MATCHERS=$3 timeout 600 python bench.py --no-cpu-baseline --no-parity --steps 10 > gpurun_out/bench_w_$1_$2_$3.json 2> gpurun_out/bench_w_$1_$2_$3.err; tail -c 300 gpurun_out/bench_w_$1_$2_$3.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_w_$1_$2_$3.json').read().strip().splitlines()[-1])
    print("nex=$1 chunks=$2 nmatch=$3", d['value'], d['host_ms_per_batch'], d['e2e']['value'])
except Exception as e: print('parse failed',e)
PY
done
