#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_d.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_d.log )
tail -4 gpurun_out/r2_gputests_d.log
timeout 300 python tools/prof_kernels.py --what exchange1 --iters 20 --warmup 5 > gpurun_out/r2_exchange1.json 2>&1; tail -1 gpurun_out/r2_exchange1.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; tail -c 300 gpurun_out/r2_bench_d.err
for cfg in "2 8" "3 8" "2 4" "2 16" "3 16"; do set -- $cfg
  ORBFE_E2E_EXTRACTORS=$1 ORBFE_CHUNKS=$2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_e2e_$1_$2.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r2_e2e_$1_$2.json').read().strip().splitlines()[-1]); print('extractors $1 chunks $2: value %.2f e2e %.2f' % (d['value'], d['e2e']['value']))"
done
