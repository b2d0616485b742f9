#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_comm.py -x -q > gpurun_out/r2_gputests_e.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_e.log ); tail -3 gpurun_out/r2_gputests_e.log
timeout 300 python tools/prof_kernels.py --what exchange1 --iters 20 --warmup 5 > gpurun_out/r2_exchange1.json 2>&1; tail -1 gpurun_out/r2_exchange1.json
for cfg in "2 4" "2 2" "2 1" "2 3" "1 4"; do set -- $cfg
  ORBFE_E2E_EXTRACTORS=$1 ORBFE_CHUNKS=$2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_e2e_$1_$2.json 2>/dev/null
  python -c "
import json; d=json.loads(open('gpurun_out/r2_e2e_$1_$2.json').read().strip().splitlines()[-1]); print('extractors $1 chunks $2: value %.2f e2e %.2f' % (d['value'], d['e2e']['value']))"
done
