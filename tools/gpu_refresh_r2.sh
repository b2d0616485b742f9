#!/bin/bash
# Round-2 evidence refresh on ONE B200 (through gpurun): GPU tests, the bench line, per-kernel CUDA-event numbers, the ncu launch
# list and full captures of every stream kernel.  Export the raw pages afterwards with tools/export_profiles.sh.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_final.log 2>&1; echo "rc=$?" >> gpurun_out/r2_gputests_final.log ); tail -3 gpurun_out/r2_gputests_final.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 400 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','parity_checked')}, d['e2e'], d['roofline']['stage_ms_per_batch_extractor_alone'])
except Exception as e: print('parse failed',e)
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 600 gpurun_out/r2_bench_ref.json
FASTARC_VARIANTS=-1:4,12:4,16:3 timeout 600 python tools/prof_kernels.py --what config3,config5,small,matchers,fastarc > gpurun_out/r2_prof_final.json 2> gpurun_out/r2_prof_final.err; cut -c1-900 gpurun_out/r2_prof_final.json; tail -2 gpurun_out/r2_prof_final.err
# launch list: shares of GPU time (numbers under ncu are never bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-200
# one 64-frame device-resident batch, every kernel of the library: 7 resize + fast + quota + cell_select + level_select + describe + sbp
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'resize_level|fast_nms|cell_quota|cell_select|level_select|describe_fused|sbp_device' -s 13 -c 13 -f -o gpurun_out/r2_stream_batch \
    python bench.py --frames 64 --repeat 1 --steps 1 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_stream_batch.log 2>&1; tail -1 gpurun_out/ncu_stream_batch.log | cut -c1-200
# matcher kernels at the reference's own call shape: one pair per launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbp_device -s 6 -c 1 -f -o gpurun_out/r2_sbp0_single python tools/prof_kernels.py --what matchers > gpurun_out/ncu_sbp0.log 2>&1; tail -1 gpurun_out/ncu_sbp0.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sbp_device -s 81 -c 1 -f -o gpurun_out/r2_sbp2_single python tools/prof_kernels.py --what matchers > gpurun_out/ncu_sbp2.log 2>&1; tail -1 gpurun_out/ncu_sbp2.log | cut -c1-200
# config 3 (4K, 12 levels): one 8-frame batch = 11 resize + 5 others
timeout 600 ncu --set full --clock-control none -k regex:'resize_level|fast_nms|cell_quota|cell_select|level_select|describe_fused' -s 32 -c 16 -f -o gpurun_out/r2_config3 python tools/prof_kernels.py --what config3 > gpurun_out/ncu_config3.log 2>&1; tail -1 gpurun_out/ncu_config3.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
