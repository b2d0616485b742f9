#!/bin/bash
# Re-captures the ncu evidence under profiles/ (run through gpurun on ONE B200; see profiles/README.md):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash tools/refresh_profiles.sh'
# then, back in the dev container, export the raw pages with
#   ncu -i gpurun_out/r1_<kernel>.ncu-rep --page raw --csv > profiles/r1_<kernel>_raw.csv
# Numbers printed under ncu are never bench values (kernels are serialised and replayed).
set -e
mkdir -p gpurun_out
# launch list of the final kernels: shares of GPU time (the first 60 launches are warm-up)
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r1_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# one full capture per kernel of interest (-s skips the warm-up launches of that kernel)
for k in fast_nms_tma describe_fused resize_level cell_select sbp_device; do
    ncu --set full --clock-control none --import-source on -k regex:$k -s 9 -c 1 -o gpurun_out/r1_$k \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1 || true
done
ls -la gpurun_out/*.ncu-rep
