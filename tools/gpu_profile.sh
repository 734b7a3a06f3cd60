#!/bin/bash
# Run on the GPU box (through gpurun): launch list and ncu --set full captures of the step's kernels and of one BA window.
# usage: tools/gpu_profile.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 500 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --reps 1 --groups 1 --no-cpu --no-single --no-others > gpurun_out/${tag}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'fmat_ransac|pnp_ransac|mindist_fast|region_select|depth_innovation|lk_track_kernel_v4|corner_response|ba_kernel|reprj_inlier|ingest_l1|hist_kernel|scharr' \
  -s 80 -c 18 -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --reps 1 --groups 1 --no-cpu --no-single --no-others > gpurun_out/${tag}_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ba_kernel -s 1 -c 1 -o gpurun_out/${tag}_ba_w10 python tools/ba_profile.py 10 1 > gpurun_out/${tag}_ba_w10.log 2>&1
ls -la gpurun_out | tail -6
