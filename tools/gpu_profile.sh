#!/bin/bash
# Run on the GPU box (through gpurun): GPU tests, launch list and one ncu --set full capture of the step's kernels.
# usage: tools/gpu_profile.sh <tag> [skip_tests]
tag=${1:-r02}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1
  tail -3 gpurun_out/${tag}_gputests.log
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --reps 1 --groups 1 --no-cpu --no-single > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'fmat_ransac|pnp_ransac|mindist_fast|region_select|depth_innovation|lk_track_kernel_v4|corner_response|ba_kernel|reprj_inlier' \
  -s 60 -c 14 -o gpurun_out/${tag}_full python bench.py --steps 2 --warmup 3 --reps 1 --groups 1 --no-cpu --no-single > gpurun_out/${tag}_full.log 2>&1
ls -la gpurun_out | tail -5
