timeout 600 ncu --set full --clock-control none --import-source on -k regex:ba_kernel -s 1 -c 1 -o gpurun_out/ba_v5 -f python tools/ba_profile.py 10 1 2>&1 | tail -5
