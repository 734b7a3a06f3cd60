timeout 300 python -m pytest tests/test_ransac_gpu.py -m gpu -q -k fundamental 2>&1 | grep -B2 -A12 "^>" | head -60 > gpurun_out/ransac_fail.txt
