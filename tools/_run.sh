timeout 900 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -k "identity or ba_full" 2>&1 | grep -B30 "Error" | head -90 > gpurun_out/fullsize_fail.txt
