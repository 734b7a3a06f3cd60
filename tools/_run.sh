timeout 900 python -m pytest tests/test_ransac_gpu.py tests/test_pipeline_gpu.py tests/test_fullsize_gpu.py -m gpu -q 2>&1 | grep -v "^  \|Warning\|^$" | tail -40 > gpurun_out/ransac_test.txt
tail -3 gpurun_out/ransac_test.txt
