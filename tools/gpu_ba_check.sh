#!/bin/bash
# BA kernel checks on the GPU box: parity tests and cycle counters for every cluster size
mkdir -p gpurun_out
for c in 4 2 1; do
  echo "== FLV_BA_CLUSTER=$c"
  FLV_BA_CLUSTER=$c timeout 600 python -m pytest tests/test_ba_gpu.py tests/test_localmap.py -m gpu -x -q 2>&1 | tail -4
  FLV_BA_CLUSTER=$c timeout 120 python tools/ba_profile.py 10 1 2>&1 | tail -14
  FLV_BA_CLUSTER=$c timeout 120 python tools/ba_profile.py 20 1 2>&1 | head -1
done
