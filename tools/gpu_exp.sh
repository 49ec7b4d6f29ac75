#!/bin/bash
python -m pytest tests/test_gpu_msm.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
echo "=== default rule"; MSM=20,22,24,26 python tools/probe.py 16 2>&1 | grep MSM
for c in 15 16 17; do echo "== c=$c"; P2B_MSM_C=$c MSM=20,22 python tools/probe.py 16 2>&1 | grep MSM; done
for c in 16 17 19 20; do echo "== c=$c"; P2B_MSM_C=$c MSM=24,26 python tools/probe.py 16 2>&1 | grep MSM; done
