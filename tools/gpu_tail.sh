#!/bin/bash
python -m pytest tests/test_gpu_msm.py tests/test_gpu_msm_pair.py -x -q 2>&1 | tail -2
for t in 0 1; do echo "== P2B_MSM_TAIL=$t"; P2B_MSM_TAIL=$t python tools/e2e_probe.py 26 25 2>&1 | grep -E "max chunk|device-resident [0-9]"; done
echo "== adaptive (no override)"; python tools/e2e_probe.py 26 25 2>&1 | grep -E "max chunk"
