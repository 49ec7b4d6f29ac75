#!/bin/bash
python -m pytest tests/test_gpu_msm.py tests/test_gpu_msm_pair.py "tests/test_gpu_config_sizes.py::test_msm_2p20_matches_oracle" -x -q 2>&1 | tail -3
for v in 0 9; do echo "== G2 accumulate variant $v (0 = two lanes per bucket, 9 = one thread per bucket)"; P2B_ACC_VARIANT_G2=$v MSM=20 MSM_G2=20,22,24,25 python tools/probe.py 10 2>&1 | grep -E "MSM G2"; done
