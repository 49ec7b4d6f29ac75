#!/bin/bash
# G2 with lazy reduction in Fq2 (variant library) vs the default build: parity on the G2 test subset, then timings
for L in "" phase2_bn254_b200/libp2b_lazy.so; do
  if [ -n "$L" ]; then export P2B_LIB=$PWD/$L; echo "== lazy Fq2"; else unset P2B_LIB; echo "== default"; fi
  python -m pytest tests/test_gpu_batch_mul.py tests/test_gpu_msm.py tests/test_gpu_msm_pair.py tests/test_gpu_group_fft.py tests/test_gpu_transform.py -x -q -k "not 2p20" 2>&1 | tail -2
  MSM=20 MSM_G2=20,22,24 python tools/probe.py 18 20 2>&1 | grep -E "G2|MSM"
done
