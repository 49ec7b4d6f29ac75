#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== main"; MSM=20,22,24,26 MSM_G2=20,22 python tools/probe.py 20 22 2>&1 | tail -14
echo "=== variant (G1 128x3, acc 128x4)"; P2B_LIB=$PWD/phase2_bn254_b200/libp2b_v.so MSM=22,26 python tools/probe.py 20 22 2>&1 | grep -E "G1|MSM"
echo "=== L2 fetch 32"; P2B_L2_FETCH=32 MSM=22,26 python tools/probe.py 16 2>&1 | grep -E "MSM"
echo "=== L2 fetch 64"; P2B_L2_FETCH=64 MSM=22,26 python tools/probe.py 16 2>&1 | grep -E "MSM"
for c in 12 13 14 15 16 17; do echo "== c=$c"; P2B_MSM_C=$c MSM=20,22 python tools/probe.py 16 2>&1 | grep MSM; done
for c in 16 17 18 19 20; do echo "== c=$c"; P2B_MSM_C=$c MSM=24,26 python tools/probe.py 16 2>&1 | grep MSM; done
