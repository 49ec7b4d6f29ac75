#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== main"; python tools/probe.py 20 22 2>&1 | grep G1
echo "=== variant (G1 128x3 + carveout)"; P2B_LIB=$PWD/phase2_bn254_b200/libp2b_v.so python tools/probe.py 20 22 2>&1 | grep -E "G1"
