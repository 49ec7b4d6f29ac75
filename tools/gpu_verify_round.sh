#!/bin/bash
# GPU session: the verifier tests, then phase-2 contribute at the north-star size (one GPU).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_verify.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_verify.log; cat gpurun_out/pytest_verify.log
timeout 300 python tools/contribute_scale.py --log-m 22 --verify > gpurun_out/contribute_2p22.json 2> gpurun_out/contribute_2p22.err; tail -3 gpurun_out/contribute_2p22.err; cat gpurun_out/contribute_2p22.json
timeout 420 python tools/contribute_scale.py --log-m 26 --verify > gpurun_out/contribute_2p26.json 2> gpurun_out/contribute_2p26.err; tail -3 gpurun_out/contribute_2p26.err; cat gpurun_out/contribute_2p26.json
