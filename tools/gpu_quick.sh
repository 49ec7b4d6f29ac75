#!/bin/bash
# GPU session: MSM / FFT / C++ mirror parity after a kernel change, then a short bench (no CPU legs).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_gpu_fft.py tests/test_cpp_mirror.py tests/test_gpu_golden.py -x -q 2>&1 | tail -15 > gpurun_out/pytest_quick.log; cat gpurun_out/pytest_quick.log
timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -5 gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json
