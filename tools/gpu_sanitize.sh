#!/bin/bash
# compute-sanitizer over the small parity tests (memcheck: all kernels; racecheck: the shared-memory FFT / MSM reduction)
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_golden.py "tests/test_gpu_group_fft.py::test_group_fft_vs_definition" -x -q > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_memcheck.log
tail -5 gpurun_out/sanitize_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_fft.py::test_fft_golden" "tests/test_gpu_golden.py::test_batch_mul_golden" -x -q > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_racecheck.log
tail -5 gpurun_out/sanitize_racecheck.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
