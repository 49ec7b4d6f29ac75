#!/bin/bash
# compute-sanitizer over the small parity tests (memcheck: all kernel families; racecheck: shared-memory kernels)
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_golden.py "tests/test_gpu_group_fft.py::test_group_fft_vs_definition" \
  "tests/test_gpu_msm.py::test_msm_skewed_scalars" "tests/test_gpu_msm.py::test_msm_streamed_host_path" "tests/test_gpu_transform.py::test_bulk_codec" -x -q > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_memcheck.log; tail -4 gpurun_out/sanitize_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_fft.py::test_fft_golden" "tests/test_gpu_golden.py::test_batch_mul_golden" \
  "tests/test_gpu_msm.py::test_msm_skewed_scalars" -x -q > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_racecheck.log; tail -4 gpurun_out/sanitize_racecheck.log
compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest "tests/test_gpu_golden.py::test_batch_mul_golden" -x -q > gpurun_out/sanitize_synccheck.log 2>&1
echo "synccheck rc=$?" | tee -a gpurun_out/sanitize_synccheck.log; tail -3 gpurun_out/sanitize_synccheck.log
