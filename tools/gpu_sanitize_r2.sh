#!/bin/bash
# round 2: compute-sanitizer over the NEW kernels / host paths (pair MSM, device ChaCha20, staged host I/O, gfft stage, raw codec)
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest \
  "tests/test_gpu_msm_pair.py::test_random_scalars_match_chacha20" "tests/test_gpu_msm_pair.py::test_msm_pair_given_scalars" \
  "tests/test_gpu_msm_pair.py::test_power_pairs_matches_two_msms_and_the_ratio" "tests/test_gpu_msm_pair.py::test_pair_streamed_chunks_and_errors" \
  "tests/test_gpu_hostio.py::test_msm_streamed_from_memory_map" "tests/test_gpu_raw_enc.py::test_raw_batch_mul_per_point_and_broadcast" \
  "tests/test_gpu_group_fft.py::test_gfft_stage_matches_oracle" "tests/test_gpu_msm.py::test_msm_matches_oracle" \
  "tests/test_gpu_transform.py::test_phase2_contribute_sharded_equals_whole" -x -q > gpurun_out/sanitize_r2_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_r2_memcheck.log; tail -4 gpurun_out/sanitize_r2_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_msm_pair.py::test_msm_pair_given_scalars" \
  "tests/test_gpu_msm.py::test_msm_edge_cases" -x -q > gpurun_out/sanitize_r2_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_r2_racecheck.log; tail -3 gpurun_out/sanitize_r2_racecheck.log
compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest "tests/test_gpu_msm_pair.py::test_msm_pair_device_coefficients" \
  "tests/test_gpu_msm.py::test_msm_matches_oracle" -x -q > gpurun_out/sanitize_r2_initcheck.log 2>&1
echo "initcheck rc=$?" | tee -a gpurun_out/sanitize_r2_initcheck.log; tail -3 gpurun_out/sanitize_r2_initcheck.log
