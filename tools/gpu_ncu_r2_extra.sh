#!/bin/bash
# ncu --set full details of the kernels that changed most in round 2: G1 batch_exp (fused doubling / addition), the two-lane G2 accumulation
mkdir -p gpurun_out
cap() { name=$1; rx=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > /dev/null 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  rm -f gpurun_out/$name.ncu-rep; }
cap prof_batch_mul_g1 k_batch_mul 1 python tools/ncu_targets.py g1
cap prof_batch_mul_g2 k_batch_mul 1 python tools/ncu_targets.py g2
MSM_LOG=22 cap prof_msm_accumulate_g2x2 k_msm_accumulate_g2x2 1 python tools/ncu_targets.py msm_g2
grep -E "Duration|Compute \(SM\)|Registers Per|Achieved Occ|DRAM Throughput" gpurun_out/prof_batch_mul_g1_details.txt gpurun_out/prof_batch_mul_g2_details.txt gpurun_out/prof_msm_accumulate_g2x2_details.txt
