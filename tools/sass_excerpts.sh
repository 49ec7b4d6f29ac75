#!/bin/bash
# Commits ~40-line SASS excerpts of the hot kernels (from the in-tree objects) under profiles/<round>/sass/: the TMA bulk copy +
# mbarrier of k_batch_mul, the IMAD.WIDE carry chains of the Montgomery multiplier, the prefetch loads of k_msm_accumulate.
out=${1:-profiles/r2a/sass}
mkdir -p $out
cs=phase2_bn254_b200/csrc
ex() {  # object, function regex, grep pattern, output name
  fn=$(cuobjdump -sass $cs/$1 | grep -o "Function : .*" | grep -E "$2" | head -1 | sed 's/Function : //')
  { echo "# $1  $fn"; echo "# (cuobjdump -sass $cs/$1; lines matching /$3/ with 2 lines of context, first 60 lines)";
    cuobjdump -sass -fun "$fn" $cs/$1 | grep -E -B2 -A2 "$3" | head -60; echo "# counts:";
    cuobjdump -sass -fun "$fn" $cs/$1 | grep -oE "IMAD\.WIDE\.U32|IMAD\.WIDE|UBLKCP[.A-Z]*|SYNCS[.A-Z0-9]*|LDG\.E[.0-9A-Z]*|STG\.E[.0-9A-Z]*|SHFL[.A-Z]*|ATOMG[.A-Z0-9]*|LDS[.0-9A-Z]*|STS[.0-9A-Z]*|IADD3[.A-Z]*" | sort | uniq -c | sort -rn | head -12; } > $out/$4.txt
}
ex batch_mul_g1.o "k_batch_mulINS_2FpINS_3FqPEEELi256ELb1ELb0" "UBLKCP|SYNCS" batch_mul_g1_tma
ex batch_mul_g1.o "k_batch_mulINS_2FpINS_3FqPEEELi256ELb1ELb0" "IMAD.WIDE.U32.X" batch_mul_g1_imad_chain
ex msm_g1.o "k_msm_accumulateINS_2FpINS_3FqPEEELi2E" "LDG" msm_accumulate_prefetch
ex msm_g1.o "k_msm_accumulateINS_2FpINS_3FqPEEELi2E" "IMAD.WIDE.U32.X" msm_accumulate_imad_chain
ex msm_g1.o "k_msm_accumulate_pairINS_2FpINS_3FqPEEE" "LDG" msm_accumulate_pair_loads
ex fft.o "k_fft_pass" "IMAD.WIDE.U32.X|LDS|STS" fft_pass
ls -la $out
