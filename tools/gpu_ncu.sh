#!/bin/bash
# ncu evidence: launch list of the bench command + full captures of the dominant kernels.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/$name "$@" > /dev/null 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/${name}_source.csv
  rm -f gpurun_out/$name.ncu-rep
}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
cap prof_msm_accumulate k_msm_accumulate 3 3 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu
cap prof_batch_mul_g1 k_batch_mul 1 1 python tools/ncu_targets.py g1
cap prof_batch_mul_g2 k_batch_mul 1 1 python tools/ncu_targets.py g2
cap prof_fft k_fft_pass 0 3 python tools/ncu_targets.py fft
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_kernels.csv python tools/ncu_targets.py all > /dev/null 2>&1
# DRAM bytes of the accumulate kernel with the default and the 32-byte L2 fetch granularity (2^24 terms)
for g in default 32; do
  if [ $g = 32 ]; then export P2B_L2_FETCH=32; fi
  MSM_LOG=24 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_msm_accumulate --csv \
      --log-file gpurun_out/acc_dram_$g.csv python tools/ncu_targets.py msm > /dev/null 2>&1
done
unset P2B_L2_FETCH
du -sh gpurun_out
