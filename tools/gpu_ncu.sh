#!/bin/bash
# ncu evidence: launch list of the bench command + full captures of the dominant kernels.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
cap() {  # name, kernel regex, skip, count, command...
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o gpurun_out/$name "$@" > /dev/null 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv 2>/dev/null | head -c 6000000 > gpurun_out/${name}_source.csv
  ls -la gpurun_out/$name.ncu-rep
  sz=$(stat -c %s gpurun_out/$name.ncu-rep); if [ "$sz" -gt 12000000 ]; then rm gpurun_out/$name.ncu-rep; fi
}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
cap prof_msm_accumulate k_msm_accumulate 1 1 python bench.py --steps 1 --warmup 1 --no-extras --no-cpu
cap prof_batch_mul_g1 k_batch_mul 1 1 python tools/ncu_targets.py g1
cap prof_batch_mul_g2 k_batch_mul 1 1 python tools/ncu_targets.py g2
cap prof_fft k_fft_pass 0 3 python tools/ncu_targets.py fft
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_kernels.csv python tools/ncu_targets.py all > /dev/null 2>&1
for c in 13 14 15 16 17 18 20; do echo "== c=$c"; P2B_MSM_C=$c MSM=22,26 python tools/probe.py 2>&1 | grep MSM; done > gpurun_out/msm_c_sweep.log 2>&1
python tools/probe.py 20 2>&1 | grep -v MSM > gpurun_out/probe_batch.log
du -sh gpurun_out; ls -la gpurun_out
