#!/bin/bash
# round 2, session B: new pair-MSM tests + whole suite, verifier timings at scale, DRAM traffic of the accumulation
mkdir -p gpurun_out
python -m pytest tests/test_gpu_msm_pair.py tests/test_gpu_verify.py tests/test_gpu_msm.py -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python tools/contribute_scale.py --log-m 22 --verify 2>&1 | tail -2 | tee gpurun_out/contribute_2p22.json
python tools/contribute_scale.py --log-m 26 --verify 2>&1 | tail -2 | tee gpurun_out/contribute_2p26.json
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -3 gpurun_out/bench_b.err; cat gpurun_out/bench_b.json
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct"
for g in default 32 64; do
  if [ $g != default ]; then export P2B_L2_FETCH=$g; fi
  MSM_LOG=26 ncu --metrics $M --clock-control none -k regex:k_msm_accumulate --csv --log-file gpurun_out/acc_dram_$g.csv python tools/ncu_targets.py msm > /dev/null 2>&1
done
unset P2B_L2_FETCH
tail -4 gpurun_out/acc_dram_*.csv
