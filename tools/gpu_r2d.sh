#!/bin/bash
# round 2, last session (G2 batch subgroup probe): what was run on the B200, one gpurun call per block (GPU minutes were short).
mkdir -p gpurun_out
# 1. raw pipe rates (is the FP64 pipe a second multiplier?)      -> profiles/r2b/pipe_bench.txt
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/pipe_bench tools/pipe_bench.cu && ./tools/bin/pipe_bench > gpurun_out/pipe_bench.txt 2>&1
# 2. probe tests + G2 batch_exp timings                            -> profiles/r2b/pytest_gpu_probe.log, g2_probe_bench.json
python -m pytest tests/test_gpu_g2_probe.py tests/test_gpu_batch_mul.py -x -q > gpurun_out/probe_tests.log 2>&1
python tools/g2_probe_bench.py > gpurun_out/g2_probe_bench.json 2> gpurun_out/g2_probe_bench.err
# 3. full suite + smoke                                            -> profiles/r2b/pytest_gpu_final.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2b.log 2>&1
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
# 4. bench                                                         -> profiles/r2b/bench_final.json
python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
# 5. accumulate capture of one whole 2^26 MSM, then `python tools/update_traffic.py profiles/r2b/ncu_accumulate_2p26_metrics.csv 26`
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
MSM_LOG=26 ncu --metrics $M --clock-control none -k regex:k_msm_accumulate --csv --log-file gpurun_out/ncu_accumulate_2p26_metrics.csv python tools/ncu_targets.py msm > /dev/null 2>&1
