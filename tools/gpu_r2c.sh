#!/bin/bash
# round 2, final session: full GPU suite, smoke, both bench arms, ncu launch list + captures of the dominant kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -14 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -3 gpurun_out/bench_final.err; head -c 1500 gpurun_out/bench_final.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; cat gpurun_out/bench_ref_final.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active"
MSM_LOG=26 ncu --metrics $M --clock-control none -k regex:k_msm_accumulate --csv --log-file gpurun_out/ncu_accumulate_2p26_metrics.csv python tools/ncu_targets.py msm > /dev/null 2>&1
MSM_LOG=26 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 1 -c 1 -f -o gpurun_out/prof_acc python tools/ncu_targets.py msm > /dev/null 2>&1
ncu -i gpurun_out/prof_acc.ncu-rep --page details > gpurun_out/prof_msm_accumulate_details.txt 2>/dev/null
ncu -i gpurun_out/prof_acc.ncu-rep --page raw --csv > gpurun_out/prof_msm_accumulate_raw.csv 2>/dev/null
rm -f gpurun_out/prof_acc.ncu-rep
tail -3 gpurun_out/ncu_accumulate_2p26_metrics.csv | cut -c1-300
