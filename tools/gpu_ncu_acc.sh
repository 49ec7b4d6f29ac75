#!/bin/bash
# refresh of the accumulate capture + launch list only
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
name=prof_msm_accumulate
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 3 -c 1 -f -o gpurun_out/$name python bench.py --steps 1 --warmup 1 --no-extras --no-cpu > /dev/null 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
rm -f gpurun_out/$name.ncu-rep
grep -E "dram__bytes_read.sum |dram__bytes_write.sum |gpu__time_duration.sum|fmaheavy_cycles_active.avg.pct" gpurun_out/${name}_details.txt | head; 
