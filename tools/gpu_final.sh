#!/bin/bash
# Round-end GPU session: full parity suite, smoke, both bench arms, then the ncu launch list of the bench command and a
# full capture of the FFT pass kernel (the kernel that changed last).  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
./tools/gpu_round.sh
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-extras --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
name=prof_fft
ncu --set full --clock-control none --import-source on -k regex:k_fft_pass -s 0 -c 3 -f -o gpurun_out/$name python tools/ncu_targets.py fft > /dev/null 2>&1
ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/${name}_details.txt 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
rm -f gpurun_out/$name.ncu-rep
du -sh gpurun_out
