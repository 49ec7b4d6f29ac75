#!/bin/bash
# round 2, session A: full GPU suite, accumulate variants, bench (both arms)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3
EXTRA_C=20 python tools/msm_variants.py 20 22 26 2>&1 | tee gpurun_out/msm_variants.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
