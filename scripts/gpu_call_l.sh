#!/bin/bash
mkdir -p gpurun_out
python scripts/variants.py cfg4 5000000 D=2 D=2,FQD_COMPARE_DENSE=0 D=1 D=1,FQD_COMPARE_DENSE=1 > gpurun_out/l_var_cfg4.log 2>&1; cat gpurun_out/l_var_cfg4.log | cut -c1-330
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_gputest.log
for d in 1 2; do python bench.py --steps 10 --warmup 3 --config cfg4 --distance $d > gpurun_out/r02_bench_cfg4_d$d.json 2> gpurun_out/r02_bench_cfg4_d$d.err; echo "bench cfg4 d$d rc=$?"; done
