#!/bin/bash
# GPU call B: bench lines of configs 1-4 and ncu --set full captures (config 5 at full size, config 4 d=2)
mkdir -p gpurun_out
b() { tag=$1; shift; python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err; echo "bench $tag rc=$?"; }
b cfg1 --config cfg1
b cfg2_adjacency --config cfg2 --method adjacency
b cfg2_highest_count --config cfg2 --method highest_count
b cfg3 --config cfg3
b cfg4_d1 --config cfg4 --distance 1
b cfg4_d2 --config cfg4 --distance 2
ncu --set full --import-source on --clock-control none --launch-count 15 -f -o gpurun_out/r02_cfg5_full python scripts/run_once.py 100000000 1 cfg5 > gpurun_out/ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"
ncu --set full --import-source on --clock-control none --launch-count 40 -f -o gpurun_out/r02_cfg4_full python scripts/run_once.py 5000000 1 cfg4 > gpurun_out/ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
ls -la gpurun_out/
