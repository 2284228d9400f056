#!/bin/bash
# GPU call A: parity suite, the config-5 bench line, launch list of the same command
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gputest.log
tail -3 gpurun_out/r02_gputest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
FQD_CPU_SAMPLE=100000 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_cfg5.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
head -c 1500 gpurun_out/r02_bench_n1.json
