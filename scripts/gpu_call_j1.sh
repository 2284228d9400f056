#!/bin/bash
# final evidence, part 1: ncu captures of the final code, CLI wall-time split, smoke, reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j_smoke.log 2>&1; tail -1 gpurun_out/j_smoke.log
ncu --set full --import-source on --clock-control none --launch-count 20 -f -o gpurun_out/r02_cfg5_full_final python scripts/run_once.py 100000000 1 cfg5 > gpurun_out/j_ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"
FQD_CPU_SAMPLE=100000 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_cfg5_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/j_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
python scripts/cli_split.py 10000000 > gpurun_out/r02_cli_split.log 2>&1; echo "cli_split rc=$?"; tail -15 gpurun_out/r02_cli_split.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/j_ref_arm.json 2> gpurun_out/j_ref_arm.err; echo "ref arm rc=$?"; head -c 600 gpurun_out/j_ref_arm.json
