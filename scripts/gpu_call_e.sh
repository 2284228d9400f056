#!/bin/bash
mkdir -p gpurun_out
python scripts/variants.py cfg4 5000000 D=2 D=1 > gpurun_out/e_var_cfg4.log 2>&1; cat gpurun_out/e_var_cfg4.log
python -m pytest tests -m gpu -x -q -k "edit or leven or cfg4 or reference_suite" > gpurun_out/e_gputest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/e_gputest.log
FQD_TRACE=1 python scripts/virtual_sharded.py cfg5 40000000 2 2 > gpurun_out/e_virtual.log 2>&1; tail -4 gpurun_out/e_virtual.log
ncu --set full --import-source on --clock-control none -k regex:"apply_edges|best_candidate|candidates_kernel|select_own" --launch-count 8 -f -o gpurun_out/r02_sharded_virtual2 python scripts/virtual_sharded.py cfg5 40000000 2 1 > gpurun_out/e_ncu.log 2>&1; echo "ncu rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"compare_dense" --launch-count 3 -f -o gpurun_out/r02_cfg4_d2_compare python scripts/variants.py cfg4 5000000 D=2 > gpurun_out/e_ncu2.log 2>&1; echo "ncu2 rc=$?"
