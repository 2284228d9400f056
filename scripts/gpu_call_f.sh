#!/bin/bash
mkdir -p gpurun_out
python scripts/variants.py cfg5 100000000 - FQD_TILE_SPLIT=0 > gpurun_out/f_var_cfg5.log 2>&1; cat gpurun_out/f_var_cfg5.log
python -m pytest tests -m gpu -x -q tests/test_gpu_partitioned.py > gpurun_out/f_gputest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/f_gputest.log
FQD_TILE_SPLIT=1 python -m pytest tests -m gpu -x -q tests/test_gpu_partitioned.py > gpurun_out/f_gputest_split.log 2>&1; echo "pytest split rc=$?"; tail -2 gpurun_out/f_gputest_split.log
