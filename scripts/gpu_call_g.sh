#!/bin/bash
mkdir -p gpurun_out
python scripts/variants.py cfg5 100000000 - FQD_NO_TILE_FOREST=1 FQD_TILE_SPLIT=0 FQD_TILE_SPLIT=0,FQD_NO_TILE_FOREST=1 > gpurun_out/g_var_cfg5.log 2>&1; cat gpurun_out/g_var_cfg5.log
python -m pytest tests -m gpu -x -q > gpurun_out/g_gputest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/g_gputest.log
