#!/bin/bash
mkdir -p gpurun_out
python scripts/variants.py cfg5 100000000 - FQD_TILE_FILL_PCT=50 FQD_TILE_FILL_PCT=55 FQD_TILE_FILL_PCT=65 FQD_TILE_FILL_PCT=70 FQD_TILE_FILL_PCT=65,FQD_TILE_SPLIT=0 > gpurun_out/i_var_cfg5.log 2>&1; cat gpurun_out/i_var_cfg5.log
