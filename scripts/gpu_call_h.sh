#!/bin/bash
mkdir -p gpurun_out
python scripts/variants.py cfg5 100000000 - > gpurun_out/h_var_cfg5.log 2>&1; cat gpurun_out/h_var_cfg5.log
python -m pytest tests -m gpu -x -q > gpurun_out/h_gputest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/h_gputest.log
