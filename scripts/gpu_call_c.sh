#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/c_gputest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c_gputest.log
python scripts/variants.py cfg4 5000000 D=2 D=2,FQD_COMPARE_V1=1 D=1 D=1,FQD_COMPARE_V1=1 > gpurun_out/c_var_cfg4.log 2>&1; cat gpurun_out/c_var_cfg4.log
python scripts/variants.py cfg5 100000000 - FQD_PARTITION_V1=1 FQD_PARTITION_BPS=3 FQD_PARTITION_BPS=5 FQD_PARTITION_BPS=6 FQD_PARTITION_BPS=8 > gpurun_out/c_var_cfg5.log 2>&1; cat gpurun_out/c_var_cfg5.log
