#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "compare_kernels_agree or split_launches or packed_and_raw" > gpurun_out/k_gputest_new.log 2>&1; echo "new tests rc=$?"; tail -3 gpurun_out/k_gputest_new.log
python scripts/ref_parity_at_scale.py cfg4d1 cfg4d2 cfg5 cfg2 cfg3 > gpurun_out/r02_ref_parity_at_scale.log 2>&1; echo "parity rc=$?"; tail -8 gpurun_out/r02_ref_parity_at_scale.log | cut -c1-200
