#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --config cfg4 --distance 2 > gpurun_out/r02_bench_cfg4_d2.json 2> gpurun_out/r02_bench_cfg4_d2.err; echo "bench cfg4 d2 rc=$?"
python -m pytest tests -m gpu -x -q -k "compare_kernels_agree or cfg4" > gpurun_out/m_gputest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/m_gputest.log
python -c "import json; d=json.load(open('gpurun_out/r02_bench_cfg4_d2.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
