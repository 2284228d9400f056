#!/bin/bash
mkdir -p gpurun_out
python scripts/variants.py cfg4 5000000 D=2 D=2,FQD_COMPARE_DENSE=0 D=1 D=1,FQD_COMPARE_DENSE=1 > gpurun_out/d_var_cfg4.log 2>&1; cat gpurun_out/d_var_cfg4.log
python -m pytest tests -m gpu -x -q -k "edit or leven or cfg4 or reference_suite or golden" > gpurun_out/d_gputest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/d_gputest.log
FQD_CPU_SAMPLE=200000 python bench.py --steps 5 --warmup 3 > gpurun_out/d_bench_hybrid.json 2> gpurun_out/d_bench_hybrid.err; echo "rc=$?"
FQD_HOST_PACK_HYBRID=0 FQD_CPU_SAMPLE=200000 python bench.py --steps 5 --warmup 3 > gpurun_out/d_bench_packall.json 2> gpurun_out/d_bench_packall.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("hybrid","packall"):
    try:
        d=json.load(open(f"gpurun_out/d_bench_{f}.json")); print(f, d["ms_per_step"], d["e2e"])
    except Exception as e: print(f, "ERR", e)
PY
