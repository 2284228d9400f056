#!/bin/bash
# final evidence, part 2: the bench lines (config 5 = the driver's default invocation, configs 1-4 labelled)
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench cfg5 rc=$?"
b() { tag=$1; shift; python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err; echo "bench $tag rc=$?"; }
b cfg1 --config cfg1
b cfg2_adjacency --config cfg2 --method adjacency
b cfg2_highest_count --config cfg2 --method highest_count
b cfg3 --config cfg3
b cfg4_d1 --config cfg4 --distance 1
b cfg4_d2 --config cfg4 --distance 2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try:
        d=json.load(open(f)); print(f.split('bench_')[1], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), round(d['roofline']['whole_path']['frac'],3))
    except Exception as e: print(f, 'ERR', e)
PY
