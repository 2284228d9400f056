"""BASELINE.json configs at their full sizes: the streaming plan (shared-memory tiles) against the
single-table / counting-sort plan on the same device-resident input -- two independent GPU
implementations of the same job must agree on every counter and on the keep bitmap bit for bit.
(Both are checked against the CPU oracle at sizes the oracle can finish: tests/.)

    python scripts/full_size_check.py [cfg1 cfg2 ...] > gpurun_out/full_size.log
"""
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from bench import generate_into
from fastqdedup_b200 import _native, synth
from fastqdedup_b200.clustering import cluster_device

names = sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
ctx = _native.Context(0)
FIELDS = ("total_records", "discarded_records", "number_of_sequences", "number_of_uniques", "number_of_clusters",
          "number_selected")
for name in names:
    cfg = synth.CONFIGS[name]
    n, L = cfg.n_reads, cfg.key_length
    keys = np.empty((n, L), dtype=np.uint8)
    quals = np.empty((n, L), dtype=np.uint8) if cfg.quality_mix else None
    t0 = time.time()
    generate_into(cfg, 0, n, keys, quals)
    kp = ctx.upload(keys)
    qp = ctx.upload(quals) if quals is not None else None
    del keys, quals
    words = (n + 31) // 32
    bp = ctx.device_alloc(words * 4)
    methods = ["adjacency", "highest_count"] if name == "cfg2" else [cfg.method]
    dists = [1, 2] if name == "cfg4" else [cfg.max_distance]
    for method in methods:
        for d in dists:
            res = {}
            for plan, env in (("streaming", {}), ("single-table", {"FQD_NO_PARTITION": "1"})):
                for k in ("FQD_NO_PARTITION",):
                    os.environ.pop(k, None)
                os.environ.update(env)
                for _ in range(2):
                    st = cluster_device(ctx, n, kp, L, quals_ptr=qp, qual_length=L, max_distance=d,
                                        use_edit_distance=cfg.use_edit_distance, method=method,
                                        max_average_error_rate=cfg.max_average_error_rate, bitmap_ptr=bp)
                bm = ctx.download(bp, words * 4, np.uint32)
                res[plan] = (st.as_dict(), bm)
            a, b = res["streaming"], res["single-table"]
            same = all(a[0][f] == b[0][f] for f in FIELDS) and np.array_equal(a[1], b[1])
            print(f"{name} n={n} L={L} d={d} edit={cfg.use_edit_distance} {method}: "
                  f"{'IDENTICAL' if same else 'MISMATCH'} | U {a[0]['number_of_uniques']} clusters {a[0]['number_of_clusters']} "
                  f"selected {a[0]['number_selected']} discarded {a[0]['discarded_records']} | streaming {a[0]['ms_total']:.2f} ms "
                  f"(flags {a[0]['plan_flags']}) vs single-table {b[0]['ms_total']:.2f} ms | popcount {int(np.unpackbits(a[1].view(np.uint8)).sum())}",
                  flush=True)
            if not same:
                for f in FIELDS:
                    print("   ", f, a[0][f], b[0][f])
    ctx.device_free(kp)
    if qp:
        ctx.device_free(qp)
    ctx.device_free(bp)
