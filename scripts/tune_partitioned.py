"""Sweep of the partitioned plan's knobs (env vars read per job by csrc/pipeline.cu) on the
device-resident config-5 job.  Scratch tool, not the bench:

    python scripts/tune_partitioned.py [n_reads] > gpurun_out/tune.log
"""
import itertools
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np

from bench import generate_into
from fastqdedup_b200 import _native, synth
from fastqdedup_b200.clustering import cluster_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
cfg = synth.CONFIGS["cfg5"].scaled(n)
L = cfg.key_length
keys = np.empty((n, L), dtype=np.uint8)
t0 = time.time()
generate_into(cfg, 0, n, keys)
print(f"generated {n} in {time.time() - t0:.1f}s", flush=True)
ctx = _native.Context(0)
kp = ctx.upload(keys)
del keys

KNOBS = ("FQD_NO_NEXT_EMIT", "FQD_NO_LEAN", "FQD_TILE_FILL_PCT", "FQD_NO_FUSED_PASS0", "FQD_NO_PARTITION", "FQD_NO_PARTITION_PASSES", "FQD_NO_SWAR")


def run(tag, **env):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    best = None
    for it in range(3):
        st = cluster_device(ctx, n, kp, L, max_distance=1, use_edit_distance=False, method="directional",
                            max_average_error_rate=1.0)
        d = st.as_dict()
        if it and (best is None or d["ms_total"] < best["ms_total"]):
            best = d
    d = best
    print(f"{tag:42s} total {d['ms_total']:7.2f}  part {d['ms_partition_kernel']:6.2f} dedupe {d['ms_dedupe_kernel']:6.2f} "
          f"ingest {d['ms_ingest']:6.2f} | build {d['ms_bucket_build']:6.2f} compare {d['ms_compare']:6.2f} | select {d['ms_select']:5.2f} "
          f"| U {d['number_of_uniques']} C {d['number_of_clusters']} S {d['number_selected']} flags {d['plan_flags']}", flush=True)


run("legacy (no partition)", FQD_NO_PARTITION=1)
run("default")
run("pass 1 partitioned separately", FQD_NO_NEXT_EMIT=1)
run("no fused pass 0", FQD_NO_FUSED_PASS0=1)
for fill in (50, 70):
    run(f"fill={fill}", FQD_TILE_FILL_PCT=fill)
ctx.device_free(kp)
