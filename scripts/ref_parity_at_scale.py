"""Parity against the UNMODIFIED reference (oracle/_ref) at BASELINE scale.

For every case the compiled reference runs the loops of ``deduplicate_cluster``
(reference ``src/fastqdedup/__init__.py:240-276``: ``Trie.add_sequence`` per kept read,
``pop_cluster`` until empty, the dissection function per cluster) on the host, and the GPU
library runs the same reads through ``fqd_cluster`` with its DEFAULT thresholds (no
``FQD_PARTITION_MIN``: at these sizes the streaming plan, the spill path and the 32-bit
cursors see real load).  Compared bit for bit: the six counters of the log lines, the
per-unique ``first`` / ``count`` / ``label`` / ``selected`` arrays and the keep bitmap.

    python scripts/ref_parity_at_scale.py [case ...] > gpurun_out/ref_parity.log

Cases (default: all): cfg5 (10 M reads), cfg2 (10 M, adjacency + highest_count),
cfg3 (20 M, filter on, d=2), cfg4d1 / cfg4d2 (5 M, Levenshtein).  ``FQD_PARITY_SCALE=0.1``
shrinks every case (smoke run).  This is a checker: it may import oracle/.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np

from bench import generate_into
from fastqdedup_b200 import _native, synth
from fastqdedup_b200.clustering import cluster_keys

CASES = {
    # name: (config, reads, distances, methods)
    "cfg5": ("cfg5", 10_000_000, [1], ["directional"]),
    "cfg2": ("cfg2", 10_000_000, [1], ["adjacency", "highest_count"]),
    "cfg3": ("cfg3", 20_000_000, [2], ["directional"]),
    "cfg4d1": ("cfg4", 5_000_000, [1], ["directional"]),
    "cfg4d2": ("cfg4", 5_000_000, [2], ["directional"]),
}
FIELDS = ("total_records", "discarded_records", "number_of_sequences", "number_of_uniques",
          "number_of_clusters", "number_selected")


def reference_run(ref, keys, quals, d, edit, methods, max_err):
    """-> {method: result dict}, timings.  One trie build + one pop loop serve all methods."""
    n, L = keys.shape
    filter_on = max_err < 1.0 and quals is not None
    t0 = time.perf_counter()
    kbytes = keys.tobytes()
    qbytes = quals.tobytes() if filter_on else None
    trie = ref.Trie(alphabet="ACGTN")
    err = ref.fastq_average_error_rate
    first = {}
    discarded = 0
    for t in range(n):
        key = kbytes[t * L:(t + 1) * L].decode("latin-1")
        first.setdefault(key, t)           # pass 2 sees every record (__init__.py:201-206)
        if filter_on and err(qbytes[t * L:(t + 1) * L].decode("latin-1")) > max_err:
            discarded += 1
            continue
        trie.add_sequence(key)
    nseq = trie.number_of_sequences
    t1 = time.perf_counter()
    clusters = []
    while trie.number_of_sequences:
        clusters.append(trie.pop_cluster(d, edit))
    t2 = time.perf_counter()
    counts, label = {}, {}
    for cl in clusters:
        root = min(first[k] for _, k in cl)
        for c, k in cl:
            counts[k] = c
            label[k] = root
    ukeys = sorted(counts, key=first.__getitem__)
    out = {}
    for method in methods:
        func = ref.CLUSTER_DISSECTION_METHODS[method]
        selected = set()
        for cl in clusters:
            selected.update(func(cl, d, edit))
        res = dict(total_records=n, discarded_records=discarded, number_of_sequences=nseq,
                   number_of_uniques=len(ukeys), number_of_clusters=len(clusters),
                   number_selected=len(selected))
        res["first"] = np.fromiter((first[k] for k in ukeys), dtype=np.uint64, count=len(ukeys))
        res["count"] = np.fromiter((counts[k] for k in ukeys), dtype=np.uint32, count=len(ukeys))
        res["label"] = np.fromiter((label[k] for k in ukeys), dtype=np.uint64, count=len(ukeys))
        res["selected"] = np.fromiter((k in selected for k in ukeys), dtype=bool, count=len(ukeys))
        out[method] = res
    t3 = time.perf_counter()
    return out, {"add_s": t1 - t0, "pop_s": t2 - t1, "dissect_s": t3 - t2}


def main():
    import ref_loader
    ref = ref_loader.load_reference()
    if ref is None:
        raise SystemExit("oracle/_ref is not built")
    scale = float(os.environ.get("FQD_PARITY_SCALE", "1"))
    names = sys.argv[1:] or list(CASES)
    ctx = _native.Context(0)
    failures = 0
    for name in names:
        cfg_name, n, dists, methods = CASES[name]
        n = max(1000, int(n * scale))
        cfg = synth.CONFIGS[cfg_name].scaled(n)
        L = cfg.key_length
        keys = np.empty((n, L), dtype=np.uint8)
        quals = np.empty((n, L), dtype=np.uint8) if cfg.quality_mix else None
        generate_into(cfg, 0, n, keys, quals)
        for d in dists:
            want_all, tim = reference_run(ref, keys, quals, d, cfg.use_edit_distance, methods,
                                          cfg.max_average_error_rate)
            for method in methods:
                want = want_all[method]
                got = cluster_keys(keys, quals, d, cfg.use_edit_distance, method,
                                   cfg.max_average_error_rate, context=ctx)
                bad = [f for f in FIELDS if getattr(got, f) != want[f]]
                bad += [f for f in ("first", "count", "label", "selected")
                        if not np.array_equal(getattr(got, f), want[f])]
                keep = np.nonzero(got.keep_mask())[0].astype(np.uint64)
                if not np.array_equal(keep, want["first"][want["selected"]]):
                    bad.append("keep_bitmap")
                failures += bool(bad)
                print(f"{name} n={n} L={L} d={d} edit={cfg.use_edit_distance} {method}: "
                      f"{'IDENTICAL to the reference' if not bad else 'MISMATCH ' + ','.join(bad)} | "
                      f"U {want['number_of_uniques']} clusters {want['number_of_clusters']} "
                      f"selected {want['number_selected']} discarded {want['discarded_records']} | "
                      f"reference add {tim['add_s']:.1f}s pop {tim['pop_s']:.1f}s dissect(all methods) "
                      f"{tim['dissect_s']:.1f}s (1 core) | GPU {got.stats['ms_total']:.2f} ms "
                      f"plan_flags {got.stats['plan_flags']}", flush=True)
                for f in bad:
                    if f in FIELDS:
                        print("    ", f, getattr(got, f), want[f], flush=True)
            del want_all
        del keys, quals
    ctx.close()
    print("ALL IDENTICAL" if not failures else f"{failures} MISMATCHING CASES", flush=True)
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
