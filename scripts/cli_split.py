"""Wall-time split of the drop-in ``deduplicate_cluster`` on a synthetic FASTQ (default 10 M reads x 150 bp with a
12-nt UMI in front, ``--check-lengths 36``): native pass 1 (read + parse + key slices), the GPU job, native pass 2
(emission), next to the Python per-record loops the reference runs for the same two passes (timed on a prefix and
extrapolated).  Also checks that the native passes and the Python loops produce the same keys / output bytes.

    python scripts/cli_split.py [reads] > gpurun_out/r02_cli_split.log
"""
import gzip
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from fastqdedup_b200 import _native, frontend, synth
from fastqdedup_b200.clustering import cluster_keys


def write_fastq(path, cfg, n, read_len=150):
    """Reads = synthetic key (UMI + prefix, duplicates and errors as in bench.py) + random tail; fixed-width names."""
    src = synth.SynthSource(cfg)
    L = cfg.key_length
    rng = np.random.default_rng(99)
    step = 1 << 20
    with open(path, "wb") as fh:
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            m = hi - lo
            keys, _, _ = src.reads(lo, hi)
            rec = np.empty((m, 1 + 12 + 1 + read_len + 3 + read_len + 1), dtype=np.uint8)
            rec[:, 0] = ord("@")
            ids = np.char.zfill(np.arange(lo, hi).astype("U"), 11)
            rec[:, 1:13] = np.frombuffer(("".join("r" + s for s in ids)).encode(), dtype=np.uint8).reshape(m, 12)
            rec[:, 13] = 10
            rec[:, 14:14 + L] = keys
            rec[:, 14 + L:14 + read_len] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(m, read_len - L))]
            p = 14 + read_len
            rec[:, p] = 10; rec[:, p + 1] = ord("+"); rec[:, p + 2] = 10
            rec[:, p + 3:p + 3 + read_len] = ord("I")
            rec[:, p + 3 + read_len] = 10
            fh.write(rec.tobytes())


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    cfg = synth.CONFIGS["cfg5"].scaled(n)
    tmp = tempfile.mkdtemp(prefix="fqd_cli_")
    fq = os.path.join(tmp, "in.fastq")
    t0 = time.perf_counter()
    write_fastq(fq, cfg, n)
    size = os.path.getsize(fq)
    print(f"input: {n} reads x 150 bp, {size / 1e9:.2f} GB FASTQ (generated in {time.perf_counter() - t0:.0f}s); "
          f"host has {os.cpu_count()} logical cores", flush=True)
    slices = frontend.length_string_to_slices("36")

    for out_name in ("out.fastq", "out.fastq.gz"):
        out = os.path.join(tmp, out_name)
        cluster_keys(np.frombuffer(b"ACGTACGTACGA" * 3, dtype=np.uint8).reshape(1, 36))     # context + first-call costs
        t0 = time.perf_counter()
        scan = _native.FastqScan([fq], slices)
        t1 = time.perf_counter()
        res = cluster_keys(scan.keys, None, 1, False, "directional", 1.0, want_uniques=False)
        t2 = time.perf_counter()
        n_rec = scan.n_records
        scan.close()
        written = _native.fastq_emit([fq], [out], res.keep_bitmap, n_rec)
        t3 = time.perf_counter()
        print(f"native, output {out_name}: pass 1 (read + parse + keys) {t1 - t0:.2f}s = {size / (t1 - t0) / 1e9:.2f} GB/s | "
              f"cluster (host buffers -> keep bitmap) {t2 - t1:.3f}s, device {res.stats['ms_total']:.1f} ms | "
              f"pass 2 (emit {written} records{', gzip -1 in the worker threads' if out_name.endswith('.gz') else ''}) {t3 - t2:.2f}s | "
              f"total {t3 - t0:.2f}s | U {res.number_of_uniques} selected {res.number_selected}", flush=True)

    # the Python per-record loops of the reference (its own readers are dnaio/xopen: not installable here, the stand-in
    # parser of fastq_io.py plays their part) on a prefix
    m = min(n, 500_000)
    small = os.path.join(tmp, "small.fastq")
    with open(fq, "rb") as src, open(small, "wb") as dst:
        dst.write(src.read(m * (size // n)))
    join = frontend.joinfunc_from_check_slices(slices)
    t0 = time.perf_counter()
    keys = [join(r.sequence for r in recs) for recs in frontend.fastq_files_to_records([small])]
    t1 = time.perf_counter()
    with _native.FastqScan([small], slices) as scan:
        same_keys = [bytes(r).decode() for r in scan.keys] == keys
        res = cluster_keys(scan.keys, None, 1, False, "directional", 1.0, want_uniques=False)
    keep = res.keep_mask()
    t2 = time.perf_counter()
    frontend.filter_fastq_files_on_bitmap([small], [os.path.join(tmp, "py.fastq")], keep)
    t3 = time.perf_counter()
    _native.fastq_emit([small], [os.path.join(tmp, "nat.fastq")], res.keep_bitmap, m)
    same_out = open(os.path.join(tmp, "py.fastq"), "rb").read() == open(os.path.join(tmp, "nat.fastq"), "rb").read()
    _native.fastq_emit([small], [os.path.join(tmp, "nat.fastq.gz")], res.keep_bitmap, m)
    same_gz = gzip.open(os.path.join(tmp, "nat.fastq.gz"), "rb").read() == open(os.path.join(tmp, "py.fastq"), "rb").read()
    print(f"python loops on {m} reads: pass 1 {t1 - t0:.2f}s, pass 2 {t3 - t2:.2f}s -> extrapolated to {n} reads: "
          f"pass 1 {(t1 - t0) * n / m:.0f}s, pass 2 {(t3 - t2) * n / m:.0f}s | keys identical: {same_keys}, "
          f"output identical: {same_out}, gzip output identical after decompression: {same_gz}", flush=True)
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)
    sys.exit(0 if same_keys and same_out and same_gz else 1)


if __name__ == "__main__":
    main()
