"""Host-only throughput of the two native FASTQ passes (csrc/fastq_native.cpp), no GPU needed:
python scripts/fastq_passes_bench.py [reads] > profiles/rNN_fastq_passes_cpu.log
Pass 1 = fqd_fastq_scan_open (read + parse + key slices), pass 2 = fqd_fastq_emit with a random keep bitmap (19 % kept),
plain and gzip output; the synthetic FASTQ is the one scripts/cli_split.py writes (150-bp reads, 36-nt keys)."""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np

import cli_split
from fastqdedup_b200 import _native, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "in.fastq")
    cfg = synth.CONFIGS["cfg5"].scaled(n)
    t0 = time.time()
    cli_split.write_fastq(path, cfg, n)
    size = os.path.getsize(path)
    print(f"input: {n} reads x 150 bp, {size / 1e9:.2f} GB FASTQ (generated in {time.time() - t0:.0f}s); host has {os.cpu_count()} logical cores")
    for it in range(3):
        t0 = time.perf_counter()
        with _native.FastqScan([path], [slice(0, 36)], want_quals=False) as scan:
            nrec = scan.n_records
        dt = time.perf_counter() - t0
        print(f"pass 1 (read + parse + 36-nt keys): {dt:.3f}s = {size / dt / 1e9:.2f} GB/s, {nrec} records")
    rng = np.random.default_rng(0)
    keep = rng.random(n) < 0.19
    words = np.packbits(keep, bitorder="little")
    words = np.concatenate([words, np.zeros((-len(words)) % 4, dtype=np.uint8)]).view(np.uint32)
    for name in ("out.fastq", "out.fastq.gz"):
        out = os.path.join(tmp, name)
        for it in range(2):
            t0 = time.perf_counter()
            _native.fastq_emit([path], [out], words, n)
            dt = time.perf_counter() - t0
            print(f"pass 2 -> {name}: {dt:.3f}s = {size / dt / 1e9:.2f} GB/s of input, {os.path.getsize(out) / 1e6:.0f} MB written")
