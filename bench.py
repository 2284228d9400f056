#!/usr/bin/env python3
"""bench.py -- unique keys clustered per second on BASELINE.json's headline workload
(config 5: 100 M reads, 12-nt UMI + 24-nt prefix = 36-nt key, Hamming d=1, directional).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path
    python bench.py --config cfg3 [--distance D] [--method M]  # the other BASELINE configs (1 GPU)

One "step" is one complete pass of the hot path (quality filter off for this config, pack,
exact dedupe, neighbour search, components, dissection) over the whole synthetic batch.
``value`` is measured with the batch resident in HBM (CUDA events inside the library, on its
stream); ``e2e`` repeats it through the same C-ABI call with pinned HOST buffers, H2D and
D2H inside the timed region.  Inputs (3.6 GB) are far larger than the 126 MB L2, so no
explicit flush is needed between iterations.  The JSON line is printed by rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "unique keys/s"


def metric_name(cfg):
    """BASELINE.json's metric, qualified by what the configuration actually runs."""
    dist = f"Levenshtein d={cfg.max_distance}" if cfg.use_edit_distance else f"d={cfg.max_distance}"
    filt = ", quality filter on" if cfg.max_average_error_rate < 1.0 else ""
    return f"unique UMIs clustered/sec ({dist}, {cfg.method}{filt})"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------

def make_config(args):
    from fastqdedup_b200 import synth
    from dataclasses import replace
    cfg = synth.CONFIGS[args.config]
    n = env_int("FQD_BENCH_READS", 0) or args.reads or cfg.n_reads
    if n != cfg.n_reads:
        cfg = cfg.scaled(n)
    if args.distance is not None:
        cfg = replace(cfg, max_distance=args.distance)
    if args.method:
        cfg = replace(cfg, method=args.method)
    return cfg


def generate_into(cfg, start, stop, out_keys, out_quals=None, threads=8):
    """Fill out_keys[(stop-start), L] with reads [start, stop) of the synthetic data set."""
    from concurrent.futures import ThreadPoolExecutor

    from fastqdedup_b200 import synth
    src = synth.SynthSource(cfg)
    step = synth.CHUNK
    bounds = []
    t = start
    while t < stop:
        hi = min(stop, (t // step + 1) * step)
        bounds.append((t, hi))
        t = hi

    def work(b):
        lo, hi = b
        k, _, q = src.reads(lo, hi)
        out_keys[lo - start:hi - start] = k
        if out_quals is not None and q is not None:
            out_quals[lo - start:hi - start] = q

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, bounds))


class PinnedArray:
    """uint8 [n, L] array in cudaHostAlloc'ed memory (so H2D copies can stream)."""

    def __init__(self, lib, shape):
        self.lib = lib
        self.nbytes = int(np.prod(shape))
        p = ctypes.c_void_p()
        from fastqdedup_b200 import _native
        _native.check(lib.fqd_host_alloc(max(self.nbytes, 16), ctypes.byref(p)))
        self.ptr = p.value
        buf = (ctypes.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.fqd_host_free(ctypes.c_void_p(self.ptr))
            self.ptr = None


# ---------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
# roofline bookkeeping (DESIGN.md "Algorithmic bytes")
# ---------------------------------------------------------------------------------------

def survey_key_bytes(L):
    """W of SURVEY.md section 8: 2-bit plane + N-mask plane, in bytes."""
    return 4 * ((L + 15) // 16) + 4 * ((L + 31) // 32)


def survey_passes(cfg, hamming_passes):
    """P of SURVEY.md section 8(d): Hamming d+1; Levenshtein (d+1)(2s+1) shifted-block passes with
    s = floor(d/2) for keys of one length (the synthetic configs) -- the library runs the shifts of one
    block as variants inside one pass, so its pass count is d+1 either way."""
    if not cfg.use_edit_distance:
        return hamming_passes
    return hamming_passes * (2 * (cfg.max_distance // 2) + 1)


def algorithmic_bytes(cfg, n, n_ok, u, passes, filter_on):
    W = survey_key_bytes(cfg.key_length)
    R = W + 4
    total = (n * cfg.key_length if filter_on else 0) + n_ok * W + u * (R + 4) + 3 * passes * u * R + 9 * u
    ingest = (n * cfg.key_length if filter_on else 0) + n_ok * W + u * (R + 4)
    return total, ingest


def bitmap_digest(parts):
    """sha256 of the keep bitmap of the whole job in global record order; `parts` = [(uint32 words, records)]
    per rank (bit t%32 of word t/32 over the rank's LOCAL record index).  The same reads give the same digest
    whatever the GPU count: compare the lines of a scaling run."""
    import hashlib
    bits = [np.unpackbits(np.ascontiguousarray(w).view(np.uint8), bitorder="little")[:n] for w, n in parts]
    packed = np.packbits(np.concatenate(bits) if len(bits) > 1 else bits[0], bitorder="little")
    return hashlib.sha256(packed.tobytes()).hexdigest()[:32]


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------
# CPU reference timing (oracle/_ref: the unmodified reference, single thread by construction)
# ---------------------------------------------------------------------------------------

def time_reference(cfg, keys, quals):
    """Exactly the loops of deduplicate_cluster (reference __init__.py:240-276) without file
    I/O.  Returns (uniques, seconds_total, split dict)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    ref = ref_loader.load_reference()
    kind = "reference"
    if ref is None:
        # no compiled reference on this box: the C port of the same algorithm (brute-force
        # neighbour search, so only a small sample finishes in seconds)
        import oracle
        cap = 40_000
        keys = keys[:cap]
        quals = None if quals is None else quals[:cap]
        t0 = time.perf_counter()
        r = oracle.cluster(keys, quals, cfg.max_distance, cfg.use_edit_distance, cfg.method,
                           cfg.max_average_error_rate)
        dt = time.perf_counter() - t0
        return r["number_of_uniques"], dt, {"total_s": dt}, "port"
    key_strs = [bytes(row).decode("latin-1") for row in keys]
    filter_on = cfg.max_average_error_rate < 1.0 and quals is not None
    qual_strs = [bytes(row).decode("latin-1") for row in quals] if filter_on else None
    func = ref.CLUSTER_DISSECTION_METHODS[cfg.method]
    t0 = time.perf_counter()
    trie = ref.Trie(alphabet="ACGTN")
    for t, key in enumerate(key_strs):
        if filter_on and ref.fastq_average_error_rate(qual_strs[t]) > cfg.max_average_error_rate:
            continue
        trie.add_sequence(key)
    t1 = time.perf_counter()
    clusters = []
    while trie.number_of_sequences:
        clusters.append(trie.pop_cluster(cfg.max_distance, cfg.use_edit_distance))
    t2 = time.perf_counter()
    selected = set()
    uniques = 0
    for cl in clusters:
        uniques += len(cl)
        for k in func(cl, cfg.max_distance, cfg.use_edit_distance):
            selected.add(hash(k))
    t3 = time.perf_counter()
    split = {"add_s": round(t1 - t0, 3), "pop_s": round(t2 - t1, 3), "dissect_s": round(t3 - t2, 3)}
    return uniques, t3 - t0, split, kind


# reference speed per read on one host core (add + pop_cluster + dissection),
# used only to size the bounded sample of the reference arm
REF_US_PER_READ = {"cfg1": 1.0, "cfg2": 3.5, "cfg3": 11.0, "cfg4": 3.0, "cfg5": 3.5}


def reference_sample_reads(cfg, args):
    """Reads per step of the reference arm: a prefix of the same synthetic workload sized so that the
    whole --steps/--warmup run ends within a few minutes (FQD_REF_SAMPLE overrides)."""
    forced = env_int("FQD_REF_SAMPLE", 0)
    if forced:
        return min(cfg.n_reads, forced)
    us = REF_US_PER_READ.get(cfg.name, 3.5) * (6.0 if cfg.use_edit_distance and cfg.max_distance >= 2 else 1.0)
    budget_s = 240.0 / max(1, args.steps + args.warmup)
    return int(min(cfg.n_reads, 4_000_000, max(250_000, budget_s / (us * 1e-6))))


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cfg = make_config(args)
    sample = reference_sample_reads(cfg, args)
    from fastqdedup_b200 import synth      # pure Python: the package resolves its native names lazily
    src = synth.SynthSource(cfg)
    keys, _, quals = src.reads(0, sample)
    times, uniques = [], 0
    for it in range(args.warmup + args.steps):
        uniques, dt, split, kind = time_reference(cfg, keys, quals)
        if it >= args.warmup:
            times.append(dt)
        log(f"[reference] step {it}: {uniques} uniques in {dt:.2f}s {split}")
    ms = 1e3 * sum(times) / len(times)
    value = uniques / (ms / 1e3)
    config = workload_config(cfg, args.gpus)
    config["reference_sample"] = {
        "reads_per_step": sample, "unique_keys_per_step": uniques,
        "note": "the reference's per-unique cost grows with the trie, so a prefix sample flatters it; "
                "the value is an extrapolation to the full workload, not a full-size run"}
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"first {sample} of {cfg.n_reads} reads per step "
                                   f"({uniques} unique keys); add+pop_cluster+dissection, "
                                   f"single thread (the reference has no threading)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


WORKLOADS = {
    "cfg1": "BASELINE config 1: single-end, 12-nt prepended UMI, --check-lengths 12 -E",
    "cfg2": "BASELINE config 2: paired-end + UMI file, --check-lengths 16,8,12 -E (36-nt key)",
    "cfg3": "BASELINE config 3: single-end, 16-nt UMI + 32-nt prefix (48-nt key), quality filter on (-e 0.001)",
    "cfg4": "BASELINE config 4: single-end, 24-nt key, --edit -E",
    "cfg5": "BASELINE config 5: 12-nt UMI + 24-nt prefix (36-nt key), -E",
}


def workload_config(cfg, gpus):
    dist = ("Levenshtein" if cfg.use_edit_distance else "Hamming") + f" d={cfg.max_distance}"
    filt = "off (-E)" if cfg.max_average_error_rate >= 1.0 else f"on (-e {cfg.max_average_error_rate})"
    return {"workload": f"{WORKLOADS.get(cfg.name, cfg.name)}: {cfg.n_reads} reads, {cfg.key_length}-nt key, "
                        f"{dist}, {cfg.method}, quality filter {filt}",
            "name": cfg.name,
            "reads": cfg.n_reads, "key_length": cfg.key_length, "molecules": cfg.n_molecules,
            "max_distance": cfg.max_distance, "edit_distance": bool(cfg.use_edit_distance),
            "method": cfg.method, "quality_filter": cfg.max_average_error_rate < 1.0,
            "sharding": "1 GPU" if gpus == 1 else f"{gpus} GPUs, reads split contiguously",
            "l2": "inputs (reads x key bytes) exceed the 126 MB L2; no flush needed"
                  if cfg.n_reads * cfg.key_length > (126 << 20) else
                  "inputs fit the 126 MB L2: a 256 MB buffer is overwritten between timed steps"}


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------

def run_ours(args, rank, world, local_rank):
    from fastqdedup_b200 import _native
    from fastqdedup_b200.clustering import cluster_device
    from fastqdedup_b200._native import ClusterJob, METHODS, MEM_HOST

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    cfg = make_config(args)
    if world > 1:
        return run_ours_sharded(args, cfg, rank, world, local_rank, dist)

    lib = _native.load()
    if lib.fqd_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    ctx = _native.Context(local_rank)
    n, L = cfg.n_reads, cfg.key_length
    filter_on = cfg.max_average_error_rate < 1.0

    t0 = time.time()
    host_keys = PinnedArray(lib, (n, L))
    host_quals = PinnedArray(lib, (n, L)) if cfg.quality_mix else None
    generate_into(cfg, 0, n, host_keys.array, None if host_quals is None else host_quals.array)
    log(f"[bench] generated {n} reads x {L} nt in {time.time() - t0:.1f}s")

    d_keys = ctx.upload(host_keys.array)
    d_quals = ctx.upload(host_quals.array) if host_quals is not None else None
    bitmap_words = (n + 31) // 32
    d_bitmap = ctx.device_alloc(bitmap_words * 4)

    # inputs smaller than the L2 (config 1): overwrite a 256 MB buffer between timed steps
    flush_bytes = (256 << 20) if n * L * (2 if host_quals is not None else 1) <= (126 << 20) else 0
    d_flush = ctx.device_alloc(flush_bytes) if flush_bytes else None

    def device_step():
        if d_flush:
            ctx.memset(d_flush, 0x5A, flush_bytes)
        return cluster_device(ctx, n, d_keys, L, quals_ptr=d_quals, qual_length=L,
                              max_distance=cfg.max_distance, use_edit_distance=cfg.use_edit_distance,
                              method=cfg.method, max_average_error_rate=cfg.max_average_error_rate,
                              bitmap_ptr=d_bitmap)

    for _ in range(args.warmup):
        st = device_step()
    ctx.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    stats = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        stats.append(device_step())
    ctx.synchronize()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()

    dev_ms = [s.ms_total for s in stats]
    ms_per_step = float(np.mean(dev_ms))
    st = stats[-1]
    U = st.number_of_uniques
    value = U / (ms_per_step / 1e3)

    # ---- e2e: same call with pinned host buffers, copies inside the timed region ----
    host_bitmap = PinnedArray(lib, (bitmap_words * 4,))
    job = ClusterJob()
    job.n_records = n
    job.keys = host_keys.ptr
    job.key_stride = job.key_length = L
    if host_quals is not None:
        job.quals = host_quals.ptr
        job.qual_stride = job.qual_length = L
    job.max_distance = cfg.max_distance
    job.use_edit_distance = int(cfg.use_edit_distance)
    job.method = METHODS[cfg.method]
    job.memory_space = MEM_HOST
    job.max_average_error_rate = cfg.max_average_error_rate
    job.phred_offset = 33
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):                         # warm-up (the first call grows the arena)
        ctx.cluster(job, host_bitmap.ptr)
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        est = ctx.cluster(job, host_bitmap.ptr)
    e2e_s = (time.perf_counter() - e0) / e2e_steps
    host_in = n * L * (2 if host_quals is not None else 1)    # what the caller hands over
    h2d = int(est.h2d_bytes)                                   # what the library copied over PCIe (keys packed on the host)
    d2h = bitmap_words * 4
    # the e2e result is the same set as the device-resident one
    dev_bitmap = ctx.download(d_bitmap, bitmap_words * 4, np.uint32)
    assert np.array_equal(dev_bitmap, host_bitmap.array.view(np.uint32)), "e2e and device-resident results differ"
    assert est.number_selected == st.number_selected
    digest = bitmap_digest([(dev_bitmap, n)])

    # ---- roofline of the dominant kernel ----
    # Every term of SURVEY.md section 8(d)'s B_total is charged to the kernel (group) that does that
    # work; times are CUDA-event averages over the timed steps, taken inside the library on its stream.
    P = survey_passes(cfg, st.n_passes)
    total_b, ingest_b = algorithmic_bytes(cfg, n, st.number_of_sequences, U, P, filter_on)
    peak, peak_src = measured_peak()
    W = survey_key_bytes(L)
    R = W + 4
    mean = lambda f: float(np.mean([getattr(s_, f) for s_ in stats]))
    streaming = bool(st.plan_flags & 1)
    tiled_passes = bool(st.plan_flags & 2)
    fused = bool(st.plan_flags & 4)
    emitted = bool(st.plan_flags & 8)      # the dedupe tiles also wrote the bucket-ordered entries of pass 1
    K_, PW_ = st.key_bits, st.key_words // max(st.key_bits, 1)
    inst = f"<{K_},{PW_}>"
    lean = streaming and not filter_on and K_ == 3 and L % 4 == 0
    passes_left = st.n_passes - (1 if fused else 0)
    per_pass = P // max(st.n_passes, 1)    # SURVEY passes per library pass (Levenshtein shifts)
    read_b = (n * L if filter_on else 0) + st.number_of_sequences * W
    if cfg.use_edit_distance:
        pass_kernels = "sig_count + scan + scatter + compare_kernel (Myers bit-vector verify)"
    elif tiled_passes:
        pass_kernels = ("" if emitted else "bucket_partition_kernel + ") + f"bucket_tile_kernel{inst} + apply_edges_kernel"
    else:
        pass_kernels = f"sig_count + scan + scatter_fat + compare_fat_kernel{inst}"
    select_kernels = {"directional": "root_best_kernel + select_kernel", "highest_count": "root_best_kernel + select_kernel",
                      "adjacency": "adj_edge/adj_node rounds + select_kernel"}[cfg.method] + " (dissection + keep bitmap)"
    if streaming:
        groups = [
            ((f"partition_dna_kernel<{PW_},{L // 4}>" if lean else f"ingest_kernel{inst} partition mode") +
             " (" + ("quality filter + " if filter_on else "") + "pack + hash + partition into tiles)",
             mean("ms_partition_kernel"), read_b),
            (f"dedupe_tile_kernel{inst}" + (" fused with pass 0" if fused else "") + (" + pass-1 tiles" if emitted else "") +
             " (exact dedupe in shared-memory tiles)",
             mean("ms_dedupe_kernel"), U * (R + 4) + (3 * U * R * per_pass if fused else 0) + (U * R if emitted else 0)),
            (f"{passes_left} pass(es): " + pass_kernels,
             mean("ms_neighbour"), 3 * passes_left * per_pass * U * R - (U * R if emitted else 0)),
            (select_kernels, mean("ms_select"), 9 * U),
        ]
    else:
        groups = [
            (f"ingest_kernel{inst} (filter + pack + exact dedupe, HBM table)", mean("ms_ingest"), ingest_b),
            (f"{st.n_passes} pass(es): " + pass_kernels, mean("ms_neighbour"), 3 * P * U * R),
            (select_kernels, mean("ms_select"), 9 * U),
        ]
    kernels = [{"kernel": k, "ms": ms, "algorithmic_bytes": int(b), "achieved_gbs": b / (ms / 1e3) / 1e9 if ms > 0 else None,
                "frac": b / (ms / 1e3) / 1e9 / peak if ms > 0 else None} for k, ms, b in groups]
    dom = max(kernels, key=lambda k: k["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    from fastqdedup_b200 import synth as _synth
    if os.path.exists(tpath) and cfg.n_reads == _synth.CONFIGS[cfg.name].n_reads:   # (captured at the config's full size)
        try:
            for name, val in json.load(open(tpath)).get(cfg.name, {}).items():
                if dom["kernel"].startswith(name):
                    traffic = val
        except Exception:
            traffic = None
    # compare phase against the integer-issue roofline: per candidate pair
    # 4*ceil(L/16) + 3*ceil(L/32) + 2 integer operations (Hamming), ~17 L (Myers); INT peak measured by
    # fqd_int_peak (dependent-free LOP3 + POPC streams) on this GPU
    ops_pair = 17 * L if cfg.use_edit_distance else 4 * ((L + 15) // 16) + 3 * ((L + 31) // 32) + 2
    int_peak = ctx.int_peak()
    t_int_ms = 1e3 * st.candidate_pairs * ops_pair / (int_peak["mixed_ops_per_s"] or 1.0)
    roofline = {"bound": "hbm", "kernel": dom["kernel"],
                "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes"],
                "kernel_ms": dom["ms"],
                "whole_path": {"algorithmic_bytes": total_b,
                               "bytes_per_unique": total_b / max(U, 1),
                               "survey_passes": P,
                               "achieved_gbs": total_b / (ms_per_step / 1e3) / 1e9,
                               "frac": total_b / (ms_per_step / 1e3) / 1e9 / peak},
                "kernels": kernels,
                "integer": {"candidate_pairs": int(st.candidate_pairs), "int_ops_per_pair": ops_pair,
                            "int_peak_ops_per_s": int_peak["mixed_ops_per_s"],
                            "lop3_peak_ops_per_s": int_peak["lop3_ops_per_s"],
                            "popc_peak_ops_per_s": int_peak["popc_ops_per_s"],
                            "t_int_ms": t_int_ms,
                            "compare_phase_ms": mean("ms_compare") if not fused else None,
                            "note": "t_int = candidate pairs x ops / measured INT peak: the compare arithmetic itself "
                                    "is far from the integer-issue limit; the passes are bound by tile staging and "
                                    "shared-memory probe latency (DESIGN.md section 4)"},
                "plan": {"streaming_dedupe": streaming, "pass0_fused": fused, "pass1_tiles_from_dedupe": emitted,
                         "streaming_passes": tiled_passes}}

    # ---- CPU baseline on a bounded sample of the same workload ----
    # about 15-25 s of single-core reference work, whatever the configuration
    us_read = REF_US_PER_READ.get(cfg.name, 3.5) * (6.0 if cfg.use_edit_distance and cfg.max_distance >= 2 else 1.0)
    sample = min(n, env_int("FQD_CPU_SAMPLE", 0) or int(min(4_000_000, max(250_000, 14.0 / (us_read * 1e-6)))))
    uniq_s, secs, split, kind = time_reference(
        cfg, host_keys.array[:sample], None if host_quals is None else host_quals.array[:sample])
    cpu = {"value": uniq_s / secs, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"first {sample} of {n} reads ({uniq_s} unique keys) in {secs:.1f}s {split}; "
                     f"single thread (the reference has no threading), "
                     f"host has {os.cpu_count()} logical cores"}

    line = {
        "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(cfg, 1),
        "unique_keys": int(U), "clusters": int(st.number_of_clusters),
        "selected": int(st.number_selected), "candidate_pairs": int(st.candidate_pairs),
        "keep_bitmap_sha256_32": digest,
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "e2e": {"value": U / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
                "steps": e2e_steps, "h2d_ms": est.ms_h2d, "host_input_bytes_per_step": int(host_in),
                "host_packing": "keys packed to 3 bits/symbol (plane streams per chunk) by the library's host threads inside the timed region, some chunks sent raw when the link would idle"
                                if h2d < host_in else "none"},
        "gpu_launches": int(sum(s.launches for s in stats)),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    ctx.device_free(d_keys)
    if d_quals:
        ctx.device_free(d_quals)
    ctx.device_free(d_bitmap)
    host_bitmap.free()
    host_keys.free()
    if host_quals is not None:
        host_quals.free()
    ctx.close()


def run_ours_sharded(args, cfg, rank, world, local_rank, dist):
    """One process per GPU (torchrun): strong scaling of the same 100 M-read job.  Each rank
    holds a contiguous slice of the reads; the library exchanges keys / forests over NCCL.
    Device time is the max over ranks of the CUDA-event time of the whole job."""
    import torch
    from fastqdedup_b200 import _native
    from fastqdedup_b200.multigpu import ShardComm, shard_bounds

    lib = _native.load()
    ctx = _native.Context(local_rank)
    comm = ShardComm.from_torch_distributed(ctx, dist)
    n, L = cfg.n_reads, cfg.key_length
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    nloc = hi - lo
    t0 = time.time()
    host_keys = PinnedArray(lib, (nloc, L))
    host_quals = PinnedArray(lib, (nloc, L)) if cfg.quality_mix else None
    generate_into(cfg, lo, hi, host_keys.array, None if host_quals is None else host_quals.array,
                  threads=max(2, 16 // world))
    log(f"[bench r{rank}] generated reads [{lo}, {hi}) in {time.time() - t0:.1f}s")
    d_keys = ctx.upload(host_keys.array)
    d_quals = ctx.upload(host_quals.array) if host_quals is not None else None
    words = (nloc + 31) // 32
    d_bitmap = ctx.device_alloc(max(words, 1) * 4)

    def step():
        return comm.cluster_device(nloc, lo, d_keys, L, quals_ptr=d_quals, max_distance=cfg.max_distance,
                                   use_edit_distance=cfg.use_edit_distance, method=cfg.method,
                                   max_average_error_rate=cfg.max_average_error_rate, bitmap_ptr=d_bitmap)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    per_step, stats = [], []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        st = step()
        stats.append(st)
        per_step.append(max_over_ranks(st.ms_total))
    dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    ms_per_step = float(np.mean(per_step))
    st = stats[-1]
    U = st.number_of_uniques

    # e2e: same call with pinned host buffers (H2D of the shard + D2H of its bitmap inside)
    host_bitmap = PinnedArray(lib, (max(words, 1) * 4,))
    def e2e_step():
        return comm.cluster_host(host_keys.array, lo, None if host_quals is None else host_quals.array,
                                 cfg.max_distance, cfg.use_edit_distance, cfg.method,
                                 cfg.max_average_error_rate, bitmap=host_bitmap.array.view(np.uint32))
    for _ in range(2):
        e2e_step()
    dist.barrier()
    e2e_steps = max(1, min(args.steps, 5))
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        est = e2e_step()
    dist.barrier()
    e2e_s = max_over_ranks((time.perf_counter() - e0) / e2e_steps)
    dev_bitmap = ctx.download(d_bitmap, max(words, 1) * 4, np.uint32)
    assert np.array_equal(dev_bitmap[:words], host_bitmap.array.view(np.uint32)[:words])
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((dev_bitmap[:words].copy(), nloc), gathered, dst=0)
    launches = int(sum(s.launches for s in stats))
    lt = torch.tensor([launches, int(est.h2d_bytes)], dtype=torch.int64, device=f"cuda:{local_rank}")
    dist.all_reduce(lt)
    if rank == 0:
        total_b, _ = algorithmic_bytes(cfg, n, st.number_of_sequences, U, st.n_passes,
                                       cfg.max_average_error_rate < 1.0)
        peak, peak_src = measured_peak()
        line = {
            "metric": metric_name(cfg), "value": U / (ms_per_step / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": workload_config(cfg, world),
            "unique_keys": int(U), "clusters": int(st.number_of_clusters),
            "selected": int(st.number_selected), "candidate_pairs": int(st.candidate_pairs),
            "keep_bitmap_sha256_32": bitmap_digest(gathered),
            "plan": "tile-sharded (peer-memory tile fetch)" if st.plan_flags & 16 else "replicated unique set",
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "e2e": {"value": U / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(lt[1].item()),
                    "d2h_bytes_per_step": int(((n + 31) // 32) * 4), "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps,
                    "host_input_bytes_per_step": int(n * L * (2 if host_quals is not None else 1))},
            "gpu_launches": int(lt[0].item()),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole sharded job (per-kernel split: see the 1-GPU line)",
                         "achieved": total_b / (ms_per_step / 1e3) / 1e9, "peak": peak * world, "unit": "GB/s",
                         "frac": total_b / (ms_per_step / 1e3) / 1e9 / (peak * world), "traffic": None,
                         "peak_source": peak_src + f" x {world} GPUs"},
            "rank0_stage_ms": {"ingest_kernel": st.ms_ingest_kernel, "compare": st.ms_compare},
        }
        print(json.dumps(line), flush=True)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--config", default="cfg5")
    ap.add_argument("--reads", type=int, default=0, help="override the read count (debugging)")
    ap.add_argument("--distance", type=int, default=None, help="override the config's max distance (cfg4: 1 or 2)")
    ap.add_argument("--method", default="", choices=("", "directional", "adjacency", "highest_count"),
                    help="override the config's dissection method (cfg2: adjacency or highest_count)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if env_int("FQD_WATCHDOG", 0):
        import faulthandler
        faulthandler.dump_traceback_later(env_int("FQD_WATCHDOG", 0), exit=True)
    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        log("[bench] note: fewer than 3 warm-up steps requested; using 3")
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
