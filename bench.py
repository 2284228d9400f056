#!/usr/bin/env python3
"""bench.py -- unique keys clustered per second on BASELINE.json's headline workload
(config 5: 100 M reads, 12-nt UMI + 24-nt prefix = 36-nt key, Hamming d=1, directional).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

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

METRIC = "unique UMIs clustered/sec (d=1, directional)"
UNIT = "unique keys/s"


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
    cfg = synth.CONFIGS[args.config]
    n = env_int("FQD_BENCH_READS", 0) or args.reads or cfg.n_reads
    if n != cfg.n_reads:
        cfg = cfg.scaled(n)
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


def algorithmic_bytes(cfg, n, n_ok, u, passes, filter_on):
    W = survey_key_bytes(cfg.key_length)
    R = W + 4
    total = (n * cfg.key_length if filter_on else 0) + n_ok * W + u * (R + 4) + 3 * passes * u * R + 9 * u
    ingest = (n * cfg.key_length if filter_on else 0) + n_ok * W + u * (R + 4)
    return total, ingest


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


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cfg = make_config(args)
    sample = min(cfg.n_reads, env_int("FQD_REF_SAMPLE", 1_000_000))
    from fastqdedup_b200 import synth
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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": workload_config(cfg, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                         "sample": f"first {sample} of {cfg.n_reads} reads per step "
                                   f"({uniques} unique keys); add+pop_cluster+dissection, "
                                   f"single thread (the reference has no threading)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(cfg, gpus):
    return {"workload": f"BASELINE config 5: {cfg.n_reads} reads, {cfg.key_length}-nt key "
                        f"(12-nt UMI + 24-nt prefix), Hamming d={cfg.max_distance}, {cfg.method}, "
                        f"quality filter off (-E)",
            "reads": cfg.n_reads, "key_length": cfg.key_length, "molecules": cfg.n_molecules,
            "max_distance": cfg.max_distance, "method": cfg.method,
            "sharding": "1 GPU" if gpus == 1 else f"{gpus} GPUs, reads split contiguously",
            "l2": "inputs (reads x key bytes) exceed the 126 MB L2; no flush needed"}


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

    def device_step():
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
    h2d = n * L * (2 if host_quals is not None else 1)
    d2h = bitmap_words * 4
    # the e2e result is the same set as the device-resident one
    dev_bitmap = ctx.download(d_bitmap, bitmap_words * 4, np.uint32)
    assert np.array_equal(dev_bitmap, host_bitmap.array.view(np.uint32)), "e2e and device-resident results differ"
    assert est.number_selected == st.number_selected

    # ---- roofline of the dominant kernel ----
    # Every term of SURVEY.md section 8(d)'s B_total is charged to the kernel (group) that does that
    # work; times are CUDA-event averages over the timed steps, taken inside the library on its stream.
    total_b, ingest_b = algorithmic_bytes(cfg, n, st.number_of_sequences, U, st.n_passes, filter_on)
    peak, peak_src = measured_peak()
    W = survey_key_bytes(L)
    R = W + 4
    mean = lambda f: float(np.mean([getattr(s_, f) for s_ in stats]))
    streaming = bool(st.plan_flags & 1)
    fused = bool(st.plan_flags & 4)
    emitted = bool(st.plan_flags & 8)      # the dedupe tiles also wrote the bucket-ordered entries of pass 1
    passes_left = st.n_passes - (1 if fused else 0)
    read_b = (n * L if filter_on else 0) + st.number_of_sequences * W
    if streaming:
        groups = [
            (("partition_dna_kernel<2,9>" if not filter_on else "ingest_kernel<3,2> partition mode") + " (pack + hash + partition into tiles)",
             mean("ms_partition_kernel"), read_b),
            ("dedupe_tile_kernel<3,2>" + (" fused with pass 0" if fused else "") + (" + pass-1 tiles" if emitted else "") +
             " (exact dedupe in shared-memory tiles)",
             mean("ms_dedupe_kernel"), U * (R + 4) + (3 * U * R if fused else 0) + (U * R if emitted else 0)),
            (f"{passes_left} pass(es): " + ("" if emitted else "bucket_partition_kernel + ") + "bucket_tile_kernel + apply_edges_kernel",
             mean("ms_neighbour"), 3 * passes_left * U * R - (U * R if emitted else 0)),
            ("root_best_kernel + select_kernel (dissection + keep bitmap)", mean("ms_select"), 9 * U),
        ]
    else:
        groups = [
            ("ingest_kernel<3,2> (filter + pack + exact dedupe, HBM table)", mean("ms_ingest"), ingest_b),
            (f"{st.n_passes} pass(es): sig_count + scan + scatter_fat + compare_fat", mean("ms_neighbour"), 3 * st.n_passes * U * R),
            ("root_best_kernel + select_kernel (dissection + keep bitmap)", mean("ms_select"), 9 * U),
        ]
    kernels = [{"kernel": k, "ms": ms, "algorithmic_bytes": int(b), "achieved_gbs": b / (ms / 1e3) / 1e9 if ms > 0 else None,
                "frac": b / (ms / 1e3) / 1e9 / peak if ms > 0 else None} for k, ms, b in groups]
    dom = max(kernels, key=lambda k: k["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            for name, val in json.load(open(tpath)).items():
                if not name.startswith("_") and dom["kernel"].startswith(name):
                    traffic = val
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom["kernel"],
                "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes"],
                "kernel_ms": dom["ms"],
                "whole_path": {"algorithmic_bytes": total_b,
                               "bytes_per_unique": total_b / max(U, 1),
                               "achieved_gbs": total_b / (ms_per_step / 1e3) / 1e9,
                               "frac": total_b / (ms_per_step / 1e3) / 1e9 / peak},
                "kernels": kernels,
                "plan": {"streaming_dedupe": streaming, "pass0_fused": fused, "pass1_tiles_from_dedupe": emitted,
                         "streaming_passes": bool(st.plan_flags & 2)}}

    # ---- CPU baseline on a bounded sample of the same workload ----
    sample = min(n, env_int("FQD_CPU_SAMPLE", 4_000_000))
    uniq_s, secs, split, kind = time_reference(
        cfg, host_keys.array[:sample], None if host_quals is None else host_quals.array[:sample])
    cpu = {"value": uniq_s / secs, "unit": UNIT, "cores": 1, "kind": kind,
           "sample": f"first {sample} of {n} reads ({uniq_s} unique keys) in {secs:.1f}s {split}; "
                     f"single thread (the reference has no threading), "
                     f"host has {os.cpu_count()} logical cores"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": workload_config(cfg, 1),
        "unique_keys": int(U), "clusters": int(st.number_of_clusters),
        "selected": int(st.number_selected), "candidate_pairs": int(st.candidate_pairs),
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "e2e": {"value": U / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
                "steps": e2e_steps, "h2d_ms": est.ms_h2d},
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
    launches = int(sum(s.launches for s in stats))
    lt = torch.tensor([launches], dtype=torch.int64, device=f"cuda:{local_rank}")
    dist.all_reduce(lt)
    if rank == 0:
        total_b, _ = algorithmic_bytes(cfg, n, st.number_of_sequences, U, st.n_passes,
                                       cfg.max_average_error_rate < 1.0)
        peak, peak_src = measured_peak()
        line = {
            "metric": METRIC, "value": U / (ms_per_step / 1e3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic", "config": workload_config(cfg, world),
            "unique_keys": int(U), "clusters": int(st.number_of_clusters),
            "selected": int(st.number_selected), "candidate_pairs": int(st.candidate_pairs),
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "e2e": {"value": U / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(n * L * (2 if host_quals is not None else 1)),
                    "d2h_bytes_per_step": int(((n + 31) // 32) * 4), "ms_per_step": 1e3 * e2e_s, "steps": e2e_steps},
            "gpu_launches": int(lt.item()),
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
