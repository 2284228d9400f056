"""The streaming plan of large jobs (csrc/partitioned.cuh: partition by hash into shared-memory
sized tiles, one thread block per tile -- for the exact dedupe and for the Hamming passes) forced
onto small inputs with FQD_PARTITION_MIN, so the oracle can check it -- and its spill path /
fallback to the single-table / counting-sort plan when a partition outgrows its tile."""
import ctypes
import os
from dataclasses import replace

import numpy as np
import pytest

from fastqdedup_b200 import synth
from fastqdedup_b200.clustering import cluster_keys
from fastqdedup_b200.multigpu import cluster_keys_sharded_local
from test_gpu_cluster import METHODS, assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def force_partitioned():
    old = os.environ.get("FQD_PARTITION_MIN")
    os.environ["FQD_PARTITION_MIN"] = "1"
    yield
    if old is None:
        del os.environ["FQD_PARTITION_MIN"]
    else:
        os.environ["FQD_PARTITION_MIN"] = old


@pytest.mark.parametrize("name,n,d", [("cfg1", 40000, 1), ("cfg5", 60000, 1), ("cfg3", 30000, 2),
                                      ("cfg4", 8000, 2), ("cfg2", 30000, 1)])
def test_partitioned_plan_vs_oracle(gpu_ctx, oracle, name, n, d):
    cfg = synth.CONFIGS[name].scaled(n)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    for method in METHODS:
        want = oracle.cluster(keys, quals, d, cfg.use_edit_distance, method, cfg.max_average_error_rate)
        got = cluster_keys(keys, quals, d, cfg.use_edit_distance, method, cfg.max_average_error_rate,
                           context=gpu_ctx)
        assert_same(got, want, f"{name}/{method}")
        assert got.stats["plan_flags"] & 1          # partitioned dedupe
        if not cfg.use_edit_distance and name != "cfg1":
            assert got.stats["plan_flags"] & 2      # partitioned Hamming passes


def test_partitioned_plan_vs_reference_midsize(gpu_ctx, oracle, reference):
    cfg = synth.CONFIGS["cfg5"].scaled(600_000)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    want = oracle.ref_cluster(keys, None, 1, False, "directional", 1.0)
    got = cluster_keys(keys, None, 1, False, "directional", 1.0, context=gpu_ctx)
    assert_same(got, want, "cfg5 600k")


def test_partitioned_filter_ragged_pad_and_alphabet(gpu_ctx, oracle):
    cfg = replace(synth.CONFIGS["cfg3"].scaled(9000), truncate_frac=0.05, key_length=24, max_distance=1)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    kin, qin = synth.to_ragged(keys, lens), synth.to_ragged(quals, lens)
    for edit in (False, True):
        want = oracle.cluster(kin, qin, 1, edit, "directional", 0.001)
        assert want["discarded_records"] > 0
        assert_same(cluster_keys(kin, qin, 1, edit, "directional", 0.001, context=gpu_ctx), want, "ragged")
        assert_same(cluster_keys(keys, quals, 1, edit, "directional", 0.001, lengths=lens, context=gpu_ctx), want, "rows")
    rng = np.random.default_rng(4)
    reads = [bytes(rng.choice(list(b"ACGTNacgtRY"), size=8).astype(np.uint8)) for _ in range(5000)]
    assert_same(cluster_keys(reads, None, 1, False, "adjacency", 1.0, context=gpu_ctx),
                oracle.cluster(reads, None, 1, False, "adjacency", 1.0), "alphabet growth")
    # first occurrence filtered, kept copy later; a key that only ever appears filtered
    k = [b"ACGTACGTACGT", b"ACGTACGTACGT", b"ACGTACGTACGA", b"TTTTTTTTTTTT", b"CCCCCCCCCCCC"]
    q = [b"?" * 12, b"I" * 12, b"I" * 12, b"?" * 12, b"I" * 12]
    want = oracle.cluster(k, q, 1, False, "directional", 0.001)
    assert_same(cluster_keys(k, q, 1, False, "directional", 0.001, context=gpu_ctx), want, "filtered first")
    assert 0 in want["selected_first"].tolist() and want["number_of_uniques"] == 3


def test_partition_overflow_falls_back(gpu_ctx, oracle):
    """One key dominating the input overflows its partition: the single-table plan takes over."""
    rng = np.random.default_rng(9)
    reads = [b"GGGGGGGGGGGG"] * 600000 + [bytes(rng.choice(list(b"ACGT"), size=12).astype(np.uint8)) for _ in range(3000)]
    perm = rng.permutation(len(reads))
    reads = [reads[i] for i in perm[:200000]]
    want = oracle.cluster(reads, None, 1, False, "directional", 1.0)
    got = cluster_keys(reads, None, 1, False, "directional", 1.0, context=gpu_ctx)
    assert_same(got, want, "skewed")
    assert not (got.stats["plan_flags"] & 1)     # not the partitioned dedupe


def test_partitioned_plan_in_sharded_jobs(gpu_ctx, oracle):
    from fastqdedup_b200 import _native
    ctxs = [_native.Context(0) for _ in range(3)]
    try:
        cfg = synth.CONFIGS["cfg3"].scaled(12000)
        keys, lens, quals = synth.SynthSource(cfg).reads()
        for method in METHODS:
            want = oracle.cluster(keys, quals, 2, False, method, 0.001)
            got = cluster_keys_sharded_local(keys, quals, 2, False, method, 0.001, world=3, contexts=ctxs)
            assert_same(got, want, method)
    finally:
        for c in ctxs:
            c.close()


def test_bucket_partition_overflow_falls_back(gpu_ctx, oracle):
    """Most keys share their first pigeonhole block: that pass's partition overflows and is redone
    by the counting-sort plan; the other pass stays partitioned."""
    rng = np.random.default_rng(11)
    tails = rng.choice(list(b"ACGT"), size=(6000, 8)).astype(np.uint8)
    reads = [b"ACGTACGT" + bytes(t) for t in tails] + \
            [bytes(rng.choice(list(b"ACGT"), size=16).astype(np.uint8)) for _ in range(3000)]
    reads = [reads[i] for i in rng.integers(0, len(reads), size=30000)]
    for method in METHODS:
        want = oracle.cluster(reads, None, 1, False, method, 1.0)
        got = cluster_keys(reads, None, 1, False, method, 1.0, context=gpu_ctx)
        assert_same(got, want, f"skewed block/{method}")
        assert got.stats["plan_flags"] & 1 and not (got.stats["plan_flags"] & 2)


def test_oversize_partitions_take_the_spill_path(gpu_ctx, oracle):
    """A few keys with thousands of copies outgrow their tiles: those partitions and the spilled
    records are deduplicated by the single-table kernels, everything else stays in tiles."""
    cfg = synth.CONFIGS["cfg5"].scaled(50000)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    rng = np.random.default_rng(5)
    heavy = np.concatenate([np.repeat(keys[7:8], 5000, axis=0), np.repeat(keys[11:12], 2500, axis=0),
                            np.repeat(keys[500:501], 1500, axis=0)])
    allk = np.concatenate([keys, heavy])
    allk = allk[rng.permutation(len(allk))]
    for method in METHODS:
        want = oracle.cluster(allk, None, 1, False, method, 1.0)
        got = cluster_keys(allk, None, 1, False, method, 1.0, context=gpu_ctx)
        assert_same(got, want, f"oversize/{method}")
        assert got.stats["plan_flags"] & 1


@pytest.mark.parametrize("L", [12, 24, 36, 48])
def test_lean_partition_kernel_rows_and_foreign_bytes(gpu_ctx, oracle, L):
    """Fixed-stride ACGTN rows of 12/24/36/48 nt take the lean partition kernel; a foreign byte anywhere
    makes the job re-run with the grown alphabet (through the general kernel) and still match."""
    rng = np.random.default_rng(L)
    mol = rng.choice(list(b"ACGT"), size=(400, L)).astype(np.uint8)
    rows = mol[rng.integers(0, len(mol), size=6000)].copy()
    err = rng.random(rows.shape) < 0.01
    rows[err] = rng.choice(list(b"ACGTN"), size=int(err.sum())).astype(np.uint8)
    for method in METHODS:
        want = oracle.cluster(rows, None, 1, False, method, 1.0)
        got = cluster_keys(rows, None, 1, False, method, 1.0, context=gpu_ctx)
        assert_same(got, want, f"lean/{L}/{method}")
        assert got.stats["plan_flags"] & 1
    dirty = rows.copy()
    dirty[17, L - 1] = ord("a")
    dirty[4000, 0] = ord("R")
    want = oracle.cluster(dirty, None, 1, False, "directional", 1.0)
    assert_same(cluster_keys(dirty, None, 1, False, "directional", 1.0, context=gpu_ctx), want, f"lean/{L}/foreign bytes")


def test_many_spilled_uniques_redo_pass0_as_a_ranged_pass(gpu_ctx, oracle):
    """Dozens of heavy keys: thousands of uniques leave the dedupe stage through the spill path; the
    fused pass 0 is completed by a streaming pass over just those uniques (not by brute force)."""
    cfg = synth.CONFIGS["cfg5"].scaled(50000)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    rng = np.random.default_rng(8)
    heavy = np.concatenate([np.repeat(keys[i:i + 1], 500, axis=0) for i in range(100, 200)])
    allk = np.concatenate([keys, heavy])
    allk = allk[rng.permutation(len(allk))]
    for method in ("directional", "highest_count"):
        want = oracle.cluster(allk, None, 1, False, method, 1.0)
        got = cluster_keys(allk, None, 1, False, method, 1.0, context=gpu_ctx)
        assert_same(got, want, f"many spilled/{method}")
        assert got.stats["plan_flags"] & 5 == 5      # streaming dedupe with fused pass 0


def test_pass1_tiles_from_the_dedupe_stage_overflow_and_are_redone(gpu_ctx, oracle, reference):
    """Hardly any duplicates: far more uniques than the dedupe stage guessed (n/2) overflow the pass-1
    tiles it fills; pass 1 is partitioned again the ordinary way and the result still matches."""
    rng = np.random.default_rng(21)
    mol = rng.choice(list(b"ACGT"), size=(150000, 24)).astype(np.uint8)
    extra = mol[rng.integers(0, len(mol), size=30000)].copy()
    pos = rng.integers(0, 24, size=len(extra))
    extra[np.arange(len(extra)), pos] = rng.choice(list(b"ACGT"), size=len(extra)).astype(np.uint8)
    rows = np.concatenate([mol, extra])
    rows = rows[rng.permutation(len(rows))]
    want = oracle.ref_cluster(rows, None, 1, False, "directional", 1.0)
    got = cluster_keys(rows, None, 1, False, "directional", 1.0, context=gpu_ctx)
    assert_same(got, want, "pass-1 tile overflow")
    assert got.stats["plan_flags"] & 7 == 7


def test_streaming_plan_with_record_multiplicities(gpu_ctx, oracle):
    """record_counts through the streaming plan: weights accumulate in the tiles, duplicates of a
    pre-counted key still merge, the statistics count reads."""
    rng = np.random.default_rng(3)
    strings = list({bytes(rng.choice(list(b"ACGT"), size=12).astype(np.uint8)) for _ in range(3000)})
    strings = strings + strings[:500]                       # some keys appear twice in the pre-counted list
    counts = rng.integers(1, 40, size=len(strings)).astype(np.uint32)
    expanded = [s for s, c in zip(strings, counts) for _ in range(int(c))]
    for method in METHODS:
        want = oracle.cluster(expanded, None, 1, False, method, 1.0)
        got = cluster_keys(strings, None, 1, False, method, 1.0, counts=counts, context=gpu_ctx)
        assert got.stats["plan_flags"] & 1
        assert got.number_of_sequences == want["number_of_sequences"]
        assert got.number_of_uniques == want["number_of_uniques"]
        assert got.number_of_clusters == want["number_of_clusters"]
        assert got.number_selected == want["number_selected"]
        sel_got = sorted(strings[i] for i in got.selected_first.tolist())
        sel_want = sorted(expanded[i] for i in want["selected_first"].tolist())
        assert sel_got == sel_want


def test_more_oversize_partitions_than_a_launch_grid_has_rows(gpu_ctx):
    """57 M records of 110 k keys with 520 copies each over ~560 k sparsely filled tiles: every occupied 512-record
    tile is oversize -- about 100 k of them, more than the 65535 gridDim.y could hold when the spill launch still put
    the partition index there -- while the surplus (8 records of most keys) stays well inside the spill buffer, so the
    streaming plan has to see it through.  Checked against the known answer (keys built so that no two are within
    distance 1: every key its own cluster, all selected, first occurrence = its index in the first round)."""
    os.environ["FQD_TILE_FILL_PCT"] = "20"
    try:
        n_keys, copies, half = 110_000, 520, 10
        digits = ((np.arange(n_keys)[:, None] >> (2 * np.arange(half)[None, :])) & 3).astype(np.uint8)
        base = np.frombuffer(b"ACGT", dtype=np.uint8)[np.concatenate([digits, digits], axis=1)]   # two keys differ in >= 2 symbols
        keys = np.tile(base, (copies, 1))
        got = cluster_keys(keys, None, 1, False, "directional", 1.0, context=gpu_ctx, want_uniques=True)
    finally:
        del os.environ["FQD_TILE_FILL_PCT"]
    n = n_keys * copies
    assert got.total_records == n and got.number_of_uniques == n_keys
    assert got.number_of_clusters == n_keys and got.number_selected == n_keys
    assert np.array_equal(got.first, np.arange(n_keys, dtype=np.uint64))
    assert np.all(got.count == copies) and np.all(got.selected)
    assert np.array_equal(np.nonzero(got.keep_mask())[0], np.arange(n_keys))
    assert got.stats["plan_flags"] & 1


@pytest.mark.parametrize("split", [0, 1])
def test_tile_kernels_split_launches(gpu_ctx, oracle, split):
    """The tile kernels as one full-size launch and as the small-tile + full-size pair (FQD_TILE_SPLIT), on inputs whose
    tiles fall on both sides of the 384-record boundary (fill 70 %) and include oversize tiles (a key with 3000
    copies): dedupe with fused pass 0, the pass-1 tiles it emits, and a d = 2 job with an ordinary third pass."""
    os.environ["FQD_TILE_SPLIT"] = str(split)
    os.environ["FQD_TILE_FILL_PCT"] = "70"
    try:
        for name, n, d in (("cfg5", 90000, 1), ("cfg3", 40000, 2)):
            cfg = synth.CONFIGS[name].scaled(n)
            keys, lens, quals = synth.SynthSource(cfg).reads()
            keys = np.concatenate([keys, np.tile(keys[:1], (3000, 1))])
            if quals is not None:
                quals = np.concatenate([quals, np.tile(quals[:1], (3000, 1))])
            for method in ("directional", "highest_count"):
                want = oracle.cluster(keys, quals, d, False, method, cfg.max_average_error_rate)
                got = cluster_keys(keys, quals, d, False, method, cfg.max_average_error_rate, context=gpu_ctx)
                assert_same(got, want, f"split{split}/{name}/{method}")
                assert got.stats["plan_flags"] & 1
    finally:
        del os.environ["FQD_TILE_SPLIT"]
        del os.environ["FQD_TILE_FILL_PCT"]


@pytest.mark.parametrize("hybrid", [0, 1])
def test_host_jobs_packed_and_raw_chunks(gpu_ctx, oracle, hybrid):
    """HOST jobs with fixed-length ACGTN keys: the keys cross PCIe packed by the host threads, or (hybrid) as a mix of
    packed and raw ASCII chunks when the caller's buffer is page-locked; several 4 M-record chunks, the result must be
    the device-resident one.  A foreign byte in a late chunk sends the job down the ASCII path with a grown alphabet."""
    from fastqdedup_b200 import _native
    from fastqdedup_b200._native import ClusterJob, METHODS as METHOD_IDS, MEM_HOST
    from fastqdedup_b200.clustering import cluster_device
    lib = _native.load()
    os.environ["FQD_HOST_PACK_HYBRID"] = str(hybrid)
    try:
        n, L = 9_000_000, 12
        cfg = synth.CONFIGS["cfg1"].scaled(n)
        keys, _, _ = synth.SynthSource(cfg).reads()
        for foreign in (False, True):
            if foreign:
                keys[n - 5, 3] = ord("R")
            hptr = ctypes.c_void_p()
            assert lib.fqd_host_alloc(n * L, ctypes.byref(hptr)) == 0
            host = np.ctypeslib.as_array(ctypes.cast(hptr, ctypes.POINTER(ctypes.c_uint8)), shape=(n, L))
            host[:] = keys
            words = (n + 31) // 32
            bm = np.zeros(words, dtype=np.uint32)
            job = ClusterJob()
            job.n_records = n
            job.keys = hptr.value
            job.key_stride = job.key_length = L
            job.max_distance = 1
            job.method = METHOD_IDS["directional"]
            job.memory_space = MEM_HOST
            job.max_average_error_rate = 1.0
            job.phred_offset = 33
            st = gpu_ctx.cluster(job, bm.ctypes.data)
            dkeys = gpu_ctx.upload(keys)
            dbm = gpu_ctx.device_alloc(words * 4)
            ref = cluster_device(gpu_ctx, n, dkeys, L, max_distance=1, method="directional", bitmap_ptr=dbm)
            want = gpu_ctx.download(dbm, words * 4, np.uint32)
            gpu_ctx.device_free(dkeys)
            gpu_ctx.device_free(dbm)
            lib.fqd_host_free(hptr)
            assert st.number_of_uniques == ref.number_of_uniques and st.number_selected == ref.number_selected
            assert np.array_equal(bm, want), f"hybrid{hybrid}/foreign{foreign}"
            assert st.h2d_bytes > 0              # (a job that falls back to the single-table plan uploads twice)
    finally:
        del os.environ["FQD_HOST_PACK_HYBRID"]
