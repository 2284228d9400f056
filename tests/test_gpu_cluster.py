"""Parity of the CUDA clustering path (through the C ABI) with the CPU oracle and with the
unmodified reference, on seeded inputs at sizes the checkers finish in seconds.
Bit-exact: this is integer / byte / index work (the filter's doubles included)."""
import os

import numpy as np
import pytest

from fastqdedup_b200 import synth
from fastqdedup_b200.clustering import cluster_keys

pytestmark = pytest.mark.gpu

STAT_FIELDS = ("total_records", "discarded_records", "number_of_sequences",
               "number_of_uniques", "number_of_clusters", "number_selected")
METHODS = ("directional", "adjacency", "highest_count")


def assert_same(got, want, tag=""):
    for f in STAT_FIELDS:
        assert getattr(got, f) == want[f], (tag, f, getattr(got, f), want[f])
    for f in ("first", "count", "label", "selected"):
        assert np.array_equal(getattr(got, f), want[f]), (tag, f)
    assert np.array_equal(got.selected_first, want["selected_first"]), tag
    # the bitmap is the same set
    assert np.array_equal(np.nonzero(got.keep_mask())[0].astype(np.uint64), want["selected_first"]), tag


def run_both(checker, keys, quals=None, d=1, edit=False, method="directional", err=1.0,
             lengths=None, ctx=None, tag=""):
    kin, qin = keys, quals
    if lengths is not None:
        kin = synth.to_ragged(keys, lengths)
        qin = None if quals is None else synth.to_ragged(quals, lengths)
    want = checker(kin, qin, d, edit, method, err)
    got = cluster_keys(keys, quals, d, edit, method, err, lengths=lengths, context=ctx)
    assert_same(got, want, tag)
    if lengths is not None:   # the ragged (offsets) input form must agree too
        got2 = cluster_keys(kin, qin, d, edit, method, err, context=ctx)
        assert_same(got2, want, tag + "/ragged")
    return got


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("name,n,d", [
    ("cfg1", 30000, 1), ("cfg2", 20000, 1), ("cfg5", 20000, 1),
    ("cfg3", 12000, 2), ("cfg4", 6000, 1), ("cfg4", 6000, 2), ("cfg5", 8000, 0),
    ("cfg5", 6000, 3),
])
def test_configs_vs_oracle(gpu_ctx, oracle, name, n, d, method):
    cfg = synth.CONFIGS[name].scaled(n)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    run_both(oracle.cluster, keys, quals, d, cfg.use_edit_distance, method,
             cfg.max_average_error_rate, ctx=gpu_ctx, tag=f"{name}/{method}/d{d}")


@pytest.mark.parametrize("name,n,d,method", [
    ("cfg1", 1_000_000, 1, "directional"),       # BASELINE config 1 at full size
    ("cfg2", 400_000, 1, "adjacency"),
    ("cfg2", 400_000, 1, "highest_count"),
    ("cfg3", 300_000, 2, "directional"),
    ("cfg4", 150_000, 1, "directional"),
    ("cfg4", 150_000, 2, "directional"),
    ("cfg5", 500_000, 1, "directional"),
])
def test_configs_vs_reference(gpu_ctx, oracle, reference, name, n, d, method):
    """Mid-size runs against the compiled, unmodified reference (oracle/_ref)."""
    cfg = synth.CONFIGS[name].scaled(n)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    run_both(oracle.ref_cluster, keys, quals, d, cfg.use_edit_distance, method,
             cfg.max_average_error_rate, ctx=gpu_ctx, tag=f"{name}/{method}/d{d}/ref")


@pytest.mark.parametrize("edit", [False, True])
@pytest.mark.parametrize("method", METHODS)
def test_truncated_reads_mixed_lengths(gpu_ctx, oracle, edit, method):
    """Reads shorter than the check length give shorter keys: never linked under Hamming,
    linked under Levenshtein (appendix C.2)."""
    from dataclasses import replace
    cfg = replace(synth.CONFIGS["cfg4"].scaled(5000), truncate_frac=0.05, use_edit_distance=edit,
                  indel_rate=0.002 if edit else 0.0)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    assert lens is not None and lens.min() < cfg.key_length
    for d in (1, 2):
        run_both(oracle.cluster, keys, None, d, edit, method, 1.0, lengths=lens, ctx=gpu_ctx,
                 tag=f"trunc/{edit}/{method}/d{d}")


def _rand_strings(rng, n, alphabet, lo, hi):
    out = []
    for _ in range(n):
        ln = int(rng.integers(lo, hi + 1))
        out.append(bytes(rng.choice(list(alphabet), size=ln).astype(np.uint8)))
    return out


@pytest.mark.parametrize("seed", range(6))
def test_random_short_keys_all_modes(gpu_ctx, oracle, seed):
    """Dense little graphs: short keys over tiny alphabets percolate into big components,
    with ties, prefixes, N and lower-case bytes (appendix C.1-5, C.11)."""
    rng = np.random.default_rng(1000 + seed)
    alphabet = [b"AC", b"ACGT", b"ACGTN", b"ACGTNacgt", b"AT", b"ACGTNRYKM"][seed]
    lo, hi = [(3, 6), (4, 4), (5, 7), (4, 6), (1, 8), (6, 6)][seed]
    strings = _rand_strings(rng, 3000, alphabet, lo, hi)
    # heavy duplication so counts vary
    idx = rng.integers(0, len(strings), size=9000)
    reads = [strings[i] for i in idx]
    for edit in (False, True):
        for d in (0, 1, 2, 3):
            for method in METHODS:
                want = oracle.cluster(reads, None, d, edit, method, 1.0)
                got = cluster_keys(reads, None, d, edit, method, 1.0, context=gpu_ctx)
                assert_same(got, want, f"rand{seed}/{edit}/d{d}/{method}")


def test_empty_and_tiny_inputs(gpu_ctx, oracle):
    got = cluster_keys([], None, 1, False, "directional", context=gpu_ctx)
    assert got.total_records == 0 and got.number_of_uniques == 0 and got.number_selected == 0
    for reads in ([b"ACGT"], [b"", b""], [b"", b"A", b"AA"], [b"A"] * 5):
        for edit in (False, True):
            for method in METHODS:
                want = oracle.cluster(reads, None, 1, edit, method, 1.0)
                got = cluster_keys(reads, None, 1, edit, method, 1.0, context=gpu_ctx)
                assert_same(got, want, f"tiny/{reads}/{edit}/{method}")


def test_long_keys(gpu_ctx, oracle):
    """Whole 150-bp reads as keys (no --check-lengths) and 2x150 pairs."""
    rng = np.random.default_rng(7)
    for L in (150, 300):
        mol = rng.integers(0, 4, size=(300, L), dtype=np.uint8)
        ids = rng.integers(0, 300, size=3000)
        k = mol[ids].copy()
        err = rng.random(k.shape) < 0.004
        k[err] = (k[err] + 1) & 3
        keys = np.frombuffer(b"ACGT", dtype=np.uint8)[k]
        for edit, d in ((False, 1), (False, 2), (True, 1), (True, 2)):
            run_both(oracle.cluster, keys, None, d, edit, "directional", ctx=gpu_ctx,
                     tag=f"long{L}/{edit}/d{d}")


def test_filter_edge_cases(gpu_ctx, oracle):
    """Appendix C.6-7: all-Q30 x12 sums to 0.0010000000000000002 > 0.001 (discarded);
    'I' kept; empty quality => NaN => kept; the first occurrence of a selected key may
    itself be a filtered record and is still the one emitted."""
    keys = [b"ACGTACGTACGT", b"ACGTACGTACGT", b"ACGTACGTACGA", b"TTTTTTTTTTTT", b"GGGGGGGGGGGG",
            b"GGGGGGGGGGGG", b"CCCCCCCCCCCC"]
    quals = [b"?" * 12, b"I" * 12, b"I" * 12, b"?" * 12, b"I" * 10 + b"55", b"I" * 12, b""]
    for method in METHODS:
        want = oracle.cluster(keys, quals, 1, False, method, 0.001)
        got = cluster_keys(keys, quals, 1, False, method, 0.001, context=gpu_ctx)
        assert_same(got, want, method)
    assert want["discarded_records"] == 3            # both '?'*12 and the Q20 tail record
    assert 0 in want["selected_first"].tolist()      # record 0 was filtered but is emitted
    # -E (threshold 1.0) switches the filter off whatever the qualities are
    want = oracle.cluster(keys, quals, 1, False, "directional", 1.0)
    got = cluster_keys(keys, quals, 1, False, "directional", 1.0, context=gpu_ctx)
    assert_same(got, want, "filter off")
    assert got.discarded_records == 0


def test_bad_phred_aborts(gpu_ctx, oracle):
    from fastqdedup_b200._native import FqdPhredError
    keys = [b"ACGT"] * 6
    quals = [b"IIII", b"IIII", b"II I", b"IIII", b"I\x7fII", b"IIII"]
    with pytest.raises(FqdPhredError) as e:
        cluster_keys(keys, quals, 1, False, "directional", 0.001, context=gpu_ctx)
    assert "outside of valid phred range" in str(e.value)
    assert e.value.record == 2 and e.value.char == ord(" ")   # first offending record, like the reference
    with pytest.raises(oracle.PhredError):
        oracle.cluster(keys, quals, 1, False, "directional", 0.001)


def test_pre_counted_records(gpu_ctx, oracle):
    """record_counts: a (count, sequence) list equals the expanded read list."""
    rng = np.random.default_rng(3)
    strings = list({bytes(rng.choice(list(b"ACGT"), size=8).astype(np.uint8)) for _ in range(1500)})
    counts = rng.integers(1, 40, size=len(strings)).astype(np.uint32)
    expanded = [s for s, c in zip(strings, counts) for _ in range(int(c))]
    for method in METHODS:
        want = oracle.cluster(expanded, None, 1, False, method, 1.0)
        got = cluster_keys(strings, None, 1, False, method, 1.0, counts=counts, context=gpu_ctx)
        assert got.number_of_sequences == want["number_of_sequences"]
        assert got.number_of_clusters == want["number_of_clusters"]
        assert got.number_selected == want["number_selected"]
        sel_got = sorted(strings[i] for i in got.selected_first.tolist())
        sel_want = sorted(expanded[i] for i in want["selected_first"].tolist())
        assert sel_got == sel_want


def test_negative_distance_rejected(gpu_ctx):
    with pytest.raises(ValueError, match="non-negative"):
        cluster_keys([b"AC"], None, -1, False, "directional", context=gpu_ctx)


@pytest.fixture
def env():
    """Set library switches for one test and restore them afterwards."""
    saved = {}

    def set_(**kv):
        for k, v in kv.items():
            saved.setdefault(k, os.environ.get(k))
            os.environ[k] = str(v)
    yield set_
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize("dense", [0, 1])
@pytest.mark.parametrize("method", METHODS)
def test_levenshtein_compare_kernels_agree_with_oracle(gpu_ctx, oracle, env, dense, method):
    """Both compare kernels of the Levenshtein passes (per-entry walk, dense tiles) forced onto the same oracle-sized
    inputs: short blocks with big buckets (24-nt keys, d = 2 and 3), truncated reads (ragged keys, PAD in play), and
    keys beyond 31 / 63 symbols (64-bit diagonals, Myers fallback)."""
    from dataclasses import replace
    env(FQD_COMPARE_DENSE=dense)
    cfg = replace(synth.CONFIGS["cfg4"].scaled(7000), truncate_frac=0.04, indel_rate=0.003)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    for d in (1, 2, 3):
        run_both(oracle.cluster, keys, None, d, True, method, 1.0, lengths=lens, ctx=gpu_ctx, tag=f"lev/dense{dense}/{method}/d{d}")
    rng = np.random.default_rng(7 + dense)
    for L in (40, 70):
        base = rng.choice(list(b"ACGT"), size=(300, L)).astype(np.uint8)
        reads = base[rng.integers(0, 300, size=3000)].copy()
        hit = rng.random(reads.shape) < 0.01
        reads[hit] = rng.choice(list(b"ACGTN"), size=int(hit.sum())).astype(np.uint8)
        run_both(oracle.cluster, reads, None, 2, True, method, 1.0, ctx=gpu_ctx, tag=f"lev/dense{dense}/{method}/L{L}")
