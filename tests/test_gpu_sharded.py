"""The sharded (multi-rank) plans, driven as virtual ranks inside one process on the one GPU a
test box has: every kernel and every exchange step of the multi-GPU path runs -- the tile kernels
fetch the fragments of their tiles through "peer" pointers that here are pointers into the other
contexts' arenas, the collectives are device copies instead of NCCL.  Results must equal the oracle
bit for bit, like the single-rank plan.  Both plans are covered: the tile-sharded plan (what Hamming
jobs take) and the replicated-set plan (Levenshtein jobs, long keys, and the fallback for skew),
forced with FQD_SHARD_REPLICATED=1."""
import os
from dataclasses import replace

import numpy as np
import pytest

from fastqdedup_b200 import synth
from fastqdedup_b200.multigpu import cluster_keys_sharded_local, shard_bounds
from test_gpu_cluster import METHODS, assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def contexts(gpu_ctx):
    from fastqdedup_b200 import _native
    ctxs = [_native.Context(0) for _ in range(8)]
    yield ctxs
    for c in ctxs:
        c.close()


PLAN_TILES = 16   # FQD_PLAN_SHARD_TILES


@pytest.fixture(params=["tiles", "replicated"])
def plan(request):
    """Which sharded plan runs: the default choice of the library, or the replicated-set plan forced."""
    old = os.environ.pop("FQD_SHARD_REPLICATED", None)
    if request.param == "replicated":
        os.environ["FQD_SHARD_REPLICATED"] = "1"
    yield request.param
    os.environ.pop("FQD_SHARD_REPLICATED", None)
    if old is not None:
        os.environ["FQD_SHARD_REPLICATED"] = old


def took_tiles(got):
    return bool(got.stats["plan_flags"] & PLAN_TILES)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("name,n,d", [("cfg5", 20000, 1), ("cfg1", 30000, 1), ("cfg3", 12000, 2),
                                      ("cfg4", 6000, 2), ("cfg2", 15000, 1), ("cfg5", 5000, 0), ("cfg5", 6000, 3)])
def test_sharded_equals_oracle(contexts, oracle, plan, name, n, d, world):
    cfg = synth.CONFIGS[name].scaled(n)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    for method in METHODS:
        want = oracle.cluster(keys, quals, d, cfg.use_edit_distance, method, cfg.max_average_error_rate)
        got = cluster_keys_sharded_local(keys, quals, d, cfg.use_edit_distance, method,
                                         cfg.max_average_error_rate, world=world,
                                         contexts=contexts[:world])
        assert_same(got, want, f"{name}/{method}/d{d}/world{world}/{plan}")
        # Hamming jobs with keys of up to 6 packed words take the tile-sharded plan (cfg1's 30 k reads of 12 nt
        # percolate into dense buckets: there the plan may hand over, which is the fallback being exercised)
        if plan == "tiles" and not cfg.use_edit_distance and name != "cfg1":
            assert took_tiles(got), f"{name}/{method}: the tile-sharded plan did not run"
        if plan == "replicated" or cfg.use_edit_distance:
            assert not took_tiles(got)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_tile_sharded_skewed_families(contexts, oracle, world):
    """Key families larger than a tile (oversize tiles -> spill path on the tile's owner, its uniques compared
    among themselves and re-emitted for pass 1), spread over all ranks."""
    rng = np.random.default_rng(11)
    L = 36
    base = rng.integers(0, 4, size=(60, L), dtype=np.uint8)
    rows = [np.repeat(base[:3], 700, axis=0)]                         # three families of 700 copies
    fam = np.repeat(base[3:9], 300, axis=0)                           # six families with many distinct error keys
    err = rng.random(fam.shape) < 0.02
    fam = np.where(err, (fam + rng.integers(1, 4, size=fam.shape, dtype=np.uint8)) & 3, fam)
    rows.append(fam)
    rows.append(rng.integers(0, 4, size=(3000, L), dtype=np.uint8))   # background
    codes = np.concatenate(rows)
    rng.shuffle(codes)
    keys = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    for method in METHODS:
        for d in (1, 2):
            want = oracle.cluster(keys, None, d, False, method, 1.0)
            got = cluster_keys_sharded_local(keys, None, d, False, method, 1.0, world=world, contexts=contexts[:world])
            assert_same(got, want, f"skew/{method}/d{d}/world{world}")


@pytest.mark.parametrize("edit", [False, True])
def test_sharded_mixed_lengths_and_filter(contexts, oracle, plan, edit):
    """Truncated reads (PAD in play), ragged input, quality filter with discarded first
    occurrences living on another rank than the kept copies."""
    cfg = replace(synth.CONFIGS["cfg3"].scaled(9000), truncate_frac=0.05, use_edit_distance=edit,
                  key_length=24, max_distance=1)
    keys, lens, quals = synth.SynthSource(cfg).reads()
    kin, qin = synth.to_ragged(keys, lens), synth.to_ragged(quals, lens)
    for method in METHODS:
        want = oracle.cluster(kin, qin, 1, edit, method, 0.001)
        assert want["discarded_records"] > 0
        for world in (2, 4):
            got = cluster_keys_sharded_local(kin, qin, 1, edit, method, 0.001, world=world,
                                             contexts=contexts[:world])
            assert_same(got, want, f"ragged/{edit}/{method}/{world}")
            got = cluster_keys_sharded_local(keys, quals, 1, edit, method, 0.001, lengths=lens,
                                             world=world, contexts=contexts[:world])
            assert_same(got, want, f"rows/{edit}/{method}/{world}")


def test_sharded_filtered_first_occurrence_on_other_rank(contexts, oracle, plan):
    # record 0 (rank 0) is filtered, its kept duplicate sits on rank 1: record 0 must be emitted
    keys = [b"ACGTACGTACGT", b"TTTTTTTTTTTT", b"GGGGGGGGGGGG", b"ACGTACGTACGT", b"ACGTACGTACGA", b"CCCCCCCCCCCC"]
    quals = [b"?" * 12, b"I" * 12, b"I" * 12, b"I" * 12, b"I" * 12, b"?" * 12]
    for method in METHODS:
        want = oracle.cluster(keys, quals, 1, False, method, 0.001)
        got = cluster_keys_sharded_local(keys, quals, 1, False, method, 0.001, world=2, contexts=contexts[:2])
        assert_same(got, want, method)
    assert 0 in want["selected_first"].tolist() and want["discarded_records"] == 2
    assert want["number_of_uniques"] == 4          # CCCCCCCCCCCC only ever appears filtered


def test_sharded_unknown_alphabet_and_errors(contexts, oracle, plan):
    from fastqdedup_b200._native import FqdPhredError
    rng = np.random.default_rng(2)
    reads = [bytes(rng.choice(list(b"ACGTNacgtRY"), size=8).astype(np.uint8)) for _ in range(4000)]
    want = oracle.cluster(reads, None, 1, False, "directional", 1.0)
    got = cluster_keys_sharded_local(reads, None, 1, False, "directional", 1.0, world=3, contexts=contexts[:3])
    assert_same(got, want, "alphabet growth")
    quals = [b"IIIIIIII"] * len(reads)
    quals[3001] = b"III\x1fIIII"
    with pytest.raises(FqdPhredError) as e:
        cluster_keys_sharded_local(reads, quals, 1, False, "directional", 0.001, world=3, contexts=contexts[:3])
    assert e.value.record == 3001 and e.value.char == 0x1f      # global index, whatever rank saw it


def test_sharded_more_ranks_than_records(contexts, oracle, plan):
    for reads in ([b"ACGT"], [b"ACGT", b"ACGA"], []):
        got = cluster_keys_sharded_local(reads, None, 1, False, "directional", 1.0, world=4, contexts=contexts[:4])
        want = oracle.cluster(reads, None, 1, False, "directional", 1.0)
        assert_same(got, want, str(reads))


def test_shard_bounds():
    assert shard_bounds(10, 3) == [0, 4, 7, 10]
    assert shard_bounds(2, 4) == [0, 1, 2, 2, 2]
    assert shard_bounds(0, 2) == [0, 0, 0]
