"""Pins the CPU oracle (oracle/fqd_oracle.c): (a) the known answers of the reference's own
tests for this path, restated as data; (b) the unmodified reference compiled into
oracle/_ref, on randomised inputs.  CPU only."""
import math

import numpy as np
import pytest

METHODS = ("directional", "adjacency", "highest_count")

# /root/reference/tests/test__distance.py:25-30
HAMMING_KATS = [("AAAA", "AAAA", 0, True), ("AAAA", "AAA", 3, False), ("AAAA", "AAAC", 1, True),
                ("AAAA", "AAAC", 0, False), ("AACA", "AAAC", 2, True), ("AACC", "CCAA", 3, False)]
# /root/reference/tests/test__distance.py:40-55
EDIT_KATS = [("AAAA", "AAAA", 0, True), ("AAAA", "AAA", 1, True), ("AAAA", "A", 3, True),
             ("AAA", "C", 2, False), ("AAA", "C", 3, True), ("AAAA", "AAAC", 1, True),
             ("AAAA", "AAAC", 0, False), ("AACA", "AAAC", 2, True), ("AACC", "CCAA", 3, False),
             ("GATTACA", "GATTAA", 1, True), ("GATTACA", "GATTAA", 0, False), ("GC", "AAAGC", 3, True),
             ("AAAGC", "GC", 3, True), ("GC", "AAAGC", 2, False), ("ABCDE", "ABDE", 1, True),
             ("ABCDE", "ABDEF", 2, True)]


@pytest.mark.parametrize("a,b,d,want", HAMMING_KATS)
def test_hamming_kats(oracle, a, b, d, want):
    assert oracle.within_distance(a.encode(), b.encode(), d) is want


@pytest.mark.parametrize("a,b,d,want", EDIT_KATS)
def test_edit_kats(oracle, a, b, d, want):
    assert oracle.within_distance(a.encode(), b.encode(), d, True) is want


def test_average_error_rate_kats(oracle):
    # /root/reference/tests/test__fastq.py:8, :12 -- exact equality on doubles
    assert oracle.average_error_rate(bytes([10, 30]), 0) == 0.0505
    assert oracle.average_error_rate(bytes([43, 63])) == 0.0505
    # SURVEY appendix C.6
    assert oracle.average_error_rate(b"?" * 12) == 0.0010000000000000002
    assert oracle.average_error_rate(b"I" * 12) <= 0.001
    assert math.isnan(oracle.average_error_rate(b""))


@pytest.mark.parametrize("i", list(range(33)) + [127, 128, 255])
def test_average_error_rate_out_of_range(oracle, i):
    # /root/reference/tests/test__fastq.py:15-19
    with pytest.raises(ValueError, match="outside of valid phred range"):
        oracle.average_error_rate(bytes([i]))


def test_lut_equals_reference_for_every_score(oracle, reference):
    from fastqdedup._fastq import average_error_rate as ref_rate
    for c in range(33, 127):
        assert oracle.average_error_rate(bytes([c])) == ref_rate(chr(c)) == 10 ** -((c - 33) / 10)
    rng = np.random.default_rng(0)
    for _ in range(2000):
        q = bytes(rng.integers(33, 127, size=int(rng.integers(1, 80))).astype(np.uint8))
        assert oracle.average_error_rate(q) == ref_rate(q.decode())   # same sequential double sum


def _labels_to_sets(reads, res):
    """clusters as sets of (count, key) like the reference's tests compare them"""
    by_label = {}
    for f, c, l in zip(res["first"].tolist(), res["count"].tolist(), res["label"].tolist()):
        by_label.setdefault(l, set()).add((c, reads[f].decode()))
    return sorted(by_label.values(), key=sorted)


TRIE_READS = [b"AAAA", b"AAAA", b"AAAC", b"AAGC", b"AGGC", b"CCCG", b"CCCG", b"TTCA", b"TTCC",
              b"TTTA", b"TTT", b"TTC"]


def test_pop_cluster_goldens(oracle):
    # /root/reference/tests/test_trie.py:75-106 (Hamming) and :109-136 (edit)
    ham = _labels_to_sets(TRIE_READS, oracle.cluster(TRIE_READS, None, 1, False, "highest_count"))
    assert ham == sorted([{(2, "AAAA"), (1, "AAGC"), (1, "AAAC"), (1, "AGGC")}, {(2, "CCCG")},
                          {(1, "TTCA"), (1, "TTCC"), (1, "TTTA")}, {(1, "TTT"), (1, "TTC")}], key=sorted)
    edit = _labels_to_sets(TRIE_READS, oracle.cluster(TRIE_READS, None, 1, True, "highest_count"))
    assert edit == sorted([{(2, "AAAA"), (1, "AAGC"), (1, "AAAC"), (1, "AGGC")}, {(2, "CCCG")},
                           {(1, "TTCA"), (1, "TTCC"), (1, "TTTA"), (1, "TTT"), (1, "TTC")}], key=sorted)


def _expand(cluster):
    return [s.encode() for c, s in cluster for _ in range(c)]


def _selected(reads, res):
    return {reads[i].decode() for i in res["selected_first"].tolist()}


def test_dissection_goldens(oracle):
    # /root/reference/tests/test_fastqdedup.py:38-64
    cluster = [(3, "AAAGT"), (10, "AAAAT"), (50, "AACAA"), (60, "AAAAA"), (10, "CAAAA"), (30, "CTAAA")]
    reads = _expand(cluster)
    assert _selected(reads, oracle.cluster(reads, None, 1, False, "highest_count")) == {"AAAAA"}
    assert _selected(reads, oracle.cluster(reads, None, 1, False, "adjacency")) == {"AAAAA", "CTAAA", "AAAGT"}
    assert _selected(reads, oracle.cluster(reads, None, 1, False, "directional")) == {"AACAA", "AAAAA", "CTAAA"}
    # :73-92 long singleton chain
    chain = [(100, "GGGGGG"), (1, "GGGTGG"), (1, "GGGTTG"), (1, "GGCTTG"), (1, "GACTTG"), (2, "AACTTG")]
    reads = _expand(chain)
    assert _selected(reads, oracle.cluster(reads, None, 1, False, "directional")) == {"GGGGGG", "AACTTG"}
    # :94-97
    reads = _expand([(7, "AAAA"), (1, "AAAT"), (1, "CAAA")])
    for m in METHODS:
        assert _selected(reads, oracle.cluster(reads, None, 1, False, m)) == {"AAAA"}


def test_tie_break_is_ascii_largest(oracle):
    # equal counts -> ASCII-largest key wins, whatever the insertion order
    for reads in ([b"AAAA", b"AAAT"], [b"AAAT", b"AAAA"]):
        for m in METHODS:
            assert _selected(reads, oracle.cluster(reads, None, 1, False, m)) == {"AAAT"}
    reads = [b"AAAN", b"AAAT", b"AAAG"]
    assert _selected(reads, oracle.cluster(reads, None, 1, False, "highest_count")) == {"AAAT"}
    reads = [b"AAAN", b"AAAG"]
    assert _selected(reads, oracle.cluster(reads, None, 1, False, "highest_count")) == {"AAAN"}


def _same(a, b):
    for f in ("total_records", "discarded_records", "number_of_sequences", "number_of_uniques",
              "number_of_clusters", "number_selected"):
        assert a[f] == b[f], f
    for f in ("first", "count", "label", "selected"):
        assert np.array_equal(a[f], b[f]), f


@pytest.mark.parametrize("seed", range(12))
def test_oracle_equals_reference_on_random_inputs(oracle, reference, seed):
    rng = np.random.default_rng(seed)
    alphabet = [b"AC", b"ACG", b"ACGT", b"ACGTN"][seed % 4]
    lo, hi = [(3, 3), (3, 6), (4, 5), (2, 7)][seed % 4]
    strings = [bytes(rng.choice(list(alphabet), size=int(rng.integers(lo, hi + 1))).astype(np.uint8))
               for _ in range(400)]
    reads = [strings[i] for i in rng.integers(0, len(strings), size=1500)]
    quals = [bytes(rng.choice([ord("I"), ord("?"), ord("5"), ord("+")], p=[0.9, 0.05, 0.03, 0.02],
                              size=len(r)).astype(np.uint8)) for r in reads]
    for edit in (False, True):
        for d in (0, 1, 2, 3):
            for m in METHODS:
                _same(oracle.cluster(reads, quals, d, edit, m, 0.001),
                      oracle.ref_cluster(reads, quals, d, edit, m, 0.001))
                _same(oracle.cluster(reads, None, d, edit, m, 1.0),
                      oracle.ref_cluster(reads, None, d, edit, m, 1.0))


def test_distance_predicates_equal_reference(oracle, reference):
    from fastqdedup._distance import within_distance as ref_within
    rng = np.random.default_rng(99)
    for _ in range(20000):
        al = [b"AC", b"ACGT", b"ACGTN"][int(rng.integers(0, 3))]
        a = bytes(rng.choice(list(al), size=int(rng.integers(0, 10))).astype(np.uint8))
        b = bytearray(a)
        for _ in range(int(rng.integers(0, 5))):
            op = int(rng.integers(0, 3))
            if op == 0 and b:
                b[int(rng.integers(0, len(b)))] = al[int(rng.integers(0, len(al)))]
            elif op == 1:
                b.insert(int(rng.integers(0, len(b) + 1)), al[int(rng.integers(0, len(al)))])
            elif b:
                del b[int(rng.integers(0, len(b)))]
        b = bytes(b)
        d = int(rng.integers(0, 5))
        for e in (False, True):
            assert oracle.within_distance(a, b, d, e) == ref_within(a.decode(), b.decode(), d, e), (a, b, d, e)
