"""The reference's own unit tests (tests/test_trie.py, test__distance.py, test__fastq.py,
test_fastqdedup.py under /root/reference: 91 cases) restated against the drop-in package:
same calls, same expected values, same exception types and messages, but every answer is
computed by the CUDA library behind the shims."""
import pytest

from test_oracle import EDIT_KATS, HAMMING_KATS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg(gpu_ctx):
    import fastqdedup_b200
    return fastqdedup_b200


# ---- tests/test_trie.py ------------------------------------------------------------------

def test_trie_one_seq(pkg):                                   # :22-30
    trie = pkg.Trie()
    trie.add_sequence("GATTACA")
    assert trie.contains_sequence("GATTACA", 0)
    assert trie.contains_sequence("AATTACA", 1)
    assert trie.contains_sequence("GATTACC", 1)
    assert trie.contains_sequence("GACCACA", 2)
    assert not trie.contains_sequence("GACCACA", 1)
    assert not trie.contains_sequence("GATTACC", 0)


def test_trie_one_seq_edit_distance(pkg):                     # :33-44
    trie = pkg.Trie()
    trie.add_sequence("GATTACA")
    yes = [("GATTACA", 0), ("AATTACA", 1), ("GATTACC", 1), ("GACCACA", 2), ("GATTAA", 1),
           ("GATTAC", 1), ("ATTAC", 2)]
    for s, d in yes:
        assert trie.contains_sequence(s, max_distance=d, use_edit_distance=True)
    assert not trie.contains_sequence("GACCACA", max_distance=1, use_edit_distance=True)
    assert not trie.contains_sequence("GATTACC", max_distance=0, use_edit_distance=True)


def test_trie_subseq(pkg):                                    # :47-53
    trie = pkg.Trie()
    trie.add_sequence("GATTACA")
    trie.add_sequence("GATTA")
    assert trie.contains_sequence("GATTA")
    assert trie.contains_sequence("GATTACA")
    assert not trie.contains_sequence("GATTAC")


@pytest.mark.parametrize(["sequence", "distance", "result"], [   # :56-72
    ("GATTA", 0, True), ("GATTACA", 0, True), ("GATTAC", 1, True), ("G", 4, True),
    ("GATTAT", 2, True), ("UU", 4, False), ("UU", 5, True), ("UUUUU", 3, False), ("ATTAC", 2, True)])
def test_trie_subseq_edit_distance(pkg, sequence, distance, result):
    trie = pkg.Trie()
    trie.add_sequence("GATTACA")
    trie.add_sequence("GATTA")
    assert trie.contains_sequence(sequence, distance, use_edit_distance=True) is result


POP_INPUT = ["AAAA", "AAAA", "AAAC", "AAGC", "AGGC", "CCCG", "CCCG", "TTCA", "TTCC", "TTTA", "TTT", "TTC"]


def test_trie_pop_cluster(pkg):                               # :75-106
    trie = pkg.Trie()
    for s in POP_INPUT:
        trie.add_sequence(s)
    cluster_list = []
    while True:
        try:
            cluster_list.append(trie.pop_cluster(1))
        except LookupError:
            break
    cluster_set = [set(c) for c in cluster_list]
    expected = [{(2, "AAAA"), (1, "AAGC"), (1, "AAAC"), (1, "AGGC")}, {(2, "CCCG")},
                {(1, "TTCA"), (1, "TTCC"), (1, "TTTA")}, {(1, "TTT"), (1, "TTC")}]
    for e in expected:
        assert e in cluster_set
        cluster_set.remove(e)
    assert not cluster_set


def test_trie_pop_cluster_edit_distance(pkg):                 # :109-136
    trie = pkg.Trie()
    for s in POP_INPUT:
        trie.add_sequence(s)
    cluster_list = []
    while trie.number_of_sequences:
        cluster_list.append(trie.pop_cluster(max_distance=1, use_edit_distance=True))
    cluster_set = [set(c) for c in cluster_list]
    expected = [{(2, "AAAA"), (1, "AAGC"), (1, "AAAC"), (1, "AGGC")}, {(2, "CCCG")},
                {(1, "TTCA"), (1, "TTCC"), (1, "TTTA"), (1, "TTT"), (1, "TTC")}]
    for e in expected:
        assert e in cluster_set
        cluster_set.remove(e)
    assert not cluster_set


def test_trie_new_with_alphabet(pkg):                         # :139-141
    assert pkg.Trie(alphabet="acd").alphabet == "acd"


def test_trie_alphabet_repeated(pkg):                         # :144-147
    with pytest.raises(ValueError) as error:
        pkg.Trie(alphabet="abcc")
    error.match("c was repeated")


def test_trie_alphabet_during_adding(pkg):                    # :150-158
    trie = pkg.Trie()
    trie.add_sequence("abc")
    trie.add_sequence("badabccdaafacb")
    assert trie.alphabet == "ab"
    trie.add_sequence("bcadac")
    assert trie.alphabet == "abc"


def test_trie_number_of_sequences(pkg):                       # :161-172
    trie = pkg.Trie()
    for s in ("abc", "ab", "abcd"):
        trie.add_sequence(s)
    assert trie.number_of_sequences == 3
    while True:
        try:
            trie.pop_cluster(0)
        except LookupError:
            break
    assert trie.number_of_sequences == 0


def test_trie_same_hand_out_order_and_errors(pkg, reference):
    """Beyond the reference's tests: cluster order, seed-first lists, interleaved adds and
    the error paths agree with the real trie."""
    seqs = ["GATTACA", "GATTACC", "TTTT", "TTTA", "TTT", "ACGT", "ACGA", "CCCC", "AC", "A", "ACG"]
    for d, edit in ((1, False), (1, True), (2, False), (0, False)):
        mine, ref = pkg.Trie(alphabet="ACGTN"), reference.Trie(alphabet="ACGTN")
        for s in seqs + seqs[:3]:
            mine.add_sequence(s); ref.add_sequence(s)
        got, want = [], []
        for k in range(3):
            got.append(mine.pop_cluster(d, edit)); want.append(ref.pop_cluster(d, edit))
        mine.add_sequence("TTTC"); ref.add_sequence("TTTC")     # legal after pop_cluster
        while ref.number_of_sequences:
            got.append(mine.pop_cluster(d, edit)); want.append(ref.pop_cluster(d, edit))
        assert mine.number_of_sequences == 0
        assert [c[0] for c in got] == [c[0] for c in want]          # same seeds in the same order
        assert [set(c) for c in got] == [set(c) for c in want]
        with pytest.raises(LookupError, match="No sequences left in Trie."):
            mine.pop_cluster(d, edit)
    t = pkg.Trie()
    t.add_sequence("AC")
    with pytest.raises(ValueError, match="non-negative"):
        t.pop_cluster(-1)
    with pytest.raises(TypeError, match="Sequence must be a str, got bytes"):
        t.add_sequence(b"AC")
    with pytest.raises(ValueError, match="ASCII"):
        t.add_sequence("é")
    assert pkg.Trie().contains_sequence("A") is False            # the reference segfaults here


# ---- tests/test__distance.py ---------------------------------------------------------------

@pytest.mark.parametrize(["string1", "string2", "max_distance", "result"], HAMMING_KATS)
def test_within_distance_hamming(pkg, string1, string2, max_distance, result):
    assert pkg.within_distance(string1, string2, max_distance) is result


@pytest.mark.parametrize(["string1", "string2", "max_distance", "result"], EDIT_KATS)
def test_within_distance_levenshtein(pkg, string1, string2, max_distance, result):
    assert pkg.within_distance(string1, string2, max_distance, use_edit_distance=True) is result


def test_within_distance_random_vs_reference(pkg, reference, gpu_ctx):
    import numpy as np
    from fastqdedup._distance import within_distance as ref_within
    rng = np.random.default_rng(5)
    a_list, b_list = [], []
    for _ in range(4000):
        a = bytes(rng.choice(list(b"ACGTN"), size=int(rng.integers(0, 40))).astype(np.uint8))
        b = bytearray(a)
        for _ in range(int(rng.integers(0, 4))):
            op = int(rng.integers(0, 3))
            if op == 0 and b:
                b[int(rng.integers(0, len(b)))] = b"ACGTN"[int(rng.integers(0, 5))]
            elif op == 1:
                b.insert(int(rng.integers(0, len(b) + 1)), b"ACGTN"[int(rng.integers(0, 5))])
            elif b:
                del b[int(rng.integers(0, len(b)))]
        a_list.append(a); b_list.append(bytes(b))
    for d in (0, 1, 2, 3):
        for edit in (False, True):
            got = gpu_ctx.within_distance(a_list, b_list, d, edit)
            want = [ref_within(a.decode(), b.decode(), d, edit) for a, b in zip(a_list, b_list)]
            assert got.tolist() == want


# ---- tests/test__fastq.py ------------------------------------------------------------------

def test_average_error_rate(pkg):                              # :6-8
    from fastqdedup_b200._fastq import average_error_rate
    assert average_error_rate(chr(10) + chr(30), phred_offset=0) == 0.0505


def test_average_error_rate_with_default_offset(pkg):          # :11-12
    from fastqdedup_b200._fastq import average_error_rate
    assert average_error_rate(chr(43) + chr(63)) == 0.0505


@pytest.mark.parametrize("i", list(range(33)) + [127])          # :15-19
def test_average_error_rate_out_of_range(pkg, i):
    from fastqdedup_b200._fastq import average_error_rate
    with pytest.raises(ValueError) as error:
        average_error_rate(chr(i))
    error.match(f"{chr(i)} outside of valid phred range")


def test_average_error_rate_non_ascii(pkg):                    # :22-25
    from fastqdedup_b200._fastq import average_error_rate
    with pytest.raises(ValueError) as error:
        average_error_rate(chr(128))
    error.match("phred_scores must be ASCII encoded")


def test_average_error_rate_batch_bit_exact(gpu_ctx, oracle):
    import math
    import numpy as np
    rng = np.random.default_rng(8)
    strings = [bytes(rng.integers(33, 127, size=int(rng.integers(0, 200))).astype(np.uint8)) for _ in range(5000)]
    got = gpu_ctx.average_error_rate(strings)
    for s, g in zip(strings, got.tolist()):
        w = oracle.average_error_rate(s)
        assert (math.isnan(w) and math.isnan(g)) or w == g, s


# ---- tests/test_fastqdedup.py --------------------------------------------------------------

TEST_CLUSTER = [(3, "AAAGT"), (10, "AAAAT"), (50, "AACAA"), (60, "AAAAA"), (10, "CAAAA"), (30, "CTAAA")]


def test_most_reads(pkg):                                      # :51-54
    dissected = list(pkg.cluster_dissection_highest_count(TEST_CLUSTER))
    assert dissected == ["AAAAA"]


def test_adjacency(pkg):                                       # :56-59
    dissected = list(pkg.cluster_dissection_adjacency(TEST_CLUSTER))
    assert len(dissected) == 3 and set(dissected) == {"AAAAA", "CTAAA", "AAAGT"}


def test_directional(pkg):                                     # :61-64
    dissected = list(pkg.cluster_dissection_directional(TEST_CLUSTER))
    assert len(dissected) == 3 and set(dissected) == {"AACAA", "AAAAA", "CTAAA"}


@pytest.mark.parametrize("name", ["directional", "adjacency", "highest_count"])
def test_no_list_aliasing(pkg, name):                          # :66-71
    cluster = TEST_CLUSTER[:]
    old = cluster[:]
    list(pkg.CLUSTER_DISSECTION_METHODS[name](cluster))
    assert old == cluster


def test_directional_long_chain(pkg):                          # :73-92
    cluster = [(100, "GGGGGG"), (1, "GGGTGG"), (1, "GGGTTG"), (1, "GGCTTG"), (1, "GACTTG"), (2, "AACTTG")]
    assert set(pkg.cluster_dissection_directional(cluster)) == {"GGGGGG", "AACTTG"}


@pytest.mark.parametrize("name", ["directional", "adjacency", "highest_count"])
def test_all_reads_same_cluster(pkg, name):                    # :94-97
    cluster = [(7, "AAAA"), (1, "AAAT"), (1, "CAAA")]
    assert set(pkg.CLUSTER_DISSECTION_METHODS[name](cluster)) == {"AAAA"}


def test_dissection_yield_order_equals_reference(pkg, reference):
    import numpy as np
    rng = np.random.default_rng(21)
    for trial in range(30):
        keys = list({"".join(rng.choice(list("ACGT"), size=5)) for _ in range(40)})
        cluster = [(int(rng.integers(1, 30)), k) for k in keys]
        for name in ("directional", "adjacency", "highest_count"):
            for d, e in ((1, False), (2, False), (1, True)):
                assert list(pkg.CLUSTER_DISSECTION_METHODS[name](cluster, d, e)) == \
                    list(reference.CLUSTER_DISSECTION_METHODS[name](cluster, d, e)), (name, d, e)


def test_trie_stats_report_equals_reference(pkg, reference):
    """raw_stats / memory_size (-v report, reference __init__.py:133-157): the shim has no nodes, it derives the node
    layout the reference's trie holds for the same sequences."""
    import random
    rng = random.Random(5)
    for alphabet, lo, hi, n in (("ACGTN", 4, 12, 400), ("ACGTN", 1, 6, 300), ("", 2, 9, 200), ("acgtRYKM", 3, 10, 250)):
        letters = alphabet or "ACGT"
        seqs = ["".join(rng.choice(letters) for _ in range(rng.randint(lo, hi))) for _ in range(n)]
        seqs += seqs[: n // 5] + [s[: len(s) // 2] for s in seqs[: n // 10] if len(s) > 1]     # duplicates, proper prefixes
        ours, ref = pkg.Trie(alphabet=alphabet), reference.Trie(alphabet=alphabet)
        for s in seqs:
            ours.add_sequence(s)
            ref.add_sequence(s)
        assert ours.alphabet == ref.alphabet
        assert ours.raw_stats() == ref.raw_stats()
        assert ours.memory_size() == ref.memory_size()
        assert pkg.trie_stats(ours) == reference.trie_stats(ref)
