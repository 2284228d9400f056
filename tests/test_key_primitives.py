"""The packed-key primitives the kernels are made of (fastqdedup_b200/csrc/key.cuh),
compiled for the host and checked against the oracle / plain Python on the CPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "probe", "key_probe.cpp")
OUT = os.path.join(ROOT, "tests", "probe", "build", "key_probe.so")

OPS = dict(hamming=0, myers=1, less=2, equal=3, length=4, blockeq=5, hash=6, symbol=7, swar=8, fixed=9,
           block0eq=10, hash32=11, shd=12, edit=13)


@pytest.fixture(scope="module")
def probe():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "fastqdedup_b200", "csrc", "key.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++",
                        "-I" + os.path.join(ROOT, "fastqdedup_b200", "csrc"), SRC, "-o", OUT], check=True)
    lib = ctypes.CDLL(OUT)
    lib.key_probe.argtypes = [ctypes.c_int] * 3 + [ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p,
                                                  ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                                  ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                                  ctypes.POINTER(ctypes.c_uint64)]

    def call(K, PW, op, alphabet, varlen, a, b, max_len, d=0, p0=0, p1=0, p2=0):
        out = ctypes.c_uint64()
        r = lib.key_probe(K, PW, OPS[op], alphabet, len(alphabet), int(varlen), a, len(a), b, len(b),
                          max_len, d, p0, p1, p2, ctypes.byref(out))
        assert r >= 0 or (op in ("swar", "fixed") and r == -2), r
        return out.value if op in ("hash", "hash32") else r
    return call


def mutate(rng, s, alphabet, nedits, indel):
    s = list(s)
    for _ in range(nedits):
        op = rng.integers(0, 3 if indel else 1)
        if op == 0 and s:
            s[rng.integers(0, len(s))] = alphabet[rng.integers(0, len(alphabet))]
        elif op == 1:
            s.insert(rng.integers(0, len(s) + 1), alphabet[rng.integers(0, len(alphabet))])
        elif s:
            del s[rng.integers(0, len(s))]
    return bytes(s)


CASES = [(3, 1, b"ACGTN", 32), (3, 2, b"ACGTN", 64), (3, 2, b"ACGTN", 36), (3, 3, b"ACGTN", 90),
         (3, 5, b"ACGTN", 150), (4, 2, b"ACGTNacgtn", 48), (8, 1, bytes(range(33, 127)), 24),
         (8, 2, bytes(range(33, 127)), 50)]


@pytest.mark.parametrize("K,PW,alphabet,L", CASES)
def test_hamming_and_myers_match_oracle(probe, oracle, K, PW, alphabet, L):
    rng = np.random.default_rng(K * 100 + PW + L)
    for it in range(1500):
        la = L if it % 3 else int(rng.integers(0, L + 1))
        a = bytes(rng.choice(list(alphabet), size=la).astype(np.uint8))
        b = mutate(rng, a, alphabet, int(rng.integers(0, 5)), indel=it % 2 == 0)[:L]
        d = int(rng.integers(0, 4))
        varlen = True
        assert probe(K, PW, "hamming", alphabet, varlen, a, b, L, d) == int(oracle.within_distance(a, b, d, False)), (a, b, d)
        assert probe(K, PW, "myers", alphabet, varlen, a, b, L, d) == int(oracle.within_distance(a, b, d, True)), (a, b, d)
        if len(a) == len(b):   # fixed-length mode (no PAD)
            assert probe(K, PW, "hamming", alphabet, False, a, b, len(a), d) == int(oracle.within_distance(a, b, d, False))
            assert probe(K, PW, "myers", alphabet, False, a, b, len(a), d) == int(oracle.within_distance(a, b, d, True))


@pytest.mark.parametrize("K,PW,alphabet,L", CASES)
def test_order_equality_length(probe, K, PW, alphabet, L):
    """key_less is Python's str order (ASCII, proper prefix first: sorted() on (count, str)
    at reference __init__.py:68); equality/length see the PAD padding."""
    rng = np.random.default_rng(K * 1000 + PW + L)
    for it in range(1500):
        la = int(rng.integers(0, L + 1))
        a = bytes(rng.choice(list(alphabet), size=la).astype(np.uint8))
        if it % 4 == 0:
            b = a[:int(rng.integers(0, la + 1))]
        elif it % 4 == 1:
            b = mutate(rng, a, alphabet, 1, False)
        else:
            b = bytes(rng.choice(list(alphabet), size=int(rng.integers(0, L + 1))).astype(np.uint8))
        assert probe(K, PW, "less", alphabet, True, a, b, L) == int(a < b), (a, b)
        assert probe(K, PW, "less", alphabet, True, b, a, L) == int(b < a), (a, b)
        assert probe(K, PW, "equal", alphabet, True, a, b, L) == int(a == b)
        assert probe(K, PW, "length", alphabet, True, a, b, L) == len(a)
        if la:
            p = int(rng.integers(0, la))
            code = probe(K, PW, "symbol", alphabet, True, a, b, L, p0=p)
            if alphabet == b"ACGTN":
                assert code == (a[p] >> 1) & 7          # the table-free DNA code
            else:
                assert alphabet[code] == a[p]


@pytest.mark.parametrize("K,PW,alphabet,L", CASES[:5])
def test_block_hash_is_position_independent(probe, K, PW, alphabet, L):
    """The shifted blocks of the Levenshtein pigeonhole hash equal wherever they sit."""
    rng = np.random.default_rng(5)
    for it in range(800):
        a = bytes(rng.choice(list(alphabet), size=L).astype(np.uint8))
        ln = int(rng.integers(0, L // 2 + 1))
        p0 = int(rng.integers(0, L - ln + 1))
        p1 = int(rng.integers(0, L - ln + 1))
        b = bytearray(rng.choice(list(alphabet), size=L).astype(np.uint8))
        b[p1:p1 + ln] = a[p0:p0 + ln]
        assert probe(K, PW, "blockeq", alphabet, False, a, bytes(b), L, p0=p0, p1=p1, p2=ln) == 1
        if ln >= 4:
            q = (p1 + 1) % (L - ln + 1)
            same = bytes(b[q:q + ln]) == a[p0:p0 + ln]
            assert probe(K, PW, "blockeq", alphabet, False, a, bytes(b), L, p0=p0, p1=q, p2=ln) == int(same)


def test_hash_separates_keys(probe):
    rng = np.random.default_rng(11)
    seen = {}
    for _ in range(20000):
        a = bytes(rng.choice(list(b"ACGTN"), size=36).astype(np.uint8))
        h = probe(3, 2, "hash", b"ACGTN", False, a, a, 36)
        assert seen.setdefault(h, a) == a


@pytest.mark.parametrize("PW,L", [(1, 12), (1, 32), (2, 36), (2, 48), (2, 64), (3, 90), (5, 150)])
def test_table_free_dna_packing(probe, PW, L):
    """pack_key_acgtn (4 bytes per step, validity by bit algebra) == the table packer for every
    length / padding, and rejects every byte that is not one of ACGTN."""
    rng = np.random.default_rng(PW * 100 + L)
    for it in range(1500):
        la = L if it % 2 else int(rng.integers(0, L + 1))
        a = bytes(rng.choice(list(b"ACGTN"), size=la).astype(np.uint8))
        for varlen in (False, True):
            assert probe(3, PW, "swar", b"ACGTN", varlen, a, a, L) == 1, (a, varlen)
    for byte in range(256):
        for pos in (0, 1, 2, 3, 5, L - 1):
            a = bytearray(b"ACGT" * 40)[:L]
            a[pos] = byte
            r = probe(3, PW, "swar", b"ACGTN", False, bytes(a), bytes(a), L)
            assert r == (1 if byte in b"ACGTN" else -2), (byte, pos, r)


@pytest.mark.parametrize("PW,L", [(1, 12), (1, 24), (2, 24), (2, 36), (2, 48)])
def test_straight_line_dna_packing(probe, PW, L):
    """pack_key_acgtn_fixed (the lean partition kernel's packer) == the table packer on keys of exactly
    12/24/36/48 symbols, and rejects every foreign byte at every position class."""
    rng = np.random.default_rng(7 * PW + L)
    for _ in range(1500):
        a = bytes(rng.choice(list(b"ACGTN"), size=L).astype(np.uint8))
        assert probe(3, PW, "fixed", b"ACGTN", False, a, a, L) == 1, a
    for byte in range(256):
        for pos in (0, 1, 2, 3, 4, 7, L - 5, L - 1):
            a = bytearray(b"ACGT" * 12)[:L]
            a[pos] = byte
            r = probe(3, PW, "fixed", b"ACGTN", False, bytes(a), bytes(a), L)
            assert r == (1 if byte in b"ACGTN" else -2), (byte, pos, r)


def test_partition_hash_depends_on_the_leading_block_only(probe):
    """Records are partitioned by block0_hash: keys that share their first pigeonhole block must share
    a tile (the fused pass 0 relies on it), keys that differ there should not collide."""
    rng = np.random.default_rng(11)
    seen = {}
    for _ in range(2000):
        a = bytes(rng.choice(list(b"ACGTN"), size=36).astype(np.uint8))
        b = a[:18] + bytes(rng.choice(list(b"ACGTN"), size=18).astype(np.uint8))
        assert probe(3, 2, "block0eq", b"ACGTN", False, a, b, 36, p2=18) == 1
        c = bytearray(a)
        pos = int(rng.integers(0, 18))
        c[pos] = ord("A") if c[pos] != ord("A") else ord("C")
        assert probe(3, 2, "block0eq", b"ACGTN", False, a, bytes(c), 36, p2=18) == 0
        h = probe(3, 2, "hash32", b"ACGTN", False, a, a, 36)
        assert seen.setdefault(h, a) == a      # 2000 keys, 32 bits: a collision would be a bug, not bad luck


@pytest.mark.parametrize("K,PW,alphabet,L", CASES)
def test_shifted_hamming_filter_never_rejects_a_neighbour(probe, oracle, K, PW, alphabet, L):
    """The prefilter of the Levenshtein verify is a necessary condition: whatever is within the distance passes it
    (mutated pairs, all lengths, PAD in play); and it does turn most random pairs away."""
    rng = np.random.default_rng(K * 1000 + PW * 10 + L)
    rejected = total = 0
    for it in range(2500):
        la = L if it % 3 else int(rng.integers(0, L + 1))
        a = bytes(rng.choice(list(alphabet), size=la).astype(np.uint8))
        d = int(rng.integers(0, 5))
        if it % 5 == 0:
            b = bytes(rng.choice(list(alphabet), size=int(rng.integers(max(0, la - 2), min(L, la + 2) + 1))).astype(np.uint8))
        else:
            b = mutate(rng, a, alphabet, int(rng.integers(0, d + 2)), indel=True)[:L]
        for varlen in ((True, False) if len(a) == len(b) else (True,)):
            maxlen = L if varlen else len(a)
            passed = probe(K, PW, "shd", alphabet, varlen, a, b, maxlen, d)
            if oracle.within_distance(a, b, d, True):
                assert passed == 1, (a, b, d, varlen)
            elif it % 5 == 0 and la >= 12 and d <= 2:
                total += 1
                rejected += passed == 0
    if total > 50 and len(alphabet) <= 10:
        assert rejected > 0.7 * total, (rejected, total)


@pytest.mark.parametrize("K,PW,alphabet,L", CASES)
def test_edit_within_matches_oracle(probe, oracle, K, PW, alphabet, L):
    """edit_within (furthest-reaching diagonals for keys of up to 63 symbols and d <= 3, Myers otherwise) is the
    reference's within_edit_distance: mutated and random pairs, all lengths including empty, PAD in play."""
    rng = np.random.default_rng(K * 977 + PW * 13 + L)
    hits = 0
    for it in range(4000):
        la = L if it % 3 else int(rng.integers(0, L + 1))
        a = bytes(rng.choice(list(alphabet), size=la).astype(np.uint8))
        d = int(rng.integers(0, 5))
        if it % 5 == 0:
            b = bytes(rng.choice(list(alphabet), size=int(rng.integers(max(0, la - 3), min(L, la + 3) + 1))).astype(np.uint8))
        else:
            b = mutate(rng, a, alphabet, int(rng.integers(0, d + 3)), indel=True)[:L]
        want = bool(oracle.within_distance(a, b, d, True))
        hits += want
        for varlen in ((True, False) if len(a) == len(b) else (True,)):
            maxlen = L if varlen else len(a)
            got = probe(K, PW, "edit", alphabet, varlen, a, b, maxlen, d)
            assert got == int(want), (a, b, d, varlen, got, want)
            assert probe(K, PW, "edit", alphabet, varlen, b, a, maxlen if varlen else len(b), d) == int(want), (b, a, d)
    assert 400 < hits < 3600


def test_edit_within_short_keys_exhaustive(probe, oracle):
    """All pairs of strings of length <= 4 over two letters, d = 1..3: every boundary case of the diagonal form."""
    import itertools
    words = [bytes(w) for n in range(0, 5) for w in itertools.product(b"AC", repeat=n)]
    for a in words:
        for b in words:
            for d in (1, 2, 3):
                want = int(bool(oracle.within_distance(a, b, d, True)))
                assert probe(3, 1, "edit", b"ACGTN", True, a, b, 8, d) == want, (a, b, d)
