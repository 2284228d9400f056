"""The host-side key packer (csrc/host_pack.cpp, ``fqd_pack_keys``): 3 bits per symbol, the three code-bit planes of
a key back to back.  The bit layout is checked on the CPU against numpy; that a HOST job gives the same answer with
and without it is a GPU test."""
import ctypes
import os

import numpy as np
import pytest

from fastqdedup_b200 import _native


def pack(keys2d, stride=None):
    lib = _native.load()
    n, L = keys2d.shape
    stride = stride or L
    rows = np.zeros((n, stride), dtype=np.uint8)
    rows[:, :L] = keys2d
    rw = (3 * L + 31) // 32
    out = np.zeros((n, rw), dtype=np.uint32)
    bad = ctypes.c_uint64()
    rc = lib.fqd_pack_keys(rows.ctypes.data, n, L, stride, out.ctypes.data, ctypes.byref(bad))
    return rc, out, int(bad.value)


def numpy_pack(keys2d):
    n, L = keys2d.shape
    code = (keys2d >> 1) & 7                      # A 0, C 1, T 2, G 3, N 7 (key.cuh)
    bits = np.concatenate([(code >> p) & 1 for p in range(3)], axis=1).astype(np.uint8)    # plane p at bits [pL, (p+1)L)
    rw = (3 * L + 31) // 32
    padded = np.zeros((n, rw * 32), dtype=np.uint8)
    padded[:, :3 * L] = bits
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint32)


@pytest.mark.parametrize("L", [1, 4, 7, 12, 24, 31, 32, 33, 36, 48, 63, 64])
def test_packed_rows_bit_layout(L):
    rng = np.random.default_rng(L)
    keys = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, size=(20_001, L))]
    for stride in (L, L + 5):
        rc, got, bad = pack(keys, stride)
        assert rc == 0 and bad == len(keys)
        assert np.array_equal(got, numpy_pack(keys)), (L, stride)


def test_first_foreign_byte_is_reported():
    rng = np.random.default_rng(1)
    keys = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(300_000, 36))].copy()
    for row, col, byte in ((299_999, 35, ord("a")), (123_456, 0, ord("R")), (123_456, 7, 0), (5, 3, 0x4F)):
        keys[row, col] = byte
        rc, _, bad = pack(keys)
        assert rc == _native.ERR_UNSUPPORTED and bad == row
    with pytest.raises(NotImplementedError):
        _native.check(rc)


def test_argument_errors():
    lib = _native.load()
    buf = np.zeros(128, dtype=np.uint8)
    bad = ctypes.c_uint64()
    assert lib.fqd_pack_keys(buf.ctypes.data, 1, 65, 65, buf.ctypes.data, ctypes.byref(bad)) == _native.ERR_ARG
    assert lib.fqd_pack_keys(buf.ctypes.data, 1, 8, 4, buf.ctypes.data, ctypes.byref(bad)) == _native.ERR_ARG
    assert lib.fqd_pack_keys(None, 0, 8, 8, None, ctypes.byref(bad)) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("L", [12, 24, 36, 48])
def test_host_jobs_agree_with_and_without_the_packer(gpu_ctx, L):
    """A HOST job large enough to be streamed chunk by chunk: packed on the host (default) and as ASCII rows."""
    from dataclasses import replace

    from fastqdedup_b200 import synth
    from fastqdedup_b200.clustering import cluster_keys
    cfg = replace(synth.CONFIGS["cfg5"].scaled(4_300_000), key_length=L)
    keys, _, _ = synth.SynthSource(cfg).reads()
    runs = {}
    for name, env in (("packed", None), ("ascii", "1")):
        os.environ.pop("FQD_NO_HOST_PACK", None)
        if env:
            os.environ["FQD_NO_HOST_PACK"] = env
        runs[name] = cluster_keys(keys, None, 1, False, "directional", 1.0, context=gpu_ctx)
    os.environ.pop("FQD_NO_HOST_PACK", None)
    a, b = runs["packed"], runs["ascii"]
    assert a.stats["plan_flags"] & 1 and b.stats["plan_flags"] & 1
    for f in ("number_of_uniques", "number_of_clusters", "number_selected", "number_of_sequences"):
        assert getattr(a, f) == getattr(b, f), f
    for f in ("first", "count", "label", "selected", "keep_bitmap"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.gpu
def test_foreign_bytes_fall_back_to_the_ascii_path(gpu_ctx, oracle):
    from fastqdedup_b200 import synth
    from fastqdedup_b200.clustering import cluster_keys
    cfg = synth.CONFIGS["cfg5"].scaled(4_200_000)
    keys, _, _ = synth.SynthSource(cfg).reads()
    keys = keys.copy()
    keys[4_199_000, 5] = ord("a")            # in the last chunk: the packed chunks before it are thrown away
    keys[17, 30] = ord("R")
    got = cluster_keys(keys, None, 1, False, "directional", 1.0, context=gpu_ctx)
    os.environ["FQD_NO_HOST_PACK"] = "1"
    try:
        want = cluster_keys(keys, None, 1, False, "directional", 1.0, context=gpu_ctx)
    finally:
        del os.environ["FQD_NO_HOST_PACK"]
    assert got.stats["key_bits"] == 3 or got.stats["key_bits"] == 4     # the alphabet grew by two symbols
    for f in ("first", "count", "label", "selected", "keep_bitmap"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f


# ---- the plane streams fqd_cluster sends for HOST jobs (pack_planes_parallel; internal, reached through a host-only probe) ----

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def planes():
    import subprocess
    src = os.path.join(ROOT, "tests", "probe", "pack_probe.cpp")
    out = os.path.join(ROOT, "tests", "probe", "build", "pack_probe.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src, os.path.join(ROOT, "fastqdedup_b200", "csrc", "host_pack.cpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I" + os.path.join(ROOT, "include"),
                        src, "-o", out], check=True)
    lib = ctypes.CDLL(out)
    lib.pack_planes_probe.restype = ctypes.c_uint64
    lib.pack_planes_probe.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p]
    lib.plane_stream_words_probe.restype = ctypes.c_uint64
    lib.plane_stream_words_probe.argtypes = [ctypes.c_uint64, ctypes.c_uint32]

    def run(keys2d):
        n, L = keys2d.shape
        words = int(lib.plane_stream_words_probe(n, L))
        dst = np.full(3 * words, 0xDEADBEEFDEADBEEF, dtype=np.uint64)
        keys2d = np.ascontiguousarray(keys2d)
        bad = int(lib.pack_planes_probe(keys2d.ctypes.data, n, L, dst.ctypes.data))
        return bad, dst.reshape(3, words)
    return run


@pytest.mark.parametrize("n,L", [(1, 12), (5, 36), (16, 36), (64, 36), (1000, 36), (4099, 24), (100_003, 48), (7, 12),
                                 (300_000, 36)])
def test_plane_streams_bit_layout(planes, n, L):
    """Bit t of plane p = code bit p of symbol t of the chunk (rows back to back), the words behind the stream zero:
    what partition_planes_kernel cuts a row's L bits out of with a funnel shift."""
    rng = np.random.default_rng(n + L)
    keys = np.frombuffer(b"ACGTN", dtype=np.uint8)[rng.integers(0, 5, size=(n, L))]
    bad, streams = planes(keys)
    assert bad == n
    assert streams.shape[1] == (n * L + 63) // 64 + 1
    flat = keys.reshape(-1)
    nb = (n * L + 7) // 8
    for p in range(3):
        want = np.packbits((flat >> (p + 1)) & 1, bitorder="little")
        got = streams[p].view(np.uint8)
        assert np.array_equal(got[:nb], want), (n, L, p)
        assert not got[nb:].any(), (n, L, p)
    # what the GPU does for row r: two words and a funnel shift
    for r in rng.integers(0, n, size=min(n, 200)):
        bit = int(r) * L
        w, sh = bit >> 6, bit & 63
        for p in range(3):
            lo, hi = int(streams[p][w]), int(streams[p][w + 1])
            bits = ((lo >> sh) | (hi << (64 - sh) if sh else 0)) & ((1 << L) - 1)
            code = (keys[r] >> (p + 1)) & 1
            assert bits == sum(int(b) << i for i, b in enumerate(code)), (r, p)


def test_plane_streams_report_the_first_foreign_row(planes):
    rng = np.random.default_rng(5)
    n, L = 50_000, 36
    keys = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, L))]
    for byte in (ord("R"), ord("a"), 0, 255, ord("B"), ord("U")):
        k = keys.copy()
        r = int(rng.integers(0, n - 1))
        k[r, int(rng.integers(0, L))] = byte
        k[int(rng.integers(r, n)), 0] = ord("x")      # a later one does not matter
        bad, _ = planes(k)
        assert bad == r, (byte, bad, r)


def test_every_foreign_byte_value_is_rejected(planes):
    """All 251 byte values outside ACGTN, in both packers (0xFF once slipped through the validity table: its filler)."""
    base = np.frombuffer(b"ACGTN", dtype=np.uint8)[np.arange(40 * 36) % 5].reshape(40, 36)
    for b in range(256):
        k = base.copy()
        k[17, 20] = b
        rc, _, bad = pack(k)
        bad_planes, _ = planes(k)
        if bytes([b]) in (b"A", b"C", b"G", b"T", b"N"):
            assert rc == 0 and bad == 40 and bad_planes == 40, b
        else:
            assert rc == _native.ERR_UNSUPPORTED and bad == 17 and bad_planes == 17, b
