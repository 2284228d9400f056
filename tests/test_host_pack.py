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
