"""ctypes view of ``libfqd_b200.so`` (``include/fqd_b200.h``): the one place the Python
host code meets the C ABI.  No torch, no numpy-side arithmetic: arrays are only handed
over as pointers.

The library is loaded from this directory (in-tree build, ``fastqdedup_b200/build.py``).
There is no CPU fallback: if the library is missing ``load()`` raises, and if no CUDA
device is usable every compute call raises ``FqdCudaError``.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32,
                    c_size_t, c_uint8, c_uint32, c_uint64, c_void_p)

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfqd_b200.so")

FQD_OK = 0
ERR_ARG, ERR_PHRED, ERR_CUDA, ERR_NOMEM, ERR_LOOKUP, ERR_UNSUPPORTED, ERR_NCCL, ERR_FASTQ, ERR_IO = range(1, 10)
MEM_HOST, MEM_DEVICE = 0, 1
METHODS = {"highest_count": 0, "adjacency": 1, "directional": 2}


class FqdError(RuntimeError):
    pass


class FqdCudaError(FqdError):
    pass


class FqdPhredError(ValueError):
    """Byte outside '!'..'~' in a quality string (reference _fastqmodule.c:65-70)."""

    def __init__(self, msg, record=None, char=None):
        super().__init__(msg)
        self.record = record
        self.char = char


class FqdFastqError(Exception):
    """Malformed FASTQ input / inputs not in sync (the frontend re-raises it as dnaio's FastqFormatError)."""


class Slice(Structure):
    """include/fqd_b200.h fqd_slice: a Python slice object, has_x == 0 meaning None."""
    _fields_ = [("start", ctypes.c_int64), ("stop", ctypes.c_int64), ("step", ctypes.c_int64),
                ("has_start", c_uint8), ("has_stop", c_uint8), ("has_step", c_uint8), ("reserved", c_uint8 * 5)]

    @classmethod
    def from_python(cls, slc):
        out = cls()
        for name in ("start", "stop", "step"):
            v = getattr(slc, name)
            if v is not None:
                setattr(out, name, int(v))
                setattr(out, "has_" + name, 1)
        return out


class ClusterJob(Structure):
    _fields_ = [
        ("n_records", c_uint64),
        ("keys", c_void_p), ("key_offsets", c_void_p), ("key_lengths", c_void_p),
        ("key_stride", c_uint32), ("key_length", c_uint32),
        ("quals", c_void_p), ("qual_offsets", c_void_p), ("qual_lengths", c_void_p),
        ("qual_stride", c_uint32), ("qual_length", c_uint32),
        ("max_distance", c_int32), ("use_edit_distance", c_int32),
        ("method", c_int32), ("memory_space", c_int32),
        ("max_average_error_rate", c_double),
        ("phred_offset", c_uint8), ("reserved", c_uint8 * 7),
        ("alphabet", c_char_p),
        ("record_counts", c_void_p),
    ]


class ClusterStats(Structure):
    _fields_ = [
        ("total_records", c_uint64), ("discarded_records", c_uint64),
        ("number_of_sequences", c_uint64), ("number_of_uniques", c_uint64),
        ("number_of_clusters", c_uint64), ("number_selected", c_uint64),
        ("candidate_pairs", c_uint64), ("bad_record", c_uint64),
        ("bad_char", c_uint32), ("key_bits", c_uint32), ("key_words", c_uint32),
        ("n_passes", c_uint32),
        ("ms_total", c_float), ("ms_ingest", c_float), ("ms_gather", c_float),
        ("ms_neighbour", c_float), ("ms_select", c_float), ("ms_h2d", c_float),
        ("ms_compare", c_float), ("ms_ingest_kernel", c_float), ("ms_table_clear", c_float),
        ("ms_bucket_build", c_float), ("launches", c_uint32), ("plan_flags", c_uint32),
        ("ms_partition_kernel", c_float), ("ms_dedupe_kernel", c_float),
        ("own_uniques", c_uint64), ("h2d_bytes", c_uint64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


# every symbol include/fqd_b200.h declares (tests/test_abi.py checks the library exports them)
EXPORTS = [
    "fqd_last_error", "fqd_device_count", "fqd_context_create", "fqd_context_destroy",
    "fqd_default_context", "fqd_device_alloc", "fqd_device_free", "fqd_device_upload",
    "fqd_device_download", "fqd_device_memset", "fqd_context_synchronize", "fqd_host_alloc", "fqd_host_free",
    "fqd_cluster", "fqd_cluster_fetch", "fqd_cluster_fetch_selected",
    "fqd_nccl_unique_id", "fqd_comm_create", "fqd_comm_destroy", "fqd_cluster_sharded",
    "fqd_cluster_sharded_local",
    "fqd_average_error_rate", "fqd_within_distance", "fqd_int_peak",
    "fqd_fastq_scan_open", "fqd_fastq_scan_records", "fqd_fastq_scan_keys", "fqd_fastq_scan_quals",
    "fqd_fastq_scan_free", "fqd_fastq_emit", "fqd_pack_keys",
    "fqd_trie_new", "fqd_trie_free", "fqd_trie_add_sequence", "fqd_trie_contains_sequence",
    "fqd_trie_pop_cluster", "fqd_trie_cluster_item", "fqd_trie_number_of_sequences",
    "fqd_trie_alphabet", "fqd_trie_memory_size", "fqd_trie_raw_stats",
]

_lib = None


def load():
    """Load the native library; fails loudly when it was never built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FqdError(f"{LIB_PATH} is missing: build it with `python -m fastqdedup_b200.build` "
                       "(the CUDA path has no Python/CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.fqd_last_error.restype = c_char_p
    lib.fqd_device_count.restype = c_int
    lib.fqd_context_create.argtypes = [c_int, POINTER(c_void_p)]
    lib.fqd_context_destroy.argtypes = [c_void_p]
    lib.fqd_context_destroy.restype = None
    lib.fqd_default_context.argtypes = [POINTER(c_void_p)]
    lib.fqd_device_alloc.argtypes = [c_void_p, c_size_t, POINTER(c_void_p)]
    lib.fqd_device_free.argtypes = [c_void_p, c_void_p]
    lib.fqd_device_upload.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
    lib.fqd_device_download.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
    lib.fqd_device_memset.argtypes = [c_void_p, c_void_p, c_int, c_size_t]
    lib.fqd_context_synchronize.argtypes = [c_void_p]
    lib.fqd_int_peak.argtypes = [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_double)]
    lib.fqd_host_alloc.argtypes = [c_size_t, POINTER(c_void_p)]
    lib.fqd_host_free.argtypes = [c_void_p]
    lib.fqd_cluster.argtypes = [c_void_p, POINTER(ClusterJob), POINTER(ClusterStats), c_void_p]
    lib.fqd_nccl_unique_id.argtypes = [c_void_p]
    lib.fqd_comm_create.argtypes = [c_void_p, c_int, c_int, c_void_p, POINTER(c_void_p)]
    lib.fqd_comm_destroy.argtypes = [c_void_p]
    lib.fqd_comm_destroy.restype = None
    lib.fqd_cluster_sharded.argtypes = [c_void_p, c_void_p, POINTER(ClusterJob), c_uint64,
                                        POINTER(ClusterStats), c_void_p]
    lib.fqd_cluster_sharded_local.argtypes = [POINTER(c_void_p), c_int, POINTER(ClusterJob),
                                              POINTER(c_uint64), POINTER(ClusterStats),
                                              POINTER(c_void_p)]
    lib.fqd_cluster_fetch.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.fqd_cluster_fetch_selected.argtypes = [c_void_p, c_void_p]
    lib.fqd_average_error_rate.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_uint8,
                                           c_void_p, POINTER(c_uint64), POINTER(c_uint32)]
    lib.fqd_within_distance.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_uint64, c_int32, c_int32, c_void_p]
    lib.fqd_fastq_scan_open.argtypes = [POINTER(c_char_p), c_int, POINTER(Slice), c_int, c_int, POINTER(c_void_p)]
    lib.fqd_fastq_scan_records.argtypes = [c_void_p]
    lib.fqd_fastq_scan_records.restype = c_uint64
    lib.fqd_fastq_scan_keys.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_uint32)]
    lib.fqd_fastq_scan_quals.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_uint32)]
    lib.fqd_fastq_scan_free.argtypes = [c_void_p]
    lib.fqd_fastq_scan_free.restype = None
    lib.fqd_fastq_emit.argtypes = [POINTER(c_char_p), POINTER(c_char_p), c_int, c_void_p, c_uint64, c_int,
                                   POINTER(c_uint64)]
    lib.fqd_pack_keys.argtypes = [c_void_p, c_uint64, c_uint32, c_uint32, c_void_p, POINTER(c_uint64)]
    _lib = lib
    return lib


def check(rc, stats=None):
    if rc == FQD_OK:
        return
    msg = load().fqd_last_error().decode("latin-1")
    if rc == ERR_PHRED:
        raise FqdPhredError(msg, None if stats is None else stats.bad_record,
                            None if stats is None else stats.bad_char)
    if rc == ERR_ARG:
        raise ValueError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    if rc == ERR_LOOKUP:
        raise LookupError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_CUDA:
        raise FqdCudaError(msg)
    if rc == ERR_FASTQ:
        raise FqdFastqError(msg)
    if rc == ERR_IO:
        raise OSError(msg)
    raise FqdError(msg)


class FastqScan:
    """Pass 1 over the input files in native code (``fqd_fastq_scan_open``): per record tuple the key (and quality)
    rows ``deduplicate_cluster`` builds, as numpy views on the library's buffers."""

    def __init__(self, paths, slices=None, want_quals=False, threads=0):
        self.lib = load()
        n = len(paths)
        c_paths = (c_char_p * n)(*[os.fsencode(p) for p in paths])
        c_slices = None
        if slices:
            c_slices = (Slice * n)(*[Slice.from_python(s) for s in slices])
        h = c_void_p()
        check(self.lib.fqd_fastq_scan_open(c_paths, n, c_slices, int(bool(want_quals)), int(threads), byref(h)))
        self.handle = h
        self.n_records = int(self.lib.fqd_fastq_scan_records(h))
        self.keys = self._rows(self.lib.fqd_fastq_scan_keys)
        self.quals = self._rows(self.lib.fqd_fastq_scan_quals) if want_quals else None

    def _rows(self, getter):
        data, off, stride = c_void_p(), c_void_p(), c_uint32()
        check(getter(self.handle, byref(data), byref(off), byref(stride)))
        n = self.n_records
        if off.value:
            offsets = np.ctypeslib.as_array(ctypes.cast(off, POINTER(c_uint64)), shape=(n + 1,))
            nbytes = int(offsets[n]) if n else 0
            flat = np.ctypeslib.as_array(ctypes.cast(data, POINTER(c_uint8)), shape=(max(nbytes, 1),))
            return flat, offsets
        width = int(stride.value)
        if n == 0 or width == 0:
            return np.zeros(1, dtype=np.uint8), np.zeros(n + 1, dtype=np.uint64)
        return np.ctypeslib.as_array(ctypes.cast(data, POINTER(c_uint8)), shape=(n, width))

    def close(self):
        if self.handle:
            self.keys = self.quals = None
            self.lib.fqd_fastq_scan_free(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fastq_emit(in_paths, out_paths, keep_bitmap, n_records, threads=0):
    """Pass 2 in native code (``fqd_fastq_emit``): writes the record tuples whose bit is set; returns their number."""
    lib = load()
    n = len(in_paths)
    c_in = (c_char_p * n)(*[os.fsencode(p) for p in in_paths])
    c_out = (c_char_p * n)(*[os.fsencode(p) for p in out_paths])
    words = np.ascontiguousarray(keep_bitmap, dtype=np.uint32)
    written = c_uint64()
    check(lib.fqd_fastq_emit(c_in, c_out, n, words.ctypes.data if len(words) else None, int(n_records), int(threads),
                             byref(written)))
    return int(written.value)


class Context:
    """One GPU context (stream + memory pool + last result)."""

    def __init__(self, device=0):
        self.lib = load()
        h = c_void_p()
        check(self.lib.fqd_context_create(int(device), byref(h)))
        self.handle = h
        self.device = device

    def close(self):
        if self.handle:
            self.lib.fqd_context_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device buffers (inputs resident in HBM) ----
    def device_alloc(self, nbytes):
        p = c_void_p()
        check(self.lib.fqd_device_alloc(self.handle, nbytes, byref(p)))
        return p.value

    def device_free(self, ptr):
        check(self.lib.fqd_device_free(self.handle, c_void_p(ptr)))

    def upload(self, array):
        array = np.ascontiguousarray(array)
        ptr = self.device_alloc(array.nbytes)
        check(self.lib.fqd_device_upload(self.handle, c_void_p(ptr), array.ctypes.data, array.nbytes))
        return ptr

    def download(self, ptr, nbytes, dtype=np.uint8):
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        check(self.lib.fqd_device_download(self.handle, out.ctypes.data, c_void_p(ptr), nbytes))
        return out

    def memset(self, ptr, value, nbytes):
        check(self.lib.fqd_device_memset(self.handle, c_void_p(ptr), int(value), nbytes))

    def synchronize(self):
        check(self.lib.fqd_context_synchronize(self.handle))

    def int_peak(self):
        """Measured integer-issue peak of this GPU (thread-level operations per second)."""
        a, b, c = c_double(), c_double(), c_double()
        check(self.lib.fqd_int_peak(self.handle, byref(a), byref(b), byref(c)))
        return {"lop3_ops_per_s": a.value, "popc_ops_per_s": b.value, "mixed_ops_per_s": c.value}

    # ---- the batched job ----
    def cluster(self, job, keep_bitmap_ptr=None):
        stats = ClusterStats()
        rc = self.lib.fqd_cluster(self.handle, byref(job), byref(stats),
                                  c_void_p(keep_bitmap_ptr) if keep_bitmap_ptr else None)
        check(rc, stats)
        return stats

    def fetch(self, n_uniques):
        first = np.empty(n_uniques, dtype=np.uint64)
        count = np.empty(n_uniques, dtype=np.uint32)
        label = np.empty(n_uniques, dtype=np.uint64)
        sel = np.empty(n_uniques, dtype=np.uint8)
        check(self.lib.fqd_cluster_fetch(self.handle, first.ctypes.data, count.ctypes.data,
                                         label.ctypes.data, sel.ctypes.data))
        return first, count, label, sel

    def fetch_selected(self, n_selected):
        idx = np.empty(n_selected, dtype=np.uint64)
        check(self.lib.fqd_cluster_fetch_selected(self.handle, idx.ctypes.data))
        return idx

    # ---- function-level entry points ----
    def average_error_rate(self, strings, phred_offset=33):
        flat, off = _flatten(strings)
        out = np.empty(len(strings), dtype=np.float64)
        bad_i, bad_c = c_uint64(), c_uint32()
        rc = self.lib.fqd_average_error_rate(self.handle, flat.ctypes.data, off.ctypes.data,
                                             len(strings), phred_offset, out.ctypes.data,
                                             byref(bad_i), byref(bad_c))
        if rc == ERR_PHRED:
            raise FqdPhredError(self.lib.fqd_last_error().decode("latin-1"), bad_i.value, bad_c.value)
        check(rc)
        return out

    def within_distance(self, a_strings, b_strings, max_distance, use_edit_distance=False):
        a, ao = _flatten(a_strings)
        b, bo = _flatten(b_strings)
        out = np.empty(len(a_strings), dtype=np.uint8)
        check(self.lib.fqd_within_distance(self.handle, a.ctypes.data, ao.ctypes.data,
                                           b.ctypes.data, bo.ctypes.data, len(a_strings),
                                           int(max_distance), int(bool(use_edit_distance)),
                                           out.ctypes.data))
        return out.astype(bool)


def _flatten(strings):
    lens = np.fromiter((len(s) for s in strings), dtype=np.uint64, count=len(strings))
    off = np.zeros(len(strings) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    joined = b"".join(strings)
    flat = np.frombuffer(joined, dtype=np.uint8) if joined else np.zeros(1, dtype=np.uint8)
    return flat, off


_default = None


def default_context():
    """Process-wide context on the device the library picks ($FQD_DEVICE / $LOCAL_RANK / 0);
    the same one the CPython shims use."""
    global _default
    if _default is None:
        lib = load()
        h = c_void_p()
        check(lib.fqd_default_context(byref(h)))
        ctx = Context.__new__(Context)
        ctx.lib, ctx.handle, ctx.device = lib, h, None
        ctx.close = lambda: None      # owned by the library
        _default = ctx
    return _default
