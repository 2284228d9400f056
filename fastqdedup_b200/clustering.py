"""Host-side front end of the batched clustering job (``fqd_cluster`` in
``include/fqd_b200.h``): what ``deduplicate_cluster`` calls instead of the reference's
per-record ``add_sequence`` loop and per-cluster ``pop_cluster`` + dissection loop
(reference ``src/fastqdedup/__init__.py:242-252`` and ``:272-276``).

Only marshals arrays into the C ABI and results back; every comparison, count and
selection is made by the CUDA kernels.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _native
from ._native import METHODS, ClusterJob, MEM_DEVICE, MEM_HOST


@dataclass
class ClusterResult:
    total_records: int
    discarded_records: int
    number_of_sequences: int
    number_of_uniques: int
    number_of_clusters: int
    number_selected: int
    stats: dict
    keep_bitmap: Optional[np.ndarray] = None     # uint32 words, bit t%32 of word t/32
    first: Optional[np.ndarray] = None           # per unique, ascending
    count: Optional[np.ndarray] = None
    label: Optional[np.ndarray] = None
    selected: Optional[np.ndarray] = None

    @property
    def selected_first(self):
        """Ascending record indices pass 2 has to emit."""
        if self.first is not None:
            return self.first[self.selected]
        words = self.keep_bitmap
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:self.total_records]
        return np.nonzero(bits)[0].astype(np.uint64)

    def keep_mask(self):
        bits = np.unpackbits(self.keep_bitmap.view(np.uint8), bitorder="little")
        return bits[:self.total_records].astype(bool)


def _as_rows(x, lengths):
    """-> (flat uint8, offsets|None, lengths|None, stride, length)"""
    if isinstance(x, tuple):                      # (flat, offsets)
        flat = np.ascontiguousarray(x[0], dtype=np.uint8)
        off = np.ascontiguousarray(x[1], dtype=np.uint64)
        if flat.size == 0:
            flat = np.zeros(1, dtype=np.uint8)
        return flat, off, None, 0, 0, len(off) - 1
    if isinstance(x, np.ndarray) and x.ndim == 2:
        arr = np.ascontiguousarray(x, dtype=np.uint8)
        n, stride = arr.shape
        lens = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.uint32)
        flat = arr.reshape(-1)
        if flat.size == 0:
            flat = np.zeros(1, dtype=np.uint8)
        return flat, None, lens, stride, stride, n
    strings = list(x)
    flat, off = _native._flatten(strings)
    return flat, off, None, 0, 0, len(strings)


def cluster_keys(keys, quals=None, max_distance=1, use_edit_distance=False,
                 method="directional", max_average_error_rate=1.0, phred_offset=33,
                 lengths=None, qual_lengths=None, counts=None, alphabet=None,
                 context=None, want_bitmap=True, want_uniques=True) -> ClusterResult:
    """Cluster the keys of N records on the GPU.

    ``keys`` / ``quals``: a 2-D uint8 array (one row per record, optional per-row
    ``lengths``), a ``(flat, offsets)`` pair, or a sequence of ``bytes``.  ``counts``
    optionally gives each record a multiplicity.  Semantics: SURVEY.md appendix B.
    """
    ctx = context or _native.default_context()
    if method not in METHODS:
        raise ValueError(f"unknown cluster dissection method {method!r}")
    kflat, koff, klens, kstride, klen, n = _as_rows(keys, lengths)
    job = ClusterJob()
    job.n_records = n
    job.keys = kflat.ctypes.data
    job.key_offsets = None if koff is None else koff.ctypes.data
    job.key_lengths = None if klens is None else klens.ctypes.data
    job.key_stride, job.key_length = kstride, klen
    hold = [kflat, koff, klens]
    if quals is not None:
        qflat, qoff, qlens, qstride, qlen, qn = _as_rows(quals, qual_lengths if qual_lengths is not None else lengths)
        if qn != n:
            raise ValueError("keys and quals describe a different number of records")
        job.quals = qflat.ctypes.data
        job.qual_offsets = None if qoff is None else qoff.ctypes.data
        job.qual_lengths = None if qlens is None else qlens.ctypes.data
        job.qual_stride, job.qual_length = qstride, qlen
        hold += [qflat, qoff, qlens]
    job.max_distance = int(max_distance)
    job.use_edit_distance = int(bool(use_edit_distance))
    job.method = METHODS[method]
    job.memory_space = MEM_HOST
    job.max_average_error_rate = float(max_average_error_rate)
    job.phred_offset = int(phred_offset)
    job.alphabet = None if alphabet is None else alphabet.encode("latin-1")
    if counts is not None:
        cnt = np.ascontiguousarray(counts, dtype=np.uint32)
        job.record_counts = cnt.ctypes.data
        hold.append(cnt)
    bitmap = np.zeros((n + 31) // 32, dtype=np.uint32) if want_bitmap else None
    stats = ctx.cluster(job, bitmap.ctypes.data if (want_bitmap and n) else None)
    res = ClusterResult(stats.total_records, stats.discarded_records, stats.number_of_sequences,
                        stats.number_of_uniques, stats.number_of_clusters, stats.number_selected,
                        stats.as_dict(), bitmap)
    if want_uniques:
        first, count, label, sel = ctx.fetch(stats.number_of_uniques)
        order = np.argsort(first, kind="stable")
        res.first, res.count = first[order], count[order]
        res.label, res.selected = label[order], sel[order].astype(bool)
    del hold
    return res


def cluster_device(ctx, n_records, keys_ptr, key_length, key_stride=None, quals_ptr=None,
                   qual_length=0, qual_stride=None, max_distance=1, use_edit_distance=False,
                   method="directional", max_average_error_rate=1.0, phred_offset=33,
                   bitmap_ptr=None):
    """The same job with fixed-stride inputs already resident in HBM (raw device
    pointers).  Returns the library's stats structure; results stay on the device."""
    job = ClusterJob()
    job.n_records = n_records
    job.keys = keys_ptr
    job.key_stride = key_stride or key_length
    job.key_length = key_length
    if quals_ptr:
        job.quals = quals_ptr
        job.qual_stride = qual_stride or qual_length
        job.qual_length = qual_length
    job.max_distance = int(max_distance)
    job.use_edit_distance = int(bool(use_edit_distance))
    job.method = METHODS[method]
    job.memory_space = MEM_DEVICE
    job.max_average_error_rate = float(max_average_error_rate)
    job.phred_offset = int(phred_offset)
    return ctx.cluster(job, bitmap_ptr)
