"""Sharding the clustering job over the GPUs of one box (BASELINE.json config 5).

Records are split contiguously over the ranks (so a global record index is the shard's base
plus the local index).  Two ways to drive the same native plan (``run_sharded`` in
``csrc/pipeline.cu``):

* :func:`cluster_keys_sharded_local` -- every rank lives in THIS process (one context each,
  possibly on the same GPU); exchanges are device copies.  This is how a single-GPU box tests
  the sharded algorithm, and a way to use several GPUs from one process.
* :class:`ShardComm` -- one process per GPU (``torchrun``): the exchanges are NCCL
  collectives over NVLink inside the library; ``torch.distributed`` is only used to hand the
  NCCL unique id from rank 0 to the others.

Host code here only splits arrays and stitches bitmaps.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_uint64, c_void_p

import numpy as np

from . import _native
from ._native import METHODS, MEM_DEVICE, MEM_HOST, ClusterJob, ClusterStats

PLAN_SHARD_TILES = 16   # include/fqd_b200.h FQD_PLAN_SHARD_TILES
from .clustering import ClusterResult, _as_rows


def shard_bounds(n_records: int, world: int):
    """Contiguous split: rank r gets records [b[r], b[r+1])."""
    base, rem = divmod(n_records, world)
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def unpack_bitmap(words: np.ndarray, n: int) -> np.ndarray:
    return np.unpackbits(words.view(np.uint8), bitorder="little")[:n].astype(bool)


def _slice_rows(flat, off, lens, stride, lo, hi):
    """Rows [lo, hi) of a (flat, offsets | lengths, stride) description, as host arrays."""
    if off is not None:
        start, stop = int(off[lo]), int(off[hi])
        sub = flat[start:stop] if stop > start else np.zeros(1, dtype=np.uint8)
        return np.ascontiguousarray(sub), np.ascontiguousarray(off[lo:hi + 1] - off[lo]), None
    sub = flat[lo * stride:hi * stride] if hi > lo and stride else np.zeros(1, dtype=np.uint8)
    sub_l = None if lens is None else np.ascontiguousarray(lens[lo:hi])
    return np.ascontiguousarray(sub), None, sub_l


def _fill_job(job, keys_part, quals_part, kstride, klen, qstride, qlen, n, max_distance,
              use_edit_distance, method, max_average_error_rate, phred_offset, counts, alphabet):
    kflat, koff, klens = keys_part
    job.n_records = n
    job.keys = kflat.ctypes.data
    job.key_offsets = None if koff is None else koff.ctypes.data
    job.key_lengths = None if klens is None else klens.ctypes.data
    job.key_stride, job.key_length = kstride, klen
    if quals_part is not None:
        qflat, qoff, qlens = quals_part
        job.quals = qflat.ctypes.data
        job.qual_offsets = None if qoff is None else qoff.ctypes.data
        job.qual_lengths = None if qlens is None else qlens.ctypes.data
        job.qual_stride, job.qual_length = qstride, qlen
    job.max_distance = int(max_distance)
    job.use_edit_distance = int(bool(use_edit_distance))
    job.method = METHODS[method]
    job.memory_space = MEM_HOST
    job.max_average_error_rate = float(max_average_error_rate)
    job.phred_offset = int(phred_offset)
    job.alphabet = alphabet
    if counts is not None:
        job.record_counts = counts.ctypes.data


def cluster_keys_sharded_local(keys, quals=None, max_distance=1, use_edit_distance=False,
                               method="directional", max_average_error_rate=1.0, phred_offset=33,
                               lengths=None, counts=None, alphabet=None, world=2, contexts=None,
                               devices=None, want_uniques=True) -> ClusterResult:
    """Same contract as ``cluster_keys`` but executed as a `world`-rank sharded job driven by
    this process.  ``devices`` maps ranks to GPU ordinals (default: all on GPU 0)."""
    lib = _native.load()
    own = contexts is None
    if own:
        devices = devices or [0] * world
        contexts = [_native.Context(d) for d in devices]
    try:
        kflat, koff, klens, kstride, klen, n = _as_rows(keys, lengths)
        q = None
        if quals is not None:
            q = _as_rows(quals, lengths)
            if q[5] != n:
                raise ValueError("keys and quals describe a different number of records")
        bounds = shard_bounds(n, world)
        jobs = (ClusterJob * world)()
        stats = (ClusterStats * world)()
        bases = (c_uint64 * world)(*bounds[:world])
        handles = (c_void_p * world)(*[c.handle for c in contexts])
        bitmaps, hold = [], []
        bm_ptrs = (c_void_p * world)()
        alpha = None if alphabet is None else alphabet.encode("latin-1")
        cnts = None if counts is None else np.ascontiguousarray(counts, dtype=np.uint32)
        for r in range(world):
            lo, hi = bounds[r], bounds[r + 1]
            kp = _slice_rows(kflat, koff, klens, kstride, lo, hi)
            qp = None if q is None else _slice_rows(q[0], q[1], q[2], q[3], lo, hi)
            cp = None if cnts is None else np.ascontiguousarray(cnts[lo:hi])
            _fill_job(jobs[r], kp, qp, kstride, klen, 0 if q is None else q[3], 0 if q is None else q[4],
                      hi - lo, max_distance, use_edit_distance, method, max_average_error_rate,
                      phred_offset, cp, alpha)
            bm = np.zeros(max((hi - lo + 31) // 32, 1), dtype=np.uint32)
            bitmaps.append(bm)
            bm_ptrs[r] = bm.ctypes.data
            hold += [kp, qp, cp]
        rc = lib.fqd_cluster_sharded_local(handles, world, jobs, bases, stats, bm_ptrs)
        _native.check(rc, stats[0])
        st = stats[0]
        keep = np.concatenate([unpack_bitmap(bitmaps[r], bounds[r + 1] - bounds[r]) for r in range(world)]) \
            if n else np.zeros(0, dtype=bool)
        words = np.packbits(keep, bitorder="little")
        words = np.concatenate([words, np.zeros((-len(words)) % 4, dtype=np.uint8)]).view(np.uint32)
        res = ClusterResult(st.total_records, st.discarded_records, st.number_of_sequences,
                            st.number_of_uniques, st.number_of_clusters, st.number_selected,
                            st.as_dict(), words)
        res.per_rank_stats = [stats[r].as_dict() for r in range(world)]
        if want_uniques and st.plan_flags & PLAN_SHARD_TILES:
            # tile-sharded plan: every rank returns its own keys; `label` is the cluster's root in the job-wide id
            # space -> smallest `first` of the cluster
            parts = [c.fetch(stats[r].own_uniques) for r, c in enumerate(contexts)]
            first = np.concatenate([p[0] for p in parts])
            count = np.concatenate([p[1] for p in parts])
            roots = np.concatenate([p[2] for p in parts])
            sel = np.concatenate([p[3] for p in parts])
            assert len(first) == st.number_of_uniques
            if len(first):
                uroots, inv = np.unique(roots, return_inverse=True)
                minfirst = np.full(len(uroots), np.iinfo(np.uint64).max, dtype=np.uint64)
                np.minimum.at(minfirst, inv, first)
                label = minfirst[inv]
            else:
                label = roots
            order = np.argsort(first, kind="stable")
            res.first, res.count = first[order], count[order]
            res.label, res.selected = label[order], sel[order].astype(bool)
        elif want_uniques:
            first, count, label, sel = contexts[0].fetch(st.number_of_uniques)
            for c in contexts[1:]:       # a rank decides only the keys whose first record is its own
                sel |= c.fetch(st.number_of_uniques)[3]
            order = np.argsort(first, kind="stable")
            res.first, res.count = first[order], count[order]
            res.label, res.selected = label[order], sel[order].astype(bool)
        del hold
        return res
    finally:
        if own:
            for c in contexts:
                c.close()


class ShardComm:
    """One rank of a multi-process sharded job (one process per GPU)."""

    def __init__(self, ctx, rank, world, unique_id: bytes):
        self.ctx, self.rank, self.world = ctx, rank, world
        # the ranks of a box share its cores: the host-side key packer of a rank gets its share of them
        # (read once, when the library first packs; an explicit FQD_PACK_THREADS wins)
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world) or world)
        os.environ.setdefault("FQD_PACK_THREADS", str(max(1, (os.cpu_count() or 4) // max(1, local_world))))
        self.lib = _native.load()
        h = c_void_p()
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        _native.check(self.lib.fqd_comm_create(ctx.handle, rank, world, buf, byref(h)))
        self.handle = h

    @staticmethod
    def new_unique_id() -> bytes:
        lib = _native.load()
        buf = (ctypes.c_uint8 * 128)()
        _native.check(lib.fqd_nccl_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch_distributed(cls, ctx, dist):
        """`dist` is an initialised torch.distributed: rank 0's NCCL id reaches the others
        through one broadcast_object_list (plumbing only)."""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.new_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(ctx, rank, world, box[0])

    def close(self):
        if self.handle:
            self.lib.fqd_comm_destroy(self.handle)
            self.handle = None

    def cluster_device(self, n_local, index_base, keys_ptr, key_length, quals_ptr=None,
                       max_distance=1, use_edit_distance=False, method="directional",
                       max_average_error_rate=1.0, bitmap_ptr=None):
        """This rank's fixed-stride records are already in HBM."""
        job = ClusterJob()
        job.n_records = n_local
        job.keys = keys_ptr
        job.key_stride = job.key_length = key_length
        if quals_ptr:
            job.quals = quals_ptr
            job.qual_stride = job.qual_length = key_length
        job.max_distance = int(max_distance)
        job.use_edit_distance = int(bool(use_edit_distance))
        job.method = METHODS[method]
        job.memory_space = MEM_DEVICE
        job.max_average_error_rate = float(max_average_error_rate)
        job.phred_offset = 33
        return self._run(job, index_base, bitmap_ptr)

    def cluster_host(self, keys2d, index_base, quals2d=None, max_distance=1, use_edit_distance=False,
                     method="directional", max_average_error_rate=1.0, bitmap=None):
        """This rank's records as host arrays (H2D inside the call)."""
        job = ClusterJob()
        keys2d = np.ascontiguousarray(keys2d, dtype=np.uint8)
        n, L = keys2d.shape
        job.n_records = n
        job.keys = keys2d.ctypes.data
        job.key_stride = job.key_length = L
        if quals2d is not None:
            quals2d = np.ascontiguousarray(quals2d, dtype=np.uint8)
            job.quals = quals2d.ctypes.data
            job.qual_stride = job.qual_length = L
        job.max_distance = int(max_distance)
        job.use_edit_distance = int(bool(use_edit_distance))
        job.method = METHODS[method]
        job.memory_space = MEM_HOST
        job.max_average_error_rate = float(max_average_error_rate)
        job.phred_offset = 33
        return self._run(job, index_base, None if bitmap is None else bitmap.ctypes.data)

    def _run(self, job, index_base, bitmap_ptr):
        stats = ClusterStats()
        rc = self.lib.fqd_cluster_sharded(self.ctx.handle, self.handle, byref(job),
                                          int(index_base), byref(stats),
                                          c_void_p(bitmap_ptr) if bitmap_ptr else None)
        _native.check(rc, stats)
        return stats
