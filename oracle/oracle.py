"""ctypes front-end of the CPU oracle (``oracle/fqd_oracle.c``) and a driver that runs
the same job through the unmodified reference (``oracle/_ref``).

TEST INFRASTRUCTURE ONLY -- see the header of ``fqd_oracle.c``.  Importable from
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference/cpu_baseline arm,
never from ``fastqdedup_b200``.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libfqd_oracle.so")

METHODS = {"highest_count": 0, "adjacency": 1, "directional": 2}


class OracleStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in (
        "total_records", "discarded_records", "number_of_sequences", "number_of_uniques",
        "number_of_clusters", "number_selected", "bad_record", "bad_char")]


def build():
    """(Re)build the C oracle if needed; building the checker is not using it."""
    src = os.path.join(HERE, "fqd_oracle.c")
    if (not os.path.exists(LIB_PATH)
            or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)):
        subprocess.run(["make", "-C", HERE, "_build/libfqd_oracle.so", f"PY={sys.executable}"],
                       check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB_PATH)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        _lib.fqd_oracle_average_error_rate.argtypes = [
            ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint8,
            ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_size_t)]
        _lib.fqd_oracle_within_hamming.argtypes = [
            ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
        _lib.fqd_oracle_within_edit.argtypes = _lib.fqd_oracle_within_hamming.argtypes
        _lib.fqd_oracle_cluster.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_double, ctypes.c_uint8, ctypes.POINTER(OracleStats),
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        del u8p
    return _lib


class PhredError(ValueError):
    pass


def average_error_rate(phred: bytes, phred_offset: int = 33) -> float:
    out = ctypes.c_double()
    bad = ctypes.c_size_t()
    rc = lib().fqd_oracle_average_error_rate(phred, len(phred), phred_offset,
                                             ctypes.byref(out), ctypes.byref(bad))
    if rc:
        raise PhredError(f"Character {chr(phred[bad.value])} outside of valid phred range "
                         f"('{chr(phred_offset)}' to '~')")
    return out.value


def within_distance(a: bytes, b: bytes, max_distance: int, use_edit_distance=False) -> bool:
    f = lib().fqd_oracle_within_edit if use_edit_distance else lib().fqd_oracle_within_hamming
    return bool(f(a, len(a), b, len(b), max_distance))


def flatten(strings):
    """list[bytes] -> (uint8 flat array, uint64 offsets[n+1])."""
    lens = np.fromiter((len(s) for s in strings), dtype=np.uint64, count=len(strings))
    off = np.zeros(len(strings) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    flat = np.frombuffer(b"".join(strings), dtype=np.uint8).copy() if len(strings) else \
        np.zeros(0, dtype=np.uint8)
    return flat, off


def as_flat(keys):
    """Accept list[bytes], a 2-D uint8 array (fixed length) or (flat, offsets)."""
    if isinstance(keys, tuple):
        return np.ascontiguousarray(keys[0], dtype=np.uint8), \
            np.ascontiguousarray(keys[1], dtype=np.uint64)
    if isinstance(keys, np.ndarray) and keys.ndim == 2:
        n, length = keys.shape
        return np.ascontiguousarray(keys, dtype=np.uint8).reshape(-1), \
            (np.arange(n + 1, dtype=np.uint64) * np.uint64(length))
    return flatten(list(keys))


def cluster(keys, quals=None, max_distance=1, use_edit_distance=False, method="directional",
            max_average_error_rate=1.0, phred_offset=33):
    """Run the whole job (filter -> exact dedupe -> components -> dissection) on the CPU
    oracle.  Returns a dict with the stats and per-unique arrays ordered by ``first``."""
    kflat, koff = as_flat(keys)
    n = len(koff) - 1
    if quals is not None:
        qflat, qoff = as_flat(quals)
        assert len(qoff) - 1 == n
        qp, qo = qflat.ctypes.data, qoff.ctypes.data
    else:
        qflat = qoff = None
        qp = qo = None
    stats = OracleStats()
    u_first = np.zeros(max(n, 1), dtype=np.uint64)
    u_count = np.zeros(max(n, 1), dtype=np.uint32)
    u_label = np.zeros(max(n, 1), dtype=np.uint64)
    u_sel = np.zeros(max(n, 1), dtype=np.uint8)
    rc = lib().fqd_oracle_cluster(
        kflat.ctypes.data, koff.ctypes.data, qp, qo, n, int(max_distance),
        int(bool(use_edit_distance)), METHODS[method], float(max_average_error_rate),
        int(phred_offset), ctypes.byref(stats),
        u_first.ctypes.data, u_count.ctypes.data, u_label.ctypes.data, u_sel.ctypes.data)
    if rc == 1:
        raise PhredError(f"Character {chr(stats.bad_char)} outside of valid phred range "
                         f"('{chr(phred_offset)}' to '~') [record {stats.bad_record}]")
    if rc:
        raise RuntimeError(f"fqd_oracle_cluster failed with code {rc}")
    nu = stats.number_of_uniques
    res = {f: getattr(stats, f) for f, _ in OracleStats._fields_[:6]}
    res.update(first=u_first[:nu].copy(), count=u_count[:nu].copy(),
               label=u_label[:nu].copy(), selected=u_sel[:nu].astype(bool))
    res["selected_first"] = res["first"][res["selected"]]
    return res


def ref_cluster(keys, quals=None, max_distance=1, use_edit_distance=False,
                method="directional", max_average_error_rate=1.0, phred_offset=33):
    """The same job through the UNMODIFIED reference (oracle/_ref): exactly the loops of
    ``deduplicate_cluster`` (``__init__.py:240-276``) without file I/O.  Single thread."""
    sys.path.insert(0, HERE)
    try:
        import ref_loader
    finally:
        sys.path.pop(0)
    ref = ref_loader.load_reference()
    if ref is None:
        raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.py where "
                           "/root/reference exists)")
    assert phred_offset == 33
    kflat, koff = as_flat(keys)
    n = len(koff) - 1
    kbytes = kflat.tobytes()
    koff_l = koff.tolist()
    key_strs = [kbytes[koff_l[i]:koff_l[i + 1]].decode("latin-1") for i in range(n)]
    if quals is not None:
        qflat, qoff = as_flat(quals)
        qbytes = qflat.tobytes()
        qoff_l = qoff.tolist()
        qual_strs = [qbytes[qoff_l[i]:qoff_l[i + 1]].decode("latin-1") for i in range(n)]
    filter_on = max_average_error_rate < 1.0 and quals is not None
    func = ref.CLUSTER_DISSECTION_METHODS[method]
    trie = ref.Trie(alphabet="ACGTN")
    discarded = 0
    first = {}
    for t, key in enumerate(key_strs):
        first.setdefault(key, t)          # pass 2 sees every record (__init__.py:201-206)
        if filter_on and ref.fastq_average_error_rate(qual_strs[t]) > max_average_error_rate:
            discarded += 1
            continue
        trie.add_sequence(key)
    nseq = trie.number_of_sequences
    selected = set()
    nclusters = 0
    label = {}
    counts = {}
    while trie.number_of_sequences:
        cl = trie.pop_cluster(max_distance, use_edit_distance)
        nclusters += 1
        root = min(first[k] for _, k in cl)
        for c, k in cl:
            label[k] = root
            counts[k] = c
        for k in func(cl, max_distance, use_edit_distance):
            selected.add(k)
    ukeys = sorted(counts, key=lambda k: first[k])
    res = dict(total_records=n, discarded_records=discarded, number_of_sequences=nseq,
               number_of_uniques=len(ukeys), number_of_clusters=nclusters,
               number_selected=len(selected))
    res["first"] = np.array([first[k] for k in ukeys], dtype=np.uint64)
    res["count"] = np.array([counts[k] for k in ukeys], dtype=np.uint32)
    res["label"] = np.array([label[k] for k in ukeys], dtype=np.uint64)
    res["selected"] = np.array([k in selected for k in ukeys], dtype=bool)
    res["selected_first"] = res["first"][res["selected"]] if len(ukeys) else \
        np.zeros(0, dtype=np.uint64)
    return res
