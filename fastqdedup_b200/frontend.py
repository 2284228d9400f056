"""The reference's Python surface (``src/fastqdedup/__init__.py``) on top of the GPU path.

Same names, arguments, defaults, log lines and error behaviour as the reference module, so
callers (and the reference's tests) can switch packages unchanged:

* ``deduplicate_cluster`` keeps its signature (reference ``:209-217``); its two inner
  loops (``:242-252`` filter + ``Trie.add_sequence`` per record, ``:272-276``
  ``pop_cluster`` + dissection per cluster) are replaced by ONE batched GPU job
  (``clustering.cluster_keys``), and pass 2 (``:189-206``) emits the records whose ordinal
  is set in the job's keep bitmap instead of re-hashing every key.
* the three ``cluster_dissection_*`` callables (``:60-122``) still accept a
  ``[(count, sequence), ...]`` list and yield the selected sequences, evaluated on the GPU.

Host code here is I/O and marshalling only.
"""
import argparse
import contextlib
import datetime
import io
import logging
import resource
import time
from typing import Callable, Dict, Iterable, Iterator, List, Optional, Tuple

import numpy as np

from . import _native, fastq_io
from ._trie import Trie
from .clustering import cluster_keys

DEFAULT_PREFIX = "fastqdedup_R"
DEFAULT_MAX_DISTANCE = 1
DEFAULT_CLUSTER_DISSECTION = "directional"
DEFAULT_MAX_AVERAGE_ERROR_RATE = 0.001


class Timer:
    def __init__(self):
        self.start_time = time.time()

    def get_difference(self) -> datetime.timedelta:
        now = time.time()
        delta = datetime.timedelta(seconds=round(now - self.start_time))
        self.start_time = now
        return delta


# ---------------------------------------------------------------------------------------
# cluster dissection callables
# ---------------------------------------------------------------------------------------

def _dissect(cluster, max_distance, use_edit_distance, method) -> List[str]:
    """Run one (count, sequence) list through the GPU job with pre-counted records and
    return the selected sequences in descending (count, sequence) order -- the order the
    reference yields them in."""
    if not cluster:
        return []
    seqs = [s.encode("latin-1") for _, s in cluster]
    counts = np.fromiter((c for c, _ in cluster), dtype=np.uint32, count=len(cluster))
    res = cluster_keys(seqs, None, max_distance, use_edit_distance, method, 1.0,
                       counts=counts, want_bitmap=True, want_uniques=False)
    keep = res.keep_mask()
    # duplicates in the list were merged by the job; map back through the first occurrence
    merged: Dict[str, int] = {}
    for c, s in cluster:
        merged[s] = merged.get(s, 0) + c
    chosen = {cluster[i][1] for i in np.nonzero(keep)[0]}
    return sorted(chosen, key=lambda s: (merged[s], s), reverse=True)


def cluster_dissection_directional(cluster: List[Tuple[int, str]],
                                   max_distance: int = DEFAULT_MAX_DISTANCE,
                                   use_edit_distance: bool = False) -> Iterator[str]:
    """Reference ``:60-91``: greedy, count-aware (an edge a->b needs count_a >= 2*count_b-1)."""
    yield from _dissect(cluster, max_distance, use_edit_distance, "directional")


def cluster_dissection_highest_count(cluster: List[Tuple[int, str]],
                                     max_distance: int = DEFAULT_MAX_DISTANCE,
                                     use_edit_distance: bool = False) -> Iterator[str]:
    """Reference ``:94-102``: the single largest (count, sequence) of the list."""
    winners = _dissect(cluster, max_distance, use_edit_distance, "highest_count")
    # the GPU job answers per connected component; a caller-made list spanning several
    # components still gets exactly one answer, the first in descending order
    yield from winners[:1]


def cluster_dissection_adjacency(cluster: List[Tuple[int, str]],
                                 max_distance: int = DEFAULT_MAX_DISTANCE,
                                 use_edit_distance: bool = False) -> Iterator[str]:
    """Reference ``:105-122``: take the largest, drop its neighbours, repeat."""
    yield from _dissect(cluster, max_distance, use_edit_distance, "adjacency")


ClusterDissectionFunc = Callable[[List[Tuple[int, str]], int, bool], Iterator[str]]
CLUSTER_DISSECTION_METHODS: Dict[str, ClusterDissectionFunc] = {
    "highest_count": cluster_dissection_highest_count,
    "adjacency": cluster_dissection_adjacency,
    "directional": cluster_dissection_directional,
}
_METHOD_OF_FUNC = {f: name for name, f in CLUSTER_DISSECTION_METHODS.items()}


def trie_stats(trie: Trie) -> str:
    """Layer table of ``Trie.raw_stats()`` in the reference's layout (``:133-157``)."""
    out = io.StringIO()
    raw = trie.raw_stats()
    width = len(trie.alphabet) + 1
    totals = [0] * (width + 1)
    out.write("layer     terminal  " + "".join(f"{i:10}" for i in range(1, width)) + "     total\n")
    for layer, row in enumerate(raw):
        row_total = sum(row)
        for j in range(width):
            totals[j] += row[j]
        totals[width] += row_total
        out.write("".join(f"{v:10}" for v in [str(layer)] + row + [row_total]) + "\n")
    out.write("".join(f"{v:10}" for v in ["total"] + totals) + "\n")
    node_bytes = sum((8 + 8 * i) * totals[i] for i in range(width))
    total_bytes = trie.memory_size()
    gib = 1024 ** 3
    out.write(f"Node memory usage: {node_bytes / gib:.2} GiB\n"
              f"Suffix memory usage: {(total_bytes - node_bytes) / gib:.2} GiB\n"
              f"Total memory usage: {total_bytes / gib:.2} GiB\n")
    return out.getvalue()


# ---------------------------------------------------------------------------------------
# FASTQ helpers (reference :160-206)
# ---------------------------------------------------------------------------------------

def joinfunc_from_check_slices(check_slices: Iterable[slice]) -> Callable[[Iterable[str]], str]:
    slices = list(check_slices)

    def joinfunc(strings: Iterable[str]) -> str:
        return "".join(s[slc] for s, slc in zip(strings, slices))
    return joinfunc


def fastq_files_to_records(input_files: List[str]):
    readers = [fastq_io.read_fastq(f) for f in input_files]
    for records in zip(*readers):
        if len(records) > 1 and not fastq_io.records_are_mates(*records):
            raise fastq_io.FastqFormatError(
                f"FASTQ files not in sync: "
                f"{', '.join(record.name for record in records)} are not mates.",
                line=None)
        yield records


def filter_fastq_files_on_bitmap(input_files: List[str], output_files: List[str],
                                 keep: np.ndarray):
    """Pass 2: emit record tuple t iff keep[t] (first occurrence of a selected key)."""
    readers = [fastq_io.read_fastq(f) for f in input_files]
    with contextlib.ExitStack() as stack:
        writers = [stack.enter_context(fastq_io.open_write(x)) for x in output_files]
        n = len(keep)
        for t, records in enumerate(zip(*readers)):
            if t < n and keep[t]:
                for out, record in zip(writers, records):
                    out.write(record.fastq_bytes())


def structure_stats(result) -> str:
    """The ``-v`` report of the structure that replaces the trie (the reference prints ``trie_stats`` here,
    ``:260-264``): the packed unique-key table, the shared-memory tiles and the pigeonhole passes of the job."""
    st = result.stats
    n, u = result.total_records, result.number_of_uniques
    key_bytes = 4 * st["key_words"]
    flags = st["plan_flags"]
    lines = [
        f"unique keys       {u:>12}  ({st['key_bits']} bit planes x {st['key_words'] // max(st['key_bits'], 1)} words"
        f" = {key_bytes} B packed key + 4 B count + 4 B first index = {(key_bytes + 8) * u / 1024 ** 3:.2} GiB)",
        f"records           {n:>12}  ({result.discarded_records} discarded by the quality filter)",
        f"exact dedupe      {'shared-memory tiles of 512 records, ~60 % full' if flags & 1 else 'one open-addressing table in HBM'}",
        f"pigeonhole passes {st['n_passes']:>12}  "
        f"({'tiles' if flags & 2 else 'counting sort by block hash'}{', pass 0 inside the dedupe tiles' if flags & 4 else ''})",
        f"candidate pairs   {st['candidate_pairs']:>12}  ({st['candidate_pairs'] / max(u, 1):.2f} per unique key, all verified)",
        f"clusters          {result.number_of_clusters:>12}",
        f"kernel launches   {st['launches']:>12}  ({st['ms_total']:.2f} ms on the device)",
    ]
    return "\n".join(lines) + "\n"


def deduplicate_cluster(
    input_files: List[str],
    output_files: List[str],
    check_slices: Optional[List[slice]],
    max_distance: int = DEFAULT_MAX_DISTANCE,
    max_average_error_rate: float = DEFAULT_MAX_AVERAGE_ERROR_RATE,
    cluster_dissection_func: ClusterDissectionFunc = cluster_dissection_directional,
    use_edit_distance: bool = False,
):
    if len(input_files) != len(output_files):
        raise ValueError(f"Amount of output files ({len(output_files)}) "
                         f"must be equal to the amount of input files "
                         f"({len(input_files)}). ")
    if check_slices and len(input_files) != len(check_slices):
        raise ValueError(f"Amount of check lengths ({len(check_slices)}) "
                         f"must be equal to the amount of input files "
                         f"({len(input_files)}). ")
    filter_on_quality = max_average_error_rate < 1.0
    timer = Timer()
    logger = logging.getLogger("fastqdedup")

    # pass 1 (host, native: csrc/fastq_native.cpp): reader + parser threads per file, mates check, the key and
    # quality slices of every record tuple written straight into the buffers of the GPU job
    try:
        scan = _native.FastqScan(input_files, check_slices or None, want_quals=filter_on_quality)
    except _native.FqdFastqError as e:
        raise fastq_io.FastqFormatError(str(e), line=None) from None
    with scan:
        keys, quals = scan.keys, scan.quals
        n_records = scan.n_records
        method = _METHOD_OF_FUNC.get(cluster_dissection_func)
        if method is not None:
            # the batched GPU job: filter + exact dedupe + neighbour search + dissection
            result = cluster_keys(keys, quals, max_distance, use_edit_distance, method,
                                  max_average_error_rate, want_uniques=False)
            if filter_on_quality:
                logger.info(
                    f"{result.discarded_records} records out of {result.total_records} "
                    f"records had an error rate higher than {max_average_error_rate} "
                    f"and were discarded.")
            logger.info(f"Processed {result.number_of_sequences} sequences. "
                        f"({timer.get_difference()})")
            if logger.level <= logging.DEBUG:
                logger.debug("\n" + structure_stats(result))
            logger.info(f"Found {result.number_selected} distinct reads "
                        f"in {result.number_of_clusters} clusters."
                        f"({timer.get_difference()})")
            bitmap = result.keep_bitmap
        else:
            keep = _deduplicate_with_callable(_as_ragged(keys), None if quals is None else _as_ragged(quals),
                                              max_distance, max_average_error_rate,
                                              cluster_dissection_func, use_edit_distance,
                                              filter_on_quality, logger, timer)
            words = np.packbits(keep, bitorder="little")
            bitmap = np.concatenate([words, np.zeros((-len(words)) % 4, dtype=np.uint8)]).view(np.uint32)
    # pass 2 (host, native): the record tuples whose bit is set; .gz outputs compressed by the worker threads
    _native.fastq_emit(input_files, output_files, bitmap, n_records)
    logger.info(f"Filtered FASTQ files based on distinct reads from each cluster. "
                f"({timer.get_difference()}) ")


def _as_ragged(rows):
    """(flat, offsets) view of the scan's rows (2-D when every row has the same length)."""
    if isinstance(rows, tuple):
        return rows
    n, width = rows.shape
    return rows.reshape(-1), np.arange(n + 1, dtype=np.uint64) * np.uint64(width)


def _deduplicate_with_callable(keys, quals, max_distance, max_average_error_rate, func,
                               use_edit_distance, filter_on_quality, logger, timer):
    """A user-supplied dissection callable cannot run inside the batched job; it gets the
    reference's own loop (``:240-276``) over the Trie shim, whose neighbour search is still
    the GPU's."""
    from ._fastq import average_error_rate
    flat, off = keys
    data = flat.tobytes()
    qdata = quals[0].tobytes() if filter_on_quality else b""
    n = len(off) - 1
    trie = Trie(alphabet="ACGTN")
    first: Dict[str, int] = {}
    discarded = 0
    for t in range(n):
        key = data[off[t]:off[t + 1]].decode("latin-1")
        first.setdefault(key, t)
        if filter_on_quality:
            q = qdata[quals[1][t]:quals[1][t + 1]].decode("latin-1")
            if average_error_rate(q) > max_average_error_rate:
                discarded += 1
                continue
        trie.add_sequence(key)
    if filter_on_quality:
        logger.info(f"{discarded} records out of {n} records had an error rate higher than "
                    f"{max_average_error_rate} and were discarded.")
    logger.info(f"Processed {trie.number_of_sequences} sequences. ({timer.get_difference()})")
    keep = np.zeros(n, dtype=bool)
    selected = 0
    clusters = 0
    while trie.number_of_sequences:
        cluster = trie.pop_cluster(max_distance, use_edit_distance)
        clusters += 1
        for key in func(cluster, max_distance, use_edit_distance):
            if not keep[first[key]]:
                keep[first[key]] = True
                selected += 1
    logger.info(f"Found {selected} distinct reads in {clusters} clusters."
                f"({timer.get_difference()})")
    return keep


# ---------------------------------------------------------------------------------------
# command line (reference :291-412) -- same flags, same log lines
# ---------------------------------------------------------------------------------------

def initiate_logger(verbose: int = 0, quiet: int = 0):
    level = logging.INFO - 10 * (verbose - quiet)
    logger = logging.getLogger("fastqdedup")
    logger.setLevel(level)
    handler = logging.StreamHandler()
    handler.setLevel(level)
    handler.setFormatter(logging.Formatter("{asctime}:{levelname}:{name}: {message}",
                                           datefmt="%m/%d/%Y %I:%M:%S", style="{"))
    logger.addHandler(handler)


def argument_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument("fastq", metavar="FASTQ", nargs="+",
                   help="Forward FASTQ and optional reverse and UMI FASTQ files.")
    p.add_argument("-l", "--check-lengths",
                   help="Comma-separated maximum check length (or slice, e.g. '4:8', '::8') "
                        "per input file; only those bases take part in the duplicate test.")
    p.add_argument("-o", "--output", action="append", required=False,
                   help="Output file; repeat once per input file.")
    p.add_argument("-p", "--prefix", default=DEFAULT_PREFIX,
                   help=f"Prefix for the output files. Default: '{DEFAULT_PREFIX}'")
    p.add_argument("-d", "--max-distance", type=int, default=DEFAULT_MAX_DISTANCE,
                   help=f"Distance at which inputs are considered different. "
                        f"Default: {DEFAULT_MAX_DISTANCE}.")
    p.add_argument("-e", "--max-average-error-rate", type=float,
                   default=DEFAULT_MAX_AVERAGE_ERROR_RATE,
                   help=f"Maximum average per base error rate of a record over the checked "
                        f"bases. Default: {DEFAULT_MAX_AVERAGE_ERROR_RATE}")
    p.add_argument("-E", "--no-average-error-rate-filter", action="store_const",
                   dest="max_average_error_rate", const=1.0,
                   help="Do not filter on average per base error rate.")
    p.add_argument("--edit", action="store_true",
                   help="Use edit (Levenshtein) distance instead of Hamming distance.")
    p.add_argument("-c", "--cluster-dissection-method",
                   choices=CLUSTER_DISSECTION_METHODS.keys(), default=DEFAULT_CLUSTER_DISSECTION,
                   help="highest_count, adjacency or directional (default).")
    p.add_argument("-v", "--verbose", action="count", default=0, help="Increase log verbosity.")
    p.add_argument("-q", "--quiet", action="count", default=0, help="Reduce log verbosity.")
    return p


def length_string_to_slices(length_string: str) -> List[slice]:
    """'8,8,8' or '8:16,8,24:8:-1' -> list of slice objects (reference :364-375)."""
    slices = []
    for part in length_string.split(","):
        fields = [None if x in ("", "None") else int(x) for x in part.split(":")]
        slices.append(slice(*fields))
    return slices


def main():
    args = argument_parser().parse_args()
    initiate_logger(args.verbose, args.quiet)
    logger = logging.getLogger("fastqdedup")
    input_files: List[str] = args.fastq
    check_slices = length_string_to_slices(args.check_lengths) if args.check_lengths else None
    output_files = args.output or [args.prefix + str(i) + ".fastq.gz"
                                   for i in range(1, len(input_files) + 1)]
    distance_name = "Levenshtein" if args.edit else "Hamming"
    timer = Timer()
    logger.info(f"Input files: {', '.join(input_files)}")
    logger.info(f"Output files: {', '.join(output_files)}")
    logger.info(f"Check lengths: {args.check_lengths}")
    logger.info(f"Maximum {distance_name} distance: {args.max_distance}")
    logger.info(f"Maximum average error rate: {args.max_average_error_rate}")
    logger.info(f"Cluster dissection method: {args.cluster_dissection_method}")
    deduplicate_cluster(input_files, output_files, check_slices, args.max_distance,
                        args.max_average_error_rate,
                        CLUSTER_DISSECTION_METHODS[args.cluster_dissection_method], args.edit)
    usage = resource.getrusage(resource.RUSAGE_SELF)
    logger.info(f"Finished. Total time: {timer.get_difference()}. "
                f"Memory usage: {usage.ru_maxrss / (1024 ** 2):.2} GiB")
