"""Host-side FASTQ reading/writing for the drop-in ``deduplicate_cluster``.

FASTQ I/O stays on the host (BASELINE.json north_star); the reference delegates it to the
third-party packages dnaio and xopen (``src/fastqdedup/__init__.py:54-57, 181-185,
197-200``).  When those are installed they are used unchanged; otherwise the small
equivalents below cover exactly the surface the reference touches, so the package also
works on machines without them (this build image has neither).
"""
import gzip
import io
from typing import Iterator, Optional

try:  # pragma: no cover - not available in the build image
    import dnaio as _dnaio
    import xopen as _xopen
except ImportError:  # pragma: no cover
    _dnaio = None
    _xopen = None


class FastqFormatError(Exception):
    def __init__(self, msg, line=None):
        super().__init__(msg)
        self.message = msg
        self.line = line


if _dnaio is not None:  # pragma: no cover
    FastqFormatError = _dnaio.FastqFormatError  # noqa: F811


class SequenceRecord:
    __slots__ = ("name", "sequence", "qualities")

    def __init__(self, name: str, sequence: str, qualities: Optional[str] = None):
        self.name = name
        self.sequence = sequence
        self.qualities = qualities

    def fastq_bytes(self) -> bytes:
        return b"".join((b"@", self.name.encode("latin-1"), b"\n",
                         self.sequence.encode("latin-1"), b"\n+\n",
                         (self.qualities or "").encode("latin-1"), b"\n"))


def open_read(path: str):
    if _xopen is not None:  # pragma: no cover
        return _xopen.xopen(path, mode="rb", threads=0)
    if str(path).endswith(".gz"):
        return gzip.open(path, "rb")
    return io.open(path, "rb")


def open_write(path: str):
    """compresslevel=1 like the reference's output opener (__init__.py:197-198)."""
    if _xopen is not None:  # pragma: no cover
        return _xopen.xopen(path, mode="wb", compresslevel=1, threads=0)
    if str(path).endswith(".gz"):
        return gzip.open(path, "wb", compresslevel=1)
    return io.open(path, "wb")


def read_fastq(path: str) -> Iterator[SequenceRecord]:
    if _dnaio is not None:  # pragma: no cover
        with _dnaio.open(path, mode="r", opener=lambda p, m="rb": open_read(p)) as reader:
            yield from reader
        return
    with open_read(path) as fh:
        readline = fh.readline
        lineno = 0
        while True:
            header = readline()
            if not header:
                return
            seq = readline()
            plus = readline()
            qual = readline()
            lineno += 4
            if not plus or not header.startswith(b"@") or not plus.startswith(b"+"):
                raise FastqFormatError(f"{path}: malformed FASTQ record", line=lineno - 4)
            yield SequenceRecord(header[1:].rstrip(b"\r\n").decode("latin-1"),
                                 seq.rstrip(b"\r\n").decode("latin-1"),
                                 qual.rstrip(b"\r\n").decode("latin-1"))


def _mate_id(name: str) -> str:
    ident = name.split(None, 1)[0] if name else name
    if len(ident) > 2 and ident[-2] == "/" and ident[-1] in "123":
        ident = ident[:-2]
    return ident


def records_are_mates(*records) -> bool:
    if _dnaio is not None:  # pragma: no cover
        return _dnaio.records_are_mates(*records)
    first = _mate_id(records[0].name)
    return all(_mate_id(r.name) == first for r in records[1:])
