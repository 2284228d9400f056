"""Test-only stand-in for the third-party ``dnaio`` package (>=0.9.0, not installed and
not installable here: no network).  Implements exactly the surface the reference uses
(``/root/reference/src/fastqdedup/__init__.py:54-57, 181-185, 206, 243-251``):
``open``, ``SequenceRecord``, ``records_are_mates``, ``FastqFormatError``.

Not part of the product; lives under ``oracle/`` and is only put on ``sys.path`` by
``oracle/ref_loader.py`` so that the unmodified reference package can be imported.
"""
import contextlib


class FastqFormatError(Exception):
    def __init__(self, msg, line=None):
        super().__init__(msg)
        self.message = msg
        self.line = line


class SequenceRecord:
    __slots__ = ("name", "sequence", "qualities")

    def __init__(self, name, sequence, qualities=None):
        self.name = name
        self.sequence = sequence
        self.qualities = qualities

    def fastq_bytes(self, two_headers=False):
        second = self.name if two_headers else ""
        return (f"@{self.name}\n{self.sequence}\n+{second}\n"
                f"{self.qualities}\n").encode("ascii")

    @property
    def id(self):
        return self.name.split(None, 1)[0] if self.name else self.name


def _mate_id(name):
    ident = name.split(None, 1)[0] if name else name
    if len(ident) > 2 and ident[-2] == "/" and ident[-1] in "123":
        ident = ident[:-2]
    elif ident and ident[-1] in "123" and False:
        ident = ident[:-1]
    return ident


def records_are_mates(*records):
    first = _mate_id(records[0].name)
    return all(_mate_id(r.name) == first for r in records[1:])


@contextlib.contextmanager
def open(filename, mode="r", opener=None, **kwargs):  # noqa: A001
    import builtins
    if opener is None:
        fh = builtins.open(filename, "rb")
    else:
        fh = opener(filename, "rb")

    def records():
        while True:
            header = fh.readline()
            if not header:
                return
            seq = fh.readline()
            plus = fh.readline()
            qual = fh.readline()
            if not qual and not plus:
                raise FastqFormatError("Premature end of file", line=None)
            if not header.startswith(b"@") or not plus.startswith(b"+"):
                raise FastqFormatError("Malformed FASTQ record", line=None)
            yield SequenceRecord(header[1:].rstrip(b"\r\n").decode("ascii"),
                                 seq.rstrip(b"\r\n").decode("ascii"),
                                 qual.rstrip(b"\r\n").decode("ascii"))
    try:
        yield records()
    finally:
        fh.close()
