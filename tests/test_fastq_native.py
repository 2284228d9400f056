"""The native FASTQ passes (csrc/fastq_native.cpp: pass 1 = reader/parser threads, mates check, key and quality
slices; pass 2 = emission from the keep bitmap with threaded gzip) against the pure-Python equivalents of the
reference's loops (``src/fastqdedup/__init__.py:160-206``).  Host code only: no GPU needed."""
import gzip
import os

import numpy as np
import pytest

from fastqdedup_b200 import _native, fastq_io, frontend


def write_fastq(path, records, plus_name=False, eol="\n", final_newline=True, members=1):
    text = "".join(f"@{n}{eol}{s}{eol}+{n if plus_name else ''}{eol}{q}{eol}" for n, s, q in records)
    if not final_newline:
        text = text[:-len(eol)]
    data = text.encode("latin-1")
    if str(path).endswith(".gz"):
        cut = [len(data) * k // members for k in range(members + 1)]
        with open(path, "wb") as fh:
            for a, b in zip(cut, cut[1:]):      # several gzip members, cut anywhere
                fh.write(gzip.compress(data[a:b], compresslevel=1))
    else:
        with open(path, "wb") as fh:
            fh.write(data)


def random_records(rng, n, lo=10, hi=80, tag=""):
    out = []
    for i in range(n):
        length = int(rng.integers(lo, hi))
        seq = "".join(rng.choice(list("ACGTN"), length))
        qual = "".join(chr(33 + int(x)) for x in rng.integers(0, 42, length))
        out.append((f"r{i}{tag} extra words", seq, qual))
    return out


def rows_of(rows, n):
    if isinstance(rows, tuple):
        flat, off = rows
        return [bytes(flat[int(off[t]):int(off[t + 1])]) for t in range(n)]
    return [bytes(r) for r in rows]


@pytest.mark.parametrize("n", [0, 1, 5, 40000, 70001])          # across the 32768-record batches
def test_scan_matches_python_slicing(tmp_path, n):
    rng = np.random.default_rng(n + 1)
    r1 = random_records(rng, n, tag="/1")
    r2 = [(name.replace("/1", "/2"), s[::-1], q[::-1]) for name, s, q in random_records(rng, n, tag="/1")]
    r3 = random_records(rng, n, 8, 20, tag="/3")
    r3 = [(a[0].replace("/1", "/3"), b[1], b[2]) for a, b in zip(r1, r3)]
    r2 = [(a[0].replace("/1", "/2"), b[1], b[2]) for a, b in zip(r1, r2)]
    p1, p2, p3 = tmp_path / "a.fastq", tmp_path / "b.fastq.gz", tmp_path / "c.fq"
    write_fastq(p1, r1)
    write_fastq(p2, r2, plus_name=True, members=3)
    write_fastq(p3, r3, eol="\r\n", final_newline=False)
    slices = [slice(16), slice(4, 12), slice(None, None, -2)]
    with _native.FastqScan([str(p1), str(p2), str(p3)], slices, want_quals=True) as scan:
        assert scan.n_records == n
        keys, quals = rows_of(scan.keys, n), rows_of(scan.quals, n)
    join = frontend.joinfunc_from_check_slices(slices)
    for t in range(0, n, max(1, n // 500)):
        assert keys[t].decode() == join([r1[t][1], r2[t][1], r3[t][1]]), t
        assert quals[t].decode() == join([r1[t][2], r2[t][2], r3[t][2]]), t


def test_scan_slices_are_python_slices(tmp_path):
    """Every slice the command line can express (length_string_to_slices, reference :364-375), negative indices
    and steps included, on sequences shorter and longer than the slice."""
    rng = np.random.default_rng(3)
    recs = random_records(rng, 300, 1, 40)
    p = tmp_path / "x.fastq"
    write_fastq(p, recs)
    for spec in ["8", "4:8", "::8", "None:None:16", "-5:", ":-3", "::-1", "20:2:-3", "-2:-30:-1", "100:", ":100", "3:3", "7:2"]:
        slc = frontend.length_string_to_slices(spec)[0]
        with _native.FastqScan([str(p)], [slc]) as scan:
            keys = rows_of(scan.keys, len(recs))
        assert [k.decode() for k in keys] == [s[slc] for _, s, _ in recs], spec
    with _native.FastqScan([str(p)]) as scan:           # no slices: whole sequences
        assert [k.decode() for k in rows_of(scan.keys, len(recs))] == [s for _, s, _ in recs]
    with pytest.raises(ValueError, match="slice step cannot be zero"):
        _native.FastqScan([str(p)], [slice(None, None, 0)])


def test_fixed_length_rows_come_back_as_a_matrix(tmp_path):
    rng = np.random.default_rng(4)
    recs = random_records(rng, 1000, 50, 51)
    p = tmp_path / "x.fastq.gz"
    write_fastq(p, recs)
    with _native.FastqScan([str(p)], [slice(12)], want_quals=True) as scan:
        assert isinstance(scan.keys, np.ndarray) and scan.keys.shape == (1000, 12) and scan.quals.shape == (1000, 12)
        assert bytes(scan.keys[7]).decode() == recs[7][1][:12]


def test_stops_at_the_shortest_input_and_checks_mates(tmp_path):
    rng = np.random.default_rng(5)
    r1 = random_records(rng, 50)
    p1, p2 = tmp_path / "a.fastq", tmp_path / "b.fastq"
    write_fastq(p1, r1)
    write_fastq(p2, r1[:33])
    with _native.FastqScan([str(p1), str(p2)]) as scan:           # zip() semantics (reference :180)
        assert scan.n_records == 33
    bad = list(r1)
    bad[20] = ("someone_else 1", bad[20][1], bad[20][2])
    write_fastq(p2, bad)
    with pytest.raises(_native.FqdFastqError) as e:
        _native.FastqScan([str(p1), str(p2)])
    # the reference's message (:182-185), names in file order
    assert str(e.value) == f"FASTQ files not in sync: {r1[20][0]}, someone_else 1 are not mates."
    # ... and the Python front end raises the same text through its own readers
    with pytest.raises(fastq_io.FastqFormatError, match="are not mates"):
        list(frontend.fastq_files_to_records([str(p1), str(p2)]))


def test_malformed_inputs(tmp_path):
    p = tmp_path / "x.fastq"
    p.write_text("@a\nACGT\n+\nIIII\n@b\nACGT\n-\nIIII\n")
    with pytest.raises(_native.FqdFastqError, match="malformed FASTQ record at line 5"):
        _native.FastqScan([str(p)])
    p.write_text("@a\nACGT\n+\nIIII\n@b\nAC")
    with pytest.raises(_native.FqdFastqError, match="premature end of file"):
        _native.FastqScan([str(p)])
    p.write_text("@a\nACGT\n+\nIII\n")
    with pytest.raises(_native.FqdFastqError, match="lengths differ"):
        _native.FastqScan([str(p)])
    with pytest.raises(OSError, match="cannot open"):
        _native.FastqScan([str(tmp_path / "missing.fastq")])
    gz = tmp_path / "t.fastq.gz"
    data = gzip.compress(b"@a\nACGT\n+\nIIII\n" * 5000)
    gz.write_bytes(data[:len(data) // 2])
    with pytest.raises(_native.FqdFastqError, match="truncated|corrupt"):
        _native.FastqScan([str(gz)])


@pytest.mark.parametrize("n,frac", [(0, 0.5), (3, 1.0), (40000, 0.3), (70001, 0.01), (500, 0.0)])
def test_emit_matches_python_writer(tmp_path, n, frac):
    rng = np.random.default_rng(n + 7)
    r1 = random_records(rng, n, tag="/1")
    r2 = [(a[0].replace("/1", "/2"), b[1], b[2]) for a, b in zip(r1, random_records(rng, n))]
    p1, p2 = tmp_path / "in1.fastq.gz", tmp_path / "in2.fastq"
    write_fastq(p1, r1, plus_name=True, members=2)
    write_fastq(p2, r2, eol="\r\n")
    keep = rng.random(n) < frac
    words = np.packbits(keep, bitorder="little")
    words = np.concatenate([words, np.zeros((-len(words)) % 4, dtype=np.uint8)]).view(np.uint32)
    o1, o2 = tmp_path / "out1.fastq", tmp_path / "out2.fastq.gz"
    written = _native.fastq_emit([str(p1), str(p2)], [str(o1), str(o2)], words, n, threads=3)
    assert written == int(keep.sum())
    w1, w2 = tmp_path / "want1.fastq", tmp_path / "want2.fastq"
    frontend.filter_fastq_files_on_bitmap([str(p1), str(p2)], [str(w1), str(w2)], keep)
    assert o1.read_bytes() == w1.read_bytes()
    assert gzip.decompress(o2.read_bytes()) == w2.read_bytes()       # gzip members differ, the bytes inside do not


def test_records_across_chunk_boundaries_and_a_giant_record(tmp_path):
    """The parser works in place on the reader's 8 MB chunks: a record cut by a chunk boundary is completed in the
    next chunk's headroom, a record longer than the headroom (64 KB) takes the copying path.  ~40 MB of long records
    (every chunk boundary cuts one) with a 300 kB record in the middle, scanned and emitted."""
    rng = np.random.default_rng(77)
    recs = random_records(rng, 24000, 800, 2400)
    giant = "".join(rng.choice(list("ACGT"), 300_000))
    recs.insert(12000, ("giant read", giant, "I" * len(giant)))
    n = len(recs)
    src = tmp_path / "long.fastq"
    write_fastq(src, recs, final_newline=False)
    with _native.FastqScan([str(src)], [slice(5, 905, 3)], want_quals=True) as scan:
        assert scan.n_records == n
        keys, quals = rows_of(scan.keys, n), rows_of(scan.quals, n)
    for t in list(range(0, n, 97)) + [11999, 12000, 12001, n - 1]:
        assert keys[t].decode() == recs[t][1][5:905:3], t
        assert quals[t].decode() == recs[t][2][5:905:3], t
    keep = rng.random(n) < 0.3
    keep[12000] = True
    words = np.packbits(keep, bitorder="little")
    words = np.concatenate([words, np.zeros((-len(words)) % 4, dtype=np.uint8)]).view(np.uint32)
    for out in (tmp_path / "out.fastq", tmp_path / "out.fastq.gz"):
        _native.fastq_emit([str(src)], [str(out)], words, n)
        data = gzip.open(out, "rb").read() if str(out).endswith(".gz") else open(out, "rb").read()
        want = "".join(f"@{nm}\n{s}\n+\n{q}\n" for (nm, s, q), k in zip(recs, keep) if k).encode("latin-1")
        assert data == want
