"""Host-side logic of the drop-in package that needs no GPU."""
import gzip

import pytest

from fastqdedup_b200 import fastq_io, frontend, length_string_to_slices


@pytest.mark.parametrize(["string", "result"], [
    ("5,6,7", [slice(5), slice(6), slice(7)]),
    ("5:8,3,-5:3:-1", [slice(5, 8), slice(3), slice(-5, 3, -1)]),
    ("None:None:16", [slice(None, None, 16)]),
    ("::16", [slice(None, None, 16)])
])
def test_length_string_to_slices(string, result):
    # /root/reference/tests/test_fastqdedup.py:27-34
    assert length_string_to_slices(string) == result


def test_joinfunc_slices_per_file_then_concatenates():
    join = frontend.joinfunc_from_check_slices([slice(4), slice(2, 4), slice(None, None, -1)])
    assert join(["ACGTACGT", "TTGGCC", "AC"]) == "ACGT" + "GG" + "CA"
    assert join(["AC", "T", ""]) == "AC"          # reads shorter than the slice give shorter keys


def test_fastq_roundtrip_and_mates(tmp_path):
    p = tmp_path / "a.fastq.gz"
    with gzip.open(p, "wb") as fh:
        fh.write(b"@r1 1\nACGT\n+\nIIII\n@r2 1\nTTTT\n+r2\n????\n")
    recs = list(fastq_io.read_fastq(str(p)))
    assert [(r.name, r.sequence, r.qualities) for r in recs] == [("r1 1", "ACGT", "IIII"), ("r2 1", "TTTT", "????")]
    assert recs[1].fastq_bytes() == b"@r2 1\nTTTT\n+\n????\n"
    a = fastq_io.SequenceRecord("x/1", "A", "I")
    assert fastq_io.records_are_mates(a, fastq_io.SequenceRecord("x/2", "C", "I"))
    assert fastq_io.records_are_mates(fastq_io.SequenceRecord("q 1:N", "A", "I"), fastq_io.SequenceRecord("q 2:N", "C", "I"))
    assert not fastq_io.records_are_mates(a, fastq_io.SequenceRecord("y/2", "C", "I"))


def test_unsynced_mates_raise(tmp_path):
    p1, p2 = tmp_path / "1.fastq", tmp_path / "2.fastq"
    p1.write_bytes(b"@a\nAC\n+\nII\n")
    p2.write_bytes(b"@b\nAC\n+\nII\n")
    with pytest.raises(fastq_io.FastqFormatError, match="FASTQ files not in sync: a, b are not mates."):
        list(frontend.fastq_files_to_records([str(p1), str(p2)]))


def test_deduplicate_cluster_validates_file_counts():
    with pytest.raises(ValueError, match="Amount of output files"):
        frontend.deduplicate_cluster(["a", "b"], ["o"], None)
    with pytest.raises(ValueError, match="Amount of check lengths"):
        frontend.deduplicate_cluster(["a", "b"], ["o", "p"], [slice(8)])


def test_cli_flags_match_reference():
    args = frontend.argument_parser().parse_args(
        ["-l", "16,8", "-o", "x", "-o", "y", "-d", "2", "-E", "--edit", "-c", "adjacency", "-vv", "r1", "r2"])
    assert (args.check_lengths, args.output, args.max_distance, args.max_average_error_rate, args.edit,
            args.cluster_dissection_method, args.verbose, args.fastq) == \
        ("16,8", ["x", "y"], 2, 1.0, True, "adjacency", 2, ["r1", "r2"])
    d = frontend.argument_parser().parse_args(["r1"])
    assert (d.max_distance, d.max_average_error_rate, d.cluster_dissection_method, d.prefix) == \
        (1, 0.001, "directional", "fastqdedup_R")
