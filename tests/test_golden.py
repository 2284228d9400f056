"""Committed golden vectors generated from the unmodified reference
(tests/golden/make_golden.py): the oracle must reproduce them on the CPU, the CUDA path on
the GPU, and the drop-in deduplicate_cluster must write byte-identical FASTQ files."""
import hashlib
import json
import logging
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
CASES = json.load(open(os.path.join(GOLD, "cluster_cases.json")))
FASTQ_CASES = json.load(open(os.path.join(GOLD, "fastq_cases.json")))
STAT_FIELDS = ("total_records", "discarded_records", "number_of_sequences", "number_of_uniques",
               "number_of_clusters", "number_selected")


def _inputs(case):
    keys = [k.encode("latin-1") for k in case["keys"]]
    quals = None if case["quals"] is None else [q.encode("latin-1") for q in case["quals"]]
    return keys, quals


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_golden(oracle, case):
    keys, quals = _inputs(case)
    r = oracle.cluster(keys, quals, case["max_distance"], case["use_edit_distance"], case["method"],
                       case["max_average_error_rate"])
    for f in STAT_FIELDS:
        assert r[f] == case["expect"][f], f
    for f in ("first", "count", "label"):
        assert r[f].tolist() == case[f], f
    assert r["selected"].astype(int).tolist() == case["selected"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_reproduces_golden(gpu_ctx, case):
    from fastqdedup_b200.clustering import cluster_keys
    keys, quals = _inputs(case)
    r = cluster_keys(keys, quals, case["max_distance"], case["use_edit_distance"], case["method"],
                     case["max_average_error_rate"], context=gpu_ctx)
    for f in STAT_FIELDS:
        assert getattr(r, f) == case["expect"][f], f
    for f in ("first", "count", "label"):
        assert getattr(r, f).tolist() == case[f], f
    assert r.selected.astype(int).tolist() == case["selected"]
    assert np.nonzero(r.keep_mask())[0].tolist() == [f for f, s in zip(case["first"], case["selected"]) if s]


@pytest.mark.gpu
@pytest.mark.parametrize("case", FASTQ_CASES, ids=[c["case"] for c in FASTQ_CASES])
def test_deduplicate_cluster_writes_identical_fastq(gpu_ctx, tmp_path, caplog, case):
    """End to end through the drop-in front end: same output bytes and the same numbers in
    the reference's log lines (__init__.py:253-259, 279-281)."""
    import re

    import fastqdedup_b200 as pkg
    cdir = os.path.join(GOLD, "fastq", case["case"])
    inputs = [os.path.join(cdir, f) for f in case["inputs"]]
    outputs = [str(tmp_path / f"out{i}.fastq") for i in range(len(inputs))]
    slices = pkg.length_string_to_slices(case["check_lengths"]) if case["check_lengths"] else None
    with caplog.at_level(logging.INFO, logger="fastqdedup"):
        pkg.deduplicate_cluster(inputs, outputs, slices, case["max_distance"], case["max_average_error_rate"],
                                pkg.CLUSTER_DISSECTION_METHODS[case["method"]], case["use_edit_distance"])
    for out, expected, sha in zip(outputs, case["outputs"], case["sha256"]):
        got = open(out, "rb").read()
        assert hashlib.sha256(got).hexdigest() == sha
        assert got == open(os.path.join(cdir, expected), "rb").read()
    log = caplog.text
    c = case["counters"]
    if "discarded_records" in c:
        m = re.search(r"(\d+) records out of (\d+) records had an error rate higher than", log)
        assert (int(m.group(1)), int(m.group(2))) == (c["discarded_records"], c["total_records"])
    assert int(re.search(r"Processed (\d+) sequences", log).group(1)) == c["number_of_sequences"]
    m = re.search(r"Found (\d+) distinct reads in (\d+) clusters", log)
    assert (int(m.group(1)), int(m.group(2))) == (c["number_selected"], c["number_of_clusters"])


@pytest.mark.gpu
def test_custom_dissection_callable_uses_trie_shim(gpu_ctx, tmp_path):
    """A user-supplied callable gets the reference's own loop over the Trie shim."""
    import fastqdedup_b200 as pkg
    case = FASTQ_CASES[0]
    cdir = os.path.join(GOLD, "fastq", case["case"])

    def my_directional(cluster, max_distance=1, use_edit_distance=False):
        yield from pkg.cluster_dissection_directional(cluster, max_distance, use_edit_distance)

    out = str(tmp_path / "o.fastq")
    pkg.deduplicate_cluster([os.path.join(cdir, case["inputs"][0])], [out],
                            pkg.length_string_to_slices(case["check_lengths"]), case["max_distance"],
                            case["max_average_error_rate"], my_directional, case["use_edit_distance"])
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == case["sha256"][0]
