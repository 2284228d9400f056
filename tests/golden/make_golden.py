#!/usr/bin/env python3
"""Generates the committed golden vectors from the UNMODIFIED reference (oracle/_ref, built
from /root/reference by oracle/build_ref.py; dnaio/xopen satisfied by oracle/stubs).

    python tests/golden/make_golden.py        # only where /root/reference exists

* cluster_cases.json : key/quality lists + parameters -> what the reference's loops
  (Trie.add_sequence / pop_cluster / cluster_dissection_*, __init__.py:240-276) produce.
* fastq/<case>/      : small FASTQ inputs + the byte-exact outputs of the reference's
  deduplicate_cluster (__init__.py:209-288) and the counters of its log lines.
"""
import hashlib
import io
import json
import logging
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import oracle          # noqa: E402
import ref_loader      # noqa: E402

METHODS = ("directional", "adjacency", "highest_count")


def rand_reads(rng, n_strings, n_reads, alphabet, lo, hi):
    strings = ["".join(rng.choice(list(alphabet), size=int(rng.integers(lo, hi + 1)))) for _ in range(n_strings)]
    return [strings[i] for i in rng.integers(0, n_strings, size=n_reads)]


def cluster_cases():
    rng = np.random.default_rng(20221)
    cases = []

    def add(name, keys, quals=None, ds=(1,), edits=(False,), err=1.0):
        for d in ds:
            for edit in edits:
                for m in METHODS:
                    r = oracle.ref_cluster([k.encode("latin-1") for k in keys],
                                           None if quals is None else [q.encode("latin-1") for q in quals],
                                           d, edit, m, err)
                    cases.append({
                        "name": f"{name}/d{d}/{'edit' if edit else 'hamming'}/{m}",
                        "keys": keys, "quals": quals, "max_distance": d, "use_edit_distance": edit,
                        "method": m, "max_average_error_rate": err,
                        "expect": {k: int(r[k]) for k in ("total_records", "discarded_records",
                                                          "number_of_sequences", "number_of_uniques",
                                                          "number_of_clusters", "number_selected")},
                        "first": r["first"].tolist(), "count": r["count"].tolist(),
                        "label": r["label"].tolist(), "selected": r["selected"].astype(int).tolist(),
                    })

    # reference test inputs (tests/test_trie.py:75-136, tests/test_fastqdedup.py:38-97)
    add("trie_pop_cluster", ["AAAA", "AAAA", "AAAC", "AAGC", "AGGC", "CCCG", "CCCG", "TTCA", "TTCC",
                             "TTTA", "TTT", "TTC"], edits=(False, True))
    tc = [(3, "AAAGT"), (10, "AAAAT"), (50, "AACAA"), (60, "AAAAA"), (10, "CAAAA"), (30, "CTAAA")]
    add("dissection_cluster", [s for c, s in tc for _ in range(c)])
    chain = [(100, "GGGGGG"), (1, "GGGTGG"), (1, "GGGTTG"), (1, "GGCTTG"), (1, "GACTTG"), (2, "AACTTG")]
    add("long_chain", [s for c, s in chain for _ in range(c)])
    # SURVEY appendix C
    add("ties_ascii_largest", ["AAAA", "AAAT", "AAAN", "AAAG", "AAAC", "AAAa"], edits=(False, True))
    add("count_boundary", ["ACGT"] * 5 + ["ACGA"] * 3 + ["ACGC"] * 4 + ["TTTT"] * 2 + ["TTTA"] * 2 + ["TTTC"])
    add("count2_next_to_singletons", ["GGGG", "GGGG", "GGGA", "GGAA", "GAAA"])
    add("mixed_lengths", ["ACGTAC", "ACGTA", "ACGT", "ACGTAC", "ACGTAG", "CGTAC", "ACGTACG", ""],
        ds=(1, 2), edits=(False, True))
    add("d0", ["ACGT", "ACGA", "ACGT", "TTTT"], ds=(0,), edits=(False, True))
    add("lowercase_iupac", ["acgt", "ACGT", "acgn", "ACGR", "ACGY", "acgt", "AcGt"], edits=(False, True))
    add("filter_threshold", ["ACGTACGTACGT", "ACGTACGTACGT", "ACGTACGTACGA", "TTTTTTTTTTTT", "GGGGGGGGGGGG",
                             "GGGGGGGGGGGG", "CCCCCCCCCCCC"],
        ["?" * 12, "I" * 12, "I" * 12, "?" * 12, "I" * 10 + "55", "I" * 12, ""], err=0.001)
    add("random_acgtn_5to7", rand_reads(rng, 300, 900, "ACGTN", 5, 7), ds=(1, 2), edits=(False, True))
    add("random_ac_percolating", rand_reads(rng, 200, 700, "AC", 6, 6), ds=(1, 2, 3))
    add("random_12mers", rand_reads(rng, 400, 1500, "ACGT", 12, 12), ds=(1, 2), edits=(False, True))
    q = ["".join(rng.choice(list("I?5+"), p=[0.9, 0.05, 0.03, 0.02], size=8)) for _ in range(800)]
    add("random_with_filter", rand_reads(rng, 200, 800, "ACGT", 8, 8), q, err=0.001, edits=(False, True))
    return cases


def write_fastq(path, records):
    with open(path, "wt") as fh:
        for name, seq, qual in records:
            fh.write(f"@{name}\n{seq}\n+\n{qual}\n")


def fastq_cases(ref):
    rng = np.random.default_rng(777)
    out = []
    base = os.path.join(HERE, "fastq")

    def make_reads(n, umi_len, read_len, n_mol, trunc=0.0, lowq=0.0):
        mols = ["".join(rng.choice(list("ACGT"), size=umi_len + read_len)) for _ in range(n_mol)]
        recs = []
        for i in range(n):
            s = list(mols[int(rng.integers(0, n_mol))])
            for p in range(len(s)):
                if rng.random() < 0.01:
                    s[p] = "ACGTN"[int(rng.integers(0, 5))]
            s = "".join(s)
            if rng.random() < trunc:
                s = s[:int(rng.integers(4, umi_len + 4))]
            ql = "I" * len(s)
            if rng.random() < lowq:
                ql = "".join(rng.choice(list("?5I"), size=len(s)))
            recs.append((f"read{i}", s, ql))
        return recs

    def run(case, files_in, check_lengths, d, err, method, edit):
        cdir = os.path.join(base, case)
        outs = [os.path.join(cdir, f"expected_R{i + 1}.fastq") for i in range(len(files_in))]
        stream = io.StringIO()
        logger = logging.getLogger("fastqdedup")
        logger.setLevel(logging.INFO)
        handler = logging.StreamHandler(stream)
        logger.addHandler(handler)
        slices = ref.length_string_to_slices(check_lengths) if check_lengths else None
        ref.deduplicate_cluster(files_in, outs, slices, d, err, ref.CLUSTER_DISSECTION_METHODS[method], edit)
        logger.removeHandler(handler)
        import re
        log = stream.getvalue()
        counters = {}
        m = re.search(r"(\d+) records out of (\d+) records", log)
        if m:
            counters["discarded_records"], counters["total_records"] = int(m.group(1)), int(m.group(2))
        counters["number_of_sequences"] = int(re.search(r"Processed (\d+) sequences", log).group(1))
        m = re.search(r"Found (\d+) distinct reads in (\d+) clusters", log)
        counters["number_selected"], counters["number_of_clusters"] = int(m.group(1)), int(m.group(2))
        out.append({"case": case, "inputs": [os.path.basename(f) for f in files_in],
                    "outputs": [os.path.basename(f) for f in outs], "check_lengths": check_lengths,
                    "max_distance": d, "max_average_error_rate": err, "method": method,
                    "use_edit_distance": edit, "counters": counters,
                    "sha256": [hashlib.sha256(open(f, "rb").read()).hexdigest() for f in outs]})

    # single end, 12-nt UMI prepended (config 1 shape), -E
    cdir = os.path.join(base, "se_umi12"); os.makedirs(cdir, exist_ok=True)
    recs = make_reads(400, 12, 40, 60)
    write_fastq(os.path.join(cdir, "in_R1.fastq"), recs)
    run("se_umi12", [os.path.join(cdir, "in_R1.fastq")], "12", 1, 1.0, "directional", False)
    # paired + UMI file (config 2 shape), adjacency
    cdir = os.path.join(base, "pe_umi"); os.makedirs(cdir, exist_ok=True)
    def noisy(seq):
        s = list(seq)
        for p in range(len(s)):
            if rng.random() < 0.01:
                s[p] = "ACGTN"[int(rng.integers(0, 5))]
        return "".join(s)
    mols = [tuple("".join(rng.choice(list("ACGT"), size=k)) for k in (40, 30, 12)) for _ in range(40)]
    picks = rng.integers(0, 40, size=300)
    r1 = [(f"pair{i}", noisy(mols[m][0]), "I" * 40) for i, m in enumerate(picks)]
    r2 = [(f"pair{i}", noisy(mols[m][1]), "I" * 30) for i, m in enumerate(picks)]
    umi = [(f"pair{i}", noisy(mols[m][2]), "I" * 12) for i, m in enumerate(picks)]
    for fn, rr, tag in (("in_R1.fastq", r1, "1"), ("in_R2.fastq", r2, "2"), ("in_R3.fastq", umi, "3")):
        write_fastq(os.path.join(cdir, fn), [(f"{n} {tag}", s, q) for n, s, q in rr])
    run("pe_umi", [os.path.join(cdir, f) for f in ("in_R1.fastq", "in_R2.fastq", "in_R3.fastq")],
        "16,8,12", 1, 1.0, "adjacency", False)
    # filter on (default -e), d=2, truncated reads -> shorter keys, low-quality first occurrences
    cdir = os.path.join(base, "se_filter_d2"); os.makedirs(cdir, exist_ok=True)
    recs = make_reads(400, 8, 24, 50, trunc=0.05, lowq=0.15)
    write_fastq(os.path.join(cdir, "in_R1.fastq"), recs)
    run("se_filter_d2", [os.path.join(cdir, "in_R1.fastq")], "24", 2, 0.001, "directional", False)
    # --edit d=1, highest_count
    cdir = os.path.join(base, "se_edit"); os.makedirs(cdir, exist_ok=True)
    recs = make_reads(300, 8, 16, 40, trunc=0.1)
    write_fastq(os.path.join(cdir, "in_R1.fastq"), recs)
    run("se_edit", [os.path.join(cdir, "in_R1.fastq")], "20", 1, 1.0, "highest_count", True)
    return out


def main():
    ref = ref_loader.load_reference()
    assert ref is not None, "needs oracle/_ref (run where /root/reference exists)"
    with open(os.path.join(HERE, "cluster_cases.json"), "wt") as fh:
        json.dump(cluster_cases(), fh, separators=(",", ":"))
    with open(os.path.join(HERE, "fastq_cases.json"), "wt") as fh:
        json.dump(fastq_cases(ref), fh, indent=1)
    print("golden vectors written")


if __name__ == "__main__":
    main()
