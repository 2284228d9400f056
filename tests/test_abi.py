"""The drop-in boundary: libfqd_b200.so loads, exports every symbol include/fqd_b200.h
declares, the ctypes structures match the header's layout, and without a GPU every compute
entry fails loudly (no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fqd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fqd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from fastqdedup_b200 import _native
    lib = _native.load()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/fqd_b200.h but not exported"
    assert sorted(_native.EXPORTS) == names


def test_struct_layouts_match_header():
    """sizeof of the two ABI structs as the C compiler sees them."""
    import subprocess
    import tempfile
    from fastqdedup_b200 import _native
    src = ('#include <stdio.h>\n#include "fqd_b200.h"\n'
           'int main(void){printf("%zu %zu\\n", sizeof(fqd_cluster_job), sizeof(fqd_cluster_stats));return 0;}\n')
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        job, stats = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert ctypes.sizeof(_native.ClusterJob) == job
    assert ctypes.sizeof(_native.ClusterStats) == stats


def test_python_shims_expose_the_reference_surface():
    import fastqdedup_b200 as pkg
    from fastqdedup_b200 import _distance, _fastq, _trie
    assert _fastq.DEFAULT_PHRED_OFFSET == 33
    assert callable(_distance.within_distance) and callable(_fastq.average_error_rate)
    for attr in ("add_sequence", "contains_sequence", "pop_cluster", "memory_size", "raw_stats",
                 "alphabet", "number_of_sequences"):
        assert hasattr(_trie.Trie, attr)
    for name in ("Trie", "within_distance", "deduplicate_cluster", "cluster_dissection_directional",
                 "cluster_dissection_adjacency", "cluster_dissection_highest_count",
                 "CLUSTER_DISSECTION_METHODS", "length_string_to_slices", "argument_parser", "main"):
        assert hasattr(pkg, name)
    assert set(pkg.CLUSTER_DISSECTION_METHODS) == {"highest_count", "adjacency", "directional"}


def test_argument_errors_need_no_gpu():
    # wrong types are rejected by the shims before any device work, like the reference
    from fastqdedup_b200 import _distance, _fastq
    with pytest.raises(TypeError):
        _distance.within_distance(b"AA", "AA", 1)
    with pytest.raises(ValueError, match="phred_scores must be ASCII encoded"):
        _fastq.average_error_rate(chr(128))


def test_no_cpu_fallback_without_a_device():
    from fastqdedup_b200 import _native
    lib = _native.load()
    if lib.fqd_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_native.FqdCudaError, match="no CPU fallback"):
        _native.Context(0)
    import fastqdedup_b200 as pkg
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.within_distance("AAAA", "AAAC", 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.Trie()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under fastqdedup_b200/ may reference it."""
    pkg = os.path.join(ROOT, "fastqdedup_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+(oracle|ref_loader|build_ref)\b", text, re.M), f
                assert "fqd_oracle" not in text, f
