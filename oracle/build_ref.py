#!/usr/bin/env python3
"""Build recipe for ``oracle/_ref``: the UNMODIFIED reference, compiled where it lies.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is on the product path; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference / cpu_baseline
arm may load what this script produces.

What it does (reference ``setup.py:52-56``):

* compiles ``/root/reference/src/fastqdedup/_{trie,distance,fastq}module.c`` with
  plain ``gcc -O2 -fPIC -shared`` (the reference passes no extra flags) straight
  from the read-only reference tree into ``oracle/_ref/fastqdedup/_X<EXT_SUFFIX>``;
* byte-compiles the reference's ``__init__.py`` (the three cluster dissection
  functions and ``deduplicate_cluster``, ``src/fastqdedup/__init__.py:60-288``) into
  a source-less ``oracle/_ref/fastqdedup/package_init.bin`` (CPython bytecode; not named
  ``.pyc`` because repository snapshots drop ``*.pyc``), loaded by ``oracle/ref_loader.py``.

Only binaries are written (``.so`` / ``.pyc``); no reference source file is copied
into this repository, and ``oracle/_ref/`` is git-ignored.  The reference tree does
not exist on the GPU box: there this script is a no-op and the prebuilt files that
travelled with the snapshot are used.
"""
import os
import py_compile
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/fastqdedup"
OUT_PKG = os.path.join(HERE, "_ref", "fastqdedup")
MODULES = ("trie", "distance", "fastq")
BYTECODE = "package_init.bin"


def ref_is_built() -> bool:
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    names = [f"_{m}{suffix}" for m in MODULES] + [BYTECODE]
    return all(os.path.exists(os.path.join(OUT_PKG, n)) for n in names)


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is usable afterwards."""
    if not os.path.isdir(REF_SRC):
        return ref_is_built()
    if ref_is_built() and not force:
        return True
    os.makedirs(OUT_PKG, exist_ok=True)
    include = sysconfig.get_paths()["include"]
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    for m in MODULES:
        out = os.path.join(OUT_PKG, f"_{m}{suffix}")
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-DNDEBUG", "-w",
               f"-I{include}", f"-I{REF_SRC}",
               os.path.join(REF_SRC, f"_{m}module.c"), "-o", out]
        subprocess.run(cmd, check=True)
    py_compile.compile(os.path.join(REF_SRC, "__init__.py"),
                       cfile=os.path.join(OUT_PKG, BYTECODE),
                       doraise=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "unavailable (no /root/reference, no prebuilt files)")
    sys.exit(0 if ok else 1)
