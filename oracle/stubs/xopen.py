"""Test-only stand-in for the third-party ``xopen`` package (reference call sites:
``/root/reference/src/fastqdedup/__init__.py:55, 197-198``).  Not part of the product."""
import builtins
import gzip


def xopen(filename, mode="r", compresslevel=6, threads=None, **kwargs):
    if "b" not in mode and "t" not in mode:
        mode = mode + "t"
    if str(filename).endswith(".gz"):
        return gzip.open(filename, mode, compresslevel=compresslevel)
    return builtins.open(filename, mode)
