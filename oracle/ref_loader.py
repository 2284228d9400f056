"""Import the compiled, unmodified reference (``oracle/_ref``) as a module object.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's reference arm).
The reference package is imported under its own name ``fastqdedup`` from
``oracle/_ref`` with the two absent third-party imports (dnaio, xopen) satisfied by
``oracle/stubs``.  Returns ``None`` when ``oracle/_ref`` was never built.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
_cached = None


def load_reference():
    global _cached
    if _cached is not None:
        return _cached
    sys.path.insert(0, HERE)
    try:
        import build_ref
    finally:
        sys.path.pop(0)
    if not build_ref.build():
        return None
    for p in (os.path.join(HERE, "stubs"), os.path.join(HERE, "_ref")):
        if p not in sys.path:
            sys.path.insert(0, p)
    _cached = importlib.import_module("fastqdedup")
    return _cached
