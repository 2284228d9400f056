"""Import the compiled, unmodified reference (``oracle/_ref``) as a module object.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's reference arm).
The reference package is imported under its own name ``fastqdedup`` from
``oracle/_ref`` with the two absent third-party imports (dnaio, xopen) satisfied by
``oracle/stubs``.  Returns ``None`` when ``oracle/_ref`` was never built.
"""
import importlib
import importlib.machinery
import importlib.util
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
    stubs = os.path.join(HERE, "stubs")
    if stubs not in sys.path:
        sys.path.insert(0, stubs)
    pkg_dir = os.path.join(HERE, "_ref", "fastqdedup")
    path = os.path.join(pkg_dir, build_ref.BYTECODE)
    loader = importlib.machinery.SourcelessFileLoader("fastqdedup", path)
    spec = importlib.util.spec_from_file_location("fastqdedup", path, loader=loader,
                                                  submodule_search_locations=[pkg_dir])
    module = importlib.util.module_from_spec(spec)
    sys.modules["fastqdedup"] = module
    try:
        spec.loader.exec_module(module)
    except BaseException:
        sys.modules.pop("fastqdedup", None)
        raise
    _cached = module
    return _cached
