import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")
    # building the product and the checker is not using them; both are quick no-ops when fresh
    import __graft_entry__
    __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle():
    import oracle as _oracle
    _oracle.build()
    return _oracle


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled into oracle/_ref (None if it was never built)."""
    import ref_loader
    ref = ref_loader.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    return ref


@pytest.fixture(scope="session")
def gpu_ctx():
    from fastqdedup_b200 import _native
    if _native.load().fqd_device_count() < 1:
        pytest.fail("no CUDA device: -m gpu tests need a B200 (the product has no CPU fallback)")
    return _native.default_context()
