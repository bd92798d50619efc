import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (always buildable: gcc + oracle/ftk_oracle.c)."""
    from oracle import pyoracle as po
    return po.OracleLib()


@pytest.fixture(scope="session")
def reflib():
    """The reference's own sources compiled in place; skipped where neither the .so nor /root/reference exists."""
    from oracle import pyoracle as po
    if not po.have_ref():
        pytest.skip("oracle/_ref/libftk_ref.so not built and /root/reference absent")
    return po.RefLib()


@pytest.fixture(scope="session")
def euroc_golden():
    return dict(np.load(os.path.join(GOLDEN, "euroc_klt_golden.npz")))


@pytest.fixture(scope="session")
def matcher_golden():
    return dict(np.load(os.path.join(GOLDEN, "matcher_golden.npz")))


@pytest.fixture(scope="session")
def ctx():
    """GPU context; the product path must fail loudly (not skip) when the CUDA library or the GPU is missing."""
    import feature_tracker_b200 as ft
    return ft.Context(0)


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and bool((a.view(np.uint32) == b.view(np.uint32)).all())
