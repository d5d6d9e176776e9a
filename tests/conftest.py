import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer CPU known-answer test")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle, build
    build(ref=True)
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled into oracle/_ref (authoring container, or prebuilt on the GPU box)."""
    from oracle.pyoracle import Reference, build
    build(ref=True)
    if not Reference.available():
        pytest.skip("oracle/_ref/libqldpc_ref.so not available")
    return Reference()


@pytest.fixture(scope="session")
def qldpc():
    import qec_ldpc_b200
    from qec_ldpc_b200 import build as b
    if not os.path.exists(qec_ldpc_b200.LIB_PATH):
        b.build_library()
    qec_ldpc_b200.load_library()
    return qec_ldpc_b200


CODES = {
    "C1": (3, 3, 6, 7, 2, 3),
    "C2": (4, 5, 10, 61, 9, 49),
    "C5": (4, 4, 8, 509, 208, 2),
}
