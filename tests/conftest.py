import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running parity sweep")


@pytest.fixture(scope="session")
def oracle():
    import t3oracle
    return t3oracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference itself, as shipped (oracle/_ref/libt3ref.so)."""
    import t3oracle
    if not t3oracle.Reference.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return t3oracle.Reference(fixed=False)


@pytest.fixture(scope="session")
def ref_fixed():
    """The reference with the 3-line arithmetic repair (oracle/_ref/libt3ref_fixed.so)."""
    import t3oracle
    if not t3oracle.Reference.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return t3oracle.Reference(fixed=True)


@pytest.fixture(scope="session")
def ref_new():
    """The reference's NEW-generation core (oracle/_ref/libt3ref_new.so), SURVEY 8(f).3."""
    import t3oracle
    try:
        return t3oracle.ReferenceNew()
    except (FileNotFoundError, OSError):
        pytest.skip("oracle/_ref/libt3ref_new.so not built and /root/reference absent")
