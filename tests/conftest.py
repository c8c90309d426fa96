import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled unmodified reference (+ lodepng, stb_image).  Built here from /root/reference/src; on the GPU box
    the prebuilt oracle/_ref/libfpng_ref.so travels with the snapshot."""
    from oracle.pyoracle import Ref
    if not Ref.available():
        pytest.skip("oracle/_ref/libfpng_ref.so not available")
    return Ref()


@pytest.fixture(scope="session")
def verifiers():
    """wuffs (verifies IDAT CRC-32 and zlib Adler-32) and pvpng: the other two independent decoders of the reference's round-trip
    check (src/fpng_test.cpp:1403-1445, 1571-1606), compiled unmodified from the reference tree (oracle/verifiers_shim.cpp)."""
    from oracle.pyoracle import Verifiers
    try:
        if not Verifiers.available():
            pytest.skip("oracle/_ref/libpng_verifiers.so not available")
        return Verifiers()
    except OSError as e:
        pytest.skip(f"libpng_verifiers.so not loadable: {e}")


@pytest.fixture(scope="session")
def gpu():
    import fpng_b200
    fpng_b200.fpng_init()
    return fpng_b200
