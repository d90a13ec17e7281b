import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def built():
    """The native libraries, built in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    from ataraxia_b200 import build
    build.build_product()
    build.build_oracle()
    return build


@pytest.fixture(scope="session")
def port(built):
    from oracle.bindings import OraclePort
    return OraclePort()


@pytest.fixture(scope="session")
def refcpu(built):
    """The unmodified reference compiled for the host; absent where oracle/ref/Makefile never ran."""
    from oracle import bindings
    if not bindings.have_reference_cpu():
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    return bindings.ReferenceCpu()


@pytest.fixture(scope="session")
def sample_scene_path():
    return GOLDEN / "sample_scene.json"
