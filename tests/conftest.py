import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The independent CPU restatement (oracle/ps_oracle.cpp), built on demand."""
    from oracle import binding
    binding.build("oracle")
    return binding.load("oracle")


@pytest.fixture(scope="session")
def ref():
    """The reference's own C++ (oracle/_ref/libps_ref.so).  Buildable only where /root/reference
    exists; on the GPU box the prebuilt library travels with the snapshot."""
    from oracle import binding
    if os.path.isdir("/root/reference/cpp"):
        binding.build("ref")
    if not binding.available("ref"):
        pytest.skip("oracle/_ref/libps_ref.so not available")
    return binding.load("ref")


@pytest.fixture(scope="session")
def drv():
    """Checker for the driver-level tests (FindMutations, Mutate, MapAlignments, ViterbiMutate, swfull): the compiled
    reference where it is available, else the restatement (which is pinned against it by the CPU tests)."""
    from oracle import binding
    if os.path.isdir("/root/reference/cpp"):
        binding.build("ref")
    if binding.available("ref") and os.environ.get("PORESEQ_TEST_CHECKER", "ref") != "oracle":
        return binding.load("ref")
    binding.build("oracle")
    return binding.load("oracle")


@pytest.fixture(scope="session")
def ctx():
    from poreseq_b200 import build, poreseqcpp
    if not os.path.exists(build.LIB):
        build.build()
    c = poreseqcpp.Context(0)
    yield c
    c.close()
