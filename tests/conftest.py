import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    # the oracle checker and the product library are built artefacts (git-ignored): make sure both exist
    if not os.path.exists(os.path.join(ROOT, "oracle", "liborc.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liborc.so"])
    if not os.path.exists(os.path.join(ROOT, "lisp_mcmc_b200", "libmfit_cuda.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "lisp_mcmc_b200", "csrc")])


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    return oracle
