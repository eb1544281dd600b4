"""Test configuration. `-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI export check, gloo multi-process logic.
`-m gpu`: parity tests proper — the CUDA library through its C ABI against the CPU oracle on identical seeded inputs."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_SO = os.path.join(ROOT, "oracle", "liblumen_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_bsdf.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """ctypes bindings of the CPU oracle (built on demand; test infrastructure only)."""
    from lumenrenderer_b200 import api
    if not os.path.exists(ORACLE_SO):
        import subprocess
        subprocess.run(["make", "liblumen_oracle.so"], cwd=os.path.join(ROOT, "oracle"), check=True)
    return api.Bindings(ctypes.CDLL(ORACLE_SO), "lo_")


@pytest.fixture(scope="session")
def gpu():
    """The product: liblumen_b200.so. Fails loudly when it is missing — there is no fallback to test instead."""
    import lumenrenderer_b200 as lr
    return lr.bindings()


def rel_l1(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).sum() / max(np.abs(b).sum(), 1e-30))
