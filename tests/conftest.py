import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"
BENCH = REPO / "nlotrajectories_b200" / "benchmarks"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def bench_yaml(prefix: str) -> Path:
    return next(BENCH.glob(f"{prefix}*.yaml"))


def close(a, b, tol=1e-5):
    """The parity bar of BASELINE.md: |a-b| <= tol * max(1, |b|), elementwise; returns the violation mask."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) > tol * np.maximum(1.0, np.abs(b))


@pytest.fixture(scope="session")
def shipped_net():
    from oracle import sdf_oracle as so
    return so.from_npz(GOLDEN / "sdf_shipped_fourier128_weights.npz")


@pytest.fixture(scope="session")
def library():
    from nlotrajectories_b200 import build, lib
    build.build_library()
    return lib.load()
