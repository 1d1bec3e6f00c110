"""pytest configuration.

Markers:
  gpu   -- needs a B200; these are the parity tests proper and call the product through its
           C ABI (libsrp_b200.so via ctypes, or the relinked reference scene executables).
Everything else runs on a CPU-only box: the oracle against the committed golden vectors,
host logic, C-ABI surface checks and world_size-2 gloo tests of the multi-GPU plumbing.

The oracle (oracle/_ref, the unmodified reference) is test infrastructure only.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(kind, name):
    z = np.load(GOLDEN / kind / f"{name}.npz")
    return z["color"], z["depth"], z["stencil"]


def assert_planes_equal(got, want, what):
    """bit-exact on all three planes (north_star: coverage, depth and stencil bit-exact;
    colour within 1 LSB -- we hold colour to 0 LSB too and report the LSB distance if not)"""
    names = ("color", "depth", "stencil")
    problems = []
    for n, a, b in zip(names, got, want):
        assert a.shape == b.shape, f"{what}: {n} shape {a.shape} != {b.shape}"
        bad = a != b
        if bad.any():
            ys, xs = np.nonzero(bad)
            extra = ""
            if n == "color":
                ch = lambda v, s: ((v >> s) & 255).astype(np.int32)
                lsb = max(int(np.abs(ch(a, s) - ch(b, s)).max()) for s in (24, 16, 8, 0))
                extra = f", max channel distance {lsb} LSB, coverage differs on {int(((a != 0) != (b != 0)).sum())} px"
            problems.append(f"{n}: {int(bad.sum())} px differ, first at x={xs[0]} y={ys[0]} "
                            f"got {int(a[ys[0], xs[0]]):#x} want {int(b[ys[0], xs[0]]):#x}{extra}")
    assert not problems, f"{what}: " + "; ".join(problems)


@pytest.fixture(scope="session")
def product():
    from srp_b200 import host
    return host.load_product()


@pytest.fixture(scope="session")
def reference():
    from oracle import refhost
    if not refhost.available():
        pytest.skip("oracle/_ref/libref_host.so not built (needs /root/reference at build time)")
    return refhost.load_oracle_reference()
