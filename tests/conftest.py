import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
warnings.filterwarnings("ignore", category=DeprecationWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout (/root/reference)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def seville_arrays():
    from invertsy_b200.env.world import Seville2009
    return Seville2009.load_world()


@pytest.fixture(scope="session")
def sparse_arrays():
    from invertsy_b200.env.world import SimpleWorld
    return SimpleWorld.load_world()


@pytest.fixture(scope="session")
def omm1000():
    import isy_oracle as o
    pitch, yaw = o.fibonacci_eye(1000)
    return pitch, yaw, o.eye_rotations(pitch, yaw)


def circ_diff(a, b):
    """|a - b| on the circle, NaN where both are NaN."""
    d = np.abs((np.asarray(a) - np.asarray(b) + np.pi) % (2 * np.pi) - np.pi)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0., d)


def pytest_sessionfinish(session, exitstatus):
    """bounds-checked builds (tools/debug_bounds.py): report the kernels' own assert count, fail the session if any fired"""
    if os.environ.get("ISY_REPORT_BOUNDS") != "1":
        return
    import ctypes
    from invertsy_b200 import _lib
    line = ctypes.c_int32(0)
    n = int(_lib.lib().isy_debug_bounds_violations(0, ctypes.byref(line)))
    print("\n[isy bounds] violations: %d%s" % (n, (" (first at csrc line %d)" % line.value) if n > 0 else
                                               (" -- NOT a bounds-checked build" if n == -1 else (" -- no CUDA device" if n < 0 else ""))))
    if n != 0:
        session.exitstatus = 3
