"""
Stand-in for compute-sanitizer (closed on the B200 pool): builds the library a second time with -DISY_DEBUG_BOUNDS
(every indexed shared-memory / output access of the kernels carries an assert, csrc/isy_common.cuh: ISY_BOUND), runs GPU
tests through it (ISY_B200_LIB) and reports how many asserts fired.

    python tools/debug_bounds.py [pytest args ...]        default: the parity, rows-mode, band, edge and familiarity suites
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from invertsy_b200 import build as b   # noqa: E402

DBG = os.path.join(ROOT, "invertsy_b200", "libisy_b200_dbg.so")


def build_debug():
    cmd = [os.environ.get("NVCC", "nvcc")] + b.NVCC_FLAGS + ["-DISY_DEBUG_BOUNDS", "-o", DBG] + b.SOURCES
    subprocess.check_call(cmd, cwd=b.CSRC)
    return DBG


if __name__ == "__main__":
    if not os.path.exists(DBG) or "--rebuild" in sys.argv:
        build_debug()
    args = [a for a in sys.argv[1:] if a != "--rebuild"] or [
        "tests/test_gpu_parity.py", "tests/test_gpu_skytab.py", "tests/test_gpu_bands.py", "tests/test_gpu_edges.py",
        "tests/test_gpu_familiarity.py", "tests/test_gpu_sensor.py", "tests/test_gpu_configs.py", "tests/test_gpu_variants.py",
        "tests/test_gpu_tiles.py"]
    env = dict(os.environ, ISY_B200_LIB=DBG, ISY_REPORT_BOUNDS="1")
    rc = subprocess.call([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] + args, cwd=ROOT, env=env)
    print("pytest exit code %d (the per-process violation count is printed by tests/conftest.py at session end)" % rc)
    sys.exit(rc)
