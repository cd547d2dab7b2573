"""Builds variants of the library with extra -D flags (tuning experiments): python tools/tune_build.py NAME=-DFLAG=V,-DFLAG2=W ...
   -> invertsy_b200/libisy_tune_NAME.so ; run a tool against one with ISY_B200_LIB=<path>."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from invertsy_b200 import build as b
procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    out = os.path.join(ROOT, "invertsy_b200", "libisy_tune_%s.so" % name)
    cmd = [os.environ.get("NVCC", "nvcc")] + b.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", out] + b.SOURCES
    procs.append((name, subprocess.Popen(cmd, cwd=b.CSRC)))
for name, p in procs:
    print(name, "rc", p.wait())
