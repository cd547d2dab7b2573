"""Times the familiarity kernels (csrc/isy_fam.cuh) on the heat-map grid's shape: Q queries x M stored x K ommatidia."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invertsy_b200 import familiarity as fam   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--Q", type=int, default=1440000)
ap.add_argument("--M", type=int, default=812)
ap.add_argument("--K", type=int, default=1000)
ap.add_argument("--steps", type=int, default=5)
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.rand((a.Q, a.K), device="cuda", generator=g)
m = torch.rand((a.M, a.K), device="cuda", generator=g)
mem = fam.PerfectMemory().store(m)
out = torch.empty(a.Q, device="cuda")
for _ in range(2):
    mem.familiarity(q, out=out)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
for s, e in ev:
    s.record()
    mem.familiarity(q, out=out)
    e.record()
torch.cuda.synchronize()
ms = float(np.median([s.elapsed_time(e) for s, e in ev]))
flop = 2.0 * a.Q * a.M * a.K
print(json.dumps({"Q": a.Q, "M": a.M, "K": a.K, "ms": ms, "tflops_tf32": flop / ms / 1e9,
                  "queries_per_s": a.Q / ms * 1e3, "query_bytes_GBps": a.Q * a.K * 4 * 2 / ms / 1e6}))
wm = fam.WillshawMemory(a.K, 40 * a.K).store(m)
Qw = min(a.Q, 148 * 64)
for _ in range(2):
    wm.familiarity(q[:Qw], out=out[:Qw])
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
wm.familiarity(q[:Qw], out=out[:Qw])
e.record()
torch.cuda.synchronize()
print(json.dumps({"willshaw": {"views": Qw, "nb_pn": a.K, "nb_kc": 40 * a.K, "ms": s.elapsed_time(e),
                               "views_per_s": Qw / s.elapsed_time(e) * 1e3}}))
