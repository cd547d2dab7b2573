"""Small drivers for ncu captures: python tools/prof_misc.py {fam|sky|config5}"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import invertsy_b200 as ib
from invertsy_b200 import familiarity as fam
what = sys.argv[1]
if what == "fam":
    g = torch.Generator(device="cuda").manual_seed(1)
    q, m = torch.rand((1440000, 1000), device="cuda", generator=g), torch.rand((812, 1000), device="cuda", generator=g)
    mem = fam.PerfectMemory().store(m)
    for _ in range(3):
        mem.familiarity(q)
elif what == "sky":
    pitch, yaw = ib.fibonacci_lattice(5000, hemisphere=True)
    empty = ib.WorldBase(np.zeros((0, 3, 3), dtype=np.float32), np.zeros((0, 3), dtype=np.float32))
    r = ib.Renderer(empty, ib.EyeLayout.from_angles(pitch, yaw), sky=ib.Sky(1.0, 0.0))
    az, el = np.deg2rad(np.arange(360.)), np.deg2rad(np.arange(90.) + .5)
    suns = np.array([(np.pi / 2 - e, (a + np.pi) % (2 * np.pi) - np.pi) for a in az for e in el])
    K = suns.shape[0]
    pos, yw, sun = torch.zeros((K, 3), dtype=torch.float64, device="cuda"), torch.zeros((K, 1), dtype=torch.float64, device="cuda"), torch.as_tensor(suns, device="cuda")
    out = {}
    for _ in range(3):
        r.render(pos, yw, sun=sun, outputs=("Y", "dop", "aop"), out=out)
elif what == "config5":
    sys.argv = [sys.argv[0], "--configs", "5", "--views5", "592"]
    exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_configs.py")).read())
torch.cuda.synchronize()
