"""One render of a slice of the heat-map grid, for ncu: python tools/prof_grid.py [--positions N] [--walk] [--resp]"""
import argparse, os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import invertsy_b200 as ib
from invertsy_b200 import _lib
ap = argparse.ArgumentParser()
ap.add_argument("--positions", type=int, default=40000)
ap.add_argument("--resp", action="store_true")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
world = ib.Seville2009()
r = ib.Renderer(world, ib.EyeLayout.from_angles(*ib.fibonacci_lattice(1000)), sky=ib.Sky(np.deg2rad(60.), np.pi))
pos, yaw = ib.views.grid_views(200, 200, 36)
sel = np.linspace(0, pos.shape[0] - 1, a.positions).astype(int)
p, y = torch.as_tensor(pos[sel]).cuda(), torch.as_tensor(np.deg2rad(yaw[sel])).cuda()
names = ("rgb", "hit", "depth", "Y", "dop", "aop") + (("resp",) if a.resp else ())
out = {}
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
ev[0].record()
for i in range(a.reps):
    ev[i + 1].record() if False else None
    r.render(p, y, outputs=names, out=out, sensor=_lib.SensorParams.make((0, 1, 0), 1., 1) if a.resp else None)
    ev[i + 1].record()
torch.cuda.synchronize()
print("ms per render:", " ".join("%.2f" % ev[i].elapsed_time(ev[i + 1]) for i in range(a.reps)))
r.check()
pc = r.phase_cycles().astype(float)
print("cycles/pos project %.0f raster %.0f shade %.0f" % tuple(pc[:3] / max(pc[3], 1)))
