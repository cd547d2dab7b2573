"""Heat-map grid through a cone-sampled eye: rows x cols positions x 36 headings, 1000 ommatidia x S samples (rho = 4 deg)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import invertsy_b200 as ib
from invertsy_b200 import views

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=50)
ap.add_argument("--cols", type=int, default=50)
ap.add_argument("--samples", type=int, default=16)
a = ap.parse_args()
rng = np.random.RandomState(2021)
N, S = 1000, a.samples
pitch, yaw = ib.fibonacci_lattice(N)
rho = np.deg2rad(4.)
rp = np.clip(pitch[:, None] + rho / 2 * rng.randn(N, S), -1.5, 1.5).ravel()
ry = ((yaw[:, None] + rho / 2 * rng.randn(N, S) + np.pi) % (2 * np.pi) - np.pi).ravel()
r = ib.Renderer(ib.Seville2009(), ib.EyeLayout.from_angles(rp, ry, samples=S), ib.Sky(np.deg2rad(60), np.pi))
pos, yaw_deg = views.grid_views(a.rows, a.cols, 36)
pos_t = torch.as_tensor(pos, device="cuda")
yw = torch.as_tensor(np.deg2rad(yaw_deg), device="cuda")
out = {}
names = ("rgb", "Y", "dop", "aop")
for _ in range(2):
    r.render(pos_t, yw, outputs=names, out=out)
torch.cuda.synchronize(); r.check()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    r.render(pos_t, yw, outputs=names, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
B = pos.shape[0] * 36
pc = r.phase_cycles().astype(float)
print(json.dumps({"views": B, "rays_per_view": N * S, "ms": ms, "views_per_s": B / ms * 1e3, "rays_per_s": B * N * S / ms * 1e3,
                  "phases": (pc[:3] / pc[:3].sum()).tolist()}))
