"""Config 5 (synthetic stress, scaled by --views): 500 k grass blades on 100 m x 100 m, 10 k ommatidia x 16 samples."""
import sys, os, json, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import invertsy_b200 as ib

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=2048)
ap.add_argument("--tris", type=int, default=500000)
ap.add_argument("--side", type=float, default=100.)
ap.add_argument("--omm", type=int, default=10000)
ap.add_argument("--samples", type=int, default=16)
a = ap.parse_args()
rng = np.random.RandomState(2021)
T, side = a.tris, a.side
base = np.c_[rng.uniform(0, side, T), rng.uniform(0, side, T), np.zeros(T)]
d = rng.uniform(-.08, .08, (T, 2))
tri = np.stack([base + np.c_[d, np.zeros(T)], base - np.c_[d, np.zeros(T)],
                base + np.c_[rng.uniform(-.05, .05, (T, 2)), rng.lognormal(np.log(.18), .6, T)]], axis=1).astype(np.float32)
col = np.c_[np.zeros(T), rng.uniform(.07, .98, T), np.zeros(T)].astype(np.float32)
world = ib.WorldBase(tri, col, horizon=2.)
N, S = a.omm, a.samples
pitch, yaw = ib.fibonacci_lattice(N)
rho = np.deg2rad(4.)
rp = np.clip(pitch[:, None] + rho / 2 * rng.randn(N, S), -1.5, 1.5).ravel()
ry = ((yaw[:, None] + rho / 2 * rng.randn(N, S) + np.pi) % (2 * np.pi) - np.pi).ravel()
r = ib.Renderer(world, ib.EyeLayout.from_angles(rp, ry, samples=S), sky=ib.Sky(np.deg2rad(60), np.pi))
B = a.views
pos = torch.as_tensor(np.c_[rng.uniform(2, side - 2, B), rng.uniform(2, side - 2, B), np.full(B, 0.01)], device="cuda")
yw = torch.as_tensor(rng.uniform(-np.pi, np.pi, (B, 1)), device="cuda")
out = {}
names = ("rgb", "Y", "dop", "aop")
for _ in range(2):
    r.render(pos, yw, outputs=names, out=out)
torch.cuda.synchronize(); r.check()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    r.render(pos, yw, outputs=names, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
pc = r.phase_cycles().astype(float)
print(json.dumps({"views": B, "rays_per_view": N * S, "tris": T, "ms": ms, "views_per_s": B / ms * 1e3,
                  "rays_per_s": B * N * S / ms * 1e3, "phases": (pc[:3] / pc[:3].sum()).tolist(),
                  "hit_frac": float((out["rgb"][:64, :, 1] > 0).float().mean())}))
