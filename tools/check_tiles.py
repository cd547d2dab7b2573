"""Tile walk against the column walk (ISY_FLAG_NO_TILES) on one scan shape: python tools/check_tiles.py RAYS HEADINGS POSITIONS"""
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import invertsy_b200 as ib
R = int(sys.argv[1]); H = int(sys.argv[2]); P = int(sys.argv[3])
world = ib.Seville2009()
pitch, yaw = ib.fibonacci_lattice(R)
r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), ib.Sky(np.deg2rad(55.), 2.5))
rng = np.random.RandomState(21)
pos = np.c_[rng.uniform(1, 9, P), rng.uniform(1, 9, P), np.full(P, 0.01)]
heads = np.deg2rad(np.arange(H) * 360. / H - 180.)
yw = np.tile(heads, (P, 1))
a = r.render(pos, yw, outputs=("hit",))["hit"].cpu().numpy()
b = r.render(pos, yw, outputs=("hit",), no_tiles=True)["hit"].cpu().numpy()
torch.cuda.synchronize(); r.check()
d = a != b
print("R", R, "H", H, "P", P, "mismatching rays", int(d.sum()), "of", d.size, "views with mismatch", int(d.any(axis=1).sum()))
if d.any():
    v, ray = np.nonzero(d)
    print("first views", v[:10], "rays", ray[:10], "tile", a[d][:10], "cols", b[d][:10])
    print("per heading", np.bincount(v % H, minlength=H), "positions", np.unique(v // H)[:20])
    print("tile-path missing (−1 where cols hit):", int(((a == -1) & (b >= 0)).sum()), " extra:", int(((a >= 0) & (b == -1)).sum()), "different tri:", int(((a >= 0) & (b >= 0) & d).sum()))
