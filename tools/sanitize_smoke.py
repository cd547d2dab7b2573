"""Small renders that touch every rasteriser / final-pass variant: for compute-sanitizer (memcheck, racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import invertsy_b200 as ib

rng = np.random.RandomState(3)
world = ib.Seville2009()
sky = ib.Sky(np.deg2rad(60), np.pi)
ALL = ("rgb", "hit", "depth", "Y", "dop", "aop")


def run(label, eye, pos, yaw, **kw):
    r = ib.Renderer(world, eye, kw.pop("sky", sky))
    out = r.render(pos, yaw, outputs=kw.pop("outputs", ALL), **kw)
    torch.cuda.synchronize()
    r.check()
    print(label, {k: tuple(v.shape) for k, v in out.items() if not k.startswith("_")}, flush=True)


def positions(n):
    return np.c_[rng.uniform(1, 9, n), rng.uniform(1, 9, n), np.full(n, 0.01)]


P = int(os.environ.get("ISY_SAN_P", "150"))
heads6 = np.tile(np.deg2rad(np.arange(6) * 60. - 180.), (P, 1))
p, y = ib.fibonacci_lattice(400)
small = ib.EyeLayout.from_angles(p, y)
run("small eye, scan (pitch-run walk)", small, positions(P), heads6)
run("small eye, single heading (ray grid)", small, positions(P), heads6[:, :1])
run("small eye, irregular headings", small, positions(P), np.sort(rng.uniform(-3.1, 3.1, (P, 7)), axis=1))
run("small eye, uniform sky, from sky", small, positions(P), heads6, sky=ib.UniformSky(10.), init_from_sky=True)
# rows mode (template rows + patches) starts at 4096 views: shared-heading Perez sky, uniform sky, no sky, and a
# Perez call whose positions have their own headings (falls back to the per-view passes inside the same launch shape)
PR = int(os.environ.get("ISY_SAN_PROWS", "700"))
headsR = np.tile(np.deg2rad(np.arange(6) * 60. - 180.), (PR, 1))
run("small eye, rows mode, sky table", small, positions(PR), headsR)
run("small eye, rows mode, uniform sky", small, positions(PR), headsR, sky=ib.UniformSky(10.))
run("small eye, rows mode, no sky", small, positions(PR), headsR, sky=None, outputs=("rgb", "hit", "depth"))
run("small eye, rows mode set up, own headings", small, positions(PR), headsR + rng.uniform(-.1, .1, headsR.shape))
run("small eye, rows mode, single heading (ray grid)", small, positions(4200), np.zeros((4200, 1)))
run("small eye, brute force", small, positions(8), heads6[:8], brute_force=True)
run("small eye, f64", small, positions(8), heads6[:8], f64=True)
p, y = ib.fibonacci_lattice(1500)
large = ib.EyeLayout.from_angles(p, y)
run("large eye, scan (bands)", large, positions(P), heads6)
run("large eye, single heading (ray grid in HBM)", large, positions(P), heads6[:, :1])
run("large eye, sun per view", large, positions(P), heads6,
    sun=np.tile(np.array([[1.0, 0.5]]), (P * 6, 1)) + rng.uniform(0, .3, (P * 6, 2)))
N, S = 120, 16
p, y = ib.fibonacci_lattice(N)
rp = np.clip(p[:, None] + 0.04 * rng.randn(N, S), -1.5, 1.5).ravel()
ry = ((y[:, None] + 0.04 * rng.randn(N, S) + np.pi) % (2 * np.pi) - np.pi).ravel()
cone = ib.EyeLayout.from_angles(rp, ry, samples=S)
run("cone eye, scan (bands, shuffle averages)", cone, positions(P), heads6)
run("cone eye, single heading", cone, positions(P), heads6[:, :1])
run("cone eye, from sky", cone, positions(P), heads6, init_from_sky=True, outputs=("rgb",))
from invertsy_b200.render import angles_to_quat
q = angles_to_quat(rng.uniform(-1.5, 1.5, 300), rng.uniform(-3.1, 3.1, 300))
anyeye = ib.EyeLayout(q)
from scipy.spatial.transform import Rotation as R
vq = R.from_euler('ZYX', np.c_[rng.uniform(-180, 180, 12), rng.uniform(-10, 10, 12), rng.uniform(-10, 10, 12)],
                  degrees=True).as_quat().reshape(6, 2, 4)
run("any orientation (bins)", anyeye, positions(6), None, quat=vq)
print("done")
