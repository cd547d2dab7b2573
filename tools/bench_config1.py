"""Config 1 (BASELINE.json configs[0]): route views, H = 1 -- single-view latency and whole-route throughput."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import invertsy_b200 as ib
from scipy.spatial.transform import Rotation as R

world = ib.Seville2009()
sky = ib.Sky(np.deg2rad(60), np.pi)
pitch, yaw = ib.fibonacci_lattice(1000)
eye = ib.EyeLayout.from_angles(pitch, yaw)
r = ib.Renderer(world, eye, sky)
route = ib.Seville2009.load_routes(degrees=True)["path"][0]
pos = torch.as_tensor(route[:, :3].copy(), device="cuda")
yw = torch.as_tensor(np.deg2rad(route[:, 3:4]), device="cuda")
names = ("rgb", "hit", "depth", "Y", "dop", "aop")
out = {}
res = {}
for label, sl in (("route_812", slice(None)), ("single_view", slice(0, 1))):
    for _ in range(5):
        r.render(pos[sl], yw[sl], outputs=names, out=out if sl == slice(None) else None)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    a.record()
    for _ in range(n):
        r.render(pos[sl], yw[sl], outputs=names, out=out if sl == slice(None) else None)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    B = pos[sl].shape[0]
    res[label] = {"ms": ms, "views_per_s": B / ms * 1e3, "phases": [float(x) for x in r.phase_cycles()]}
# drop-in callables, one view at a time (what a closed-loop simulation does)
omm = R.from_euler('ZY', np.c_[yaw, pitch])
t0 = time.perf_counter()
n = 100
for k in range(n):
    ori = R.from_euler('Z', route[k, 3], degrees=True) * omm
    y, p, a = sky(ori)
    c = world(route[k:k + 1, :3], ori, init_c=y)
res["dropin_world_plus_sky_call"] = {"ms": (time.perf_counter() - t0) / n * 1e3}
t0 = time.perf_counter()
for k in range(n):
    h = r.render_host(route[k:k + 1, :3], np.deg2rad(route[k:k + 1, 3:4]), outputs=("rgb", "Y", "dop", "aop"))
res["render_host_single_view"] = {"ms": (time.perf_counter() - t0) / n * 1e3}
# the drop-in eye of a closed-loop simulation step (simulation.py:836-838): eye.xyz / eye.ori set, eye(sky=, scene=) called
ce = ib.CompoundEye(nb_input=1000, omm_res=5., c_sensitive=[0, 0., 1., 0., 0.])
for label, make in (("eye_call_yaw_only", lambda k: R.from_euler('Z', route[k, 3], degrees=True)),
                    ("eye_call_general_orientation", lambda k: R.from_euler('ZYX', [route[k, 3], 4., -2.], degrees=True))):
    for k in range(5):
        ce.xyz, ce.ori = route[k, :3], make(k)
        ce(sky=sky, scene=world)
    t0 = time.perf_counter()
    for k in range(n):
        ce.xyz, ce.ori = route[k, :3], make(k)
        resp = ce(sky=sky, scene=world)
    res[label] = {"ms": (time.perf_counter() - t0) / n * 1e3, "what": "CompoundEye.__call__ -> isy_render_host(resp), host to host"}
print(json.dumps(res))
