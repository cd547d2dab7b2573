"""
Device-resident timings of the five BASELINE.json configurations on one GPU (CUDA events, inputs and outputs in HBM):

  1  Seville2009 single view / the whole 812-step route (H = 1)             -> ray-grid rasteriser
  2  sky sweep: 360 x 90 suns for a 5000-direction dorsal eye               -> sky only, one sun per view
  3  familiarity heat-map grid 200 x 200 x 36 (the bench.py workload)        -> pitch-run rasteriser
  4  parallel route lines: 100 offsets x 1000 steps x 72 headings, rendered in chunks of --chunk positions
  5  synthetic stress: 500 k blades, 10 k ommatidia x 16 cone samples, a --views5 sample of the 1 M views

    python tools/bench_configs.py [--configs 1,2,3,4,5] > profiles/rNN_configs.json

Under torchrun (one process per GPU) configs 4 and 5 are sharded over the ranks -- config 4 by contiguous blocks of route
steps, config 5 by contiguous chunks of its 1 M views (--views5 is then the TOTAL) -- and timed as the
max over ranks between two barriers; rank 0 prints:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_configs.py --configs 4,5 --views5 1000000
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import invertsy_b200 as ib
from invertsy_b200 import views

ALL = ("rgb", "hit", "depth", "Y", "dop", "aop")
RANK, WORLD = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))


def sync_all():
    torch.cuda.synchronize()
    if WORLD > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def max_ms(ms):
    if WORLD == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def out_bytes(out):
    return int(sum(v.numel() * v.element_size() for k, v in out.items() if not k.startswith("_")))


def config1():
    world, sky = ib.Seville2009(), ib.Sky(np.deg2rad(60), np.pi)
    pitch, yaw = ib.fibonacci_lattice(1000)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), sky)
    route = ib.Seville2009.load_routes(degrees=True)["path"][0]
    pos = torch.as_tensor(route[:, :3].copy(), device="cuda")
    yw = torch.as_tensor(np.deg2rad(route[:, 3:4]), device="cuda")
    res = {}
    for label, sl in (("route_812_views", slice(None)), ("single_view", slice(0, 1))):
        out = {}
        ms = timed(lambda: r.render(pos[sl], yw[sl], outputs=ALL, out=out), 50, 5)
        B = pos[sl].shape[0]
        res[label] = {"ms": ms, "views_per_s": B / ms * 1e3}
    r.check()
    return res


def config2():
    pitch, yaw = ib.fibonacci_lattice(5000, hemisphere=True)
    empty = ib.WorldBase(np.zeros((0, 3, 3), dtype=np.float32), np.zeros((0, 3), dtype=np.float32))
    r = ib.Renderer(empty, ib.EyeLayout.from_angles(pitch, yaw), sky=ib.Sky(1.0, 0.0))
    az, el = np.deg2rad(np.arange(360.)), np.deg2rad(np.arange(90.) + .5)
    suns = np.array([(np.pi / 2 - e, (a + np.pi) % (2 * np.pi) - np.pi) for a in az for e in el])
    K = suns.shape[0]
    pos = torch.zeros((K, 3), dtype=torch.float64, device="cuda")
    yw = torch.zeros((K, 1), dtype=torch.float64, device="cuda")
    sun = torch.as_tensor(suns, device="cuda")
    out = {}
    ms = timed(lambda: r.render(pos, yw, sun=sun, outputs=("Y", "dop", "aop"), out=out), 10)
    r.check()
    nb = out_bytes(out)
    return {"suns": K, "directions": K * 5000, "ms": ms, "directions_per_s": K * 5000 / ms * 1e3,
            "views_per_s": K / ms * 1e3, "out_GB_per_s": nb / ms / 1e6}


def config3():
    world, sky = ib.Seville2009(), ib.Sky(np.deg2rad(60), np.pi)
    pitch, yaw = ib.fibonacci_lattice(1000)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), sky)
    pos, yaw_deg = views.grid_views(200, 200, 36)
    pos_t = torch.as_tensor(pos, device="cuda")
    yw = torch.as_tensor(np.deg2rad(yaw_deg), device="cuda")
    out = {}
    ms = timed(lambda: r.render(pos_t, yw, outputs=ALL, out=out), 10, 3)
    r.check()
    B = pos.shape[0] * 36
    return {"views": B, "ms": ms, "views_per_s": B / ms * 1e3, "out_GB_per_s": out_bytes(out) / ms / 1e6}


def config4(chunk):
    world, sky = ib.Seville2009(), ib.UniformSky(10.)
    pitch, yaw = ib.fibonacci_lattice(1000)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), sky)
    route = ib.Seville2009.load_routes(degrees=True)["path"][0]
    idx = np.linspace(0, route.shape[0] - 1, 1000)      # the route resampled to 1000 steps
    rt = np.stack([np.interp(idx, np.arange(route.shape[0]), route[:, k]) for k in range(3)], axis=1)
    hd = np.rad2deg(np.unwrap(np.deg2rad(route[:, 3])))
    rt = np.c_[rt, (np.interp(idx, np.arange(route.shape[0]), hd) + 180) % 360 - 180]
    pos, yaw_deg = views.parallel_views(rt, nb_orientations=72, nb_parallel=100, disposition_step=0.02)
    P_total = pos.shape[0]
    if WORLD > 1:     # blocks of route steps: positions are ordered (line k, step m); every rank takes its steps of every line
        a, b = views.shard_positions(1000, RANK, WORLD)
        keep = np.concatenate([np.arange(k * 1000 + a, k * 1000 + b) for k in range(100)])
        pos, yaw_deg = pos[keep], yaw_deg[keep]
    P = pos.shape[0]
    pos_t = torch.as_tensor(pos, device="cuda")
    yw = torch.as_tensor(np.deg2rad(yaw_deg), device="cuda")
    out = {}

    def run():
        for s in range(0, P, chunk):
            e = min(s + chunk, P)
            r.render(pos_t[s:e], yw[s:e], outputs=ALL, out=out if e - s == chunk else None)
    run()
    sync_all()
    ms = max_ms(timed(run, 2, 0))
    r.check()
    B = P_total * 72
    return {"views": B, "positions": P_total, "ranks": WORLD, "chunk_positions": chunk, "ms": ms, "views_per_s": B / ms * 1e3,
            "out_GB_per_s": B * 1000 * 32 / ms / 1e6}


def config5(nviews):
    rng = np.random.RandomState(2021)
    T, side, N, S = 500000, 100., 10000, 16
    base = np.c_[rng.uniform(0, side, T), rng.uniform(0, side, T), np.zeros(T)]
    d = rng.uniform(-.08, .08, (T, 2))
    tri = np.stack([base + np.c_[d, np.zeros(T)], base - np.c_[d, np.zeros(T)],
                    base + np.c_[rng.uniform(-.05, .05, (T, 2)), rng.lognormal(np.log(.18), .6, T)]],
                   axis=1).astype(np.float32)
    col = np.c_[np.zeros(T), rng.uniform(.07, .98, T), np.zeros(T)].astype(np.float32)
    world = ib.WorldBase(tri, col, horizon=2.)
    pitch, yaw = ib.fibonacci_lattice(N)
    rho = np.deg2rad(4.)
    rp = np.clip(pitch[:, None] + rho / 2 * rng.randn(N, S), -1.5, 1.5).ravel()
    ry = ((yaw[:, None] + rho / 2 * rng.randn(N, S) + np.pi) % (2 * np.pi) - np.pi).ravel()
    r = ib.Renderer(world, ib.EyeLayout.from_angles(rp, ry, samples=S), sky=ib.Sky(np.deg2rad(60), np.pi))
    B_total = nviews
    all_pos = np.c_[rng.uniform(2, side - 2, B_total), rng.uniform(2, side - 2, B_total), np.full(B_total, 0.01)]
    all_yaw = rng.uniform(-np.pi, np.pi, (B_total, 1))
    a, b = views.shard_positions(B_total, RANK, WORLD)
    chunk = 4096                                        # views per call: 4096 x 10 000 ommatidia x 24 B = 0.98 GB of outputs
    out = {}

    def run():
        for s0 in range(a, b, chunk):
            e0 = min(b, s0 + chunk)
            r.render(torch.as_tensor(all_pos[s0:e0], device="cuda"), torch.as_tensor(all_yaw[s0:e0], device="cuda"),
                     outputs=("rgb", "Y", "dop", "aop"), out=out if e0 - s0 == chunk else None)
    B = B_total
    if b - a <= chunk:
        pos = torch.as_tensor(all_pos[a:b], device="cuda")
        yw = torch.as_tensor(all_yaw[a:b], device="cuda")
        run = lambda: r.render(pos, yw, outputs=("rgb", "Y", "dop", "aop"), out=out)   # noqa: E731
    run()
    sync_all()
    ms = max_ms(timed(run, 3 if B_total <= 65536 else 1, 0))
    r.check()
    pc = r.phase_cycles().astype(float)
    print("config5 cycles per position: project %.0f accel %.0f final %.0f" % tuple(pc[:3] / max(pc[3], 1)), file=sys.stderr)
    return {"views_timed": B, "ranks": WORLD, "rays_per_view": N * S, "triangles": T, "ms": ms, "views_per_s": B / ms * 1e3,
            "rays_per_s": B * N * S / ms * 1e3, "seconds_for_1M_views_one_gpu": 1e6 / (B / ms * 1e3)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3,4,5")
    ap.add_argument("--chunk", type=int, default=20000)
    ap.add_argument("--views5", type=int, default=2048)
    a = ap.parse_args()
    want = set(a.configs.split(","))
    torch.cuda.set_device(LOCAL)
    if WORLD > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
    res = {"device": torch.cuda.get_device_name(0), "ranks": WORLD}
    if "1" in want:
        res["config1_route"] = config1()
    if "2" in want:
        res["config2_sky_sweep"] = config2()
    if "3" in want:
        res["config3_heatmap_grid"] = config3()
    if "4" in want:
        res["config4_parallel_lines"] = config4(a.chunk)
    if "5" in want:
        res["config5_synthetic_stress"] = config5(a.views5)
    if RANK == 0:
        print(json.dumps(res, indent=1))
    if WORLD > 1:
        dist.destroy_process_group()
