#!/usr/bin/env python
"""
bench.py -- rendered eye views / second (1000 ommatidia, Seville2009), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): the familiarity
heat-map grid -- Seville2009 world, 200 x 200 positions x 36 headings = 1.44 M views of a
1000-ommatidia full-sphere eye, fused with the analytic sky (sun at zenith distance 60 deg,
azimuth 180 deg as in examples/test_sky.py).  One *step* renders the whole grid; every view
writes rgb + hit + depth + Y + DoP + AoP (32 B per ommatidium) to HBM once.

  value : whole-job views/s, inputs (positions, headings) resident in HBM, outputs left in HBM.
  e2e   : views/s through Renderer.render_host (C-ABI isy_render_host) with pinned HOST buffers:
          H2D of the poses and D2H of rgb, Y, DoP, AoP inside the timed region, on a bounded slice.
  roofline : dominant kernel = the render kernel (one launch per step); bound = HBM (the only
          traffic the algorithm requires is writing each view once); an "issue" sub-object gives
          the brute-force-equivalent FP32 figure of SURVEY.md 8(d) for context.
  cpu_baseline : the oracle port of the reference renderer (oracle/isy_oracle.py, bit-identical to
          the reference) on a bounded sample of the same views on this box's host cores.

--impl reference times only that CPU path (all host cores) and prints the same line shape.
N > 1: one process per GPU, each rank renders its own 1/N-cell-shifted copy of the grid (weak
scaling: fixed work per GPU, no data-path collective); time = max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rendered eye views/sec (1000 ommatidia, Seville2009)"
UNIT = "views/s"
N_OMM = 1000
GRID_ROWS, GRID_COLS, GRID_ORIS = 200, 200, 36
SUN = (np.deg2rad(60.), np.pi)
OUT_BYTES_PER_RAY = 12 + 4 + 4 + 12          # rgb f32x3, hit i32, depth f32, Y/DoP/AoP f32
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def grid_workload(rank=0, world_size=1, rows=GRID_ROWS, cols=GRID_COLS, oris=GRID_ORIS):
    from invertsy_b200 import views
    pos, yaw_deg = views.grid_views(rows, cols, oris)
    if world_size > 1:   # weak scaling: rank r renders the grid shifted by r/world_size of a cell
        pos = pos.copy()
        pos[:, 0] += (10. / cols) * rank / world_size
    return pos, np.deg2rad(yaw_deg)


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle port of the reference (the ONLY place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
_cpu_state = {}


def _cpu_init():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import isy_oracle as o
    from invertsy_b200.env.world import Seville2009
    polygons, colours = Seville2009.load_world()
    pitch, yaw = o.fibonacci_eye(N_OMM)
    _cpu_state.update(o=o, polygons=polygons, colours=colours, omm=o.eye_rotations(pitch, yaw))


def _cpu_views(args):
    """renders a list of (x, y, z, yaw_deg) views with the reference's algorithm; returns seconds"""
    if not _cpu_state:
        _cpu_init()
    o, s = _cpu_state["o"], _cpu_state
    t0 = time.perf_counter()
    for x, y, z, yaw_deg in args:
        ori = o.view_rotations(yaw_deg, s["omm"])
        c = o.world_render(s["polygons"], s["colours"], 2.0, np.array([[x, y, z]]), ori)
        Y, P, A = o.sky_render(SUN[0], SUN[1], ori)
    return time.perf_counter() - t0, float(np.nansum(c)) + float(Y.sum())


def cpu_reference_rate(views_per_core, cores, seed=2021):
    """views/s of the CPU path on `cores` processes, each rendering `views_per_core` sampled views"""
    import multiprocessing as mp
    pos, yaw = grid_workload()
    rng = np.random.RandomState(seed)
    n = views_per_core * cores
    p_idx = rng.randint(0, pos.shape[0], n)
    h_idx = rng.randint(0, yaw.shape[1], n)
    views = [(pos[p, 0], pos[p, 1], pos[p, 2], float(np.rad2deg(yaw[p, h]))) for p, h in zip(p_idx, h_idx)]
    chunks = [views[i::cores] for i in range(cores)]
    if cores == 1:
        _cpu_views(chunks[0][:2])
        t0 = time.perf_counter()
        _cpu_views(chunks[0])
        wall = time.perf_counter() - t0
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores, initializer=_cpu_init) as pool:
            pool.map(_cpu_views, [c[:1] for c in chunks])     # warm-up: import, first-touch
            t0 = time.perf_counter()
            pool.map(_cpu_views, chunks)
            wall = time.perf_counter() - t0
    return n / wall, n


# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """samples nvidia-smi clocks and throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=GRID_ROWS)
    ap.add_argument("--cols", type=int, default=GRID_COLS)
    ap.add_argument("--oris", type=int, default=GRID_ORIS)
    ap.add_argument("--cpu-views-per-core", type=int, default=12)
    ap.add_argument("--e2e-positions", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--brute-force", action="store_true")
    ap.add_argument("--per-view-sky", action="store_true",
                    help="evaluate the sky for every view instead of once per (heading, ray) of the grid's shared headings")
    ap.add_argument("--sky", default="perez", choices=["perez", "uniform", "none"],
                    help="diagnostic only: the headline workload is 'perez'")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    config = {"workload": "familiarity heat-map grid (BASELINE.json configs[2]): Seville2009, %dx%d positions x %d "
                          "headings, %d-ommatidia full-sphere Fibonacci eye, Perez sky sun=(60deg,180deg), "
                          "horizon 2 m" % (args.rows, args.cols, args.oris, N_OMM),
              "views_per_step_per_gpu": args.rows * args.cols * args.oris,
              "outputs": "rgb,hit,depth,Y,dop,aop float32/int32 (32 B per ommatidium)",
              "l2": "outputs per step far exceed L2 (126 MB); nothing is re-read between steps",
              "parallelism": "positions sharded, %d rank(s), no data-path collective" % world_size}

    # ---------------- reference arm: the CPU path only ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        rates = []
        per_core = max(2, args.cpu_views_per_core // 2)
        for _ in range(args.warmup):
            cpu_reference_rate(1, cores)
        for _ in range(args.steps):
            r, n = cpu_reference_rate(per_core, cores)
            rates.append(r)
        v = float(np.mean(rates))
        sample = "%d steps x %d sampled grid views (%d per core, seed 2021), world+sky per view" % (args.steps, n, per_core)
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * n / v, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------- CPU baseline first (fork before CUDA is initialised) ----------------
    cpu_baseline = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        v, n = cpu_reference_rate(args.cpu_views_per_core, cores)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d grid views sampled with seed 2021 (%d per core), reference algorithm "
                                  "(oracle/isy_oracle.py world_render + sky_render)" % (n, args.cpu_views_per_core)}

    import torch
    import torch.distributed as dist
    import invertsy_b200 as ib
    from invertsy_b200 import _lib

    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    world = ib.Seville2009(device=local_rank)
    pitch, yaw_e = ib.fibonacci_lattice(N_OMM)
    eye = ib.EyeLayout.from_angles(pitch, yaw_e)
    sky = {"perez": ib.Sky(SUN[0], SUN[1]), "uniform": ib.UniformSky(10.), "none": None}[args.sky]
    if args.sky != "perez":
        config["workload"] += " [DIAGNOSTIC sky=%s]" % args.sky
    if args.sky == "perez":
        config["sky_evaluation"] = ("per view (--per-view-sky)" if args.per_view_sky else
                                    "once per (heading, ray) per launch, inside the timed region: every grid position "
                                    "has the same 36 headings and sun (checked on the device each launch), each view "
                                    "still gets its own copy written")
    r = ib.Renderer(world, eye, sky=sky, device=local_rank)

    pos, yaw = grid_workload(rank, world_size, args.rows, args.cols, args.oris)
    P, H = yaw.shape
    B = P * H
    dev = torch.device("cuda", local_rank)
    pos_d = torch.as_tensor(pos, device=dev)
    yaw_d = torch.as_tensor(yaw, device=dev)
    names = ("rgb", "hit", "depth", "Y", "dop", "aop") if sky is not None else ("rgb", "hit", "depth")
    out = {}

    def step():
        r.render(pos_d, yaw_d, outputs=names, out=out, brute_force=args.brute_force, per_view_sky=args.per_view_sky)

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)      # nvidia-smi needs ~0.3 s to start streaming: start it before the warm-up
    sampler.start()
    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    r.check()
    sampler.rows.clear()                    # keep only the samples of the timed region
    launches0 = _lib.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for a, b in evs:
        a.record()
        step()
        b.record()
    t_end.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    r.check()
    elapsed_ms = t_start.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    if world_size > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world_size * B * args.steps / (elapsed_ms / 1e3)

    pc = r.phase_cycles().astype(np.float64)
    phase = {"project_frac": pc[0] / pc[:3].sum(), "raster_frac": pc[1] / pc[:3].sum(), "shade_frac": pc[2] / pc[:3].sum(),
             "cycles_per_position": pc[:3].sum() / max(pc[3], 1)}
    # sanity on the last step's result (cheap, outside the timed region)
    hit_frac = float((out["hit"][: 36 * 64] >= 0).float().mean().item())

    # ---------------- e2e: host buffers through the C-ABI host entry point ----------------
    Pe = min(args.e2e_positions, P)
    sel = np.linspace(0, P - 1, Pe).astype(np.int64)
    e_names = ("rgb", "Y", "dop", "aop") if sky is not None else ("rgb",)
    pos_h = torch.as_tensor(pos[sel]).pin_memory()
    yaw_h = torch.as_tensor(yaw[sel]).pin_memory()
    host_out = {}
    for _ in range(2):
        r.render_host(pos_h.numpy(), yaw_h.numpy(), outputs=e_names, out=host_out, per_view_sky=args.per_view_sky)
    barrier()
    t0 = time.perf_counter()
    e_steps = max(2, min(args.steps, 5))
    for _ in range(e_steps):
        r.render_host(pos_h.numpy(), yaw_h.numpy(), outputs=e_names, out=host_out, per_view_sky=args.per_view_sky)
    torch.cuda.synchronize()
    e_s = time.perf_counter() - t0
    if world_size > 1:
        t = torch.tensor([e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_s = float(t.item())
    e2e_value = world_size * Pe * H * e_steps / e_s
    h2d = int(pos_h.numel() * 8 + yaw_h.numel() * 8 + 16)
    d2h = int(sum(host_out[n].numel() * host_out[n].element_size() for n in e_names))

    # ---------------- roofline of the dominant kernel (the render kernel, one launch per step) -----
    peak, peak_src = load_peaks()
    alg_bytes = B * N_OMM * OUT_BYTES_PER_RAY + P * 24 + B * 8
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    vis_mean = 517.0   # mean visible triangles per position (brute-force-equivalent count)
    alg_flop = B * (N_OMM * vis_mean * 12 + N_OMM * 100) + P * (5000 * 30 + vis_mean * 180)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "render kernel (1 launch/step)",
                "kernel_ms": kernel_ms,
                "issue": {"note": "SURVEY.md 8(d) brute-force-equivalent FP32 work; the rasteriser skips most of it, so this may exceed the peak",
                          "achieved_tflops": alg_flop / (kernel_ms / 1e3) / 1e12, "peak_tflops": FP32_PEAK_TFLOPS}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (edge filter, polarisation) + f64 (projection, exact re-check, luminance)", "data": "synthetic",
                "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "views_per_step": int(Pe * H), "api": "Renderer.render_host -> isy_render_host"},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "check": {"hit_fraction_first_rows": hit_frac}, "phases": phase}
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
