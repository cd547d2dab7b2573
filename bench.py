#!/usr/bin/env python
"""
bench.py -- rendered eye views / second (1000 ommatidia, Seville2009), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): the familiarity heat-map grid --
Seville2009 world, 200 x 200 positions x 36 headings = 1.44 M views of a 1000-ommatidia full-sphere eye, fused with
the analytic sky (sun at zenith distance 60 deg, azimuth 180 deg as in examples/test_sky.py).

One *step* is the whole job the reference's heat-map scripts do with those views (sim/simulation.py:2675-2722,
:3061-3092), on N GPUs (STRONG scaling: the one grid is split by position):
  1. render  : rank r renders every N-th position (views.shard_positions_interleaved: all ranks see the same mix of
               dense and sparse vegetation); every view writes rgb + hit + depth +
               Y + DoP + AoP + the eye's PN response (36 B per ommatidium) to HBM once;
  2. evaluate: the PN responses of the block go through the perfect memory of the 812 route views
               (tcgen05 TF32 filter + exact fp32 re-check, csrc/isy_fam.cuh) -> one familiarity per view;
  3. gather  : NCCL all_gather of the familiarity rows (1.44 M floats in total) + an on-device transpose into view
               order, so that every rank holds the map.
All three are inside the timed region; time = max over ranks (CUDA events + barriers).

  value : whole-job views/s (poses resident in HBM, results left in HBM).
  e2e   : views/s with
          HOST buffers: the same step through the package's public call `familiarity.HeatMap.__call__` on the WHOLE
          grid -- H2D of the poses from pinned memory, the step, D2H of its result (the familiarity map) into pinned
          memory, all inside the timed region.  `e2e_responses` (poses in, the PN responses of every view out through
          isy_render_host: 4 KB per view, what the data-collection simulations log, simulation.py:2721 + agent.py:744)
          and `e2e_full_views` (rgb, Y, DoP, AoP out: 24 KB per view, bounded slice) are reported next to it.
  roofline : dominant kernel = the render kernel; bound = HBM (the only traffic the algorithm requires is writing
          each view once); `roofline_familiarity` gives the tensor-core figure of the filter kernel.
  cpu_baseline : the UNMODIFIED reference (oracle/_ref, shipped by build()) -- or its bit-identical NumPy port when
          that directory is absent -- on a bounded sample of the same views on this box's host cores.
  weak / render_only / per_view_sky : extra keys (whole grid per rank, render kernel alone; sky evaluated per view).

--impl reference times only the CPU path (all host cores) and prints the same line shape.
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
OUT_BYTES_PER_RAY = 12 + 4 + 4 + 12 + 4      # rgb f32x3, hit i32, depth f32, Y/DoP/AoP f32, PN response f32
SATURATION = 1.                               # omm_res: no saturation, so that all 812 route views stay distinct (the reference scripts use 5-10 with
                                              # InvertPy's own, unpinned transform; with ours that clips most ommatidia to 1 and leaves 203 distinct views)
C_SENSITIVE = (0., 0., 1., 0., 0.)            # green-only eye (agent.py:589-590)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1650.0}, "fallback (B200_PROFILING.md)"


def grid_workload(rows=GRID_ROWS, cols=GRID_COLS, oris=GRID_ORIS):
    from invertsy_b200 import views
    pos, yaw_deg = views.grid_views(rows, cols, oris)
    return pos, np.deg2rad(yaw_deg)


def base_config(args):
    """the same dictionary in both arms (the driver compares them)"""
    return {"workload": "familiarity heat-map grid (BASELINE.json configs[2]): Seville2009, %dx%d positions x %d "
                        "headings, %d-ommatidia full-sphere Fibonacci eye, Perez sky sun=(60deg,180deg), "
                        "horizon 2 m" % (args.rows, args.cols, args.oris, N_OMM),
            "views_per_step": args.rows * args.cols * args.oris,
            "outputs": "rgb,hit,depth,Y,dop,aop,PN response float32/int32 (36 B per ommatidium) + familiarity per view",
            "l2": "outputs per step far exceed L2 (126 MB); nothing is re-read between steps",
            "parallelism": "strong scaling: positions of the ONE grid dealt round-robin over the ranks, familiarity rows "
                           "all-gathered (NCCL) and put into view order inside the timed region"}


# ------------------------------------------------------------------------------------------------
# CPU leg: the unmodified reference from oracle/_ref, else its NumPy port (the ONLY place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
_cpu_state = {}


def _cpu_kind():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import make_ref
    return "reference" if make_ref.ref_available() else "port"


def _cpu_init():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import isy_oracle as o
    import make_ref
    pitch, yaw = o.fibonacci_eye(N_OMM)
    _cpu_state.update(o=o, omm=o.eye_rotations(pitch, yaw))
    if make_ref.ref_available():
        import ref_loader
        world_mod, sky_mod = ref_loader.load_reference(make_ref.REF_DST)
        _cpu_state.update(kind="reference", world=world_mod.Seville2009(), sky=sky_mod.Sky(SUN[0], SUN[1]))
    else:
        from invertsy_b200.env.world import Seville2009
        polygons, colours = Seville2009.load_world()
        _cpu_state.update(kind="port", polygons=polygons, colours=colours)


def _cpu_views(args):
    """renders a list of (x, y, z, yaw_deg) views with the reference's own code (or its port); returns seconds"""
    if not _cpu_state:
        _cpu_init()
    o, s = _cpu_state["o"], _cpu_state
    t0 = time.perf_counter()
    c = Y = None
    for x, y, z, yaw_deg in args:
        ori = o.view_rotations(yaw_deg, s["omm"])
        if s["kind"] == "reference":
            c = s["world"](np.array([[x, y, z]]), ori=ori)                 # env/world.py:55
            Y, P, A = s["sky"](ori)                                         # env/sky.py:269
        else:
            c = o.world_render(s["polygons"], s["colours"], 2.0, np.array([[x, y, z]]), ori)
            Y, P, A = o.sky_render(SUN[0], SUN[1], ori)
    return time.perf_counter() - t0, float(np.nansum(c)) + float(np.sum(Y))


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
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the weak / per-view-sky / secondary e2e figures")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    config = base_config(args)

    # ---------------- reference arm: the CPU path only ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        rates = []
        per_core = max(2, args.cpu_views_per_core // 2)
        for _ in range(args.warmup):
            cpu_reference_rate(1, cores)
        n = 0
        for _ in range(args.steps):
            r, n = cpu_reference_rate(per_core, cores)
            rates.append(r)
        v = float(np.mean(rates))
        kind = _cpu_kind()
        sample = ("%d steps x %d sampled grid views (%d per core, seed 2021), world() + sky() per view through %s"
                  % (args.steps, n, per_core, "the unmodified reference modules (oracle/_ref)" if kind == "reference"
                     else "the NumPy port (oracle/isy_oracle.py)"))
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * n / v, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "ms_per_step is per sample of %d views, not per grid; the familiarity evaluation of the GPU arm's "
                        "step is not part of this arm (the reference's memory model is InvertPy's, not vendored)" % n}
        print(json.dumps(line))
        return 0

    # ---------------- CPU baseline first (fork before CUDA is initialised) ----------------
    cpu_baseline = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        v, n = cpu_reference_rate(args.cpu_views_per_core, cores)
        kind = _cpu_kind()
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                        "sample": "%d grid views sampled with seed 2021 (%d per core), world() + sky() per view through %s"
                                  % (n, args.cpu_views_per_core, "the unmodified reference modules (oracle/_ref)"
                                     if kind == "reference" else "the NumPy port (oracle/isy_oracle.py)")}

    import torch
    import torch.distributed as dist
    import invertsy_b200 as ib
    from invertsy_b200 import _lib, distributed as isy_dist, familiarity as fam, views

    torch.cuda.set_device(local_rank)
    numa = isy_dist.bind_to_gpu_numa(local_rank)     # before any pinned allocation
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    world = ib.Seville2009(device=local_rank)
    pitch, yaw_e = ib.fibonacci_lattice(N_OMM)
    eye = ib.EyeLayout.from_angles(pitch, yaw_e)
    sky = ib.Sky(SUN[0], SUN[1])
    r = ib.Renderer(world, eye, sky=sky, device=local_rank)
    sensor = _lib.SensorParams.make(C_SENSITIVE[1:4], SATURATION, 1)

    # the memory of the training route (Ant1_Route1, 812 views), set up once outside the timed region
    route = ib.Seville2009.load_routes(degrees=True)["path"][0]
    stored = r.render(route[:, :3], np.deg2rad(route[:, 3:4]), outputs=("resp",), sensor=sensor)["resp"]
    memory = fam.PerfectMemory(device=local_rank).store(stored)

    pos, yaw = grid_workload(args.rows, args.cols, args.oris)
    P, H = yaw.shape
    B = P * H
    names = ("rgb", "hit", "depth", "Y", "dop", "aop", "resp")
    # the job as the package's public API runs it: positions dealt round-robin over the ranks (every rank sees the same mix
    # of dense and sparse vegetation), render -> familiarity -> all_gather + transpose into view order
    hm = fam.HeatMap(r, memory, sensor, P, H, outputs=names)
    mine = hm.positions()
    Pl, Bl = len(mine), len(mine) * H
    pos_d = torch.as_tensor(pos[mine], device=dev)
    yaw_d = torch.as_tensor(yaw[mine], device=dev)
    out = hm.views

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step(marks=None):
        if marks:
            marks[0].record()
        r.render(pos_d, yaw_d, outputs=names, out=out, sensor=sensor)
        if marks:
            marks[1].record()
        memory.familiarity(out["resp"], out=hm.local[:Bl])
        if marks:
            marks[2].record()
        isy_dist.all_gather_interleaved(hm.local, P, H, out=hm.full, scratch=hm.scratch)
        if marks:
            marks[3].record()

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world_size == 1:
            return float(x)
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)      # nvidia-smi needs ~0.3 s to start streaming: start it before the warm-up
    sampler.start()
    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    r.check()
    sampler.rows.clear()                    # keep only the samples of the timed region
    launches0 = _lib.launch_count()
    marks = [[ev() for _ in range(4)] for _ in range(args.steps)]
    barrier()
    t_start, t_end = ev(), ev()
    t_start.record()
    for m in marks:
        step(m)
    t_end.record()
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop()
    r.check()
    elapsed_ms = max_over_ranks(t_start.elapsed_time(t_end))
    render_ms = max_over_ranks(float(np.mean([m[0].elapsed_time(m[1]) for m in marks])))
    fam_ms = max_over_ranks(float(np.mean([m[1].elapsed_time(m[2]) for m in marks])))
    gather_ms = max_over_ranks(float(np.mean([m[2].elapsed_time(m[3]) for m in marks])))
    value = B * args.steps / (elapsed_ms / 1e3)

    pc = r.phase_cycles().astype(np.float64)
    phase = {"project_frac": pc[0] / pc[:3].sum(), "raster_frac": pc[1] / pc[:3].sum(), "shade_frac": pc[2] / pc[:3].sum(),
             "cycles_per_position": pc[:3].sum() / max(pc[3], 1)}
    # sanity on the last step's result (cheap, outside the timed region)
    hit_frac = float((out["hit"][: 36 * 64] >= 0).float().mean().item())
    fam_local = hm.local[:Bl]
    unresolved = memory.unresolved()
    check = {"hit_fraction_first_rows": hit_frac, "familiarity_min": float(fam_local.min().item()),
             "familiarity_max": float(fam_local.max().item()), "familiarity_unresolved_queries": unresolved}
    if world_size > 1:      # every rank holds the whole map after the gather: this rank's rows are where they belong
        got = hm.full[:B].view(P, H)[torch.as_tensor(mine, device=dev)].reshape(-1)
        check["gathered_rows_match"] = bool(torch.equal(got, fam_local))

    # ---------------- extras: render kernel alone on the whole grid per rank (weak), sky per view ------------
    extras = {}
    if not args.no_extras:
        pos_w = pos.copy()
        pos_w[:, 0] += (10. / args.cols) * rank / world_size      # rank r: the grid shifted by r/N of a cell
        pos_wd, yaw_wd = torch.as_tensor(pos_w, device=dev), torch.as_tensor(yaw, device=dev)
        wout = {}
        six = names[:6]

        def timed(fn, k):
            for _ in range(2):
                fn()
            barrier()
            s, e = ev(), ev()
            s.record()
            for _ in range(k):
                fn()
            e.record()
            barrier()
            return max_over_ranks(s.elapsed_time(e)) / k

        k = max(3, min(args.steps, 10))
        ms = timed(lambda: r.render(pos_wd, yaw_wd, outputs=six, out=wout), k)
        extras["weak"] = {"what": "round-1 workload: every rank renders its own copy of the WHOLE grid (6 outputs, 32 B per "
                                  "ommatidium), render kernel only, no collective", "ms_per_step": ms,
                          "value": world_size * B / (ms / 1e3), "unit": UNIT,
                          "hbm_frac": (B * N_OMM * 32 / (ms / 1e3) / 1e9) / load_peaks()[0]["hbm_gbs"]}
        if world_size == 1:
            ms = timed(lambda: r.render(pos_wd, yaw_wd, outputs=six, out=wout, per_view_sky=True), k)
            extras["per_view_sky"] = {"what": "the same with the sky evaluated for every view (ISY_FLAG_PER_VIEW_SKY) instead "
                                              "of once per (heading, ray) of the grid's shared headings", "ms_per_step": ms,
                                      "value": B / (ms / 1e3), "unit": UNIT,
                                      "hbm_frac": (B * N_OMM * 32 / (ms / 1e3) / 1e9) / load_peaks()[0]["hbm_gbs"]}
        del wout, pos_wd, yaw_wd
        torch.cuda.empty_cache()

    # ---------------- e2e: the same step through the public API with HOST buffers ----------------
    # poses of this rank's positions in pinned host memory -> H2D; render (all 7 outputs) + familiarity + all_gather +
    # transpose; the familiarity map (the step's result) -> D2H into pinned host memory, on every rank
    pos_h = torch.as_tensor(pos[mine]).pin_memory()
    yaw_h = torch.as_tensor(yaw[mine]).pin_memory()
    hm(pos_h, yaw_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        fam_host = hm(pos_h, yaw_h)
    e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = int(P * 24 + B * 8)
    d2h = int(B * 4 * world_size)
    e2e = {"value": B * args.e2e_steps / e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "views_per_step": int(B), "api": "familiarity.HeatMap.__call__: host poses in -> Renderer.render (7 outputs) -> "
           "PerfectMemory.familiarity -> all_gather -> familiarity map out (pinned host, every rank)"}
    check["e2e_matches_device"] = bool(torch.equal(fam_host[:B], (hm.full if world_size > 1 else hm.local)[:B].cpu()))

    if not args.no_extras:
        # the data-collection path: PN responses (4 KB per view) to the host through isy_render_host, whole grid
        host_out = {"resp": torch.empty((Bl, N_OMM), dtype=torch.float32, pin_memory=True)}

        def resp_step():
            r.render_host(pos_h.numpy(), yaw_h.numpy(), outputs=("resp",), out=host_out, sensor=sensor)

        resp_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            resp_step()
        torch.cuda.synchronize()
        e_r = max_over_ranks(time.perf_counter() - t0)
        extras["e2e_responses"] = {"value": B * args.e2e_steps / e_r, "unit": UNIT, "h2d_bytes_per_step": h2d,
                                   "d2h_bytes_per_step": int(B * N_OMM * 4), "pcie_GBps_per_gpu": Bl * N_OMM * 4 * args.e2e_steps / e_r / 1e9,
                                   "what": "Renderer.render_host(outputs=('resp',)) -> isy_render_host: poses in, the PN responses of every "
                                           "view out (4 KB per view, what the data-collection simulations log); the box exposes one NUMA "
                                           "node (%s), the ranks share its host-memory path" % numa}
        check["e2e_responses_match_device"] = bool(torch.equal(host_out["resp"][:1000], out["resp"][:1000].cpu()))
        del host_out
        Pe = min(2048, Pl)
        full_names = ("rgb", "Y", "dop", "aop")
        full_out = {}
        r.render_host(pos[mine[:Pe]], yaw[mine[:Pe]], outputs=full_names, out=full_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            r.render_host(pos[mine[:Pe]], yaw[mine[:Pe]], outputs=full_names, out=full_out)
        e_v = max_over_ranks(time.perf_counter() - t0)
        extras["e2e_full_views"] = {"value": world_size * Pe * H * args.e2e_steps / e_v, "unit": UNIT,
                                    "d2h_bytes_per_step": int(world_size * Pe * H * N_OMM * 24),
                                    "what": "round-1 e2e: rgb, Y, DoP, AoP of a slice of %d positions per rank (24 KB per view)" % Pe}

    # ---------------- roofline of the dominant kernel (the render kernel, one launch per step) -----
    peaks, peak_src = load_peaks()
    peak = float(peaks["hbm_gbs"])
    alg_bytes = Bl * N_OMM * OUT_BYTES_PER_RAY + Pl * 24 + Bl * 8
    achieved = alg_bytes / (render_ms / 1e3) / 1e9
    vis_mean = 517.0   # mean visible triangles per position (brute-force-equivalent count)
    alg_flop = Bl * (N_OMM * vis_mean * 12 + N_OMM * 100) + Pl * (5000 * 30 + vis_mean * 180)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "render_kernel<false,kEyeSmall> (1 launch/step, "
                "timed with its 7 set-up launches)", "kernel_ms": render_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "issue": {"note": "SURVEY.md 8(d) brute-force-equivalent FP32 work; the rasteriser skips most of it, so this may exceed the peak",
                          "achieved_tflops": alg_flop / (render_ms / 1e3) / 1e12, "peak_tflops": FP32_PEAK_TFLOPS}}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and world_size == 1:
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
    tf32_peak = float(peaks.get("bf16_tflops", 1650.0)) / 2
    fam_flop = 2.0 * Bl * stored.shape[0] * N_OMM
    roofline_fam = {"bound": "tensor", "achieved": fam_flop / (fam_ms / 1e3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": fam_flop / (fam_ms / 1e3) / 1e12 / tf32_peak, "kernel": "fam_filter_kernel + fam_recheck_kernel "
                    "(2 launches/step, timed together)", "kernel_ms": fam_ms,
                    "peak_source": "TF32 = half the measured dense bf16 rate (%s)" % peak_src}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (edge filter, polarisation, responses) + f64 (projection, exact re-check, luminance) + tf32 (familiarity filter)",
                "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(launches), "roofline": roofline, "roofline_familiarity": roofline_fam,
                "cpu_baseline": cpu_baseline,
                "step_ms": {"render": render_ms, "familiarity": fam_ms, "all_gather": gather_ms,
                            "note": "max over ranks of the mean per step, CUDA events between the phases"},
                "sky_evaluation": "once per (heading, ray) per launch, inside the timed region: every grid position has the "
                                  "same 36 headings and sun (checked on the device each launch), each view still gets its own "
                                  "copy written; see per_view_sky for the other mode",
                "check": check, "phases": phase}
        line.update(extras)
        print(json.dumps(line))
    if world_size > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
