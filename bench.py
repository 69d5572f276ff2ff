#!/usr/bin/env python
"""bench.py — Gcell-updates/s of the world tick (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S] [--workload mixed|column|sparse]

A "step" is one world::tick() (cell_iter = 3 automaton iterations over the whole tickZone) on a synthetic world.
Default workload at N = 1 = BASELINE.json configs[1]: 8192x8192 mixed powders/liquids/gases with fire and the registered
interacting materials (worldgen.bench_table).  N > 1 runs configs[2]: the SAME 32768 x 33024 world (tickZone 32512 x 32768)
cut into N horizontal strips, one rank per GPU, halo rows over NCCL ("strong" scaling); --scaling weak keeps a fixed strip
height per GPU instead.  --workload bodies is configs[3] (2000 rigid bodies rastered, erased and outlined every tick).

The device-timed `value` is world::tick alone (BASELINE.json metric).  `e2e` is one whole game tick through the C ABI with host
buffers, in the reference's order (game.cpp:1711-2159): body raster, chunk merges from pinned host memory, world::tick,
tickCells, body erase, tickTemperature on tick % 4 == 2, dirty -> texture planes with the movingTiles histogram read back,
dirty clear.

One JSON line is printed by rank 0; see DESIGN.md §6 for every key.
"""
import argparse
import ctypes as C
import functools
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CELL_ITER = 3
ALGO_BYTES_PER_CELL_UPDATE = 36  # 18 B state read + 18 B written per cell per iteration
KERNEL_OF = {"default": "one colour phase = fse::tick_pass_kernel<1> + tick_pass_kernel<2> + tick_pass3_kernel", "rows_fused": "fse::tick_rows_kernel"}
KERNEL_OF["rows"] = KERNEL_OF["default"]
METRIC = "Gcell-updates/sec (device-timed) at 1/2/4/8 B200; % HBM roofline"
UNIT = "Gcell-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=0, help="world width; 0 = 8192 on one GPU, 32768 on several (strong scaling)")
    ap.add_argument("--height", type=int, default=0, help="world height (per-GPU strip height with --scaling weak); 0 = width (+256 rows for the 32768 world)")
    ap.add_argument("--workload", default="mixed", choices=["mixed", "column", "sparse", "air", "bodies"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="--gpus N > 1: one 32768 x 33024 world cut N ways, or a fixed strip per GPU")
    ap.add_argument("--bodies", type=int, default=2000, help="rigid bodies of --workload bodies")
    ap.add_argument("--seed", type=int, default=1337)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--schedule", default="default", choices=["default", "rows", "rows_fused"], help="in-row schedule of the tick kernel")
    ap.add_argument("--active", type=int, default=-1, help="active-chunk tracking: 1 on, 0 off, -1 = on for --workload sparse")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def band_fn(args, table, extra):
    from falling_sand_engine_b200 import worldgen as G

    if args.workload == "mixed":
        return functools.partial(G.mixed_band, table, seed=args.seed, extra=list(extra.values()))
    if args.workload == "bodies":  # generator 2 at 60 % AIR
        return functools.partial(G.mixed_band, table, seed=args.seed, extra=list(extra.values()), air_frac=0.6)
    if args.workload == "column":
        return functools.partial(G.column_drop_band, table, seed=args.seed)
    if args.workload == "air":
        return functools.partial(G.air_band, table, seed=args.seed)
    return functools.partial(G.sparse_band, table, seed=args.seed)


def make_table():
    """Stock materials + the three registered interacting powders of config 2 (the 'Lua material table')."""
    from falling_sand_engine_b200 import materials as M
    from falling_sand_engine_b200 import worldgen as G

    return G.bench_table(M.default_materials(1337))


def workload_name(args, W, H, n):
    mixed = ("8192x8192 mixed powders/liquids/gases with fire and Lua-table reactions (BASELINE configs[1])" if (W, H) == (8192, 8192)
             else "32768x32768 mixed-material world strip-partitioned with NVLink halo exchange (BASELINE configs[2]; tickZone 32512x32768, generator of configs[1])"
             if W == 32768 else "mixed powders/liquids/gases with fire and Lua-table reactions, generator of BASELINE configs[1]")
    base = {"mixed": mixed,
            "column": "sand/water/stone column drop (BASELINE configs[0])",
            "sparse": "mostly-settled sparse-activity world (BASELINE configs[4])",
            "air": "empty world (AIR inside the tickZone): the do-nothing floor of the tick",
            "bodies": f"8192x8192 world with {args.bodies} Box2D rigid bodies: pixel raster/erase, CCL + marching-squares outline each tick (BASELINE configs[3])"}[args.workload]
    return f"{base}; world {W}x{H} cells over {n} GPU(s)"


# ------------------------------------------------------------------------------------------------------------------
def cpu_tick_rate(args, table, extra, threads, budget_s, steps=None, warmup=1, game_ticks=0):
    """Reference-schedule CPU tick (oracle, 16-worker pool as world.cpp:59, libc rand(), AoS cells) on the same
    workload.  Returns (Gcell-updates/s, description, seconds per tick)."""
    from falling_sand_engine_b200 import worldgen as G
    from oracle import pyoracle as O

    # probe throughput on a 1024^2 crop, then pick the largest power-of-two square <= args.size that fits the budget
    probe = O.OracleWorld(1024, 1024, table)
    G.fill_world(probe, band_fn(args, table, extra), 1024, 1024, band_rows=512)
    O.lib().fseo_srand(1)
    probe.tick(0, schedule=O.REFERENCE, rng=O.RNG_LIBC, threads=threads)
    t = probe.tick(1, schedule=O.REFERENCE, rng=O.RNG_LIBC, threads=threads)
    rate = CELL_ITER * (1024 - 256) ** 2 / max(t, 1e-6)
    probe.close()
    n_ticks = (steps + warmup) if steps else 4
    size = args.size or 8192
    while size > 1024 and CELL_ITER * (size - 256) ** 2 * n_ticks / rate > budget_s:
        size //= 2
    if steps is None:
        n_ticks = max(2, min(16, int(budget_s * rate / (CELL_ITER * (size - 256) ** 2))))
    w = O.OracleWorld(size, size, table)
    G.fill_world(w, band_fn(args, table, extra), size, size, band_rows=512)
    times = []
    for i in range(n_ticks):
        times.append(w.tick(i, schedule=O.REFERENCE, rng=O.RNG_LIBC, threads=threads))
    game = None
    if game_ticks:  # the whole game tick on the CPU, in the order bench.py's e2e leg uses on the GPU (game.cpp:1711-2159)
        import numpy as np

        planes = [np.zeros((size, size, 4), dtype=np.uint8) for _ in range(3)]
        t0 = time.perf_counter()
        for i in range(n_ticks, n_ticks + game_ticks):
            w.tick(i, schedule=O.REFERENCE, rng=O.RNG_LIBC, threads=threads)
            w.particles_tick(schedule=O.REFERENCE)
            if i % 4 == 2:
                w.tick_temperature()
            O.render_dirty(w, planes)
            w.clear_dirty()
        game = (time.perf_counter() - t0) / game_ticks
    w.close()
    timed = times[warmup:] if len(times) > warmup else times
    per_tick = sum(timed) / len(timed)
    val = CELL_ITER * (size - 256) ** 2 / per_tick / 1e9
    desc = (f"oracle reference-schedule tick (4 colours x 128^2 chunks, {threads}-thread pool, libc rand, 40-byte AoS cells), "
            f"{len(timed)} ticks of the same generator at {size}x{size} after {min(warmup, len(times) - 1)} warm-up")
    if game_ticks:
        return val, desc, per_tick, size, len(timed), game
    return val, desc, per_tick, size, len(timed)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    table, extra = make_table()
    threads = args.cpu_threads or 16  # world.cpp:59: the reference always uses 16 tick workers
    cores = os.cpu_count() or 1
    t0 = time.time()
    val, desc, per_tick, size, n, game_s = cpu_tick_rate(args, table, extra, threads, budget_s=150.0, steps=args.steps, warmup=max(1, min(args.warmup, 3)),
                                                         game_ticks=3)
    out = {
        "impl": "reference",
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_tick * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args, size, size, 1), "cell_iter": CELL_ITER, "cpu_world": f"{size}x{size}"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": min(threads, cores), "threads": threads, "host_cores": cores,
                         "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        # the whole game tick on the CPU (world tick + tickCells + tickTemperature on tick % 4 == 2 + dirty -> textures + dirty clear), the
        # loop bench.py's own e2e leg runs on the GPU: 3 ticks right after the timed ones
        "game_loop": {"value": CELL_ITER * (size - 256) ** 2 / game_s / 1e9, "unit": UNIT, "ms_per_step": game_s * 1e3, "steps": 3},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------------------
def make_bodies(table, n_bodies, W, H, seed=7):
    """configs[3]: OBSIDIAN masks of 16..32 cells a side with hashed holes (~80 % fill), random pose inside the tickZone."""
    import numpy as np

    from falling_sand_engine_b200 import worldgen as G

    rng = np.random.default_rng(seed)
    bodies, masks = [], []
    for b in range(n_bodies):
        bw, bh = int(rng.integers(16, 33)), int(rng.integers(16, 33))
        hh = G.hash2(b + 1, np.arange(bw, dtype=np.uint32)[None, :], np.arange(bh, dtype=np.uint32)[:, None])
        m = np.where((hh % np.uint32(100)) < 80, 22, 0).astype(np.uint16)
        bodies.append(G.cells_from_mat(table, np.broadcast_to(m, (bh, bw)).copy(), 0, 0, b))
        mk = np.zeros((32, 32), dtype=np.uint8)
        mk[:bh, :bw] = m != 0
        masks.append(mk)
    xf = np.stack([rng.uniform(200, W - 200, n_bodies), rng.uniform(200, H - 200, n_bodies), rng.uniform(-3.1, 3.1, n_bodies)], axis=1).astype(np.float32)
    return bodies, np.stack(masks), xf


def run_ours(args):
    import numpy as np
    import torch

    import falling_sand_engine_b200 as fse
    from falling_sand_engine_b200 import types as T
    from falling_sand_engine_b200 import worldgen as G

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = max(args.gpus, world_size)
    dist = None
    if world_size > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fse library has no CPU fallback")

    table, extra = make_table()
    ctx = fse.Context(local_rank, table)
    strong = world_size > 1 and args.scaling == "strong"
    if world_size > 1:
        from falling_sand_engine_b200 import strips

        if strong:  # BASELINE configs[2]: one 32768-wide world with a 32768-row tickZone, cut world_size ways
            W = args.size or 32768
            Htot = args.height or (W + 2 * T.FSE_CHUNK)
        else:       # fixed tickZone rows per GPU
            W = args.size or 8192
            rows = args.height or W
            Htot = 2 * T.FSE_CHUNK + (rows - 2 * T.FSE_CHUNK) * world_size
        world = strips.StripWorld(ctx, W, Htot, rank, world_size, dist)
        H = Htot
    else:
        W = args.size or 8192
        H = args.height or W
        world = fse.World(ctx, W, H)
    zone_cells_total = (W - 2 * T.FSE_CHUNK) * (H - 2 * T.FSE_CHUNK)
    world.particles_reserve(1 << 25)
    if args.schedule != "default":
        world.set_schedule({"rows": 1, "rows_fused": 2}[args.schedule])
    use_active = (args.active == 1 or (args.active < 0 and args.workload == "sparse")) and world_size == 1
    if use_active:
        world.active_enable(True)

    fn = band_fn(args, table, extra)
    y_lo, y_hi = world.owned_rows() if world_size > 1 else (0, H)
    G.fill_world(world, fn, W, H, band_rows=1024, y_lo=y_lo, y_hi=y_hi)
    world.sync()

    def barrier():
        world.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def reduce_max(x):
        if not dist:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tick_no = 0
    for _ in range(max(args.warmup, 3)):
        world.tick(tick_no, seed=args.seed, cell_iter=CELL_ITER)
        tick_no += 1
    world.particles_clear()  # loose particles are integrated by fse_particles_tick, not part of this metric
    barrier()

    # ---- device-timed region: K ticks, state resident in HBM (1.1 GB per 8192^2, far larger than the 126 MB L2) ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    world.kernel_timing(True)
    barrier()
    world.timer_start()
    for _ in range(args.steps):
        world.tick(tick_no, seed=args.seed, cell_iter=CELL_ITER)
        tick_no += 1
    ms = world.timer_stop()
    barrier()
    clocks = sampler.stop()
    timeline = None
    if world_size > 1 and os.environ.get("FSE_STRIP_TIMELINE") == "1":
        tl = world.strip_timeline_read()
        if len(tl) >= 4 * CELL_ITER * args.steps:  # the timed ticks are the last ones recorded
            tl = np.asarray(tl[-4 * CELL_ITER * args.steps:], dtype=np.float64)
            mine = [float(tl[:, q].mean()) for q in range(3)] + [float(np.maximum(tl[:, 1], tl[:, 2]).mean())]
            tt = torch.tensor(mine, device="cuda", dtype=torch.float64)
            tmax = tt.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tt, op=dist.ReduceOp.SUM)
            timeline = {"what": "mean over the timed colour phases, ms after the phase start on the main stream (CUDA events; nsys is not in this image)",
                        "boundary_chunks_done_ms": {"mean_over_ranks": float(tt[0]) / world_size, "max_over_ranks": float(tmax[0])},
                        "halo_exchange_done_ms": {"mean_over_ranks": float(tt[1]) / world_size, "max_over_ranks": float(tmax[1])},
                        "interior_chunks_done_ms": {"mean_over_ranks": float(tt[2]) / world_size, "max_over_ranks": float(tmax[2])},
                        "phase_done_ms": {"mean_over_ranks": float(tt[3]) / world_size, "max_over_ranks": float(tmax[3])}}
    ph_ms = world.kernel_timing_phases()
    k_ms, k_launches = world.kernel_timing_read()
    world.kernel_timing(False)
    launches = ctx.launch_count() - launches0
    ms = reduce_max(ms)
    value = CELL_ITER * zone_cells_total * args.steps / (ms * 1e-3) / 1e9
    n_particles = world.particles_count()

    # ---- what the timed ticks computed: hash of the whole grid + per-material cell counts (the same numbers the oracle prints;
    #      tests/test_gpu_parity.py::test_benched_kernel_path_at_more_than_one_wave checks them against it on this kernel path).
    #      Strips: every rank hashes the rows it owns; the hash is a sum over cells keyed on global coordinates. ----
    st = world.stats_owned() if world_size > 1 else world.stats()
    h_own = np.array([st.hash], dtype=np.uint64).view(np.int64)
    cnt_own = np.array(list(st.count), dtype=np.int64)
    if dist:
        th = torch.from_numpy(np.concatenate([h_own, cnt_own, [int(st.n_dirty), int(st.n_moved)]]).astype(np.int64)).cuda()
        dist.all_reduce(th, op=dist.ReduceOp.SUM)  # int64 sums wrap like the uint64 hash does
        allv = th.cpu().numpy()
        h_tot, cnt_tot, n_dirty, n_moved = allv[:1].view(np.uint64)[0], allv[1:1 + len(cnt_own)], int(allv[-2]), int(allv[-1])
    else:
        h_tot, cnt_tot, n_dirty, n_moved = h_own.view(np.uint64)[0], cnt_own, int(st.n_dirty), int(st.n_moved)
    state = {"hash": f"{int(h_tot):016x}", "ticks": tick_no, "seed": args.seed, "n_dirty": n_dirty, "n_moved": n_moved,
             "counts": {str(i): int(c) for i, c in enumerate(cnt_tot) if c}}
    ref_hashes = os.path.join(ROOT, "profiles", "state_hashes.json")
    if os.path.exists(ref_hashes):  # hashes of the same (world, seed, ticks) from a 1-GPU run and from the oracle, when recorded
        try:
            with open(ref_hashes) as f:
                known = json.load(f).get(f"{args.workload}:{W}x{H}:seed{args.seed}:ticks{tick_no}")
            if known:
                state["recorded"] = known
                state["matches_recorded"] = all(v == state["hash"] for k, v in known.items() if k.startswith("hash"))
        except Exception:
            pass

    # ---- roofline of the dominant kernels (the chunk tick of one colour phase = classify + pass-1 + pass-2 + pass-3 kernels; one
    #      "launch" below is one phase), from CUDA events around every phase in the timed region ----
    peak, peak_src = peaks()
    own_zone_cells = zone_cells_total // max(world_size, 1)
    algo_bytes = ALGO_BYTES_PER_CELL_UPDATE * CELL_ITER * own_zone_cells * args.steps  # all launches of this rank
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic, traffic_src = None, "not measured in this run"
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.workload == "mixed" and (W, H) == (8192, 8192):
        try:
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj.get("tick_phase_bytes_per_launch")
            traffic_src = tj.get("source", "profiles/traffic.json") + " (ncu capture of this command, not of this run)"
        except Exception:
            traffic = None
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src, "peak_source": peak_src,
            "kernel": KERNEL_OF.get(args.schedule, KERNEL_OF["default"]),
            "launches": k_launches,
            "avg_launch_ms": k_ms / max(k_launches, 1),
            "algorithmic_bytes_per_launch": algo_bytes / max(k_launches, 1)}
    if len(ph_ms) == 4 * CELL_ITER * args.steps:  # mean ms of the 4 colour phases of each automaton iteration
        pm = np.asarray(ph_ms, dtype=np.float64).reshape(args.steps, CELL_ITER, 4).mean(axis=0)
        roof["phase_ms_by_iteration"] = [[round(float(v), 4) for v in row] for row in pm]
    if use_active:  # bytes actually touched: only awake chunks are streamed; their share of the zone's chunks scales the rate
        aw, tot = world.active_stats()
        if tot:
            roof["awake_chunk_fraction"] = aw / tot
            roof["achieved_touched"] = achieved * aw / tot
            roof["frac_touched"] = achieved * aw / tot / peak
    if args.workload in ("sparse", "air") or use_active:
        roof["dense_equivalent"] = True
        roof["note"] = ("settled rows / sleeping chunks are neither loaded nor stored, so `achieved` counts bytes the kernels did not move: it is the "
                        "dense-equivalent rate of SURVEY 8d(5), not a bandwidth; see config.awake_chunks and profiles/ for the DRAM bytes actually touched")

    # ---- end to end: one whole game tick through the C ABI with host buffers, in the reference's order (game.cpp:1711-2159) ----
    n_chunks = min(16, (H - 2 * T.FSE_CHUNK) // T.FSE_CHUNK)  # world::frame merges <= 16 loaded chunks per tick (world.cpp:2334-2391)
    pinned = torch.empty((n_chunks, T.FSE_CHUNK, T.FSE_CHUNK, T.CELL_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    src = fn(width=W, height=H, y0=0, rows=T.FSE_CHUNK)[:, : T.FSE_CHUNK]
    pinned_np = pinned.numpy()
    for i in range(n_chunks):
        pinned_np[i] = np.ascontiguousarray(src).view(np.uint8).reshape(T.FSE_CHUNK, T.FSE_CHUNK, -1)
    e2e_steps = max(3, min(args.steps, 10))
    zone = world.tickZone
    e2e_y0 = (world.own[0] if world_size > 1 else 0) + T.FSE_CHUNK
    single = world_size == 1
    have_bodies = single and args.workload == "bodies"
    bodies = masks = xf = None
    if have_bodies:
        bodies, masks, xf = make_bodies(table, args.bodies, W, H)
        world.bodies_upload(bodies)
    if single:
        world.pixels_enable(True)
    h2d = n_chunks * T.FSE_CHUNK * T.FSE_CHUNK * T.CELL_DTYPE.itemsize
    d2h = C.sizeof(T.RenderStats) if single else C.sizeof(T.Stats)
    if have_bodies:
        h2d += 2 * xf.nbytes + masks.nbytes
        d2h += 2 * 16 * len(bodies)  # feedback of raster and erase (outline results come back as well, size varies)
    stages = (["fse_bodies_raster"] if have_bodies else []) + [f"{n_chunks} chunk merges from pinned host memory (fse_write_rect)", "fse_tick"]
    strip_particles = (not single) and os.environ.get("FSE_E2E_STRIP_PARTICLES", "1") != "0"  # tickCells over the strips (migration + band proposals)
    stages += (["fse_particles_tick"] if single or strip_particles else []) + (["fse_bodies_erase", "fse_mask_outline of every body"] if have_bodies else [])
    stages += ["fse_tick_temperature on tick % 4 == 2"] + (["fse_render_dirty + movingTiles histogram readback", "fse_clear_dirty"] if single else ["fse_stats_rect readback"])
    moving = 0

    def game_tick(t):
        nonlocal moving
        if have_bodies:
            xf[:, 1] += 1.0   # the host's Box2D step moves the bodies a little
            xf[:, 2] += 0.02
            world.bodies_raster(xf, tick=t)
        for i in range(n_chunks):  # left border column of chunks (outside the tickZone), where scrolled-in chunks land
            world.write_rect_ptr(0, e2e_y0 + T.FSE_CHUNK * i, T.FSE_CHUNK, T.FSE_CHUNK, pinned_np[i].ctypes.data)
        world.tick(t, seed=args.seed, cell_iter=CELL_ITER)
        if single or strip_particles:
            world.particles_tick()
        if have_bodies:
            world.bodies_erase(xf)
            world.mask_outline(masks, want_labels=False, as_lists=False)
        if t % 4 == 2:
            world.tick_temperature()
        if single:
            rs = world.render_dirty(want_stats=True)
            moving = int(rs[0])
            world.clear_dirty()
        else:
            world.stats(T.Rect(zone.x, e2e_y0, zone.w, 1024))

    for _ in range(4):  # untimed: texture planes, scratch pools, body tables and (one tick in four) the temperature scratch reach their size
        game_tick(tick_no)
        tick_no += 1
    barrier()
    t0 = time.perf_counter()
    for s_ in range(e2e_steps):
        game_tick(tick_no)
        tick_no += 1
    barrier()
    e2e_s = reduce_max(time.perf_counter() - t0)
    e2e_val = CELL_ITER * zone_cells_total * e2e_steps / e2e_s / 1e9

    body_ms = None
    if have_bodies:  # per-stage wall clock of the bridge, synchronised after each call (3 ticks)
        acc = {"raster": 0.0, "erase": 0.0, "outline": 0.0}
        for _ in range(3):
            world.sync()
            a0 = time.perf_counter(); world.bodies_raster(xf, tick=tick_no); world.sync(); a1 = time.perf_counter()
            world.tick(tick_no, seed=args.seed, cell_iter=CELL_ITER); world.sync(); a2 = time.perf_counter()
            world.bodies_erase(xf); world.sync(); a3 = time.perf_counter()
            world.mask_outline(masks, want_labels=False, as_lists=False); a4 = time.perf_counter()
            acc["raster"] += a1 - a0; acc["erase"] += a3 - a2; acc["outline"] += a4 - a3
            tick_no += 1
        body_ms = {k: 1e3 * v / 3 for k, v in acc.items()}
        body_ms["bodies"] = len(bodies)
        body_ms["pixels"] = int(sum(int((b["mat"] != 0).sum()) for b in bodies))

    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        threads = args.cpu_threads or 16
        cval, cdesc, _, csize, _ = cpu_tick_rate(args, table, extra, threads, budget_s=20.0)
        cpu = {"value": cval, "unit": UNIT, "cores": min(threads, os.cpu_count() or 1), "threads": threads,
               "host_cores": os.cpu_count(), "kind": "port", "sample": cdesc}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "u8/f32", "data": "synthetic",
            "config": {"workload": workload_name(args, W, H, n), "cell_iter": CELL_ITER, "tick_zone": [W - 256, H - 256],
                       "ticks_per_s": args.steps / (ms * 1e-3), "l2": "state (17 B/cell, >=1.1 GB) is larger than the 126 MB L2",
                       "parallelism": f"strips{n}" if n > 1 else "single", "active_chunk_tracking": bool(use_active),
                       "awake_chunks": list(world.active_stats()) if use_active else None},
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "what": " + ".join(stages) + "; wall clock, max over ranks",
                    "dirty_cells_last_tick": moving},
            "state": state,
            "gpu_launches": launches,
            "clocks": clocks,
            "particles_spawned_in_timed_region": n_particles,
        }
        if body_ms:
            out["bodies"] = body_ms
        if timeline:
            out["strip_timeline"] = timeline
        if strong:
            n1 = os.path.join(ROOT, "profiles", "r2_cfg3_32768x33024_n1.json")
            if os.path.exists(n1) and (W, H) == (32768, 33024):  # the same world on one GPU, measured with this command line (--size 32768 --height 33024)
                try:
                    with open(n1) as f:
                        d1 = json.loads(f.read().strip().splitlines()[-1])
                    out["one_gpu_same_world"] = {"value": d1["value"], "ms_per_step": d1["ms_per_step"], "source": "profiles/r2_cfg3_32768x33024_n1.json"}
                except Exception:
                    pass
        print(json.dumps(out))
    world.close()
    ctx.close()
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
