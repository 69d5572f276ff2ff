#!/usr/bin/env python
"""bench.py — Gcell-updates/s of the world tick (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S] [--workload mixed|column|sparse]

A "step" is one world::tick() (cell_iter = 3 automaton iterations over the whole tickZone) on a synthetic world.
Default workload = BASELINE.json configs[1]: 8192x8192 mixed powders/liquids/gases with fire and the registered
interacting materials (worldgen.bench_table).  N > 1 partitions the world into N horizontal strips (one rank per
GPU, halo rows exchanged over NCCL) with a fixed strip height per GPU ("weak" scaling: 8192 x 8192*N... see config).

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
KERNEL_OF = {"default": "one colour phase = fse::tick_pass_kernel<1> + tick_pass_kernel<2> + tick_pass3_kernel", "rows_fused": "fse::tick_rows_kernel",
             "classes": "fse::tick_chunk_kernel"}
KERNEL_OF["rows"] = KERNEL_OF["default"]
METRIC = "Gcell-updates/sec (device-timed) at 1/2/4/8 B200; % HBM roofline"
UNIT = "Gcell-updates/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=8192, help="world width (and per-GPU strip height)")
    ap.add_argument("--height", type=int, default=0, help="world height (per-GPU strip height with --gpus N); 0 = --size")
    ap.add_argument("--workload", default="mixed", choices=["mixed", "column", "sparse"])
    ap.add_argument("--seed", type=int, default=1337)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--schedule", default="default", choices=["default", "rows", "rows_fused", "classes"], help="in-row schedule of the tick kernel")
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
    if args.workload == "column":
        return functools.partial(G.column_drop_band, table, seed=args.seed)
    return functools.partial(G.sparse_band, table, seed=args.seed)


def make_table():
    """Stock materials + the three registered interacting powders of config 2 (the 'Lua material table')."""
    from falling_sand_engine_b200 import materials as M
    from falling_sand_engine_b200 import worldgen as G

    return G.bench_table(M.default_materials(1337))


def workload_name(args, W, H, n):
    mixed = ("8192x8192 mixed powders/liquids/gases with fire and Lua-table reactions (BASELINE configs[1])" if (W, H) == (8192, 8192) or n > 1 and W == 8192
             else "mixed powders/liquids/gases with fire and Lua-table reactions, generator of BASELINE configs[1]"
             + (" (configs[2]: 32768-wide world, strip-partitioned)" if W == 32768 else ""))
    base = {"mixed": mixed,
            "column": "sand/water/stone column drop (BASELINE configs[0])",
            "sparse": "mostly-settled sparse-activity world (BASELINE configs[4])"}[args.workload]
    return f"{base}; world {W}x{H} cells over {n} GPU(s)"


# ------------------------------------------------------------------------------------------------------------------
def cpu_tick_rate(args, table, extra, threads, budget_s, steps=None, warmup=1):
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
    size = args.size
    while size > 1024 and CELL_ITER * (size - 256) ** 2 * n_ticks / rate > budget_s:
        size //= 2
    if steps is None:
        n_ticks = max(2, min(16, int(budget_s * rate / (CELL_ITER * (size - 256) ** 2))))
    w = O.OracleWorld(size, size, table)
    G.fill_world(w, band_fn(args, table, extra), size, size, band_rows=512)
    times = []
    for i in range(n_ticks):
        times.append(w.tick(i, schedule=O.REFERENCE, rng=O.RNG_LIBC, threads=threads))
    w.close()
    timed = times[warmup:] if len(times) > warmup else times
    per_tick = sum(timed) / len(timed)
    val = CELL_ITER * (size - 256) ** 2 / per_tick / 1e9
    desc = (f"oracle reference-schedule tick (4 colours x 128^2 chunks, {threads}-thread pool, libc rand, 40-byte AoS cells), "
            f"{len(timed)} ticks of the same generator at {size}x{size} after {min(warmup, len(times) - 1)} warm-up")
    return val, desc, per_tick, size, len(timed)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    table, extra = make_table()
    threads = args.cpu_threads or 16  # world.cpp:59: the reference always uses 16 tick workers
    cores = os.cpu_count() or 1
    t0 = time.time()
    val, desc, per_tick, size, n = cpu_tick_rate(args, table, extra, threads, budget_s=150.0, steps=args.steps, warmup=max(1, min(args.warmup, 3)))
    out = {
        "impl": "reference",
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": per_tick * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args, size, size, 1), "cell_iter": CELL_ITER, "cpu_world": f"{size}x{size}"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": min(threads, cores), "threads": threads, "host_cores": cores,
                         "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------------------------
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
    W = args.size
    rows = args.height or args.size
    if world_size > 1:
        from falling_sand_engine_b200 import strips

        Htot = 2 * T.FSE_CHUNK + (rows - 2 * T.FSE_CHUNK) * world_size  # fixed tickZone rows per GPU ("weak")
        world = strips.StripWorld(ctx, W, Htot, rank, world_size, dist)
        H = Htot
    else:
        H = rows
        world = fse.World(ctx, W, H)
    zone_cells_total = (W - 2 * T.FSE_CHUNK) * (H - 2 * T.FSE_CHUNK)
    world.particles_reserve(1 << 25)
    if args.schedule != "default":
        world.set_schedule({"classes": 0, "rows": 1, "rows_fused": 2}[args.schedule])
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
    k_ms, k_launches = world.kernel_timing_read()
    world.kernel_timing(False)
    launches = ctx.launch_count() - launches0
    if dist:
        tms = torch.tensor([ms], device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = CELL_ITER * zone_cells_total * args.steps / (ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernels (the chunk tick of one colour phase = pass-1 + pass-2 + pass-3 kernels; one
    #      "launch" below is one phase), from CUDA events around every phase in the timed region ----
    peak, peak_src = peaks()
    own_zone_cells = zone_cells_total // max(world_size, 1)
    algo_bytes = ALGO_BYTES_PER_CELL_UPDATE * CELL_ITER * own_zone_cells * args.steps  # all launches of this rank
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get("tick_phase_bytes_per_launch")
        except Exception:
            traffic = None
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": peak_src,
            "kernel": ("fse::tick_graph_kernel (one launch = one whole tick: 4 colours x cell_iter phases as a task graph over the chunks, passes 1-3 per chunk)"
                       if k_launches == args.steps else KERNEL_OF.get(args.schedule, KERNEL_OF["default"])),
            "launches": k_launches,
            "avg_launch_ms": k_ms / max(k_launches, 1),
            "algorithmic_bytes_per_launch": algo_bytes / max(k_launches, 1)}

    # ---- end to end through the C ABI with host buffers: per step, merge 16 freshly "loaded" chunks from pinned host
    #      memory (world::frame, world.cpp:2334-2391: <=16 chunks per tick), tick, read the statistics back ----
    n_chunks = min(16, (H - 2 * T.FSE_CHUNK) // T.FSE_CHUNK)  # 16 per tick in the reference; fewer only on tiny worlds
    pinned = torch.empty((n_chunks, T.FSE_CHUNK, T.FSE_CHUNK, T.CELL_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    src = fn(width=W, height=H, y0=0, rows=T.FSE_CHUNK)[:, : T.FSE_CHUNK]
    pinned_np = pinned.numpy()
    for i in range(n_chunks):
        pinned_np[i] = np.ascontiguousarray(src).view(np.uint8).reshape(T.FSE_CHUNK, T.FSE_CHUNK, -1)
    h2d = n_chunks * T.FSE_CHUNK * T.FSE_CHUNK * T.CELL_DTYPE.itemsize
    d2h = C.sizeof(T.Stats)
    e2e_steps = max(3, min(args.steps, 10))
    zone = world.tickZone
    n_particles = world.particles_count()
    world.particles_clear()
    e2e_y0 = (world.own[0] if world_size > 1 else 0) + T.FSE_CHUNK
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        for i in range(n_chunks):  # left border column of chunks (outside the tickZone), where scrolled-in chunks land
            world.write_rect_ptr(0, e2e_y0 + T.FSE_CHUNK * i, T.FSE_CHUNK, T.FSE_CHUNK, pinned_np[i].ctypes.data)
        world.tick(tick_no, seed=args.seed, cell_iter=CELL_ITER)
        tick_no += 1
        st = world.stats(T.Rect(zone.x, e2e_y0, zone.w, 1024))  # movingTiles-style histogram readback
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist:
        te = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
    e2e_val = CELL_ITER * zone_cells_total * e2e_steps / e2e_s / 1e9

    cpu = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        threads = args.cpu_threads or 16
        cval, cdesc, _, csize, _ = cpu_tick_rate(args, table, extra, threads, budget_s=20.0)
        cpu = {"value": cval, "unit": UNIT, "cores": min(threads, os.cpu_count() or 1), "threads": threads,
               "host_cores": os.cpu_count(), "kind": "port", "sample": cdesc}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args, W, H, n), "cell_iter": CELL_ITER, "tick_zone": [W - 256, H - 256],
                       "ticks_per_s": args.steps / (ms * 1e-3), "l2": "state (17 B/cell, >=1.1 GB) is larger than the 126 MB L2",
                       "parallelism": f"strips{n}" if n > 1 else "single", "active_chunk_tracking": bool(use_active),
                       "awake_chunks": list(world.active_stats()) if use_active else None},
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "what": f"{n_chunks} chunk merges from pinned host memory (fse_write_rect) + fse_tick + fse_stats_rect readback, wall clock"},
            "gpu_launches": launches,
            "clocks": clocks,
            "particles_spawned_in_timed_region": n_particles,
        }
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
