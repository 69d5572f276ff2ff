"""Worker of tests/test_strips_gpu.py (torchrun, one rank per GPU): ticks a StripWorld and saves its owned rows."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import falling_sand_engine_b200 as fse  # noqa: E402
from falling_sand_engine_b200 import materials as M, strips, worldgen as G  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    W, H, ticks, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    table = M.default_materials(1337)
    ctx = fse.Context(local, table)
    sw = strips.StripWorld(ctx, W, H, rank, world, dist)
    lo, hi = sw.owned_rows()
    sw.write_rect(0, lo, G.mixed_band(table, W, H, lo, hi - lo, seed=77, blob=32))
    for t in range(ticks):
        sw.tick(t, seed=1337)
        if t == 2:  # a blast and an eraser stroke across the cuts: every rank makes the call, each clears the rows it holds
            sw.explosion(W // 2, H // 2 + 5, 40, tick=t, seed=1337)
            sw.tool_erase_line(200, H // 4 - 30, 700, 3 * H // 4 + 20, 9)
        if t == 4:  # the camera moves one chunk to the right: a horizontal shift stays inside every rank's rows
            sw.scroll(-128, 0)
        if t == 5:  # ... one chunk down: rows change ranks (one message of 128 rows each way, the rest is shifted in place)
            sw.scroll(0, 128)
        if t == 6:  # ... and up and to the right by odd amounts
            sw.scroll(36, -97)
        sw.particles_tick()  # ghost refresh, migration, integration and the deposit rounds with the band proposals exchanged
        if t % 4 == 2:
            sw.tick_temperature()
    sw.sync()
    np.save(f"{out}.rank{rank}.npy", sw.read_owned())
    np.save(f"{out}.parts{rank}.npy", sw.particles_read())
    dist.barrier()
    sw.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
