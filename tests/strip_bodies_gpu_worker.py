"""Worker of tests/test_strips_gpu.py::test_strip_bodies_match_oracle (torchrun, one rank per GPU): the rigid-body bridge on a StripWorld.
Every rank makes every call with the same transforms; a body is run by the rank that holds its footprint box."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import falling_sand_engine_b200 as fse  # noqa: E402
from falling_sand_engine_b200 import materials as M, strips, worldgen as G  # noqa: E402
from tests.strip_bodies_scene import scene, stone_blocks, tool_calls, run_tool, entities  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    W, H, ticks, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    table = M.default_materials(1337)
    ctx = fse.Context(local, table)
    sw = strips.StripWorld(ctx, W, H, rank, world, dist)
    lo, hi = sw.owned_rows()
    sw.write_rect(0, lo, G.mixed_band(table, W, H, lo, hi - lo, seed=21, air_frac=0.6, blob=48))
    bodies, xf = scene(table, W, H, world)
    for (x0, y0, cells) in stone_blocks(table, H, world):  # something for the pickaxe and the hammer to work on, right on the cuts
        sw.write_rect(x0, y0, cells)
    sw.bodies_upload(bodies)
    fbs, tools = [], []
    ents = entities(H, world)
    for t in range(ticks):
        ents = sw.entities_tick(ents, tick=t)   # world::tickEntities, then the entities become OBJECT cells for the rest of the game tick
        sw.entities_stamp(ents, tick=t)
        fbs.append(sw.bodies_raster(xf, tick=t))
        sw.tick(t, seed=1337)
        if t == 1:  # tools across the cuts: every rank makes the call and gets the same answer
            for call in tool_calls(H, world, t):
                tools.append(np.asarray(run_tool(sw, call), dtype=np.int64).reshape(-1))
        sw.particles_tick()
        if t == 3:
            tools.append(np.array([sw.particles_vacuum_pull(500.0, 700.0)], dtype=np.int64))
        fe, need = sw.bodies_erase(xf)
        fbs.append(fe)
        sw.object_delete()
        xf[:, 1] += 1.5   # the host's Box2D step: bodies drift down (some change strips on the way) and turn
        xf[:, 2] += 0.05
    sw.sync()
    np.save(f"{out}.rank{rank}.npy", sw.read_owned())
    np.save(f"{out}.parts{rank}.npy", sw.particles_read())
    np.save(f"{out}.fb{rank}.npy", np.stack(fbs))
    np.save(f"{out}.tools{rank}.npy", np.concatenate(tools))
    np.save(f"{out}.ents{rank}.npy", ents)
    np.save(f"{out}.tiles{rank}.npy", np.concatenate([sw.bodies_read(i).reshape(-1) for i in range(len(bodies))]))
    dist.barrier()
    sw.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
