"""Worker of tests/test_strips_cpu.py::test_strip_box_edit_protocol_matches_single_world: world_size ranks over gloo, each holding its
window of an ORACLE world, run the rigid-body bridge and the entities with the product's ownership plan (`fse_strip_plan`, host code of
libfse_b200.so): the runner of a body / entity works on its window, the planned rectangles travel to the neighbours, what the calls
return is summed over the ranks.  The same protocol fse_bodies.cu / fse_entities.cu / fse_comm.cu implement with NCCL."""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from falling_sand_engine_b200 import api, strips, types as T, worldgen as G  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from tests.strip_bodies_scene import scene, entities  # noqa: E402


def plan(L, W, H, rank, nranks, boxes):
    n = len(boxes)
    b = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
    runner = np.zeros(max(n, 1), dtype=np.int32)
    cap = 4 * n + 8
    rects = np.zeros((4, cap, 4), dtype=np.int32)
    cnt = np.zeros(4, dtype=np.int32)
    L.fse_strip_plan.argtypes = [C.c_int32] * 4 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    rc = L.fse_strip_plan(W, H, rank, nranks, b.ctypes.data, n, runner.ctypes.data, rects.ctypes.data, cap, cnt.ctypes.data)
    assert rc == 0, L.fse_last_error()
    return runner[:n].copy(), [rects[q, :cnt[q]].copy() for q in range(4)]


def body_boxes(bodies, xf):
    """Footprint boxes like bodies_aabb_kernel's (float32 sums, +-3 cells)."""
    out = []
    for b, (x, y, ang) in zip(bodies, xf):
        h, w = b.shape
        s, c = np.float32(math.sin(np.float32(ang))), np.float32(math.cos(np.float32(ang)))
        xs, ys = [], []
        for tx in (np.float32(0), np.float32(w - 1)):
            for ty1 in (np.float32(1), np.float32(h)):
                xs.append(tx * c - ty1 * s + np.float32(x))
                ys.append(tx * s + ty1 * c + np.float32(y))
        out.append((math.floor(min(xs)) - 3, math.floor(min(ys)) - 3, math.ceil(max(xs)) + 3, math.ceil(max(ys)) + 3))
    return out


def entity_boxes(ents):
    out = []
    for e in ents:
        m = math.ceil(abs(float(e["vx"]))) + math.ceil(abs(float(e["vy"]))) + 14
        x0, y0 = int(e["x"]) - m, int(e["y"]) - m
        out.append((x0, y0, x0 + int(e["hw"]) + 2 * m, y0 + int(e["hh"]) + 2 * m))
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    W, H, ticks, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = api.load_library()
    table = O.default_materials(1337)
    own_lo, own_hi, held_lo, held_hi, _, _ = strips.strip_layout(H, rank, world)
    ow = O.OracleWorld(W, H, table)  # full-size container; only the held rows are meaningful on this rank
    full = G.mixed_band(table, W, H, 0, H, seed=21, air_frac=0.6, blob=48)
    junk = G.cells_from_mat(table, np.full((H, W), 7, dtype=np.uint16))  # rows this rank must never rely on
    junk[held_lo:held_hi] = full[held_lo:held_hi]
    ow.write_rect(0, 0, junk)
    bodies, xf = scene(table, W, H, world)
    air_tiles = [np.zeros_like(b) for b in bodies]
    for a in air_tiles:
        a["fluid"] = 2.0  # Tiles_NOTHING: mat AIR (0)
    ents = entities(H, world)

    def exchange(rects):
        """rects: [up send, up recv, down send, down recv], (x0, y0 local, w, h); one cut after the other, the upper rank sends first."""
        def send(lst, peer):
            for (x0, y0, w, h) in lst:
                buf = torch.from_numpy(np.ascontiguousarray(ow.read_rect(int(x0), int(y0) + held_lo, int(w), int(h))).view(np.uint8).copy())
                dist.send(buf, peer)

        def recv(lst, peer):
            for (x0, y0, w, h) in lst:
                buf = torch.empty((int(w) * int(h) * T.CELL_DTYPE.itemsize,), dtype=torch.uint8)
                dist.recv(buf, peer)
                ow.write_rect(int(x0), int(y0) + held_lo, buf.numpy().view(T.CELL_DTYPE).reshape(int(h), int(w)))
        if rank > 0:  # the cut above me: my upper neighbour sends first
            recv(rects[1], rank - 1)
            send(rects[0], rank - 1)
        if rank + 1 < world:
            send(rects[2], rank + 1)
            recv(rects[3], rank + 1)

    def summed(a):
        t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).astype(np.int64))
        dist.all_reduce(t)
        return t.numpy().astype(np.uint8).view(a.dtype).reshape(a.shape)

    fbs = []
    for t in range(ticks):
        # tickEntities: the runner of an entity updates its record, the others contribute zeros
        runner, rects = plan(L, W, H, rank, world, entity_boxes(ents))
        mine = ents.copy()
        upd = O.entities_tick(ow, mine[runner == rank], tick=t) if (runner == rank).any() else mine[:0]
        contrib = np.zeros_like(ents)
        contrib[runner == rank] = upd
        exchange(rects)
        ents = summed(contrib)
        # body raster / erase: a rank runs its bodies (the others are empty here: AIR tiles are skipped, indices — RNG keys, particle ids — stay)
        for erase in (False, True):
            runner, rects = plan(L, W, H, rank, world, body_boxes(bodies, xf))
            use = [bodies[i] if runner[i] == rank else air_tiles[i] for i in range(len(bodies))]
            fb = O.bodies_erase(ow, use, xf) if erase else O.bodies_raster(ow, use, xf, tick=t)
            exchange(rects)
            fbs.append(summed(fb.astype(np.int32)))
            if erase:  # the rewritten tiles of every runner on every rank
                for i in range(len(bodies)):
                    z = bodies[i] if runner[i] == rank else np.zeros_like(bodies[i])
                    bodies[i][...] = summed(np.ascontiguousarray(z))
        xf[:, 1] += 1.5
        xf[:, 2] += 0.05
    np.save(f"{out}.rank{rank}.npy", ow.read_rect(0, own_lo, W, own_hi - own_lo))
    np.save(f"{out}.parts{rank}.npy", ow.particles_read())
    np.save(f"{out}.fb{rank}.npy", np.stack(fbs))
    np.save(f"{out}.tiles{rank}.npy", np.concatenate([b.reshape(-1) for b in bodies]))
    np.save(f"{out}.ents{rank}.npy", ents)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
