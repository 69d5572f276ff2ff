"""Worker of tests/test_strips_cpu.py: world_size ranks over gloo, each ticking its strip of an ORACLE world with the
halo protocol of falling_sand_engine_b200.strips (the same rule fse_comm.cu implements with NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from falling_sand_engine_b200 import strips, types as T, worldgen as G  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from tests import helpers as Hh  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    W, H, ticks, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table = O.default_materials(1337)
    own_lo, own_hi, held_lo, held_hi, j0, j1 = strips.strip_layout(H, rank, world)
    ow = O.OracleWorld(W, H, table)  # full-size container; only the held rows are meaningful on this rank
    full = G.mixed_band(table, W, H, 0, H, seed=77, blob=32)
    junk = G.cells_from_mat(table, np.full((H, W), 7, dtype=np.uint16))  # rows this rank must never rely on
    junk[held_lo:held_hi] = full[held_lo:held_hi]
    ow.write_rect(0, 0, junk)
    zone = T.zone_of(W, H)
    nx = zone.w // T.FSE_CHUNK
    for t in range(ticks):
        for it in range(3):
            for tk in range(4):
                ofx, ofy = tk % 2, 1 - (tk // 2)
                ow.clear_visited()
                for j in range(j0, j1):
                    if j % 2 != ofy:
                        continue
                    for i in range(ofx, nx, 2):
                        ow.run_chunk(t, 1337, it, zone.x + i * T.FSE_CHUNK, zone.y + j * T.FSE_CHUNK)
                reqs, recvs = [], []
                for peer, kind, ylo, yhi in strips.phase_messages(rank, world, j0, j1, ofy, zone.y):
                    if kind == "send":
                        buf = torch.from_numpy(np.ascontiguousarray(ow.read_rect(0, ylo, W, yhi - ylo)).view(np.uint8).copy())
                        reqs.append(dist.isend(buf, peer))
                    else:
                        buf = torch.empty(((yhi - ylo) * W * T.CELL_DTYPE.itemsize,), dtype=torch.uint8)
                        reqs.append(dist.irecv(buf, peer))
                        recvs.append((buf, ylo, yhi))
                for r in reqs:
                    r.wait()
                for buf, ylo, yhi in recvs:
                    ow.write_rect(0, ylo, buf.numpy().view(T.CELL_DTYPE).reshape(yhi - ylo, W))
    np.save(f"{out}.rank{rank}.npy", ow.read_rect(0, own_lo, W, own_hi - own_lo))
    parts = ow.particles_read()
    np.save(f"{out}.parts{rank}.npy", parts)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
