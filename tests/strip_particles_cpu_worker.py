"""Worker of tests/test_strips_cpu.py::test_strip_particle_protocol: the world tick AND the loose-particle tick over horizontal
strips, on oracle worlds over gloo.  The particle protocol is the one DESIGN.md §8 lays out for the CUDA path:
  1. after the tick, ghost rows are refreshed and every particle moves to the rank that owns its row;
  2. each rank integrates its particles (oracle prt_begin) and, per deposit round, proposes cells for them (prt_propose);
  3. proposals that target the band of GHOST rows on either side of a cut are exchanged with that neighbour, both ranks commit the
     union (prt_commit): the lowest id wins a cell whoever owns the particle, and both write the band cells they hold;
  4. rounds stop when no rank has a proposal left; deposited particles are dropped (prt_end).
The result must equal the single-world tick + tick_particles_rounds bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from falling_sand_engine_b200 import strips, types as T, worldgen as G  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def exchange(rank, world, to_up, to_down, dtype):
    """Send one array to each neighbour (None where there is none), return (from_up, from_down)."""
    got = {}
    for peer, arr in ((rank - 1, to_up), (rank + 1, to_down)):
        if peer < 0 or peer >= world:
            continue
        n_out = torch.tensor([len(arr)], dtype=torch.int64)
        n_in = torch.zeros(1, dtype=torch.int64)
        ops = [dist.isend(n_out, peer), dist.irecv(n_in, peer)]
        for o in ops:
            o.wait()
        out = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).reshape(-1).copy())
        inc = torch.empty(int(n_in.item()) * dtype.itemsize, dtype=torch.uint8)
        ops = []
        if len(out):
            ops.append(dist.isend(out, peer))
        if len(inc):
            ops.append(dist.irecv(inc, peer))
        for o in ops:
            o.wait()
        got[peer] = inc.numpy().view(dtype).copy()
    empty = np.zeros(0, dtype=dtype)
    return got.get(rank - 1, empty), got.get(rank + 1, empty)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    W, H, ticks, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table = O.default_materials(1337)
    own_lo, own_hi, held_lo, held_hi, j0, j1 = strips.strip_layout(H, rank, world)
    ow = O.OracleWorld(W, H, table)
    full = G.mixed_band(table, W, H, 0, H, seed=77, blob=32)
    junk = G.cells_from_mat(table, np.full((H, W), 7, dtype=np.uint16))
    junk[held_lo:held_hi] = full[held_lo:held_hi]
    ow.write_rect(0, 0, junk)
    zone = T.zone_of(W, H)
    nx = zone.w // T.FSE_CHUNK
    GH = strips.GHOST
    cell_b = T.CELL_DTYPE

    n_exchanged = n_migrated = 0

    def rows(lo, hi):
        return np.ascontiguousarray(ow.read_rect(0, lo, W, hi - lo)).reshape(-1)

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
                        buf = torch.empty(((yhi - ylo) * W * cell_b.itemsize,), dtype=torch.uint8)
                        reqs.append(dist.irecv(buf, peer))
                        recvs.append((buf, ylo, yhi))
                for r in reqs:
                    r.wait()
                for buf, ylo, yhi in recvs:
                    ow.write_rect(0, ylo, buf.numpy().view(cell_b).reshape(yhi - ylo, W))
        # ---- loose particles ----
        # 1a. ghost rows: the owner's rows next to each cut (fse_strip_refresh)
        up = rows(own_lo, own_lo + GH) if rank > 0 else np.zeros(0, dtype=cell_b)
        down = rows(own_hi - GH, own_hi) if rank + 1 < world else np.zeros(0, dtype=cell_b)
        from_up, from_down = exchange(rank, world, up, down, cell_b)
        if rank > 0:
            ow.write_rect(0, own_lo - GH, from_up.reshape(GH, W))
        if rank + 1 < world:
            ow.write_rect(0, own_hi, from_down.reshape(GH, W))
        # 1b. every particle to the rank that owns its row
        parts = ow.particles_read()
        row = np.floor(parts["y"]).astype(np.int64)
        go_up = parts[(row < own_lo) & (rank > 0)]
        go_down = parts[(row >= own_hi) & (rank + 1 < world)]
        keep = parts[~(((row < own_lo) & (rank > 0)) | ((row >= own_hi) & (rank + 1 < world)))]
        n_migrated += len(go_up) + len(go_down)
        from_up, from_down = exchange(rank, world, go_up, go_down, T.PARTICLE_DTYPE)
        ow.particles_clear()
        ow.particles_add(np.concatenate([keep, from_up, from_down]))
        # 2-4. integrate, then deposit rounds with the band proposals exchanged
        O.prt_begin(ow, zone)
        for r in range(16):
            n = O.prt_propose(ow)
            total = torch.tensor([n], dtype=torch.int64)
            dist.all_reduce(total)
            if int(total.item()) == 0:
                break
            to_up = O.prt_get(ow, own_lo - GH, own_lo + GH) if rank > 0 else np.zeros(0, dtype=O.PROPOSAL_DTYPE)
            to_down = O.prt_get(ow, own_hi - GH, own_hi + GH) if rank + 1 < world else np.zeros(0, dtype=O.PROPOSAL_DTYPE)
            n_exchanged += len(to_up) + len(to_down)
            from_up, from_down = exchange(rank, world, to_up, to_down, O.PROPOSAL_DTYPE)
            O.prt_commit(ow, np.concatenate([from_up, from_down]), held_lo, held_hi)
        O.prt_end(ow)
    np.save(f"{out}.rank{rank}.npy", ow.read_rect(0, own_lo, W, own_hi - own_lo))
    parts = ow.particles_read()
    row = np.floor(parts["y"]).astype(np.int64)
    mine = ((row >= own_lo) | (rank == 0)) & ((row < own_hi) | (rank + 1 == world))
    np.save(f"{out}.parts{rank}.npy", parts)
    np.save(f"{out}.counts{rank}.npy", np.array([n_exchanged, n_migrated, int((~mine).sum())]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
