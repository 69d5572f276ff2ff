"""Multi-GPU worlds: one horizontal strip per rank, halo rows over NCCL (include/fse.h fse_comm_* / fse_strip_*).

`torch.distributed` is only the out-of-band channel for the 128-byte NCCL id and for reductions of statistics;
the halo exchange itself is ncclSend/ncclRecv issued by the C++ library on its own stream.

`strip_layout` is the pure host-side partition rule (also used by the CPU tests with gloo).
"""
import ctypes as C

import numpy as np

from . import api
from . import types as T

GHOST = 32


def strip_layout(height_global, rank, nranks):
    """Rows of rank `rank`: (own_lo, own_hi, held_lo, held_hi, j0, j1) with cuts at y = 128 + 128*j."""
    nz = (height_global - 2 * T.FSE_CHUNK) // T.FSE_CHUNK
    if nz < nranks:
        raise ValueError(f"{nz} chunk rows cannot be split over {nranks} ranks")
    j0, j1 = nz * rank // nranks, nz * (rank + 1) // nranks
    own_lo = 0 if rank == 0 else T.FSE_CHUNK + T.FSE_CHUNK * j0
    own_hi = height_global if rank == nranks - 1 else T.FSE_CHUNK + T.FSE_CHUNK * j1
    return own_lo, own_hi, max(0, own_lo - GHOST), min(height_global, own_hi + GHOST), j0, j1


def phase_messages(rank, nranks, j0, j1, ofy, zone_y=T.FSE_CHUNK):
    """Halo messages of one colour phase (row parity `ofy`) as (peer, 'send'|'recv', y_lo, y_hi) in global rows:
    the rank whose boundary chunk row just ran sends the rows around the cut (fse_comm.cu strip_exchange)."""
    out = []
    if rank > 0:
        cut = zone_y + j0 * T.FSE_CHUNK
        out.append((rank - 1, "send", cut - 5, cut + 10) if j0 % 2 == ofy else (rank - 1, "recv", cut - 5, cut + 5))
    if rank + 1 < nranks:
        cut = zone_y + j1 * T.FSE_CHUNK
        out.append((rank + 1, "send", cut - 5, cut + 5) if (j1 - 1) % 2 == ofy else (rank + 1, "recv", cut - 5, cut + 10))
    return out


class StripWorld(api.World):
    """One rank's strip of a `width` x `height_global` world.  Coordinates stay global."""

    def __init__(self, ctx, width, height_global, rank, nranks, dist=None):
        self.L = ctx.L
        self.ctx = ctx
        self.width, self.height = width, height_global
        self.rank, self.nranks, self.dist = rank, nranks, dist
        L = self.L
        L.fse_comm_unique_id.argtypes = [C.c_void_p]
        L.fse_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.fse_comm_destroy.argtypes = [C.c_void_p]
        L.fse_strip_create.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.fse_strip_rows.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 4
        L.fse_strip_refresh.argtypes = [C.c_void_p]
        if nranks > 1:
            import torch

            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = (C.c_char * 128)()
                api._ck(L.fse_comm_unique_id(buf))
                uid = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
            uid = uid.cuda()
            dist.broadcast(uid, src=0)
            raw = bytes(uid.cpu().numpy().tobytes())
            api._ck(L.fse_comm_init(ctx.h, rank, nranks, raw))
        self.h = C.c_void_p()
        api._ck(L.fse_strip_create(ctx.h, width, height_global, C.byref(self.h)))
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        api._ck(L.fse_strip_rows(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        self.own = (a.value, b.value)
        self.held = (c.value, d.value)
        self.tickZone = T.zone_of(width, height_global)

    def owned_rows(self):
        """Rows the generator must upload on this rank (owned + ghost)."""
        return self.held

    def write_rect(self, x, y, cells):
        """Global rect; the part outside the held rows is dropped."""
        cells = np.ascontiguousarray(cells, dtype=T.CELL_DTYPE)
        h = cells.shape[0]
        lo, hi = max(y, self.held[0]), min(y + h, self.held[1])
        if lo < hi:
            super().write_rect(x, lo, cells[lo - y: hi - y])

    def read_owned(self):
        return self.read_rect(0, self.own[0], self.width, self.own[1] - self.own[0])

    def stats_owned(self):
        return self.stats(T.Rect(0, self.own[0], self.width, self.own[1] - self.own[0]))

    def refresh(self):
        api._ck(self.L.fse_strip_refresh(self.h))

    def close(self):
        super().close()
        if self.nranks > 1 and self.ctx.h:
            self.L.fse_comm_destroy(self.ctx.h)
