"""Host logic of the calls that edit the grid from one place on multi-rank strips (bodies, entities, cracks, probed components): who
runs which box and which rectangles travel afterwards — `fse_strip_plan`, pure host code of the product library (fse_comm.cu:
strip_group_runners, strip_rects_of_box), checked on the CPU for every rank of 2 .. 8-rank layouts:
every rank computes the same runners; overlapping boxes share a runner; a runner holds its boxes; what one rank sends is, rectangle by
rectangle and in the same order, what its neighbour expects; every row of a box that another rank holds reaches that rank."""
import ctypes as C

import numpy as np
import pytest

from falling_sand_engine_b200 import api, strips

W = 2048
GHOST = strips.GHOST


def _plan(L, H, rank, nranks, boxes):
    n = len(boxes)
    b = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
    runner = np.zeros(max(n, 1), dtype=np.int32)
    cap = 4 * n + 8
    rects = np.zeros((4, cap, 4), dtype=np.int32)
    cnt = np.zeros(4, dtype=np.int32)
    L.fse_strip_plan.argtypes = [C.c_int32] * 4 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    rc = L.fse_strip_plan(W, H, rank, nranks, b.ctypes.data, n, runner.ctypes.data, rects.ctypes.data, cap, cnt.ctypes.data)
    return rc, runner[:n].copy(), [rects[q, :cnt[q]].copy() for q in range(4)]


def _boxes(rng, H, nranks, n):
    """Boxes up to 48 rows tall all over the world, a third of them placed on the cuts; some overlap."""
    cuts = [strips.strip_layout(H, r, nranks)[0] for r in range(1, nranks)]
    out = []
    for i in range(n):
        w, h = int(rng.integers(4, 60)), int(rng.integers(4, 26))
        x0 = int(rng.integers(-10, W - 20))
        if i % 3 == 0 and cuts:
            y0 = int(cuts[int(rng.integers(len(cuts)))] + rng.integers(-h, 2))
        elif i % 7 == 1 and out:  # overlaps the previous one a little below it (a chain of two stays within the ghost rows)
            x0, y0 = out[-1][0] + 3, out[-1][1] + 4
        else:
            y0 = int(rng.integers(-5, H - 10))
        out.append((x0, y0, x0 + w - 1, y0 + h - 1))
    return out


@pytest.mark.parametrize("nranks,H", [(2, 1536), (3, 2048), (4, 1536), (8, 4352)])
def test_plans_of_all_ranks_agree(nranks, H):
    L = api.load_library()
    rng = np.random.default_rng(100 + nranks)
    lay = [strips.strip_layout(H, r, nranks) for r in range(nranks)]
    for trial in range(6):
        boxes = _boxes(rng, H, nranks, 60)
        plans = [_plan(L, H, r, nranks, boxes) for r in range(nranks)]
        assert all(p[0] == 0 for p in plans), api.load_library().fse_last_error()
        runner = plans[0][1]
        for r in range(1, nranks):
            assert np.array_equal(runner, plans[r][1])  # every rank derives the same runners
        for i, a in enumerate(boxes):
            e = runner[i]
            ya, yb = max(a[1], 0), min(a[3], H - 1)
            if ya <= yb:
                assert lay[e][2] <= ya and yb < lay[e][3]  # the runner holds the whole box
            for j in range(i):
                b = boxes[j]
                if a[0] <= b[2] and b[0] <= a[2] and a[1] <= b[3] and b[1] <= a[3]:
                    assert runner[i] == runner[j]  # overlapping boxes meet on one rank
        for r in range(nranks - 1):  # the cut between r and r + 1: what one side sends is what the other side expects
            up, dn = plans[r][2], plans[r + 1][2]
            for send, recv, so, ro in ((up[2], dn[1], lay[r][2], lay[r + 1][2]), (dn[0], up[3], lay[r + 1][2], lay[r][2])):
                assert len(send) == len(recv)
                gs, gr = send.copy(), recv.copy()
                gs[:, 1] += so  # local rows -> global rows
                gr[:, 1] += ro
                assert np.array_equal(gs, gr)
                for (x0, y0, w, h) in gs:
                    assert w > 0 and h > 0 and 0 <= x0 and x0 + w <= W
                    assert lay[r][2] <= y0 and y0 + h <= lay[r][3] and lay[r + 1][2] <= y0 and y0 + h <= lay[r + 1][3]  # both sides hold the rows
        # every row of a box that a rank other than its runner holds is delivered to that rank
        got = [set() for _ in range(nranks)]
        for r in range(nranks):
            for q in (1, 3):
                for (x0, y0, w, h) in plans[r][2][q]:
                    got[r].update((x0, y0 + lay[r][2] + k, w) for k in range(h))
        for i, a in enumerate(boxes):
            xa, xb = max(a[0], 0), min(a[2], W - 1)
            if xa > xb:
                continue
            for r in range(nranks):
                if r == runner[i]:
                    continue
                for y in range(max(a[1], 0, lay[r][2]), min(a[3], H - 1, lay[r][3] - 1) + 1):
                    assert (xa, y, xb - xa + 1) in got[r], (i, r, y)


def test_a_box_that_no_strip_can_hold_is_refused_on_every_rank():
    L = api.load_library()
    H, nranks = 1536, 4
    cut = strips.strip_layout(H, 1, nranks)[0]
    boxes = [(100, 200, 140, 230), (300, cut - 60, 340, cut + 60)]  # the second: 121 rows across a cut, more than one window's ghost rows
    rcs = [_plan(L, H, r, nranks, boxes)[0] for r in range(nranks)]
    assert len(set(rcs)) == 1 and rcs[0] != 0
    assert b"does not fit" in L.fse_last_error()
    # two boxes that each fit but overlap into a group that does not
    boxes = [(300, cut - 40, 340, cut - 4), (320, cut - 8, 360, cut + 38)]
    rcs = [_plan(L, H, r, nranks, boxes)[0] for r in range(nranks)]
    assert len(set(rcs)) == 1 and rcs[0] != 0


@pytest.mark.parametrize("nranks,H", [(2, 1536), (4, 1536), (8, 4352)])
def test_planned_exchange_brings_every_window_in_step(nranks, H):
    """The plan executed on data: every rank holds its window (own + ghost rows) of a global array; the runner of a box rewrites the box in
    its window; the planned rectangles are copied from the sender's window into the receiver's, in list order.  Afterwards every window
    must equal the global array with all boxes rewritten (later boxes over earlier ones where they overlap — they share a runner)."""
    L = api.load_library()
    rng = np.random.default_rng(7 + nranks)
    lay = [strips.strip_layout(H, r, nranks) for r in range(nranks)]
    for trial in range(4):
        boxes = _boxes(rng, H, nranks, 50)
        world = rng.integers(0, 1000, size=(H, W), dtype=np.int32)
        win = [world[lay[r][2]:lay[r][3]].copy() for r in range(nranks)]
        plans = [_plan(L, H, r, nranks, boxes) for r in range(nranks)]
        assert all(p[0] == 0 for p in plans)
        runner = plans[0][1]
        for i, (x0, y0, x1, y1) in enumerate(boxes):  # the edit: box i becomes 10000 + i, on the single world and in its runner's window
            xa, xb, ya, yb = max(x0, 0), min(x1, W - 1), max(y0, 0), min(y1, H - 1)
            if xa > xb or ya > yb:
                continue
            world[ya:yb + 1, xa:xb + 1] = 10000 + i
            e = runner[i]
            win[e][ya - lay[e][2]:yb + 1 - lay[e][2], xa:xb + 1] = 10000 + i
        msgs = {}
        for r in range(nranks):  # pack: up send (0) and down send (2), from the windows as the runners left them
            for q, peer in ((0, r - 1), (2, r + 1)):
                msgs[(r, peer)] = [win[r][y0:y0 + h, x0:x0 + w].copy() for (x0, y0, w, h) in plans[r][2][q]]
        for r in range(nranks):  # unpack: from above (1) and from below (3)
            for q, peer in ((1, r - 1), (3, r + 1)):
                rects = plans[r][2][q]
                if len(rects) == 0:
                    continue
                got = msgs[(peer, r)]
                assert len(got) == len(rects)
                for (x0, y0, w, h), data in zip(rects, got):
                    assert data.shape == (h, w)
                    win[r][y0:y0 + h, x0:x0 + w] = data
        for r in range(nranks):
            assert np.array_equal(win[r], world[lay[r][2]:lay[r][3]]), (trial, r)
