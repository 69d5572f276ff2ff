"""Bodies of tests/test_strips_gpu.py::test_strip_bodies_match_oracle: random ones all over the world plus, around every strip cut, a body
just above it, one just below it, and a pair that overlaps ACROSS the cut (cross-body order matters there, so one rank must run both)."""
import numpy as np

from falling_sand_engine_b200 import strips
from falling_sand_engine_b200 import types as T
from falling_sand_engine_b200 import worldgen as G
from tests.test_bridge_cpu import make_body


def scene(table, W, H, nranks):
    rng = np.random.default_rng(7)
    bodies, xf = [], []

    def add(x, y, ang, w=None, h=None, fill=0.8):
        w = int(rng.integers(12, 30)) if w is None else w
        h = int(rng.integers(12, 30)) if h is None else h
        bodies.append(make_body(table, w, h, seed=100 + len(bodies), fill=fill))
        xf.append((float(x), float(y), float(ang)))

    for _ in range(40):
        add(rng.uniform(160, W - 160), rng.uniform(160, H - 160), rng.uniform(-3.1, 3.1))
    for r in range(1, nranks):
        cut = strips.strip_layout(H, r, nranks)[0]
        x0 = 200 + 150 * r
        add(x0, cut - 22, 0.2, 20, 18)          # above the cut, reaches a few rows below it
        add(x0 + 120, cut + 2, -0.4, 16, 20)    # starts in the lower strip's first rows, reaches above the cut when turned
        add(x0 + 300, cut - 14, 0.0, 18, 16, fill=1.0)   # a pair that overlaps across the cut
        add(x0 + 306, cut + 1, 0.1, 18, 16, fill=1.0)
        add(x0 + 480, cut - 9, 1.2, 14, 14)     # drifts across the cut during the run: its runner changes
    return bodies, np.array(xf, dtype=np.float32)


def _cuts(H, nranks):
    return [strips.strip_layout(H, r, nranks)[0] for r in range(1, nranks)]


def stone_blocks(table, H, nranks):
    """(x0, y0, cells) of an 80 x 44 STONE block across every cut (what the pickaxe and the hammer of tool_calls hit)."""
    ids = G._names(table)
    out = []
    for k, cut in enumerate(_cuts(H, nranks)):
        mat = np.full((44, 80), ids["STONE"], dtype=np.uint16)
        out.append((60, cut - 22, G.cells_from_mat(table, mat, 60, cut - 22, 5 + k)))
        # two loose pieces of stone in a pocket of AIR, both lying across the cut: 30 cells (physicsCheck cuts it out into a body) and 6 (deleted)
        mat = np.full((30, 44), ids["AIR"], dtype=np.uint16)
        mat[12:17, 12:18] = ids["STONE"]
        mat[14:16, 30:33] = ids["STONE"]
        out.append((8, cut - 15, G.cells_from_mat(table, mat, 8, cut - 15, 9 + k)))
    return out


def tool_calls(H, nranks, tick):
    """Pickaxe circles and hammer cracks whose boxes cross the cuts."""
    calls = []
    for cut in _cuts(H, nranks):
        calls.append(("pickaxe", 70, cut - 9, 19.0))
        calls.append(("hammer", 100, cut - 12, 90, cut - 24, tick))   # the crack runs away from the target: down across the cut
        calls.append(("hammer", 126, cut + 10, 134, cut + 22, tick))  # ... and up across it
        calls.append(("physcheck", 22, cut - 1))   # the 30-cell piece: probed in the upper strip, reaches into the lower one
        calls.append(("physcheck", 39, cut))       # the 6-cell crumb: probed in the lower strip
        calls.append(("physcheck", 45, cut + 8))   # AIR: nothing
        calls.append(("vacuum", 80, cut, 130, cut, tick))          # along the cut inside the pickaxe's hole: the square it empties lies on both sides
        calls.append(("vacuum", 100, cut - 70, 104, cut + 60, tick))   # the walk starts in the upper strip and crosses the cut region
        calls.append(("vacuum", 120, cut + 80, 110, cut - 40, tick))   # ... and from below
    return calls


def run_tool(world_or_oracle, call, oracle_mod=None):
    """Runs one call of tool_calls on a (Strip)World, or on an OracleWorld through oracle/pyoracle; returns the numbers it reports."""
    kind = call[0]
    if kind == "pickaxe":
        pix, n = (oracle_mod.tool_pickaxe(world_or_oracle, *call[1:]) if oracle_mod else world_or_oracle.tool_pickaxe(*call[1:]))
        return np.concatenate([np.asarray(pix, dtype=np.int64).reshape(-1), [int(n)]])
    if kind == "physcheck":
        count, action, box, tiles = (oracle_mod.physics_check(world_or_oracle, *call[1:]) if oracle_mod else world_or_oracle.physics_check(*call[1:]))
        t = np.zeros(0, dtype=np.int64) if tiles is None else np.frombuffer(np.ascontiguousarray(tiles).tobytes(), dtype=np.uint8).astype(np.int64)
        return np.concatenate([np.array([count, action, *box], dtype=np.int64), t])
    if kind == "vacuum":
        wcx, wcy, wmx, wmy, tick = call[1:]
        res = (oracle_mod.tool_vacuum(world_or_oracle, wcx, wcy, wmx, wmy, tick=tick) if oracle_mod else world_or_oracle.tool_vacuum(wcx, wcy, wmx, wmy, tick=tick))
        return np.asarray(res, dtype=np.int64).reshape(-1)
    hx, hy, x, y, tick = call[1:]
    res = (oracle_mod.tool_hammer(world_or_oracle, hx, hy, x, y, tick=tick) if oracle_mod else world_or_oracle.tool_hammer(hx, hy, x, y, tick=tick))
    return np.asarray(res, dtype=np.int64).reshape(-1)


def entities(H, nranks):
    """Entities (players / NPCs: AABBs that collide with the grid, kick sand and are stamped as OBJECT cells) all over the world and, around
    every cut: one falling through it, one walking along it, and two whose reach overlaps across it (they see each other's kicks)."""
    rng = np.random.default_rng(3)
    rows = []
    cuts = _cuts(H, nranks)
    while len(rows) < 10:  # the random ones keep clear of the cuts: what happens AT the cuts is placed by hand below
        x, y = rng.uniform(300, 860), rng.uniform(160, H - 200)
        if all(abs(y - c) > 110 for c in cuts):
            rows.append((x, y, rng.uniform(-3, 3), rng.uniform(-3, 3), 8, 14, 0, 0))
    for cut in _cuts(H, nranks):
        rows.append((150.0, cut - 20.0, 0.4, 3.5, 8, 14, 0, 0))     # falls across the cut
        rows.append((185.0, cut - 7.0, 2.5, -0.2, 8, 14, 0, 0))     # walks along it, its box on both sides
        rows.append((240.0, cut - 16.0, 1.0, 1.0, 8, 14, 0, 0))     # two whose reach boxes overlap across the cut
        rows.append((246.0, cut - 9.0, -1.0, -1.5, 8, 14, 0, 0))
    return np.array(rows, dtype=T.ENTITY_DTYPE)
