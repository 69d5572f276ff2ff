"""CPU pins of the rigid-body bridge and outline oracle (SURVEY.md §8c pins 6 and 7)."""
import numpy as np
import pytest

from falling_sand_engine_b200 import types as T
from falling_sand_engine_b200 import worldgen as G
from tests import helpers as Hh


def make_body(table, w, h, mat=22, seed=1, fill=0.7):
    """A w x h body of `mat` (OBSIDIAN) with hashed holes."""
    hh = G.hash2(seed, np.arange(w, dtype=np.uint32)[None, :], np.arange(h, dtype=np.uint32)[:, None])
    m = np.where((hh % np.uint32(100)) < int(fill * 100), mat, 0).astype(np.uint16)
    return G.cells_from_mat(table, np.broadcast_to(m, (h, w)).copy(), 0, 0, seed)


def test_marching_squares_square_and_orientation(oracle):
    """Pin 6: a filled rectangle gives one closed 4-corner contour; Douglas-Peucker keeps >= 2 points and the corners."""
    m = np.zeros((12, 16), dtype=np.uint8)
    m[3:9, 4:12] = 1
    cs = oracle.outlines(m)
    assert len(cs) == 1
    pts = cs[0]
    # the reference starts at the first pixel whose right/down/diagonal neighbours are not all set — the top-right
    # pixel (11,3) — and keeps the first and last vertex of the walk, so the start vertex stays on the top edge
    # (vertex value 12 -> West, physics_math.cpp:1939); Douglas-Peucker (tolerance 1) then drops the corner (12,3), which
    # lies 0.99 px from the chord (12,9)-(11,3)
    assert pts.tolist() == [[4.0, 3.0], [4.0, 9.0], [12.0, 9.0], [11.0, 3.0]]
    area2 = float(np.sum(pts[:, 0] * np.roll(pts[:, 1], -1) - np.roll(pts[:, 0], -1) * pts[:, 1]))
    assert abs(abs(area2) - 2 * 8 * 6) <= 2 * 3.5


def test_marching_squares_all_16_cases_close(oracle):
    """Every 2x2 neighbourhood case occurs in a random mask; each traced loop is closed, axis-aligned and simplification
    never invents points."""
    rng = np.random.default_rng(3)
    m = (rng.random((40, 40)) < 0.55).astype(np.uint8)
    cs = oracle.outlines(m)
    assert len(cs) >= 5
    for pts in cs:
        assert len(pts) >= 3
        assert (pts >= 0).all() and (pts[:, 0] <= 40).all() and (pts[:, 1] <= 40).all()
        assert np.all(pts == np.round(pts))


def test_ccl_counts_components(oracle):
    m = np.zeros((10, 10), dtype=np.uint8)
    m[0:3, 0:3] = 1
    m[5:8, 5:8] = 1
    m[3, 3] = 1   # touches the first block only diagonally: a separate 4-connected component
    labels, n = oracle.ccl(m)
    assert n == 3
    assert labels[0, 0] == 0 and labels[2, 2] == 0 and labels[3, 3] == 33 and labels[5, 5] == 55 and labels[9, 9] == -1


def test_raster_erase_round_trip_is_identity(oracle, table):
    """Pin 7: raster then erase of a static body on an empty grid restores both the grid and the body."""
    W = H = 384
    ow = oracle.OracleWorld(W, H, table)
    ow.write_rect(0, 0, Hh.empty_world_cells(table, W, H))
    before = ow.read_all()
    body = make_body(table, 24, 20, fill=1.0)
    orig = body.copy()
    for angle in (0.0, 0.4, 1.3):
        xf = [(180.0, 170.0, angle)]
        fb = oracle.bodies_raster(ow, [body], xf)
        assert fb[0, 2] > 0.9 * 24 * 20
        grid = ow.read_all()
        assert int((grid["mat"] == 22).sum()) == fb[0, 2]
        fb2 = oracle.bodies_erase(ow, [body], xf)
        assert fb2[0, 2] == fb[0, 2]
        after = ow.read_all()
        for f in ("mat", "color", "temp", "fluid"):
            assert np.array_equal(after[f], before[f]), (angle, f)
        if angle == 0.0:
            assert fb[0, 2] == 24 * 20 and body.tobytes() == orig.tobytes()
        body = orig.copy()


def test_raster_displaces_sand_into_particles(oracle, table):
    W = H = 384
    ow = oracle.OracleWorld(W, H, table)
    ow.write_rect(0, 0, Hh.empty_world_cells(table, W, H))
    ow.write_rect(170, 170, G.cells_from_mat(table, np.full((10, 40), 2, dtype=np.uint16), 170, 170))
    body = make_body(table, 16, 16, fill=1.0)
    fb = oracle.bodies_raster(ow, [body], [(180.0, 165.0, 0.0)])
    assert fb[0, 0] > 0 and ow.particles_count() == fb[0, 0]
    assert int((ow.read_all()["mat"] == 2).sum()) + ow.particles_count() == 400


def test_flood_component_sizes(oracle, table):
    W = H = 384
    ow = oracle.OracleWorld(W, H, table)
    ow.write_rect(0, 0, Hh.empty_world_cells(table, W, H))
    ow.write_rect(200, 200, G.cells_from_mat(table, np.full((5, 7), 7, dtype=np.uint16), 200, 200))
    n, bbox, pix = oracle.flood_component(ow, 203, 202)
    assert n == 35 and list(bbox) == [200, 200, 206, 204] and len(pix) == 35
    assert oracle.flood_component(ow, 150, 150)[0] == 0          # AIR seed
    assert oracle.flood_component(ow, 10, 10)[0] == 1001         # the solid border is larger than the cap


def test_explosion_conserves_what_it_throws(oracle, table):
    """world.cpp:2294-2332: inside the radius SOLID always vanishes and other cells vanish 6 times in 10; survivors and every
    non-SOLID cell of the ring out to 2r leave as particles.  AIR afterwards inside r; SOLID untouched in the ring."""
    from oracle import pyoracle as O
    W = H = 256
    ow = oracle.OracleWorld(W, H, table)
    SAND, STONE, AIR = 2, 7, 0
    mat = np.full((H, W), SAND, dtype=np.uint16)
    mat[:, 128:] = STONE
    cells = G.cells_from_mat(table, mat, 0, 0, 3)
    ow.write_rect(0, 0, cells)
    r = 20
    O.explosion(ow, 128, 128, r, tick=5, seed=9)
    after = ow.read_all()
    parts = ow.particles_read()
    ys, xs = np.mgrid[0:H, 0:W]
    d2 = (xs - 128) ** 2 + (ys - 128) ** 2
    sq = (np.abs(xs - 128 + 0.5) < 2 * r) & (np.abs(ys - 128 + 0.5) < 2 * r)  # the loop covers [c-2r, c+2r)
    inner = d2 < r * r
    ring = (~inner) & (d2 < 4 * r * r) & sq
    assert (after["mat"][inner & sq] == AIR).all()
    assert (after["mat"][ring & (mat == STONE)] == STONE).all()
    assert (after["mat"][ring & (mat == SAND)] == AIR).all()
    assert (after["mat"][~(inner | ring)] == mat[~(inner | ring)]).all()
    n_ring = int((ring & (mat == SAND)).sum())
    n_inner_sand = int((inner & sq & (mat == SAND)).sum())
    n_parts = len(parts)
    assert n_ring <= n_parts <= n_ring + n_inner_sand
    kept = n_parts - n_ring
    assert 0.3 * n_inner_sand < kept < 0.5 * n_inner_sand  # 4 in 10 survive as particles
    assert (parts["tile"]["mat"] == SAND).all()
    assert len(np.unique(parts["id"])) == n_parts


def test_render_dirty_known_answers(oracle, table):
    """game.cpp:1994-2060 on hand-made cells: byte order r, g, b, a; AIR clears all three planes; only FIRE writes the fire plane
    (other materials leave it alone); emission is Material::emitColor; clean cells are not touched; movingTiles counts per material."""
    from oracle import pyoracle as O
    W, H = 8, 4
    ow = oracle.OracleWorld(W, H, table)
    cells = Hh.empty_world_cells(table, W, H)
    cells["mat"][1, 1], cells["color"][1, 1] = 15, 0x112233   # WATER: alpha 0x80, emit 0x3000AFB5
    cells["mat"][1, 2], cells["color"][1, 2] = 25, 0xFF6432   # FIRE: alpha 255
    cells["mat"][1, 3] = 0                                    # AIR
    cells["mat"][1, 4], cells["color"][1, 4] = 2, 0xABCDEF    # SAND, written clean below
    cells["dirty"][:] = 0
    cells["dirty"][1, 1:4] = 1
    ow.write_rect(0, 0, cells)
    planes = [np.full((H, W, 4), 7, dtype=np.uint8) for _ in range(3)]
    dirty, fire, moving = O.render_dirty(ow, planes)
    assert (dirty, fire) == (3, 1) and moving[15] == 1 and moving[25] == 1 and moving[0] == 1 and moving.sum() == 3
    assert planes[0][1, 1].tolist() == [0x11, 0x22, 0x33, 0x80] and planes[2][1, 1].tolist() == [0x00, 0xAF, 0xB5, 0x30]
    assert planes[1][1, 1].tolist() == [7, 7, 7, 7]                                   # not FIRE: fire plane untouched
    assert planes[0][1, 2].tolist() == [0xFF, 0x64, 0x32, 255] == planes[1][1, 2].tolist()
    assert all(p[1, 3].tolist() == [0, 0, 0, 0] for p in planes)                      # AIR: transparent black everywhere
    assert all(p[1, 4].tolist() == [7, 7, 7, 7] for p in planes)                      # clean cell: untouched
    assert ow.read_all()["dirty"].sum() == 3                                          # dirty flags stay (game.cpp:2153 clears them later)


@pytest.mark.parametrize("dx,dy", [(3, 0), (-2, 1), (0, -3), (5, 4), (-300, 0)])
def test_scroll_is_a_shift_that_keeps_unsourced_cells(oracle, table, dx, dy):
    """world.cpp:2454-2478: every cell moves by (dx, dy); cells without a source inside the world keep their content; particles
    move along (2579-2582)."""
    from oracle import pyoracle as O
    W, H = 40, 24
    ow = oracle.OracleWorld(W, H, table)
    rng = np.random.default_rng(1)
    mat = rng.integers(0, 28, size=(H, W)).astype(np.uint16)
    cells = G.cells_from_mat(table, mat, 0, 0, 5)
    cells["temp"] = rng.integers(-100, 100, size=(H, W))
    ow.write_rect(0, 0, cells)
    parts = np.zeros(2, dtype=T.PARTICLE_DTYPE)
    parts["x"], parts["y"], parts["id"] = [5.5, 30.0], [7.25, 20.0], [1, 2]
    ow.particles_add(parts)
    before = ow.read_all()
    O.scroll(ow, dx, dy)
    after = ow.read_all()
    ys, xs = np.mgrid[0:H, 0:W]
    sy, sx = ys - dy, xs - dx
    ok = (sx >= 0) & (sx < W) & (sy >= 0) & (sy < H)
    for f in ("mat", "color", "temp", "fluid", "moved", "settle"):
        want = before[f].copy()
        want[ok] = before[f][sy[ok], sx[ok]]
        assert np.array_equal(after[f], want), f
    p = ow.particles_read()
    assert np.allclose(np.sort(p["x"]), np.sort(parts["x"] + dx)) and np.allclose(np.sort(p["y"]), np.sort(parts["y"] + dy))


def test_body_pixels_with_overlapping_footprints_are_within_the_7x7_stencil():
    """The fast path of the rigid-body bridge (fse_bodies.cu) finds the earlier pixels a pixel depends on in its 7 x 7 body
    neighbourhood.  Pin the geometry it relies on with the reference's own transform (game.cpp:1763-1764, float32, int()
    truncation): two plus-shaped footprints share a cell exactly when their centres are at Manhattan distance <= 2, and body
    pixels whose centres land that close are never more than 3 apart in either body coordinate."""
    rng = np.random.default_rng(3)
    w = h = 40
    tx, ty = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32), indexing="ij")
    plus = [(0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)]
    for trial in range(200):
        ang = np.float32(rng.uniform(-np.pi, np.pi))
        x0, y0 = np.float32(rng.uniform(50, 4000)), np.float32(rng.uniform(50, 4000))
        s, c = np.float32(np.sin(ang)), np.float32(np.cos(ang))
        wx = (tx * c - (ty + np.float32(1)) * s + x0).astype(np.int32).ravel()  # same operation order as the kernel and the oracle
        wy = (tx * s + (ty + np.float32(1)) * c + y0).astype(np.int32).ravel()
        bx, by = tx.astype(np.int32).ravel(), ty.astype(np.int32).ravel()
        # pairs whose footprints share a cell: bucket pixels by covered cell
        cover = {}
        for dx, dy in plus:
            for i, key in enumerate(zip((wx + dx).tolist(), (wy + dy).tolist())):
                cover.setdefault(key, []).append(i)
        worst = 0
        for idx in cover.values():
            if len(idx) < 2:
                continue
            a = np.array(idx)
            man = np.abs(wx[a][:, None] - wx[a][None, :]) + np.abs(wy[a][:, None] - wy[a][None, :])
            assert man.max() <= 2  # sharing a cell implies Manhattan distance <= 2 between the centres
            worst = max(worst, int(np.abs(bx[a][:, None] - bx[a][None, :]).max()), int(np.abs(by[a][:, None] - by[a][None, :]).max()))
        assert worst <= 3, (trial, float(ang), worst)
    # and the converse used by the kernel's test: Manhattan distance <= 2 means the plus shapes do intersect
    for dx in range(-3, 4):
        for dy in range(-3, 4):
            a = {(px, py) for px, py in plus}
            b = {(dx + px, dy + py) for px, py in plus}
            assert bool(a & b) == (abs(dx) + abs(dy) <= 2), (dx, dy)


def test_body_split_cracked_plate(oracle, table):
    """Fracture hand-off (world::updateRigidBodyHitbox, world.cpp:288-720) on the oracle: a 40 x 24 plate cut by a one-pixel crack and
    with a loose crumb gives three pieces in first-pixel order; each is cropped to its bounding box (305-320), the pieces' tiles
    union to the original plate pixel for pixel, the weld pixel flags exactly one piece (620) and the shift is the rotated box
    corner (350-362)."""
    plate = make_body(table, 40, 24, fill=1.0)
    plate["mat"][:, 17] = 0            # the crack
    plate["mat"][3:6, 25:28] = 0       # a hole with ...
    plate["mat"][4, 26] = 22           # ... a crumb inside
    plate["color"] = np.arange(24 * 40, dtype=np.uint32).reshape(24, 40) + 7
    angle = 0.3
    pieces = oracle.body_split(plate, angle=angle, weld=(30, 20))
    assert len(pieces) == 3
    recs = [p for p, _ in pieces]
    assert [(int(r["x0"]), int(r["y0"]), int(r["w"]), int(r["h"])) for r in recs] == [(0, 0, 17, 24), (18, 0, 22, 24), (26, 4, 1, 1)]
    assert [int(r["weld"]) for r in recs] == [0, 1, 0]
    assert [int(r["n_pixels"]) for r in recs] == [17 * 24, 22 * 24 - 9, 1]
    back = np.zeros_like(plate)
    back["fluid"] = 2.0
    for r, t in pieces:
        sel = t["mat"] != 0
        view = back[r["y0"]:r["y0"] + r["h"], r["x0"]:r["x0"] + r["w"]]
        view[sel] = t[sel]
        assert np.isclose(r["shift_x"], r["x0"] * np.cos(angle) - r["y0"] * np.sin(angle), atol=1e-5)
        assert np.isclose(r["shift_y"], r["x0"] * np.sin(angle) + r["y0"] * np.cos(angle), atol=1e-5)
    want = plate.copy()
    for f in Hh.FIELDS:
        assert np.array_equal(back[f][plate["mat"] != 0], want[f][plate["mat"] != 0]), f
    assert (back["mat"][plate["mat"] == 0] == 0).all()


def _physcheck_world(table, W=768, H=640):
    """Floating blobs over AIR: a 12 x 9 block with a hole (105 cells), a 3-cell crumb, a 40 x 30 block (1200 cells: too big) and a thin
    L of 60 cells that sprawls over a 30 x 31 box."""
    cells = Hh.empty_world_cells(table, W, H)
    mat = cells["mat"]
    mat[200:209, 300:312] = 7
    mat[203:206, 304:305] = 0
    mat[250, 160:163] = 11
    mat[300:330, 400:440] = 22
    mat[350:381, 200] = 7
    mat[380, 200:230] = 7
    cells["color"] = (np.arange(W * H, dtype=np.uint32).reshape(H, W) * np.uint32(2246822519)) & np.uint32(0xFFFFFF)
    return cells


def test_physics_check_known_answers(oracle, table):
    """world::physicsCheck (world.cpp:3330-3411): not SOLID -> nothing; 1..10 cells -> deleted; 11..1000 -> cut out (cells become
    Tiles_NOTHING and dirty, the body's tiles are OBSIDIAN with the cells' colours inside the bounding box, AIR elsewhere); more than
    1000 -> the flood is abandoned and nothing changes."""
    from oracle import pyoracle as O
    W, H = 768, 640
    ow = oracle.OracleWorld(W, H, table)
    cells = _physcheck_world(table, W, H)
    ow.write_rect(0, 0, cells)
    ow.clear_dirty()
    assert O.physics_check(ow, 150, 150)[:2] == (0, 0)                    # AIR seed
    assert O.physics_check(ow, 420, 315)[:2] == (1001, 0)                 # 1200 cells: abandoned
    assert O.physics_check(ow, 5, 5)[:2] == (1001, 0)                     # the world's border
    assert np.array_equal(ow.read_all()["mat"], cells["mat"]) and not ow.read_all()["dirty"].any()
    n, act, box, tiles = O.physics_check(ow, 161, 250)
    assert (n, act, box, tiles) == (3, 1, (160, 250, 3, 1), None)
    after = ow.read_all()
    assert (after["mat"][250, 160:163] == 0).all() and after["dirty"][250, 160:163].all() and after["dirty"].sum() == 3
    n, act, box, tiles = O.physics_check(ow, 305, 208)
    assert (n, act, box) == (12 * 9 - 3, 2, (300, 200, 12, 9)) and tiles.shape == (9, 12)
    want = cells["mat"][200:209, 300:312] != 0
    assert np.array_equal(tiles["mat"] == 22, want) and (tiles["mat"][~want] == 0).all()   # OBSIDIAN where the component was
    assert np.array_equal(tiles["color"][want], cells["color"][200:209, 300:312][want])
    assert (tiles["fluid"] == 2.0).all()
    after = ow.read_all()
    assert (after["mat"][200:209, 300:312] == 0).all() and int(after["dirty"].sum()) == 3 + 105
    n, act, box, tiles = O.physics_check(ow, 200, 360)
    assert (n, act, box) == (60, 2, (200, 350, 30, 31)) and int((tiles["mat"] == 22).sum()) == 60
    ow.write_rect(0, 0, cells)
    with pytest.raises(ValueError):
        O.physics_check(ow, 200, 360, cap_tiles=100)
    assert np.array_equal(ow.read_all()["mat"], cells["mat"])             # too small a tile buffer: nothing was changed
