"""Entities <-> grid: world::tickEntities (world.cpp:3010-3247), WorldEntitySystem::process
(game/player.cpp:173-199) and the objectDelete loop (game.cpp:2128-2139).

CPU part: pins of the oracle restatement derived from the reference source (no goldens exist: parity unpinned).
GPU part (`-m gpu`): the CUDA kernels against the oracle, bit-exact on grid, entities and kicked particles."""
import numpy as np
import pytest

from falling_sand_engine_b200 import types as T
from falling_sand_engine_b200 import worldgen as G
from tests import helpers as Hh

SAND, STONE, WATER, OBJECT = 2, 7, 15, 6


def _ents(*rows):
    e = np.zeros(len(rows), dtype=T.ENTITY_DTYPE)
    for i, (x, y, vx, vy, hw, hh) in enumerate(rows):
        e[i] = (x, y, vx, vy, hw, hh, 0, 0)
    return e


def _scene(table, W=512, H=512):
    """Stone floor with a one-cell ledge, a sand heap, a water pool, a stone ceiling with sand under it."""
    cells = Hh.empty_world_cells(table, W, H)
    mat = cells["mat"].copy()
    mat[300:304, 128:384] = STONE
    mat[299, 200:384] = STONE          # ledge: one cell up
    mat[290:299, 230:250] = SAND       # heap in the way
    mat[294:299, 300:340] = WATER
    mat[150:154, 128:384] = STONE      # ceiling
    mat[154:158, 160:200] = SAND       # grains hanging under it
    return G.cells_from_mat(table, mat, 0, 0, 3)


def test_entity_falls_lands_and_walks(oracle, table):
    """Gravity 0.25 per tick (3036), 8 sub-steps, collision halves the speed and sets `ground` (3177-3181); walking right over a
    one-cell ledge steps up (3052-3066); sand in the way is kicked into loose particles and slows the entity by 1 % per grain
    (3068-3074); velocity decays by 1 % per tick (3227-3229).  Grains are conserved (grid + particles)."""
    ow = oracle.OracleWorld(512, 512, table)
    ow.write_rect(0, 0, _scene(table))
    sand0 = int((ow.read_all()["mat"] == SAND).sum())
    e = _ents((150.0, 250.0, 0.0, 0.0, 8, 16))
    grounded = False
    for t in range(60):
        e = oracle.entities_tick(ow, e, tick=t)
        grounded |= bool(e["ground"][0])
    assert grounded and abs(float(e["y"][0]) + 16 - 300) <= 1.0  # feet on the floor at y = 300
    assert abs(float(e["vy"][0])) < 0.6 and not e["destroy"][0]
    e["vx"] = 1.5
    xs = []
    for t in range(60, 160):
        e["vx"] = 1.5  # the player keeps pushing right
        e = oracle.entities_tick(ow, e, tick=t)
        xs.append(float(e["x"][0]))
    assert xs[-1] > 260          # got past the ledge (x = 200) and through the heap (x = 230..250)
    assert float(e["y"][0]) + 16 <= 300.0  # stands on the ledge (row 299), one cell higher than the floor (row 300)
    p = ow.particles_read()
    assert len(p) > 20 and (p["tile"]["mat"] == SAND).all()
    assert int((ow.read_all()["mat"] == SAND).sum()) + len(p) == sand0
    assert len(np.unique(p["id"])) == len(p)


def test_entity_rising_kicks_sand_overhead_and_too_fast_is_destroyed(oracle, table):
    """vy < 0: sand overhead is kicked upwards (3190-3204), stone stops the entity; |v| >= 1024 destroys it (3222-3225)."""
    ow = oracle.OracleWorld(512, 512, table)
    ow.write_rect(0, 0, _scene(table))
    e = _ents((170.0, 170.0, 0.0, -6.0, 8, 10))
    for t in range(6):
        e = oracle.entities_tick(ow, e, tick=t)
    p = ow.particles_read()
    assert len(p) > 0 and (p["vy"] < 0.1).all()
    assert float(e["y"][0]) >= 154 - 0.01  # never entered the stone ceiling (rows 150..153)
    e2 = oracle.entities_tick(ow, _ents((300.0, 200.0, 2000.0, 0.0, 4, 4)), tick=9)
    assert e2["destroy"][0] == 1


def test_stamp_and_object_delete_round_trip(oracle, table):
    """AIR under an entity becomes Tiles_OBJECT (colour 0x00ff00) and goes back to Tiles_NOTHING at the end of the tick; SAND and
    SOUP are thrown up as particles first (player.cpp:183-194) — matter is conserved, the stamped cells are solid for the tick in
    between (OBJECT blocks entities, world.cpp:3022)."""
    ow = oracle.OracleWorld(512, 512, table)
    cells = _scene(table)
    ow.write_rect(0, 0, cells)
    before = ow.read_all()
    e = _ents((236.0, 285.0, 0.5, -1.0, 8, 16), (310.0, 290.0, 0.0, 0.0, 6, 8))  # one over the sand heap, one in the water
    oracle.entities_stamp(ow, e, tick=3)
    mid = ow.read_all()
    box = mid[285:301, 236:244]
    assert ((box["mat"] == OBJECT) | (box["mat"] == STONE)).all() and (box["color"][box["mat"] == OBJECT] == 0x00ff00).all()
    p = ow.particles_read()
    n_sand = int((before["mat"][285:301, 236:244] == SAND).sum())
    n_water = int((before["mat"][290:298, 310:316] == WATER).sum())
    assert n_sand > 0 and n_water > 0
    assert int((p["tile"]["mat"] == SAND).sum()) == n_sand and int((p["tile"]["mat"] == WATER).sum()) == n_water
    ow.tick(3)
    oracle.object_delete(ow)
    after = ow.read_all()
    assert not (after["mat"] == OBJECT).any()
    assert (after["mat"][285:299, 236:244] == 0).all()  # the box is empty: what was there is now a particle


@pytest.mark.gpu
def test_entities_match_oracle_on_gpu(oracle, gpu_ctx, table):
    """Game-loop order (game.cpp:1820-1838, 2128-2139): tickEntities, stamp, world tick, tickCells, object delete — several
    entities, two of them overlapping, on a mixed world; grid, particles and entity state bit-identical every tick."""
    import falling_sand_engine_b200 as fse

    W, H = 640, 512
    gpu_ctx.set_materials(table)
    gw, ow = fse.World(gpu_ctx, W, H), oracle.OracleWorld(W, H, table)
    cells = G.mixed_band(table, W, H, 0, H, seed=12, air_frac=0.55, blob=24)
    for w in (gw, ow):
        w.write_rect(0, 0, cells)
    rng = np.random.default_rng(3)
    rows = [(float(rng.uniform(140, 480)), float(rng.uniform(140, 340)), float(rng.uniform(-3, 3)), float(rng.uniform(-4, 2)), int(rng.integers(4, 15)),
             int(rng.integers(6, 27))) for _ in range(7)]
    rows.append((rows[0][0] + 3.0, rows[0][1] + 2.0, -1.0, 0.5, 10, 12))  # overlaps entity 0: array order decides
    rows.append((300.25, 200.75, 0.0, 0.0, 40, 60))                        # a big box
    eg, eo = _ents(*rows), _ents(*rows)
    lz = (1.0, -2.0)
    for t in range(12):
        eg = gw.entities_tick(eg, load_zone=lz, tick=t)
        eo = oracle.entities_tick(ow, eo, load_zone=lz, tick=t)
        assert eg.tobytes() == eo.tobytes(), (t, eg, eo)
        gw.entities_stamp(eg, load_zone=lz, tick=t)
        oracle.entities_stamp(ow, eo, load_zone=lz, tick=t)
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"stamp {t}")
        for w in (gw, ow):
            w.tick(t)
            w.particles_tick()
        gw.object_delete()
        oracle.object_delete(ow)
        Hh.assert_cells_equal(ow.read_all(), gw.read_all(), f"tick {t}")
        Hh.assert_particles_equal(ow.particles_read(), gw.particles_read(), f"tick {t}")
        if t % 3 == 0:  # the host's input: a push sideways, a jump
            for e in (eg, eo):
                e["vx"] += np.float32(1.25)
                e["vy"][::2] -= np.float32(3.0)
    assert len(ow.particles_read()) >= 0 and (np.abs(eo["vx"]) < 1024).all()
