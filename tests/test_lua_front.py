"""The scripts' front door on a real Lua VM (north_star: "Lua-defined material tables"): csrc/lua_front.c on the reference's vendored
Lua 5.4.4 (libfse_lua.so, built by csrc/Makefile.lua from /root/reference/source/libs/lua where it lies).  The reference binds
materials_init / materials_register / materials_push (game_basic.cpp:79-81) and reads its settings back from `global_def`
(cvar.cpp:57-99, data/scripts/global.lua).  CPU only."""
import os

import pytest

from falling_sand_engine_b200 import materials as M
from falling_sand_engine_b200 import types as T

pytestmark = pytest.mark.skipif(M.lua_library() is None, reason="libfse_lua.so not built (needs the reference tree: make -f csrc/Makefile.lua)")

LITERAL = """
-- materials of a mod
local GLOW = 0x40FFAA00
OnGameEngineLoad = function()
    materials_init()
    materials_register(1001, "Test Ash", "TEST_ASH", SAND, 12, 255, 6.5, 2, 0, 0, 0x555555)
    materials_register(1002, 'Glow Oil', 'GLOW_OIL', 3, 0, 0xC0, 1.2, 4, 8, GLOW, 0x332211) --[[ SOUP ]]
    materials_push()
end
"""

PROGRAM = """
global_def = {}
global_def.cell_iter = 2 + 1
global_def.brush_size = 5
global_def.tick_temperature = false
local base = { {"Ash", SAND, 12, 6.5}, {"Oil", SOUP, 0, 1.2}, {"Fog", GAS, 0, -0.5} }
local function reg(i, t)
    materials_register(1000 + i, t[1], string.upper(t[1]), t[2], t[3], 255, t[4], i + 1, 0, 0, 0x101010 * i)
end
OnGameEngineLoad = function()
    InitGraphics(); InitAudio(); textures_load("a", "b"); controls_init()   -- engine calls the material path does not need
    materials_init()
    for i, t in ipairs(base) do reg(i, t) end
    for k = 1, 4 do materials_register(2000 + k, "Grain" .. k, "GRAIN_" .. k, SAND, 8 + k, 255, 4.0 + k / 2, 2, 0, 0, 0xC0B0A0 + k) end
    materials_push()
end
"""


def test_vm_and_declarative_reader_agree_on_literal_scripts():
    vm_tbl, vm_ids = M.load_lua(LITERAL, engine="vm")
    de_tbl, de_ids = M.load_lua(LITERAL, engine="declarative")
    assert vm_ids == de_ids and vm_tbl.n == de_tbl.n == M.default_materials(1337).n + 2
    assert bytes(vm_tbl.mats) == bytes(de_tbl.mats) and list(vm_tbl.inter_offsets) == list(de_tbl.inter_offsets)
    oil = vm_tbl.mats[vm_ids[1002]]
    assert (oil.physics, oil.alpha, oil.iterations, oil.emit, oil.emit_color, oil.color) == (T.SOUP, 0xC0, 4, 8, 0x40FFAA00, 0x332211)


def test_vm_runs_loops_functions_and_engine_calls():
    """What the declarative reader cannot do: registrations from loops over tables, string functions, arithmetic — and the other
    engine functions a game script calls on the way (they resolve to no-ops)."""
    tbl, ids = M.load_lua(PROGRAM, engine="vm")
    n0 = M.default_materials(1337).n
    assert tbl.n == n0 + 7
    assert [ids[k] for k in (1001, 1002, 1003, 2001, 2004)] == [n0, n0 + 1, n0 + 2, n0 + 3, n0 + 6]
    assert ids["Grain3"] == n0 + 5 and tbl.mats[ids["Fog"]].physics == T.GAS and tbl.mats[ids[1003]].iterations == 4
    g4 = tbl.mats[ids[2004]]
    assert (g4.physics, g4.slipperyness, g4.color) == (T.SAND, 12, 0xC0B0A0 + 4) and abs(g4.density - 6.0) < 1e-6
    with pytest.raises(ValueError):  # the declarative reader sees no literal registration calls it could use
        M.load_lua(PROGRAM, engine="declarative")
    recs, res = M.lua_run(PROGRAM)
    assert (res.n_init, res.n_register, res.n_push) == (1, 7, 1) and recs[0].index_name == b"ASH"
    assert M.load_global_def(PROGRAM) == {"cell_iter": 3, "brush_size": 5, "tick_world": -1, "tick_box2d": -1, "tick_temperature": 0}


def test_lua_errors_surface():
    with pytest.raises(ValueError, match="11 arguments"):
        M.lua_run("materials_init() materials_register(1, 'x') materials_push()")
    with pytest.raises(ValueError, match="lua:"):
        M.lua_run("this is not lua")
    with pytest.raises(ValueError, match="materials_init"):
        M.load_lua("materials_push()", engine="vm")


@pytest.mark.skipif(not os.path.exists("/root/reference/data/scripts/global.lua"), reason="reference tree not present")
def test_reads_cell_iter_from_the_references_own_global_lua():
    """cvar.cpp:98-99 reads cell_iter and brush_size from the script's global_def table; data/scripts/global.lua:47 sets cell_iter = 3."""
    d = M.load_global_def("/root/reference/data/scripts/global.lua", is_file=True)
    assert d["cell_iter"] == 3 and d["brush_size"] == 5 and d["tick_world"] == 1 and d["tick_temperature"] == 1
