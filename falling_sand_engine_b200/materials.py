"""The stock material table: InitMaterials()/RegisterMaterial()/PushMaterials()
(source/engine/game_datastruct.cpp:69-300) flattened into the arguments of fse_materials_set.

The reference draws its ten `Mat_0..9` materials and their interactions from libc rand() seeded with
time(NULL) (game_utils/rng.cpp:9); here they come from a counter generator on `seed`, so a table is
reproducible.  Texture-sampled colours (TilesCreate*, gds.cpp:344-470) become POSITIONAL colours.
"""
import ctypes as C

from . import types as T

NAMES = ["GENERIC_AIR", "GENERIC_SOLID", "GENERIC_SAND", "GENERIC_LIQUID", "GENERIC_GAS", "GENERIC_PASSABLE", "GENERIC_OBJECT",
         "STONE", "GRASS", "DIRT", "SMOOTH_STONE", "COBBLE_STONE", "SMOOTH_DIRT", "COBBLE_DIRT", "SOFT_DIRT", "WATER", "LAVA",
         "CLOUD", "GOLD_ORE", "GOLD_MOLTEN", "GOLD_SOLID", "IRON_ORE", "OBSIDIAN", "STEAM", "SOFT_DIRT_SAND", "FIRE",
         "FLAT_COBBLE_STONE", "FLAT_COBBLE_DIRT"]
ID = {n: i for i, n in enumerate(NAMES)}

_M = 0xFFFFFFFF


def _mix32(v):
    v &= _M
    v ^= v >> 16
    v = (v * 0x7FEB352D) & _M
    v ^= v >> 15
    v = (v * 0x846CA68B) & _M
    v ^= v >> 16
    return v


class _Rand:
    def __init__(self, seed):
        self.st = _mix32(seed ^ 0xA5A5A5A5)

    def __call__(self):
        self.st = _mix32(self.st + 0x9E3779B9)
        return self.st >> 1


def _mat(phys, slip, alpha, dens, iters, emit=0, emit_color=0, color=0xFFFFFFFF, kind=T.COLOR_FIXED, jshift=0, jrange=0, ctemp=0):
    m = T.Material()
    m.physics, m.slipperyness, m.alpha, m.density, m.iterations = phys, slip, alpha, dens, iters
    m.emit, m.emit_color, m.color, m.color_kind = emit, emit_color, color, kind
    m.jitter_shift, m.jitter_range, m.create_temp = jshift, jrange, ctemp
    m.conduction_self = m.conduction_other = 1.0
    return m


def default_materials(seed=1337):
    """InitMaterials() (gds.cpp:117-280) -> MaterialTable."""
    FIX, JIT, POS = T.COLOR_FIXED, T.COLOR_JITTER, T.COLOR_POSITIONAL
    A, SO, SA, LI, GA, PA = T.AIR, T.SOLID, T.SAND, T.SOUP, T.GAS, T.PASSABLE
    mats = [
        _mat(A, 0, 255, 0, 0, 16, 0, 0x000000),            # 0 GENERIC_AIR       gds.cpp:72
        _mat(SO, 0, 255, 1, 0),                            # 1 GENERIC_SOLID
        _mat(SA, 20, 255, 10, 2),                          # 2 GENERIC_SAND
        _mat(LI, 0, 255, 1.5, 3),                          # 3 GENERIC_LIQUID
        _mat(GA, 0, 255, -1, 1),                           # 4 GENERIC_GAS
        _mat(PA, 0, 255, 0, 0),                            # 5 GENERIC_PASSABLE
        _mat(T.OBJECT, 0, 255, 1000.0, 0),                 # 6 GENERIC_OBJECT
        _mat(SO, 0, 255, 1, 0, color=0x808080, kind=POS),  # 7 STONE             gds.cpp:81
        _mat(SA, 20, 255, 12, 1, color=(40 << 16) + (120 << 8) + 20, kind=JIT, jshift=8, jrange=20),   # 8 GRASS gds.cpp:358
        _mat(SA, 8, 255, 15, 1, color=(60 << 16) + (40 << 8) + 20, kind=JIT, jshift=16, jrange=10),    # 9 DIRT  gds.cpp:365
        _mat(SO, 0, 255, 1, 0, color=0x888888, kind=POS),  # 10 SMOOTH_STONE
        _mat(SO, 0, 255, 1, 0, color=0x6B6B6B, kind=POS),  # 11 COBBLE_STONE
        _mat(SO, 0, 255, 1, 0, color=0x6B4A2F, kind=POS),  # 12 SMOOTH_DIRT
        _mat(SO, 0, 255, 1, 0, color=0x5A3D26, kind=POS),  # 13 COBBLE_DIRT
        _mat(SO, 0, 255, 15, 2, color=0x7A5533, kind=POS),  # 14 SOFT_DIRT
        _mat(LI, 0, 0x80, 1.5, 6, 40, 0x3000AFB5, 0x00B69F, ctemp=-1023),  # 15 WATER  gds.cpp:92,427
        _mat(LI, 0, 0xC0, 2, 1, 40, 0xFFFF6900, 0xFF7C00, ctemp=1024),     # 16 LAVA   gds.cpp:93,433
        _mat(SO, 0, 127, 1, 0, color=0xF0F0F0, kind=POS),  # 17 CLOUD
        _mat(SA, 20, 255, 20, 2, 8, 0x804000, 0xD4AF37, POS),      # 18 GOLD_ORE
        _mat(LI, 0, 255, 20, 2, 8, 0x6FFF9B40, 0xFFC84A, POS),     # 19 GOLD_MOLTEN
        _mat(SO, 0, 255, 20, 2, 8, 0, 0xFFD700, POS),              # 20 GOLD_SOLID
        _mat(SA, 20, 255, 20, 2, 8, 0x7F442F, 0x8A5A44, POS),      # 21 IRON_ORE
        _mat(SO, 0, 255, 1, 0, color=0x2A1A3A, kind=POS),  # 22 OBSIDIAN
        _mat(GA, 0, 255, -1, 1, color=0x666666),           # 23 STEAM            gds.cpp:475
        _mat(SA, 8, 255, 15, 2),                           # 24 SOFT_DIRT_SAND
        _mat(PA, 0, 255, 20, 1, color=(255 << 16) + (100 << 8) + 50, kind=JIT, jshift=8, jrange=50),   # 25 FIRE gds.cpp:477
        _mat(SO, 0, 255, 1, 0, color=0x707070, kind=POS),  # 26 FLAT_COBBLE_STONE
        _mat(SO, 0, 255, 1, 0, color=0x5E4128, kind=POS),  # 27 FLAT_COBBLE_DIRT
    ]
    # gds.cpp:124-141
    mats[ID["GENERIC_AIR"]].conduction_self = mats[ID["GENERIC_AIR"]].conduction_other = 0.8
    mats[ID["LAVA"]].conduction_self, mats[ID["LAVA"]].conduction_other, mats[ID["LAVA"]].add_temp = 0.5, 0.7, 2
    mats[ID["COBBLE_STONE"]].conduction_self, mats[ID["COBBLE_STONE"]].conduction_other = 0.01, 0.4

    rnd = _Rand(seed)
    rand0 = len(mats)
    for _ in range(10):  # gds.cpp:172-191
        rgb = rnd() % 255
        rgb = (rgb << 8) + rnd() % 255
        rgb = (rgb << 8) + rnd() % 255
        typ = (SA if rnd() % 2 == 0 else GA) if rnd() % 2 == 0 else LI
        base = {SA: 5, LI: 4, GA: 3}[typ]
        dens = C.c_float(base + (rnd() % 1000) / 1000.0).value
        alpha = 255 if typ == SA else rnd() % 192 + 63
        mats.append(_mat(typ, 10, alpha, dens, rnd() % 4 + 1, color=rgb))
    # gds.cpp:263-269 scriptable test materials (ids 38..40)
    mats.append(_mat(SA, 20, 255, 10, 2, color=(220 << 16) + (155 << 8) + 100, kind=JIT, jshift=8, jrange=30))
    mats.append(_mat(SA, 20, 255, 10, 2, color=0xDCB464, kind=POS))
    mats.append(_mat(LI, 0, 255, 1.5, 4, color=0x0000FF))
    n = len(mats)

    pair = {}
    for i in range(10):  # gds.cpp:207-239
        mid = rand0 + i
        for _ in range(rnd() % 3 + 1):
            while True:
                imat = rand0 + rnd() % 10
                if imat != mid:
                    ty = rnd() % 2 + 1
                    prod = rand0 + rnd() % 10
                    rad = rnd() % 4
                    ox = rnd() % 5 - 2
                    oy = rnd() % 5 - 2
                    pair.setdefault((mid, imat), []).append(T.Interaction(ty, prod, 0, rad, ox, oy))
                    break
    flat, offs = [], [0]
    for a in range(n):
        for b in range(n):
            flat.extend(pair.get((a, b), []))
            offs.append(len(flat))
    rx = {  # gds.cpp:246-260
        ID["LAVA"]: [T.Interaction(T.REACT_TEMPERATURE_BELOW, 512, 0, ID["OBSIDIAN"], 0, 0)],
        ID["WATER"]: [T.Interaction(T.REACT_TEMPERATURE_ABOVE, 128, 0, ID["STEAM"], 0, 0)],
        ID["GOLD_ORE"]: [T.Interaction(T.REACT_TEMPERATURE_ABOVE, 512, 0, ID["GOLD_MOLTEN"], 0, 0)],
        ID["GOLD_MOLTEN"]: [T.Interaction(T.REACT_TEMPERATURE_BELOW, 128, 0, ID["GOLD_SOLID"], 0, 0)],
    }
    rflat, roffs = [], [0]
    for m in range(n):
        rflat.extend(rx.get(m, []))
        roffs.append(len(rflat))
        if m in rx:
            mats[m].react = 1
    ids = T.SpecialIds(ID["GENERIC_AIR"], ID["FIRE"], ID["WATER"], ID["LAVA"], ID["STEAM"], ID["OBSIDIAN"])
    return T.MaterialTable((T.Material * n)(*mats), ids, (T.Interaction * max(len(flat), 1))(*flat), (C.c_int32 * len(offs))(*offs),
                           (T.Interaction * max(len(rflat), 1))(*rflat), (C.c_int32 * len(roffs))(*roffs))


def register_material(table, physics, slipperyness, alpha, density, iterations, emit=0, emit_color=0, color=0xFFFFFFFF):
    """RegisterMaterial() (gds.cpp:282-289; Lua `materials_register`, game_basic.cpp:80): append one material,
    returning (new_table, id)."""
    n0 = table.n
    mats = [T.Material.from_buffer_copy(bytes(m)) for m in table.mats]
    mats.append(_mat(physics, slipperyness, alpha, density, iterations, emit, emit_color, color))
    n = n0 + 1
    flat, offs = [], [0]
    for a in range(n):
        for b in range(n):
            if a < n0 and b < n0:
                lo, hi = table.inter_offsets[a * n0 + b], table.inter_offsets[a * n0 + b + 1]
                flat.extend(T.Interaction.from_buffer_copy(bytes(table.inter[k])) for k in range(lo, hi))
            offs.append(len(flat))
    ro = list(table.react_offsets) + [table.react_offsets[n0]]
    t = T.MaterialTable((T.Material * n)(*mats), table.ids, (T.Interaction * max(len(flat), 1))(*flat), (C.c_int32 * len(offs))(*offs),
                        table.react, (C.c_int32 * len(ro))(*ro))
    return t, n0


# ---- Lua front door -----------------------------------------------------------------------------------------------------
# The reference binds three globals for scripts (game_basic.cpp:79-81): materials_init(), materials_register(s_id, name,
# index_name, physicsType, slipperyness, alpha, density, iterations, emit, emitColor, color) and materials_push()
# (gds.cpp:117, 282, 291).  Scripts run on a real Lua 5.4 VM: csrc/lua_front.c on the reference's vendored Lua 5.4.4, built into
# libfse_lua.so by csrc/Makefile.lua (lua_run / load_lua(engine="vm")); loops, functions, tables and the `global_def` settings table
# (cvar.cpp:57-99) work as in the game.  Without that library (it is a built artefact: present wherever __graft_entry__.build() ran
# with the reference tree at hand, and on the GPU box through the snapshot) load_lua falls back to reading the declarative subset —
# the three calls with literal arguments, plain `name = literal` assignments, comments — and ignores everything else.
import os  # noqa: E402
import re  # noqa: E402

_LUA_CALL = re.compile(r"\b(materials_init|materials_register|materials_push)\s*\(([^()]*)\)")
_LUA_ASSIGN = re.compile(r"^\s*(?:local\s+)?([A-Za-z_]\w*)\s*=\s*([^=\n][^\n]*?)\s*;?\s*$", re.M)
_PHYS = {"AIR": T.AIR, "SOLID": T.SOLID, "SAND": T.SAND, "SOUP": T.SOUP, "GAS": T.GAS, "PASSABLE": T.PASSABLE, "OBJECT": T.OBJECT}


def _lua_strip_comments(src):
    src = re.sub(r"--\[\[.*?\]\]", "", src, flags=re.S)
    return re.sub(r"--[^\n]*", "", src)


def _lua_literal(tok, env):
    tok = tok.strip()
    if not tok:
        raise ValueError("empty argument")
    if tok[0] in "\"'" and tok[-1] == tok[0]:
        return tok[1:-1]
    if tok in env:
        return env[tok]
    if tok in _PHYS:
        return _PHYS[tok]
    if tok in ("true", "false"):
        return tok == "true"
    try:
        return int(tok, 0)
    except ValueError:
        return float(tok)


class LuaMaterial(C.Structure):  # fse_lua_material (csrc/lua_front.c)
    _fields_ = [("s_id", C.c_int32), ("name", C.c_char * 64), ("index_name", C.c_char * 64), ("physics_type", C.c_int32), ("slipperyness", C.c_int32),
                ("alpha", C.c_int32), ("density", C.c_float), ("iterations", C.c_int32), ("emit", C.c_int32), ("emit_color", C.c_uint32), ("color", C.c_uint32)]


class LuaResult(C.Structure):  # fse_lua_result
    _fields_ = [("n_init", C.c_int32), ("n_register", C.c_int32), ("n_push", C.c_int32), ("has_global_def", C.c_int32), ("cell_iter", C.c_int32),
                ("brush_size", C.c_int32), ("tick_world", C.c_int32), ("tick_box2d", C.c_int32), ("tick_temperature", C.c_int32), ("error", C.c_char * 256)]


_LUA_LIB = None


def lua_library():
    """libfse_lua.so (Lua 5.4 VM + front-door shim) or None when it was not built."""
    global _LUA_LIB
    if _LUA_LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfse_lua.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.fse_lua_run.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(LuaMaterial), C.c_int, C.POINTER(LuaResult)]
        L.fse_lua_version.restype = C.c_char_p
        _LUA_LIB = L
    return _LUA_LIB


def lua_run(source, is_file=False, entry="OnGameEngineLoad", cap=1024):
    """Run a script on the VM the way the reference boots its scripts (load, then call OnGameEngineLoad, game_basic.cpp:60-61).
    Returns (list of registered LuaMaterial records, LuaResult)."""
    L = lua_library()
    if L is None:
        raise RuntimeError("libfse_lua.so is missing: make -C falling_sand_engine_b200/csrc -f Makefile.lua (needs the reference tree)")
    out = (LuaMaterial * cap)()
    res = LuaResult()
    rc = L.fse_lua_run(source.encode(), 1 if is_file else 0, (entry or "").encode(), out, cap, C.byref(res))
    if rc != 0:
        raise ValueError("lua: " + res.error.decode(errors="replace"))
    if res.n_register > cap:
        raise ValueError(f"the script registers {res.n_register} materials, capacity {cap}")
    return [out[i] for i in range(res.n_register)], res


def load_global_def(source, is_file=False):
    """The settings the reference reads back from the scripts' `global_def` table (cvar.cpp:57-99): cell_iter, brush_size, tick_*."""
    _, res = lua_run(source, is_file=is_file, entry=None)
    if not res.has_global_def:
        raise ValueError("the script defines no global_def table")
    return {"cell_iter": res.cell_iter, "brush_size": res.brush_size, "tick_world": res.tick_world, "tick_box2d": res.tick_box2d,
            "tick_temperature": res.tick_temperature}


def load_lua(source, seed=1337, engine="auto"):
    """Material table from a Lua script: materials_init() gives the stock table (default_materials), every materials_register(...)
    appends one material (arguments as in gds.cpp:282; physicsType is a number or one of AIR / SOLID / SAND / SOUP / GAS / PASSABLE /
    OBJECT), materials_push() ends it.  Returns (table, {s_id or name: material id}) ready for Context.set_materials.
    engine: "vm" (the Lua VM, required), "declarative" (literal calls only, no VM) or "auto" (the VM when libfse_lua.so exists)."""
    if engine not in ("auto", "vm", "declarative"):
        raise ValueError(f"engine {engine!r}")
    if engine == "vm" or (engine == "auto" and lua_library() is not None):
        recs, res = lua_run(source)
        if res.n_init == 0:
            raise ValueError("the script never calls materials_init()")
        if res.n_push == 0:
            raise ValueError("the script never calls materials_push()")
        table, ids = default_materials(seed), {}
        for r in recs:
            table, mid = register_material(table, r.physics_type, r.slipperyness, r.alpha, r.density, r.iterations, r.emit, r.emit_color, r.color)
            ids[r.s_id] = mid
            ids[r.name.decode()] = mid
        return table, ids
    src = _lua_strip_comments(source)
    env = {}
    for m in _LUA_ASSIGN.finditer(src):
        try:
            env[m.group(1)] = _lua_literal(m.group(2), env)
        except ValueError:
            pass  # not a literal: not something a materials_register argument can use
    table, ids, pushed = None, {}, False
    for m in _LUA_CALL.finditer(src):
        fn, args = m.group(1), [a for a in m.group(2).split(",")] if m.group(2).strip() else []
        if fn == "materials_init":
            table = default_materials(seed)
        elif fn == "materials_register":
            if table is None:
                raise ValueError("materials_register before materials_init")
            if len(args) != 11:
                raise ValueError(f"materials_register takes 11 arguments, got {len(args)}: {m.group(0)}")
            s_id, name, _index, phys, slip, alpha, dens, iters, emit, emit_color, color = (_lua_literal(a, env) for a in args)
            table, mid = register_material(table, int(phys), int(slip), int(alpha), float(dens), int(iters), int(emit), int(emit_color), int(color))
            ids[s_id] = mid
            ids[name] = mid
        else:
            pushed = True
    if table is None:
        raise ValueError("the script never calls materials_init()")
    if not pushed:
        raise ValueError("the script never calls materials_push()")
    return table, ids
