/* lua_front.c — the scripts' front door on a real Lua 5.4 VM (host code; no GPU involved).
 *
 * The reference binds materials_init / materials_register / materials_push for its scripts (game_basic.cpp:79-81 ->
 * InitMaterials / RegisterMaterial / PushMaterials, game_datastruct.cpp:117, 282, 291) and reads its settings back from the
 * `global_def` table (cvar.cpp:57-99; data/scripts/global.lua).  This shim runs a script on the Lua VM the reference vendors
 * (source/libs/lua, Lua 5.4.4 — compiled where it lies by Makefile.lua, never copied), records what the script registers and
 * hands it to the host as plain C structs; the host (materials.py / a C++ `world` shim) turns the records into the flattened
 * table fse_materials_set takes.  Engine functions the material path does not need (textures_load, audio_init, create_biome ...)
 * resolve to a no-op, so an unmodified game script gets through its OnGameEngineLoad.
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "lauxlib.h"
#include "lua.h"
#include "lualib.h"

#define FSE_LUA_API __attribute__((visibility("default")))

typedef struct fse_lua_material { /* arguments of RegisterMaterial, game_datastruct.cpp:282 */
    int32_t s_id;
    char name[64];
    char index_name[64];
    int32_t physics_type, slipperyness, alpha;
    float density;
    int32_t iterations, emit;
    uint32_t emit_color, color;
} fse_lua_material;

typedef struct fse_lua_result {
    int32_t n_init, n_register, n_push; /* calls seen (registrations beyond the caller's capacity are counted, not stored) */
    int32_t has_global_def;
    int32_t cell_iter, brush_size;      /* global_def.cell_iter / .brush_size (cvar.cpp:98-99); -1 when absent */
    int32_t tick_world, tick_box2d, tick_temperature; /* cvar.cpp:88-90; -1 when absent */
    char error[256];
} fse_lua_result;

typedef struct Front {
    fse_lua_material* out;
    int cap;
    fse_lua_result* res;
} Front;

static Front* front_of(lua_State* L) {
    lua_getfield(L, LUA_REGISTRYINDEX, "fse_front");
    Front* f = (Front*)lua_touserdata(L, -1);
    lua_pop(L, 1);
    return f;
}
static int l_materials_init(lua_State* L) {
    front_of(L)->res->n_init++;
    return 0;
}
static int l_materials_push(lua_State* L) {
    front_of(L)->res->n_push++;
    return 0;
}
static uint32_t to_u32(lua_State* L, int idx) { /* colours such as 0xFFFF6900 do not fit a signed 32-bit integer */
    if (lua_isinteger(L, idx)) return (uint32_t)(uint64_t)lua_tointeger(L, idx);
    return (uint32_t)(uint64_t)(int64_t)luaL_checknumber(L, idx);
}
static int l_materials_register(lua_State* L) {
    Front* f = front_of(L);
    if (lua_gettop(L) != 11) return luaL_error(L, "materials_register takes 11 arguments, got %d", lua_gettop(L));
    fse_lua_material m;
    memset(&m, 0, sizeof m);
    m.s_id = (int32_t)luaL_checkinteger(L, 1);
    snprintf(m.name, sizeof m.name, "%s", luaL_checkstring(L, 2));
    snprintf(m.index_name, sizeof m.index_name, "%s", luaL_checkstring(L, 3));
    m.physics_type = (int32_t)luaL_checkinteger(L, 4);
    m.slipperyness = (int32_t)luaL_checkinteger(L, 5);
    m.alpha = (int32_t)luaL_checkinteger(L, 6);
    m.density = (float)luaL_checknumber(L, 7);
    m.iterations = (int32_t)luaL_checkinteger(L, 8);
    m.emit = (int32_t)luaL_checkinteger(L, 9);
    m.emit_color = to_u32(L, 10);
    m.color = to_u32(L, 11);
    if (f->res->n_register < f->cap) f->out[f->res->n_register] = m;
    f->res->n_register++;
    return 0;
}
static int l_noop(lua_State* L) {
    (void)L;
    return 0;
}
/* unknown globals read as a function that does nothing (and can be indexed / called again): the engine calls of a game script */
static int l_missing_global(lua_State* L) {
    lua_pushcfunction(L, l_noop);
    return 1;
}
static int field_int(lua_State* L, const char* name) { /* table at the top of the stack */
    int v = -1;
    lua_getfield(L, -1, name);
    if (lua_isboolean(L, -1)) v = lua_toboolean(L, -1);
    else if (lua_isnumber(L, -1)) v = (int)lua_tonumber(L, -1);
    lua_pop(L, 1);
    return v;
}

/* Run `source` (a chunk of Lua text, or a file name when is_file != 0), then call the global function `entry` if it exists (the
 * reference calls OnGameEngineLoad after loading game.lua, game_basic.cpp:60-61).  Returns 0 on success, 1 on a Lua error (message
 * in res->error). */
FSE_LUA_API int fse_lua_run(const char* source, int is_file, const char* entry, fse_lua_material* out, int cap, fse_lua_result* res) {
    memset(res, 0, sizeof *res);
    res->cell_iter = res->brush_size = res->tick_world = res->tick_box2d = res->tick_temperature = -1;
    lua_State* L = luaL_newstate();
    if (!L) {
        snprintf(res->error, sizeof res->error, "luaL_newstate failed");
        return 1;
    }
    luaL_openlibs(L);
    Front f = {out, cap, res};
    lua_pushlightuserdata(L, &f);
    lua_setfield(L, LUA_REGISTRYINDEX, "fse_front");
    lua_register(L, "materials_init", l_materials_init);
    lua_register(L, "materials_register", l_materials_register);
    lua_register(L, "materials_push", l_materials_push);
    static const char* phys[] = {"AIR", "SOLID", "SAND", "SOUP", "GAS", "PASSABLE"}; /* PhysicsType, game_datastruct.hpp:87-95 */
    for (int i = 0; i < 6; i++) {
        lua_pushinteger(L, i);
        lua_setglobal(L, phys[i]);
    }
    lua_pushinteger(L, 5);
    lua_setglobal(L, "OBJECT");
    lua_pushglobaltable(L); /* setmetatable(_G, {__index = function() return noop end}) */
    lua_newtable(L);
    lua_pushcfunction(L, l_missing_global);
    lua_setfield(L, -2, "__index");
    lua_setmetatable(L, -2);
    lua_pop(L, 1);
    int rc = is_file ? luaL_dofile(L, source) : luaL_dostring(L, source);
    if (rc == LUA_OK && entry && entry[0]) {
        lua_getglobal(L, entry);
        if (lua_isfunction(L, -1) && lua_tocfunction(L, -1) != l_noop) rc = lua_pcall(L, 0, 0, 0);
        else lua_pop(L, 1);
    }
    if (rc != LUA_OK) {
        snprintf(res->error, sizeof res->error, "%s", lua_tostring(L, -1) ? lua_tostring(L, -1) : "lua error");
        lua_close(L);
        return 1;
    }
    lua_pushglobaltable(L);
    lua_pushstring(L, "global_def");
    lua_rawget(L, -2); /* raw: the __index hook above must not invent it */
    if (lua_istable(L, -1)) {
        res->has_global_def = 1;
        res->cell_iter = field_int(L, "cell_iter");
        res->brush_size = field_int(L, "brush_size");
        res->tick_world = field_int(L, "tick_world");
        res->tick_box2d = field_int(L, "tick_box2d");
        res->tick_temperature = field_int(L, "tick_temperature");
    }
    lua_close(L);
    return 0;
}

FSE_LUA_API const char* fse_lua_version(void) { return LUA_RELEASE; }
