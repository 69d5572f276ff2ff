/*
 * fse.h — C ABI of the B200-native falling-sand world tick ("fse").
 *
 * This is the drop-in boundary for the per-tick world update of
 * cstom4994/falling_sand_engine.  The reference has no FFI seam for this path:
 * it is a set of member functions on `class world` (source/engine/world.hpp:63-193)
 * plus three Lua-bound free functions (source/engine/game_basic.cpp:79-81).  Every
 * entry point below names the reference member it replaces (file:line relative to
 * /root/reference).  The host-side C++ `fse::World` shim
 * (falling_sand_engine_b200/csrc/world.hpp) keeps the reference's method names on
 * top of this ABI.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success and
 * a negative FSE_E* code on failure (message via fse_last_error); one caller at a
 * time per fse_world (thread-compatible, like the reference's main-thread-only
 * world); device work is asynchronous on the world's internal CUDA stream unless
 * the call returns host data, and fse_sync() joins it.  There is no CPU fallback:
 * without a CUDA device fse_ctx_create fails.
 */
#ifndef FSE_H
#define FSE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSE_API __attribute__((visibility("default")))

#define FSE_CHUNK 128          /* CHUNK_W == CHUNK_H, source/engine/core/const.h:19-20 */
#define FSE_MAX_MATERIALS 256  /* device material plane is u8 (reference ships 41)      */
#define FSE_MAX_REACH 5        /* max |ofs|+radius of an interaction box (gds.cpp:221-231 generates <=5) */

/* error codes */
#define FSE_OK 0
#define FSE_EINVAL (-1)
#define FSE_ECUDA (-2)
#define FSE_ENOMEM (-3)
#define FSE_ESTATE (-4)
#define FSE_ENCCL (-5)

/* PhysicsType, source/engine/game_datastruct.hpp:87-95 (PASSABLE == OBJECT == 5) */
enum { FSE_AIR = 0, FSE_SOLID = 1, FSE_SAND = 2, FSE_SOUP = 3, FSE_GAS = 4, FSE_PASSABLE = 5, FSE_OBJECT = 5 };

/* interaction / reaction kinds, game_datastruct.hpp:97-103 */
enum {
    FSE_INTERACT_NONE = 0,
    FSE_INTERACT_TRANSFORM_MATERIAL = 1, /* data1 = product id, data2 = box radius */
    FSE_INTERACT_SPAWN_MATERIAL = 2,     /* data1 = product id, data2 = box radius */
    FSE_EXPLODE = 3,
    FSE_REACT_TEMPERATURE_BELOW = 4,     /* data1 = threshold, data2 = product id  */
    FSE_REACT_TEMPERATURE_ABOVE = 5
};

/* How TilesCreate(id,x,y) (game_datastruct.cpp:485-574) colours a freshly created
 * cell.  The reference samples texture-pack images for most solids; textures are
 * presentation assets (out of scope), so POSITIONAL substitutes a position hash. */
enum {
    FSE_COLOR_FIXED = 0,      /* color                                   (Water/Lava/Steam ...)      */
    FSE_COLOR_JITTER = 1,     /* color + ((r % jitter_range) << jitter_shift)   (Fire/Grass/Dirt ...) */
    FSE_COLOR_POSITIONAL = 2  /* color ^ (hash(x,y) & 0x0f0f0f)          (textured solids)           */
};

/* Flattened `Material` (game_datastruct.hpp:130-168). */
typedef struct fse_material {
    int32_t physics;        /* physicsType */
    float density;
    int32_t iterations;
    int32_t slipperyness;   /* >= 1 required for FSE_SAND (reference divides by it, world.cpp:1656) */
    int32_t emit;
    uint32_t emit_color;
    uint32_t color;         /* base colour of created cells */
    uint32_t add_temp;      /* Material::addTemp (u32, game_datastruct.hpp:143) */
    float conduction_self;
    float conduction_other;
    int16_t create_temp;    /* temperature of TilesCreate'd cells: WATER -1023, LAVA 1024 (gds.cpp:427-437) */
    uint8_t alpha;
    uint8_t interact;       /* Material::interact */
    uint8_t react;          /* Material::react    */
    uint8_t color_kind;     /* FSE_COLOR_*        */
    uint8_t jitter_shift;
    uint8_t jitter_range;
} fse_material;

/* `MaterialInteraction` (game_datastruct.hpp:112-118); field overloading as in the
 * reference: TRANSFORM/SPAWN: data1 = product material, data2 = radius;
 * REACT_*: data1 = threshold temperature, data2 = product material. */
typedef struct fse_interaction {
    int32_t type;
    int16_t data1;
    uint16_t _pad;
    uint32_t data2;
    int32_t ofs_x;
    int32_t ofs_y;
} fse_interaction;

/* Materials the reference's tick refers to by identity (world.cpp:1101,1519,1883). */
typedef struct fse_special_ids {
    int32_t air;       /* GENERIC_AIR  (Tiles_NOTHING) */
    int32_t fire;
    int32_t water;
    int32_t lava;
    int32_t steam;
    int32_t obsidian;
} fse_special_ids;

/* One grid cell crossing the boundary = the value fields of `MaterialInstance`
 * (game_datastruct.hpp:207-225; `id` is always `mat->id`).           */
typedef struct fse_cell {
    uint16_t mat;
    uint8_t moved;
    uint8_t settle;      /* settleCount */
    uint32_t color;
    int16_t temp;        /* temperature */
    uint8_t dirty;       /* world::dirty[] (world.hpp:131) folded into the cell */
    uint8_t _pad;
    float fluid;         /* fluidAmount (default 2.0f, gds.hpp:215) */
    float fluid_diff;    /* fluidAmountDiff */
} fse_cell;

typedef struct fse_rect {
    int32_t x, y, w, h;
} fse_rect;

/* Arguments of one world::tick() (world.cpp:1036): globaldef.cell_iter
 * (data/scripts/global.lua:47), tickZone (game.cpp:1629), and the (seed,tick) key
 * of the counter RNG that replaces libc rand(). */
typedef struct fse_tick_args {
    uint32_t tick;
    uint32_t seed;
    int32_t cell_iter;
    fse_rect tick_zone;
} fse_tick_args;

/* `CellData` (game_utils/cells.h:15-36) without the std::function killCallback;
 * `id` lets the host keep an id -> callback map (only the vacuum tool uses it). */
typedef struct fse_particle {
    fse_cell tile;
    float x, y, vx, vy, ax, ay;
    float target_x, target_y, target_force;
    int32_t lifetime;
    int32_t fade_time;
    uint8_t phase;
    uint8_t temporary;
    uint8_t in_object_state;
    uint8_t vacuum;      /* held by the vacuum tool: member of Item::vacuumCells (game.cpp:2507, 2562) */
    uint32_t _pad2;
    uint64_t id;
} fse_particle;

/* Per-world statistics used as parity metrics (the reference's movingTiles[]
 * histogram, game.cpp:1991-2000, generalised). */
typedef struct fse_stats {
    uint64_t hash;                           /* order-independent 64-bit state hash of the rect */
    uint64_t count[FSE_MAX_MATERIALS];       /* cells per material */
    double fluid_mass[FSE_MAX_MATERIALS];    /* sum(fluid + fluid_diff) over SOUP cells per material */
    uint64_t n_dirty;
    uint64_t n_moved;
} fse_stats;

typedef struct fse_ctx fse_ctx;
typedef struct fse_world fse_world;

/* ---- context ---------------------------------------------------------------- */
FSE_API int fse_ctx_create(int device, fse_ctx** out);
FSE_API void fse_ctx_destroy(fse_ctx* ctx);
FSE_API const char* fse_last_error(void);
FSE_API const char* fse_version(void);
/* sizeof() of the POD types above as compiled into the library (0 material, 1 interaction, 2 special_ids,
 * 3 cell, 4 rect, 5 tick_args, 6 particle, 7 stats): lets a foreign-language binding verify its layout. */
FSE_API int fse_abi_sizeof(int which);

/* InitMaterials/RegisterMaterial/PushMaterials (game_datastruct.cpp:117-300), the
 * Lua binds materials_init/materials_register/materials_push
 * (game_basic.cpp:79-81): flattened POD copy of the whole table; re-callable.
 * inter_offsets has n*n+1 entries: interactions of material a against material b
 * are inter[inter_offsets[a*n+b] .. inter_offsets[a*n+b+1]) (Material::interactions[b],
 * gds.hpp:151); react_offsets has n+1 entries (Material::reactions). */
FSE_API int fse_materials_set(fse_ctx* ctx, const fse_material* tbl, int n, const fse_special_ids* ids,
                              const fse_interaction* inter, const int32_t* inter_offsets,
                              const fse_interaction* react, const int32_t* react_offsets);

/* ---- world container: world::init / ~world (world.cpp:43-172, 3532-3636) ------ */
FSE_API int fse_world_create(fse_ctx* ctx, int32_t width, int32_t height, fse_world** out);
FSE_API void fse_world_destroy(fse_world* w);
FSE_API int fse_sync(fse_world* w);

/* getTile/setTile (world.cpp:999-1008), frame() merge and chunkSaveCache
 * (world.cpp:2374-2391, 2780-2792): AoS rect <-> device SoA planes. */
FSE_API int fse_write_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, const fse_cell* cells);
FSE_API int fse_read_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, fse_cell* cells);
/* memset(dirty, false, ...) (game.cpp:2153) */
FSE_API int fse_clear_dirty(fse_world* w);
FSE_API int fse_stats_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, fse_stats* out);

/* ---- the tick ------------------------------------------------------------------ */
/* In-row visiting order of the chunk tick: the reference's chunk colours, passes and bottom-up rows are kept; inside
 * a row every cell decides from the state before the row step, then the row commits ("rows" schedule, bit-exact CPU restatement in
 * oracle/rows_oracle.cpp).  The two values pick the kernels, not the results.  (Value 0 was round 1's in-place "classes" schedule;
 * it was removed — slower at every size — and fse_set_schedule refuses it.) */
#define FSE_SCHEDULE_ROWS 1    /* one kernel per pass and colour phase (default)                        */
#define FSE_SCHEDULE_ROWS_FUSED 2 /* same results as ROWS, all three passes pipelined in one kernel      */
FSE_API int fse_set_schedule(fse_world* w, int schedule);
/* world::tick() (world.cpp:1036-1948) without the physicsCheck tail (see
 * fse_flood_component).  Asynchronous. */
FSE_API int fse_tick(fse_world* w, const fse_tick_args* args);
/* world::tickTemperature() (world.cpp:1950-2004). */
FSE_API int fse_tick_temperature(fse_world* w, const fse_rect* tick_zone);

/* ---- loose particles: world::cells (world.hpp), addCell (world.cpp:2292),
 *      tickCells (world.cpp:2030-2195) ---------------------------------------- */
FSE_API int fse_particles_add(fse_world* w, const fse_particle* p, int32_t n);
FSE_API int fse_particles_tick(fse_world* w, const fse_rect* tick_zone);
FSE_API int fse_particles_count(fse_world* w, int64_t* out);
FSE_API int fse_particles_read(fse_world* w, fse_particle* out, int64_t cap, int64_t* n_out);
FSE_API int fse_particles_clear(fse_world* w);
/* capacity of the device particle pool (default 1<<20); the reference's std::vector grows unbounded.  The pool also grows on its
 * own: fse_tick keeps an eighth of it free, fse_explosion and fse_bodies_raster make room for every cell they may throw up. */
FSE_API int fse_particles_reserve(fse_world* w, int64_t capacity);
/* cumulative count of particles a kernel could not store because one call spawned more than the head-room it was given
 * (their cells are gone from the grid); 0 in normal operation.  fse_particles_count never fails on an overflowed pool. */
FSE_API int fse_particles_dropped(fse_world* w, int64_t* out);

/* ---- entities <-> grid: world::tickEntities (world.cpp:3010-3247), WorldEntitySystem::process
 *      (game/player.cpp:173-199), the objectDelete loop (game.cpp:2128-2139).  fse_entity mirrors the value fields of WorldEntity
 *      (game_datastruct.hpp:42-62); the b2Body the reference moves along (world.cpp:3227-3228) stays with the host.
 *      load_x / load_y = world::loadZone.x / .y (floats, as in the reference's MErect).  Entities are processed in array order. */
typedef struct fse_entity {
    float x, y, vx, vy;
    int32_t hw, hh;      /* hit box, cells */
    int32_t ground;      /* out: stood on something this tick */
    int32_t destroy;     /* out: |v| >= 1024, the reference destroys the entity (world.cpp:3222-3225) */
} fse_entity;
/* overlap push-out, gravity, the 8-sub-step horizontal and vertical sweeps with step-up and sand kicks, velocity decay; updates
 * `ents` in place */
FSE_API int fse_entities_tick(fse_world* w, fse_entity* ents, int32_t n, float load_x, float load_y, uint32_t tick, uint32_t seed);
/* AIR under an entity becomes Tiles_OBJECT (material `object_mat`, GENERIC_OBJECT = 6 in the stock table), SAND / SOUP is thrown up
 * as a particle first; every stamped cell is remembered for fse_object_delete */
FSE_API int fse_entities_stamp(fse_world* w, const fse_entity* ents, int32_t n, float load_x, float load_y, int32_t object_mat, uint32_t tick,
                               uint32_t seed);
/* the stamped cells become Tiles_NOTHING again (end of the game tick) */
FSE_API int fse_object_delete(fse_world* w);

/* ---- rigid-body bridge: the raster / erase loops of game::tick (game.cpp:1711-1815, 1896-1983) ---------------
 * Box2D stays on the host (north_star); the host keeps b2Body poses and sends one fse_xform per body per tick.
 * fse_body_desc mirrors RigidBody::matWidth / matHeight / tiles (game/player.hpp:15-65); AIR tiles are empty.
 * On multi-rank strips every rank uploads every body and makes every call with the same transforms (global coordinates); a body is
 * run by the rank that holds its footprint box, together with every body whose box overlaps its own, the rows shared with a
 * neighbour are brought back in step afterwards, and every rank gets the same feedback and the same tiles.  A group of overlapping
 * bodies that does not fit the rows of one strip plus its 32 ghost rows is refused (FSE_ESTATE). */
typedef struct fse_body_desc {
    int32_t w, h;
    const fse_cell* tiles; /* w*h, index tx + ty*w */
} fse_body_desc;
typedef struct fse_xform {
    float x, y, angle; /* b2Body::GetPosition() / GetAngle() */
} fse_xform;
typedef struct fse_body_feedback {
    int32_t sand_hits;  /* displaced SAND cells: host damps v *= 0.99, w *= 0.98 per hit (game.cpp:1796-1797) */
    int32_t soup_hits;  /* displaced SOUP cells: v *= 0.998, w *= 0.99 (game.cpp:1806-1807)                    */
    int32_t placed;     /* raster: pixels stamped; erase: pixels lifted back                                   */
    int32_t destroyed;  /* erase: pixels found missing (tile becomes AIR, game.cpp:1959-1965)                  */
} fse_body_feedback;
/* replace the device copy of all bodies' tiles (once, and whenever the host changes a body) */
FSE_API int fse_bodies_upload(fse_world* w, const fse_body_desc* bodies, int32_t n);
/* stamp every body pixel into the grid, displacing sand/liquid into particles */
FSE_API int fse_bodies_raster(fse_world* w, const fse_xform* xf, int32_t n, uint32_t tick, uint32_t seed, fse_body_feedback* out);
/* lift the pixels back out; tiles that are gone are cleared; needs_update[b] = 1 as in game.cpp:1982 */
FSE_API int fse_bodies_erase(fse_world* w, const fse_xform* xf, int32_t n, fse_body_feedback* out, uint8_t* needs_update);
FSE_API int fse_bodies_read(fse_world* w, int32_t body, fse_cell* tiles_out);

/* Fracture hand-off (world::updateRigidBodyHitbox, world.cpp:288-720): an uploaded body whose pixels were carved apart is cut into
 * one piece per 4-connected component of its non-AIR tiles (north_star: connected-component labelling; the reference assigns pixels
 * by nearest triangle centroid, world.cpp:587-610).  Per piece: bounding box inside the body's tile array (the crop of world.cpp:
 * 305-320), pixel count, whether it holds the weld pixel (world.cpp:620), the position shift of the new body (the rotated box
 * corner, world.cpp:350-362; angle = b2Body::GetAngle) and its own w x h tile array in tiles_out (AIR where the box covers another
 * piece).  Pieces are numbered by their first pixel in row-major order.  The host feeds each piece's mask to fse_mask_outline for
 * the TPPL / b2PolygonShape step and creates the bodies (pose, velocities and weld copied as in world.cpp:657-707). */
typedef struct fse_body_piece {
    int32_t x0, y0, w, h;
    int32_t n_pixels;
    int32_t weld;
    int32_t tile_off;         /* first tile of the piece in tiles_out */
    float shift_x, shift_y;
} fse_body_piece;
FSE_API int fse_bodies_split(fse_world* w, int32_t body, float angle, int32_t weld_x, int32_t weld_y, fse_body_piece* pieces, int32_t cap_pieces,
                             int32_t* n_pieces, fse_cell* tiles_out, int64_t cap_tiles);

/* `world::explosion(cx, cy, radius)` (world.cpp:2294-2332): every non-AIR cell within `radius` of (cx, cy) is removed — SOLID
 * cells and 6 in 10 of the others vanish, the rest leave as loose particles (colour darkened to a quarter, spawned one cell
 * lower, thrown outward) — and every non-SOLID cell of the ring out to 2*radius is thrown outward as a particle.  Cells decide
 * independently; rand() is replaced by the counter RNG keyed on (seed, tick, x, y).  On multi-rank strips every rank makes the call
 * with the same arguments: each clears the part of the blast it holds, the particles come from the rank that owns the row. */
FSE_API int fse_explosion(fse_world* w, int32_t cx, int32_t cy, int32_t radius, uint32_t tick, uint32_t seed);

/* ---- interactive tools, grid side.  Rigid-body surfaces, Box2D bodies, audio and UI stay with the host. ----------
 * erase brush (game.cpp:593-625): every non-AIR cell under a brush_size square (corners with |dx| + |dy| == brush_size left out)
 * stamped along world::forLine from (x0, y0) to (x1, y1) becomes Tiles_NOTHING */
FSE_API int fse_tool_erase_line(fse_world* w, int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t brush_size);
/* pickaxe (game.cpp:771-790): SOLID cells inside the circle of diameter break_size whose box starts at (x, y) leave the grid; their
 * colours come back as the int(break_size)^2 ARGB pixels of the rigid body the host makes from them, *n_out = cells taken */
FSE_API int fse_tool_pickaxe(fse_world* w, int32_t x, int32_t y, float break_size, uint32_t* pixels_out, int32_t* n_out);
/* hammer release (game.cpp:843-890): a crack from the hammer point away from the release point (x, y) in jittered ~10-cell segments
 * along world::forLineCornered; SOLID cells on it become `sand_mat` (GENERIC_SAND) at half brightness until the crack leaves the
 * solid.  The host then probes both sides of the crack's middle with fse_flood_component (game.cpp:892-901). */
typedef struct fse_hammer_result {
    int32_t end_x, end_y;   /* last cell changed, -1 if none */
    int32_t n_changed;
    int32_t broke;          /* the crack left the solid */
} fse_hammer_result;
FSE_API int fse_tool_hammer(fse_world* w, int32_t hammer_x, int32_t hammer_y, int32_t x, int32_t y, int32_t sand_mat, uint32_t tick, uint32_t seed,
                            fse_hammer_result* out);
/* vacuum (game.cpp:2456-2585): walk from the screen centre (wcx, wcy) towards the mouse (wmx, wmy) to the first SOLID / SAND / SOUP
 * cell, turn the matter in the 11 x 11 disc around it into `phase` particles held by the vacuum, and catch the loose particles
 * already inside it.  Nothing happens when the mouse is more than 256 cells away. */
typedef struct fse_vacuum_result {
    int32_t x, y;           /* centre of the disc, -1 when out of reach */
    int32_t n_sucked;       /* grid cells turned into particles */
    int32_t n_caught;       /* loose particles caught */
} fse_vacuum_result;
FSE_API int fse_tool_vacuum(fse_world* w, int32_t wcx, int32_t wcy, int32_t wmx, int32_t wmy, uint32_t tick, uint32_t seed, fse_vacuum_result* out);
/* the vacuumCells update (game.cpp:2640-2664): held particles whose lifetime is up fly towards (target_x, target_y); within 10 cells
 * they are collected (dropped by the next fse_particles_tick).  n_collected may be null. */
FSE_API int fse_particles_vacuum_pull(fse_world* w, float target_x, float target_y, int32_t* n_collected);

/* ---- render planes: the dirty -> texture loops of game::tick (game.cpp:1994-2126).  Every cell whose dirty
 * flag is set refreshes its texel in device-resident RGBA8 planes (byte order r, g, b, a as the reference fills
 * dpixels_ar): main (colour + material alpha; AIR = transparent black), fire (only FIRE cells write it, AIR clears it,
 * other materials leave it alone), emission (Material::emitColor) and — with fse_flow_enable — flow (game.cpp:2040-2062:
 * the smoothed flowX / flowY of dirty liquid cells, r = x, g = y, b = 0, a = 255).  movingTiles[mat] counts the dirty cells
 * per material (game.cpp:1998-2000).  Dirty flags are left set, as in the reference (fse_clear_dirty = game.cpp:2153). */
typedef struct fse_render_stats {
    int64_t dirty;                       /* hadDirty: number of dirty cells */
    int64_t fire;                        /* hadFire: dirty FIRE cells */
    int64_t moving[FSE_MAX_MATERIALS];   /* movingTiles */
    int64_t flow;                        /* hadFlow: dirty liquid cells (0 unless fse_flow_enable) */
} fse_render_stats;
enum { FSE_PIXELS_MAIN = 0, FSE_PIXELS_FIRE = 1, FSE_PIXELS_EMISSION = 2, FSE_PIXELS_FLOW = 3, FSE_PIXELS_LAYER2 = 4, FSE_PIXELS_BACKGROUND = 5 };
/* allocate (zeroed: transparent) / free the main, fire and emission planes */
FSE_API int fse_pixels_enable(fse_world* w, int enable);
/* refresh the texels of all dirty cells; `out` may be null (no host sync then) */
FSE_API int fse_render_dirty(fse_world* w, fse_render_stats* out);
/* copy a rect of one plane to the host, rw*rh*4 bytes */
FSE_API int fse_pixels_read(fse_world* w, int which, int32_t x, int32_t y, int32_t rw, int32_t rh, uint8_t* rgba);
/* device pointer of a plane (W*H RGBA8, row-major) for CUDA-GL interop or further kernels; null when that plane is not allocated */
FSE_API void* fse_pixels_device(fse_world* w, int which);

/* ---- liquid flow accumulators: world::flowX / flowY / prevFlowX / prevFlowY (world.hpp:116-119).  While enabled, pass 1 of
 * fse_tick adds every liquid flow it decides to the source cell (flowY += down, flowX -= left, flowX += right, flowY -= up;
 * world.cpp:1334, 1374, 1402, 1432) and fse_render_dirty turns them into the flow texture and resets them.  Render-only
 * state: 16 B per cell + the texture, rows schedule only, off by default. */
FSE_API int fse_flow_enable(fse_world* w, int enable);
/* which: 0 flowX, 1 flowY, 2 prevFlowX, 3 prevFlowY; rw*rh floats */
FSE_API int fse_flow_read(fse_world* w, int which, int32_t x, int32_t y, int32_t rw, int32_t rh, float* out);

/* ---- second cell layer and background colours: world::real_layer2 / background (world.hpp:112-113), written by
 * setTileLayer2 (world.cpp:1015-1019) and the chunk merge (world.cpp:2384-2389), read by the renderer only (the tick never
 * touches them).  Allocated on first use (13 B per cell + two textures).  A write sets layer2Dirty / backgroundDirty;
 * fse_render_layers (game.cpp:2068-2126 + the clears of 2154-2155) refreshes FSE_PIXELS_LAYER2 / FSE_PIXELS_BACKGROUND from the
 * dirty cells and clears the marks.  Of an fse_cell the layer keeps mat, color and temp (what a chunk file holds). */
FSE_API int fse_layer2_write_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, const fse_cell* cells);
FSE_API int fse_layer2_read_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, fse_cell* cells);
FSE_API int fse_background_write_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, const uint32_t* argb);
FSE_API int fse_background_read_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, uint32_t* argb);
/* draw_background_grid = globaldef.draw_background_grid; the counts (hadLayer2Dirty / hadBackgroundDirty as cell counts) may be null */
FSE_API int fse_render_layers(fse_world* w, int draw_background_grid, int64_t* n_layer2, int64_t* n_background);

/* ---- camera scroll: the grid shift of world::tickChunks (world.cpp:2454-2478, 2579-2582).  Every cell
 * (and its layer-2 cell and background colour, when those planes exist) moves by (dx, dy); cells whose source lies outside the
 * world keep their old content (the reference's in-place copy), dirty flags stay where they are (world::dirty is not
 * shifted), loose particles move along.  The shifted world is written into a second set of planes which then becomes the
 * world (one read + one write per cell).  Chunk load / save around the scroll stays with the host (fse_write_rect /
 * fse_read_rect).  On multi-rank strips every rank makes the call and shifts the rows it holds; for a vertical shift the |dy| rows that
 * change ranks travel as one message to each neighbour first (|dy| at most the shortest strip's own rows minus the ghost rows; the
 * layer-2 / background planes do not take vertical shifts on strips yet). */
FSE_API int fse_scroll(fse_world* w, int32_t dx, int32_t dy);

/* ---- fracture / hitbox outlines: updateRigidBodyHitbox, updateChunkMesh (world.cpp:288-720, 722-959) with
 * MarchingSquares::FindPerimeter + simplify(...,1) (physics_math.cpp:1766-1965), physicsCheck flood (world.cpp:3330-3429).
 * The device labels 4-connected components and extracts + simplifies every contour; TPPL hole removal / ear clipping
 * and b2Body creation stay on the host.  `masks` holds n_masks images of w*h bytes (non-zero = solid pixel).
 * labels (optional, n_masks*w*h int32): lowest pixel index of the pixel's component, -1 for empty pixels.
 * Contours come back flattened: contour k of mask m is pts[2*pt_off[c] .. 2*pt_off[c+1]) with c = mask_off[m] + k. */
FSE_API int fse_mask_outline(fse_world* w, const uint8_t* masks, int32_t n_masks, int32_t mw, int32_t mh, int32_t* labels,
                             int32_t* n_components, float* pts, int32_t cap_pts, int32_t* pt_off, int32_t cap_contours,
                             int32_t* mask_off);
/* SOLID mask of a world rect (updateChunkMesh's input, world.cpp:722-760), written to `mask` (rw*rh bytes) */
/* The host half of updateRigidBodyHitbox / updateChunkMesh (world.cpp:497-563) for the outlines of ONE mask as fse_mask_outline
 * returns them (pts / pt_off of the mask's contours): each outline reversed is a polygon, clockwise ones are holes, holes are bridged
 * into their outer polygons and every outer polygon is ear-clipped — the same choices, in the same order, as the reference's
 * TPPLPartition::RemoveHoles / Triangulate_EC (physics_math.cpp:297-580; csrc/polygons.hpp) — and triangles whose three x or three y
 * coincide are dropped.  tris: 6 doubles per triangle (x0 y0 x1 y1 x2 y2); group_off[g] .. group_off[g + 1]: the triangles of the
 * g-th polygon = the b2PolygonShapes of one body; *n_groups polygons kept at least one triangle.  Pure host code. */
FSE_API int fse_hitbox_triangles(const float* pts, const int32_t* pt_off, int32_t n_contours, double* tris, int32_t cap_tris,
                                 int32_t* group_off, int32_t cap_groups, int32_t* n_groups);
FSE_API int fse_solid_mask(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, uint8_t* mask);
/* physicsCheck(x,y): size / bbox {minx,miny,maxx,maxy} / sorted pixel indices (x + y*width) of the 4-connected SOLID
 * component at (x,y); *count = cap+1 when it exceeds cap, 0 when the seed is not SOLID.  pixels may be NULL. */
FSE_API int fse_flood_component(fse_world* w, int32_t x, int32_t y, int32_t cap, int32_t* count, int32_t* bbox, int32_t* pixels);
/* world::physicsCheck(x, y) (world.cpp:3330-3411) as one call.  The 4-connected SOLID component at (x, y), abandoned beyond 1000 cells
 * like the reference's flood: 1..10 cells are deleted (Tiles_NOTHING, dirty); 11..1000 cells are cut out of the grid into the tile
 * array of a new rigid body — OBSIDIAN carrying each cell's colour, AIR elsewhere in the bounding box, what makeRigidBody builds from
 * the colour surface (world.cpp:191-209; membership by component, not by the colour's alpha byte).  The host creates the b2Body at
 * (x, y) with its random velocity and calls updateRigidBodyHitbox (fse_bodies_split).  tiles_out: w * h cells, row-major, written
 * for action 2; when cap_tiles is too small the call fails with the box in *out and changes nothing. */
typedef struct fse_physcheck_result {
    int32_t count;   /* cells of the component; 1001 = abandoned (> 1000), 0 = (x, y) is not SOLID or outside the world */
    int32_t action;  /* 0 nothing, 1 deleted, 2 cut out into tiles_out */
    int32_t x, y, w, h;  /* bounding box of the component (count 1..1000) */
} fse_physcheck_result;
FSE_API int fse_physics_check(fse_world* w, int32_t x, int32_t y, fse_physcheck_result* out, fse_cell* tiles_out, int32_t cap_tiles);
/* The probe at the end of world::tick (world.cpp:1929-1934): physicsCheck(tickZone.x + rand() % tickZone.w, tickZone.y + rand() %
 * tickZone.h), with rand() replaced by the counter RNG — draws S_PROBE_X / S_PROBE_Y of cell (0, 0) under rng_key(seed, tick, 15).
 * Pure host function; the caller passes the position to fse_physics_check. */
FSE_API void fse_probe_position(uint32_t seed, uint32_t tick, const fse_rect* tick_zone, int32_t* x, int32_t* y);

/* ---- active-region tracking: world::active/lastActive (world.hpp:131-133) are allocated but dead in the reference
 * (every writer is commented out), so the only contract is "same cells as a full sweep".  When
 * enabled, each 128x128 chunk falls asleep once a pass over it changed nothing and every cell in it is provably
 * inert, is woken by any write next to it, and fse_tick launches only the compacted list of awake chunks. */
FSE_API int fse_active_enable(fse_world* w, int enable);
FSE_API int fse_active_stats(fse_world* w, int64_t* awake_chunks, int64_t* total_chunks);

/* ---- multi-GPU: horizontal strips + NCCL halo rows (no reference counterpart) -----------
 * One process per GPU.  Rank 0 makes a 128-byte id (fse_comm_unique_id), every rank gets it out of band and calls
 * fse_comm_init; fse_strip_create then gives each rank the strip of chunk rows it owns plus ghost rows.  All rect /
 * zone arguments stay in GLOBAL cell coordinates.  fse_tick on a strip world exchanges the rows around each cut after
 * every colour phase (ncclSend/ncclRecv on a side stream, overlapped with the interior chunk rows); the result is
 * bit-identical to the unpartitioned world.  The game loop is SPMD on strips: every rank makes every call (tick, particles,
 * temperature, explosion, tools, scroll, bodies, entities, physicsCheck) with the same arguments and gets the same results; calls
 * that edit the grid from one place (a body, an entity, a crack, a probed component) are run by the rank that holds the place's box
 * and the box travels to the neighbours it reaches into — such a box, with the ones it overlaps, must lie within 32 rows of one
 * strip (FSE_ESTATE otherwise). */
FSE_API int fse_comm_unique_id(void* out128);
FSE_API int fse_comm_init(fse_ctx* ctx, int rank, int nranks, const void* id128);
FSE_API int fse_comm_destroy(fse_ctx* ctx);
FSE_API int fse_strip_create(fse_ctx* ctx, int32_t width, int32_t height_global, fse_world** out);
/* owned global rows [own_lo, own_hi) and held rows (owned + ghost) [held_lo, held_hi) */
FSE_API int fse_strip_rows(fse_world* w, int32_t* own_lo, int32_t* own_hi, int32_t* held_lo, int32_t* held_hi);
/* owner-authoritative ghost refresh after edits outside fse_tick (write_rect near a cut, particles) */
FSE_API int fse_strip_refresh(fse_world* w);

/* ---- measurement helpers ------------------------------------------------------- */
/* CUDA events on the world's own stream (torch.cuda.Event only sees torch's). */
FSE_API int fse_timer_start(fse_world* w);
FSE_API int fse_timer_stop(fse_world* w, float* ms_out);
/* kernels launched by this library since fse_ctx_create (bench.py "gpu_launches"). */
FSE_API int64_t fse_launch_count(fse_ctx* ctx);
/* average device time per launch of the dominant (chunk tick) kernel since the
 * last reset, from CUDA events recorded around every launch when enabled. */
FSE_API int fse_kernel_timing_enable(fse_world* w, int enable);
FSE_API int fse_kernel_timing_read(fse_world* w, double* total_ms, int64_t* launches);
/* strip worlds created with FSE_STRIP_TIMELINE=1: per colour phase, ms after the phase start at which the cut-adjacent chunks, the halo
 * exchange and the interior chunks finished (out: n x 3 floats); resets the record */
FSE_API int fse_strip_timeline_read(fse_world* w, float* out, int64_t cap_phases, int64_t* n_out);
FSE_API int fse_kernel_timing_phases(fse_world* w, float* out_ms, int64_t cap, int64_t* n_out);
/* The host-side plan of a call that edits the grid from n boxes on multi-rank strips (bodies, entities, cracks; boxes = n x (x0, y0, x1,
 * y1), global and inclusive, applied in index order), as rank `rank` of `nranks` computes it for a `width` x `height_global` world — pure
 * host code, no device needed: runner[i] = the rank that runs box i (overlapping boxes share one), and the rectangles (x0, y0 in the
 * rank's LOCAL rows, w, h) this rank sends up / receives from above / sends down / receives from below, at most cap_rects of each in
 * rects_out[4][cap_rects][4] with their counts in n_rects[4].  FSE_ESTATE when a box does not fit the rows its runner holds. */
FSE_API int fse_strip_plan(int32_t width, int32_t height_global, int32_t rank, int32_t nranks, const int32_t* boxes, int32_t n, int32_t* runner,
                           int32_t* rects_out, int32_t cap_rects, int32_t* n_rects);
/* profiling aid: cycles each warp role of the tick kernel spent working between step barriers (out[0..3]) and chunks (out[4]) */
FSE_API int fse_debug_role_cycles(fse_world* w, int enable, unsigned long long* out);

#ifdef __cplusplus
}
#endif
#endif /* FSE_H */
