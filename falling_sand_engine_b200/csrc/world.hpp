// world.hpp — header-only C++ host shim over include/fse.h that keeps the reference's `class world` method names for
// the tick path (source/engine/world.hpp:148-192).  This is the piece a maintainer drops next to the reference's
// world.cpp: every method forwards to one C-ABI call; errors become exceptions because the
// reference's methods return void.  Links against libfse_b200.so only — no CUDA headers needed by the caller.
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fse.h"
#include "polygons.hpp"

namespace fse_host {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
inline void check(int rc) {
    if (rc != FSE_OK) throw Error(std::string("fse: ") + fse_last_error());
}

// MaterialInstance value fields (game_datastruct.hpp:207-225) <-> fse_cell
using Cell = fse_cell;

// Flattened result of InitMaterials()/RegisterMaterial()/PushMaterials() (game_datastruct.cpp:117-300).  The reference's
// Lua binds materials_init / materials_register / materials_push (game_basic.cpp:79-81) fill this table on the host.
struct MaterialTable {
    std::vector<fse_material> mats;
    fse_special_ids ids{};
    std::vector<fse_interaction> inter;   // grouped by (material, other material)
    std::vector<int32_t> inter_offsets;   // n*n+1
    std::vector<fse_interaction> react;   // grouped by material
    std::vector<int32_t> react_offsets;   // n+1

    // RegisterMaterial(s_id, name, index_name, physicsType, slipperyness, alpha, density, iterations, emit, emitColor, color)
    int materials_register(int physicsType, int slipperyness, uint8_t alpha, float density, int iterations, int emit, uint32_t emitColor,
                           uint32_t color) {
        const int n0 = (int)mats.size(), n = n0 + 1;
        fse_material m{};
        m.physics = physicsType; m.slipperyness = slipperyness; m.alpha = alpha; m.density = density; m.iterations = iterations;
        m.emit = emit; m.emit_color = emitColor; m.color = color; m.conduction_self = 1.0f; m.conduction_other = 1.0f;
        mats.push_back(m);
        std::vector<int32_t> io((size_t)n * n + 1, 0);
        std::vector<fse_interaction> flat;
        for (int a = 0; a < n; a++)
            for (int b = 0; b < n; b++) {
                if (a < n0 && b < n0 && !inter_offsets.empty())
                    for (int k = inter_offsets[a * n0 + b]; k < inter_offsets[a * n0 + b + 1]; k++) flat.push_back(inter[k]);
                io[(size_t)a * n + b + 1] = (int32_t)flat.size();
            }
        inter.swap(flat);
        inter_offsets.swap(io);
        react_offsets.resize(n + 1, react_offsets.empty() ? 0 : react_offsets.back());
        return n0;
    }
};

class Context {
public:
    explicit Context(int device = 0) { check(fse_ctx_create(device, &h_)); }
    ~Context() { fse_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    // materials_push(): hand the whole table to the device
    void materials_push(const MaterialTable& t) {
        check(fse_materials_set(h_, t.mats.data(), (int)t.mats.size(), &t.ids, t.inter.empty() ? nullptr : t.inter.data(),
                                t.inter_offsets.empty() ? nullptr : t.inter_offsets.data(), t.react.empty() ? nullptr : t.react.data(),
                                t.react_offsets.empty() ? nullptr : t.react_offsets.data()));
    }
    fse_ctx* handle() const { return h_; }
    // multi-GPU (one process per GPU, include/fse.h "multi-GPU"): rank 0 makes the 128-byte id, the host carries it to the other
    // ranks by whatever channel it has (MPI, a socket, a file), every rank joins
    static void commUniqueId(void* out128) { check(fse_comm_unique_id(out128)); }
    void commInit(int rank, int nranks, const void* id128) { check(fse_comm_init(h_, rank, nranks, id128)); }

private:
    fse_ctx* h_ = nullptr;
};

// Mirror of the reference's `class world` for the tick path.
class world {
public:
    int32_t width = 0, height = 0;   // world.hpp:121-122 has u16; the C ABI takes int32 and so does this shim (65536-wide worlds)
    fse_rect tickZone{};             // world.hpp zone; game.cpp:1629 keeps it one chunk inside the grid
    uint32_t tickCt = 0;
    uint32_t seed = 1337;            // replaces srand(time(NULL)) (game_utils/rng.cpp:9)
    int cell_iter = 3;               // globaldef.cell_iter (data/scripts/global.lua:47)

    // world::init(path, w, h, ...) (world.cpp:43-172)
    void init(Context& ctx, int w, int h) {
        check(fse_world_create(ctx.handle(), w, h, &h_));
        width = w;
        height = h;
        tickZone = {FSE_CHUNK, FSE_CHUNK, w - 2 * FSE_CHUNK, h - 2 * FSE_CHUNK};
    }
    // one strip of a w x hGlobal world on this rank (after Context::commInit); width / height / tickZone and every coordinate of every
    // method stay GLOBAL, every rank makes the same calls in the same order (gameTick is SPMD), results are those of the single world
    void initStrip(Context& ctx, int w, int hGlobal) {
        check(fse_strip_create(ctx.handle(), w, hGlobal, &h_));
        width = w;
        height = hGlobal;
        tickZone = {FSE_CHUNK, FSE_CHUNK, w - 2 * FSE_CHUNK, hGlobal - 2 * FSE_CHUNK};
        check(fse_strip_rows(h_, &ownLo, &ownHi, &heldLo, &heldHi));
    }
    int32_t ownLo = 0, ownHi = 0, heldLo = 0, heldHi = 0;  // strips: rows this rank owns / holds (owned + ghost); chunk loads go to the rank that holds them
    ~world() { fse_world_destroy(h_); }

    // world.cpp:999-1008
    Cell getTile(int x, int y) {
        Cell c{};
        if (x < 0 || x >= width || y < 0 || y >= height) {  // Tiles_TEST_SOLID
            c.mat = 1;
            c.color = 0xff0000;
            c.fluid = 2.0f;
            return c;
        }
        check(fse_read_rect(h_, x, y, 1, 1, &c));
        return c;
    }
    void setTile(int x, int y, Cell c) {
        if (x < 0 || x >= width || y < 0 || y >= height) return;
        c.dirty = 1;
        check(fse_write_rect(h_, x, y, 1, 1, &c));
    }
    // frame() merge (world.cpp:2374-2391) / chunkSaveCache (2780-2792): whole chunks in and out
    void writeChunk(int cx, int cy, const Cell* cells) { check(fse_write_rect(h_, cx, cy, FSE_CHUNK, FSE_CHUNK, cells)); }
    void readChunk(int cx, int cy, Cell* cells) { check(fse_read_rect(h_, cx, cy, FSE_CHUNK, FSE_CHUNK, cells)); }

    // a body the physicsCheck probe cut loose: the caller makes the b2Body at (res.x, res.y), gives it the random velocity of
    // world.cpp:3377 and runs updateRigidBodyHitbox on it
    struct CutOut {
        fse_physcheck_result res;
        std::vector<fse_cell> tiles;
    };
    std::vector<CutOut> cutOuts;  // since the caller last drained it

    // world::tick() (world.cpp:1036-1948), with the probe at its end (1929-1934)
    void tick() {
        fse_tick_args a{tickCt, seed, cell_iter, tickZone};
        check(fse_tick(h_, &a));
        int32_t px = 0, py = 0;
        fse_probe_position(seed, tickCt, &tickZone, &px, &py);
        tickCt++;
        CutOut c;
        if (physicsCheckCut(px, py, c.res, c.tiles)) cutOuts.push_back(std::move(c));
    }
    void tickTemperature() { check(fse_tick_temperature(h_, &tickZone)); }  // world.cpp:1950
    void tickCells() { check(fse_particles_tick(h_, &tickZone)); }          // world.cpp:2030
    void addCell(const fse_particle& p) { check(fse_particles_add(h_, &p, 1)); }  // world.cpp:2292

    // game.cpp:1711-1815 / 1896-1983; `damp` applies the velocity damping the reference does inline
    template <class Damp>
    void rasterBodies(const std::vector<fse_xform>& xf, Damp damp) {
        std::vector<fse_body_feedback> fb(xf.size());
        check(fse_bodies_raster(h_, xf.data(), (int)xf.size(), tickCt, seed, fb.data()));
        for (size_t i = 0; i < xf.size(); i++)
            damp(i, std::pow(0.99f, (float)fb[i].sand_hits) * std::pow(0.998f, (float)fb[i].soup_hits),
                 std::pow(0.98f, (float)fb[i].sand_hits) * std::pow(0.99f, (float)fb[i].soup_hits));
    }
    void eraseBodies(const std::vector<fse_xform>& xf, std::vector<uint8_t>& needsUpdate) {
        needsUpdate.assign(xf.size(), 0);
        check(fse_bodies_erase(h_, xf.data(), (int)xf.size(), nullptr, needsUpdate.data()));
    }
    // world::physicsCheck(x, y) (world.cpp:3330): returns the component size (cap+1 = too large, 0 = not solid)
    int physicsCheck(int x, int y, std::vector<int32_t>& pixels, int32_t bbox[4]) {
        int32_t n = 0;
        pixels.assign(1001, 0);
        check(fse_flood_component(h_, x, y, 1000, &n, bbox, pixels.data()));
        pixels.resize(n <= 1000 ? n : 0);
        return n;
    }
    // the whole of world::physicsCheck: crumbs are deleted, a loose component of 11..1000 cells leaves the grid as the tile array of a
    // new body (the caller makes the b2Body at (res.x, res.y) and runs updateRigidBodyHitbox on it); returns true when a body was cut out
    bool physicsCheckCut(int x, int y, fse_physcheck_result& res, std::vector<fse_cell>& tiles) {
        tiles.assign(1 << 16, fse_cell{});
        int rc = fse_physics_check(h_, x, y, &res, tiles.data(), (int32_t)tiles.size());
        if (rc == FSE_EINVAL && res.count > 10 && res.count <= 1000) {  // a sprawling component: its box did not fit, nothing was changed
            tiles.assign((size_t)res.w * res.h, fse_cell{});
            rc = fse_physics_check(h_, x, y, &res, tiles.data(), (int32_t)tiles.size());
        }
        check(rc);
        tiles.resize(res.action == 2 ? (size_t)res.w * res.h : 0);
        return res.action == 2;
    }
    // world::updateRigidBodyHitbox (world.cpp:288-720) for an uploaded body: the device cuts it into connected pieces (crop box,
    // rotated shift, weld flag, tile arrays: fse_bodies_split), traces and simplifies the outlines of every piece (fse_mask_outline)
    // and polygons.hpp turns them into the triangles of the new bodies' b2PolygonShapes.  The caller creates the b2Bodies (pose and
    // velocities copied, weld joint where `rec.weld`, world.cpp:657-707) and destroys the old one.
    struct HitboxPiece {
        fse_body_piece rec;
        std::vector<fse_cell> tiles;                         // rec.w x rec.h
        std::vector<std::vector<Triangle>> shapes;           // one group per outer polygon of the piece
    };
    std::vector<HitboxPiece> updateRigidBodyHitbox(int body, int bodyW, int bodyH, float angle, int weldX = -1, int weldY = -1) {
        std::vector<fse_body_piece> recs(1024);
        std::vector<fse_cell> tiles((size_t)4 * bodyW * bodyH + 64);
        int32_t n = 0;
        check(fse_bodies_split(h_, body, angle, weldX, weldY, recs.data(), (int32_t)recs.size(), &n, tiles.data(), (int64_t)tiles.size()));
        std::vector<HitboxPiece> out((size_t)n);
        for (int k = 0; k < n; k++) {
            HitboxPiece& p = out[(size_t)k];
            p.rec = recs[(size_t)k];
            const size_t cells = (size_t)p.rec.w * p.rec.h;
            p.tiles.assign(tiles.begin() + p.rec.tile_off, tiles.begin() + p.rec.tile_off + (long)cells);
            std::vector<uint8_t> mask(cells);
            for (size_t i = 0; i < cells; i++) mask[i] = p.tiles[i].mat != 0;  // data[] = "alpha != 0" (world.cpp:395-402); AIR is material 0
            std::vector<float> pts(4 * cells + 128);
            std::vector<int32_t> ptOff(cells / 2 + 66), maskOff(2);
            check(fse_mask_outline(h_, mask.data(), 1, p.rec.w, p.rec.h, nullptr, nullptr, pts.data(), (int32_t)(pts.size() / 2), ptOff.data(),
                                   (int32_t)ptOff.size() - 1, maskOff.data()));
            std::vector<std::vector<Vec2d>> outlines((size_t)maskOff[1]);
            for (int c = 0; c < maskOff[1]; c++)
                for (int q = ptOff[c]; q < ptOff[c + 1]; q++) outlines[(size_t)c].push_back(Vec2d{pts[2 * q], pts[2 * q + 1]});
            p.shapes = hitbox_triangles(outlines);
        }
        return out;
    }
    // world::explosion(x, y, r) (world.cpp:2294)
    void explosion(int x, int y, int r) { check(fse_explosion(h_, x, y, r, tickCt, seed)); }
    // world::tickChunks() grid + particle shift (world.cpp:2454-2478, 2579-2582); chunk load / save stays with the caller
    void scroll(int changeX, int changeY) { check(fse_scroll(h_, changeX, changeY)); }
    // the dirty -> texture loops of game::tick (game.cpp:1994-2126) followed by the dirty clears (game.cpp:2153-2155)
    void renderDirty(fse_render_stats* movingTiles = nullptr, bool drawBackgroundGrid = false) {
        check(fse_pixels_enable(h_, 1));
        check(fse_render_dirty(h_, movingTiles));
        check(fse_render_layers(h_, drawBackgroundGrid ? 1 : 0, nullptr, nullptr));
        check(fse_clear_dirty(h_));
    }
    // world::flowX / flowY (world.hpp:116-119): kept by the tick from now on, drawn into the flow texture by renderDirty
    void enableFlow(bool on = true) { check(fse_flow_enable(h_, on ? 1 : 0)); }
    // setTileLayer2 (world.cpp:1015-1019) and the background part of the chunk merge (world.cpp:2384-2389)
    void setTileLayer2(int x, int y, const fse_cell& c) { check(fse_layer2_write_rect(h_, x, y, 1, 1, &c)); }
    fse_cell getTileLayer2(int x, int y) {
        fse_cell c{};
        check(fse_layer2_read_rect(h_, x, y, 1, 1, &c));
        return c;
    }
    void setBackground(int x, int y, int w, int h, const uint32_t* argb) { check(fse_background_write_rect(h_, x, y, w, h, argb)); }
    // world::tickEntities (world.cpp:3010-3247), WorldEntitySystem::process (game/player.cpp:173-199), objectDelete (game.cpp:2128-2139)
    void tickEntities(std::vector<fse_entity>& ents) {
        if (!ents.empty()) check(fse_entities_tick(h_, ents.data(), (int)ents.size(), loadZoneX, loadZoneY, tickCt, seed));
    }
    void stampEntities(const std::vector<fse_entity>& ents, int objectMat = 6) {
        if (!ents.empty()) check(fse_entities_stamp(h_, ents.data(), (int)ents.size(), loadZoneX, loadZoneY, objectMat, tickCt, seed));
    }
    void objectDelete() { check(fse_object_delete(h_)); }
    float loadZoneX = 0, loadZoneY = 0;  // world::loadZone.x / .y

    // One game tick in the order of game::tick (game.cpp:1659-2201): body raster (1711-1815), tickEntities (1820), entity
    // stamping (1835), world tick (1838), tickCells (1878-1886, joined at 1892), body erase (1896-1983), tickTemperature on
    // tick % GameTick == 2 (2157), objectDelete (2128-2139), dirty -> textures + dirty clear (1994-2060, 2153).  `damp` receives the
    // velocity factors of the raster pass; Box2D itself (world.cpp:2264-2276) is the caller's, between two calls.
    template <class Damp>
    void gameTick(const std::vector<fse_xform>& bodies, std::vector<fse_entity>& ents, Damp damp, fse_render_stats* movingTiles = nullptr) {
        std::vector<uint8_t> needsUpdate;
        if (!bodies.empty()) rasterBodies(bodies, damp);
        tickEntities(ents);
        stampEntities(ents);
        const uint32_t t = tickCt;
        tick();  // advances tickCt
        tickCells();
        if (!bodies.empty()) eraseBodies(bodies, needsUpdate);
        if (t % 4 == 2) tickTemperature();
        objectDelete();
        renderDirty(movingTiles);
    }
    void stats(fse_stats* out) { check(fse_stats_rect(h_, 0, 0, width, height, out)); }
    void sync() { check(fse_sync(h_)); }
    fse_world* handle() const { return h_; }

private:
    fse_world* h_ = nullptr;
};

}  // namespace fse_host
