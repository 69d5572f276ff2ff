// oracle_capi.cpp — C entry points of the CPU oracle for ctypes (TEST INFRASTRUCTURE ONLY).
// Mirrors include/fse.h so the parity tests drive oracle and CUDA path with the same calls.
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "fse_oracle.hpp"

using namespace fseo;

extern "C" {

#define OAPI __attribute__((visibility("default")))

OAPI void* fseo_world_create(int w, int h) { return new World(w, h); }
OAPI void fseo_world_destroy(void* p) { delete (World*)p; }

OAPI int fseo_materials_set(void* p, const fse_material* tbl, int n, const fse_special_ids* ids, const fse_interaction* inter,
                            const int32_t* inter_offsets, const fse_interaction* react, const int32_t* react_offsets) {
    ((World*)p)->set_materials(tbl, n, *ids, inter, inter_offsets, react, react_offsets);
    return 0;
}

// Two-call export of the default table: pass null arrays to query sizes.
OAPI int fseo_default_materials(uint32_t seed, fse_material* mats, int* n, fse_special_ids* ids, fse_interaction* inter, int* n_inter,
                                int32_t* inter_offsets, fse_interaction* react, int* n_react, int32_t* react_offsets) {
    MaterialTable T = default_materials(seed);
    *n = (int)T.mats.size();
    *n_inter = (int)T.inter.size();
    *n_react = (int)T.react.size();
    if (ids) *ids = T.ids;
    if (mats) std::memcpy(mats, T.mats.data(), T.mats.size() * sizeof(fse_material));
    if (inter && !T.inter.empty()) std::memcpy(inter, T.inter.data(), T.inter.size() * sizeof(fse_interaction));
    if (inter_offsets) std::memcpy(inter_offsets, T.inter_offsets.data(), T.inter_offsets.size() * sizeof(int32_t));
    if (react && !T.react.empty()) std::memcpy(react, T.react.data(), T.react.size() * sizeof(fse_interaction));
    if (react_offsets) std::memcpy(react_offsets, T.react_offsets.data(), T.react_offsets.size() * sizeof(int32_t));
    return 0;
}

OAPI int fseo_write_rect(void* p, int x, int y, int w, int h, const fse_cell* c) {
    ((World*)p)->write_rect(x, y, w, h, c);
    return 0;
}
OAPI int fseo_read_rect(void* p, int x, int y, int w, int h, fse_cell* c) {
    ((World*)p)->read_rect(x, y, w, h, c);
    return 0;
}
OAPI int fseo_clear_dirty(void* p) {
    ((World*)p)->clear_dirty();
    return 0;
}
OAPI int fseo_stats_rect(void* p, int x, int y, int w, int h, fse_stats* out) {
    ((World*)p)->stats_rect(x, y, w, h, out);
    return 0;
}

// schedule: 0 reference, 1 partitioned; rng: 0 slot, 1 libc; returns wall seconds of tick() alone.
OAPI double fseo_tick(void* p, const fse_tick_args* a, int schedule, int rng, int threads) {
    auto t0 = std::chrono::steady_clock::now();
    ((World*)p)->tick(*a, (Schedule)schedule, rng ? RngMode::LIBC : RngMode::SLOT, threads);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// One chunk task (host logic of the multi-process strip test).
OAPI int fseo_run_chunk(void* p, const fse_tick_args* a, int iter, int cx, int cy, int schedule) {
    World* w = (World*)p;
    std::vector<Particle> out;
    w->set_iteration(a->seed, a->tick, iter, RngMode::SLOT);
    w->run_chunk(cx, cy, iter, (Schedule)schedule, out);
    for (auto& q : out) w->add_particle(q);
    return 0;
}

OAPI int fseo_clear_visited(void* p) {
    ((World*)p)->clear_visited();
    return 0;
}

OAPI int fseo_tick_temperature(void* p, const fse_rect* z) {
    Rect r{z->x, z->y, z->w, z->h};
    ((World*)p)->tick_temperature(r);
    return 0;
}

static void to_pod(const World* w, const Particle& s, fse_particle& d) {
    std::memset(&d, 0, sizeof d);
    d.tile.mat = (uint16_t)s.tile.mat->id;
    d.tile.moved = s.tile.moved;
    d.tile.settle = s.tile.settleCount;
    d.tile.color = s.tile.color;
    d.tile.temp = s.tile.temperature;
    d.tile.fluid = s.tile.fluidAmount;
    d.tile.fluid_diff = s.tile.fluidAmountDiff;
    d.x = s.x; d.y = s.y; d.vx = s.vx; d.vy = s.vy; d.ax = s.ax; d.ay = s.ay;
    d.target_x = s.targetX; d.target_y = s.targetY; d.target_force = s.targetForce;
    d.lifetime = s.lifetime; d.fade_time = s.fadeTime;
    d.phase = s.phase; d.temporary = s.temporary; d.in_object_state = s.inObjectState; d.vacuum = s.vacuum ? 1 : 0;
    d.id = s.id;
    (void)w;
}

OAPI int fseo_particles_add(void* p, const fse_particle* src, int n) {
    World* w = (World*)p;
    for (int i = 0; i < n; i++) {
        const fse_particle& s = src[i];
        Particle d;
        d.tile.mat = &w->mats[s.tile.mat];
        d.tile.id = s.tile.mat;
        d.tile.moved = s.tile.moved != 0;
        d.tile.settleCount = s.tile.settle;
        d.tile.color = s.tile.color;
        d.tile.temperature = s.tile.temp;
        d.tile.fluidAmount = s.tile.fluid;
        d.tile.fluidAmountDiff = s.tile.fluid_diff;
        d.x = s.x; d.y = s.y; d.vx = s.vx; d.vy = s.vy; d.ax = s.ax; d.ay = s.ay;
        d.targetX = s.target_x; d.targetY = s.target_y; d.targetForce = s.target_force;
        d.lifetime = s.lifetime; d.fadeTime = s.fade_time;
        d.phase = s.phase != 0; d.temporary = s.temporary != 0; d.inObjectState = s.in_object_state; d.vacuum = s.vacuum != 0;
        d.id = s.id;
        w->add_particle(d);
    }
    return 0;
}
OAPI int64_t fseo_particles_count(void* p) { return (int64_t)((World*)p)->cells.size(); }
OAPI int fseo_particles_read(void* p, fse_particle* out, int64_t cap) {
    World* w = (World*)p;
    int64_t n = std::min<int64_t>(cap, (int64_t)w->cells.size());
    for (int64_t i = 0; i < n; i++) to_pod(w, w->cells[i], out[i]);
    return (int)n;
}
OAPI int fseo_particles_clear(void* p) {
    ((World*)p)->cells.clear();
    return 0;
}
OAPI int fseo_tick_particles(void* p, const fse_rect* z) {
    Rect r{z->x, z->y, z->w, z->h};
    ((World*)p)->tick_particles(r);
    return 0;
}

OAPI int fseo_prt_begin(void* p, const fse_rect* z) {
    Rect r{z->x, z->y, z->w, z->h};
    ((World*)p)->prt_begin(r);
    return 0;
}
OAPI int fseo_prt_propose(void* p) { return ((World*)p)->prt_propose(); }
// copies the proposals whose cell lies in rows [y0, y1) into out (capacity cap); returns how many there are
OAPI int fseo_prt_get(void* p, int y0, int y1, fseo_proposal* out, int cap) {
    World* w = (World*)p;
    int n = 0;
    for (const fseo_proposal& pr : w->prt_props) {
        const int row = (int)(pr.cell / w->width);
        if (row < y0 || row >= y1) continue;
        if (n < cap) out[n] = pr;
        n++;
    }
    return n;
}
OAPI int fseo_prt_commit(void* p, const fseo_proposal* ext, int n_ext, int hold_lo, int hold_hi) {
    ((World*)p)->prt_commit(ext, n_ext, hold_lo, hold_hi);
    return 0;
}
OAPI int fseo_prt_end(void* p) {
    ((World*)p)->prt_end();
    return 0;
}
OAPI int fseo_tick_particles_rounds(void* p, const fse_rect* z, int max_rounds) {
    Rect r{z->x, z->y, z->w, z->h};
    ((World*)p)->tick_particles_rounds(r, max_rounds);
    return 0;
}

OAPI void fseo_srand(unsigned s) { srand(s); }
OAPI uint64_t fseo_cell_hash(int x, int y, const fse_cell* c) { return cell_hash(x, y, *c); }
OAPI uint32_t fseo_rng_draw(uint32_t seed, uint32_t tick, uint32_t iter, int x, int y, uint32_t slot) {
    return rng_draw(rng_cell(rng_key(seed, tick, iter), x, y), slot);
}
}
