// entity_oracle.cpp — CPU restatement of the entity <-> grid coupling.  TEST INFRASTRUCTURE ONLY (see fse_oracle.hpp).
//
//   entities_tick    world::tickEntities            (world.cpp:3010-3247)
//   entities_stamp   WorldEntitySystem::process     (game/player.cpp:173-199)
//   object_delete    the objectDelete loop          (game.cpp:2128-2139, plane cleared at game.cpp:1705)
//
// Followed line by line, including the float / double mix of the reference's arithmetic (`vy += 0.25`, `vx *= 0.99`,
// `dx += vx / 8.0` are double operations rounded back to f32; loadZone is an MErect of floats).  rand() is replaced by the
// counter RNG keyed on the CELL that is kicked or stamped (a cell is kicked / stamped at most once per call, so the key is unique).
// Box2D calls (SetTransform / SetLinearVelocity, world.cpp:3227-3228) stay with the host and are not part of this path.
#include <cmath>

#include "fse_oracle.hpp"

namespace fseo {

enum { AIR_ = 0, SOLID_ = 1, SAND_ = 2, SOUP_ = 3, OBJECT_ = 5 };
enum : uint32_t { S_ENT_VX = 76, S_ENT_VY = 77, S_STAMP_X = 78, S_STAMP_VX = 79, S_STAMP_VY = 80 };

static inline bool blocks(const World* w, int sx, int sy) {  // SOLID || SAND || OBJECT (world.cpp:3022-3023 and its seven copies)
    const int t = w->tiles[sx + (size_t)sy * w->width].mat->physicsType;
    return t == SOLID_ || t == SAND_ || t == OBJECT_;
}
static inline uint64_t entity_particle_id(uint32_t tick, int kind, int x, int y) {
    return (2ULL << 62) | (1ULL << 61) | ((uint64_t)(kind & 1) << 60) | ((uint64_t)(tick & 0xfffff) << 40) | ((uint64_t)(y & 0xfffff) << 20) | (uint64_t)(x & 0xfffff);
}

// addCell(new CellData(tp, sx, sy, vx, vy, 0, 0.1f)); real_tiles[...] = Tiles_NOTHING; dirty[...] = true (world.cpp:3070-3072 etc.)
static void kick(World* w, uint32_t rkey, uint32_t tick, int sx, int sy, float bx, float by) {
    const size_t i = sx + (size_t)sy * w->width;
    const uint32_t cb = rng_cell(rkey, sx, sy);
    Particle p;
    p.tile = w->tiles[i];
    p.x = (float)sx;
    p.y = (float)sy;
    p.vx = ((int)(rng_draw(cb, S_ENT_VX) % 10) - 5) / 10.0f + bx;
    p.vy = ((int)(rng_draw(cb, S_ENT_VY) % 10) - 5) / 10.0f + by;
    p.ax = 0;
    p.ay = 0.1f;
    p.id = entity_particle_id(tick, 0, sx, sy);
    w->cells.push_back(p);
    w->tiles[i] = w->nothing();
    w->dirty[i] = 1;
}

void entities_tick(World* w, fse_entity* ents, int n, float lzx, float lzy, uint32_t tick, uint32_t seed) {
    const uint32_t rkey = rng_key(seed, tick, 8u);
    const int width = w->width, height = w->height;
    for (int e = 0; e < n; e++) {
        fse_entity* cur = &ents[e];
        cur->destroy = 0;
        int nIntersect = 0, avInX = 0, avInY = 0;  // 3013-3030
        for (int xx = 0; xx < cur->hw; xx++)
            for (int yy = 0; yy < cur->hh; yy++) {
                const int sx = (int)((cur->x + xx) + lzx), sy = (int)((cur->y + yy) + lzy);
                if (sx < 0 || sy < 0 || sx >= width || sy >= height) continue;
                if (blocks(w, sx, sy)) {
                    nIntersect++;
                    avInX += (xx - cur->hw / 2);
                    avInY += (yy - cur->hh / 2);
                }
            }
        if (nIntersect > 0) {  // 3031-3034
            cur->x += avInX > 0 ? -1 : (avInX < 0 ? 1 : 0);
            cur->y += avInY > 0 ? -1 : (avInY < 0 ? 1 : 0);
        }
        cur->vy = (float)(cur->vy + 0.25);  // 3036

        const int dir = cur->vx > 0.001 ? 0 : (cur->vx < -0.001 ? 1 : -1);
        if (dir >= 0) {  // 3038-3094 (vx > 0.001) and 3095-3151 (vx < -0.001): the two copies differ in signs only
            const float stx = cur->x;
            for (float dx = 0; dir == 0 ? dx < cur->vx : dx > cur->vx; dx = (float)(dx + cur->vx / 8.0)) {
                const float nx = stx + dx;
                float ny = cur->y;
                bool collide = false;
                for (int xx = 0; xx < cur->hw; xx++)
                    for (int yy = 0; yy < cur->hh; yy++) {
                        const int sx = (int)((nx + xx) + lzx), sy = (int)((ny + yy) + lzy);
                        if (!(sx >= 0 && sy >= 0 && sx < width && sy < height)) continue;
                        if (!blocks(w, sx, sy)) continue;
                        if (yy == cur->hh - 1) {  // feet: step up one cell if the body fits there (3052-3066)
                            for (int xx1 = 0; xx1 < cur->hw; xx1++)
                                for (int yy1 = 0; yy1 < cur->hh; yy1++) {
                                    const int sx1 = (int)((nx + xx1) + lzx), sy1 = (int)((ny + yy1) + lzy - 1);
                                    if (sx1 >= 0 && sy1 >= 0 && sx1 < width && sy1 < height && blocks(w, sx1, sy1)) collide = true;
                                }
                            if (!collide) ny--;
                        } else if (w->tiles[sx + (size_t)sy * width].mat->physicsType == SAND_) {  // 3068-3074: kick the grain away
                            kick(w, rkey, tick, sx, sy, dir == 0 ? 0.5f : -0.5f, 0.0f);
                            cur->vx = (float)(cur->vx * 0.99);
                        } else {
                            collide = true;
                        }
                    }
                if (!collide) {
                    cur->x = nx;
                    cur->y = ny;
                } else {
                    cur->vx /= 2;
                    break;
                }
            }
        }

        cur->ground = 0;  // 3153

        if (cur->vy > 0.001) {  // 3155-3183: falling, nothing is kicked
            const float sty = cur->y;
            for (float dy = 0; dy < cur->vy; dy = (float)(dy + cur->vy / 8.0)) {
                const float ny = sty + dy, nx = cur->x;
                bool collide = false;
                for (int xx = 0; xx < cur->hw; xx++)
                    for (int yy = 0; yy < cur->hh; yy++) {
                        const int sx = (int)((nx + xx) + lzx), sy = (int)((ny + yy) + lzy);
                        if (sx >= 0 && sy >= 0 && sx < width && sy < height && blocks(w, sx, sy)) collide = true;
                    }
                if (!collide) {
                    cur->y = ny;
                } else {
                    cur->vy /= 2;
                    cur->ground = 1;
                    break;
                }
            }
        } else if (cur->vy < -0.001) {  // 3184-3220: rising, sand overhead is kicked
            const float sty = cur->y;
            for (float dy = 0; dy > cur->vy; dy = (float)(dy + cur->vy / 8.0)) {
                const float ny = sty + dy, nx = cur->x;
                bool collide = false;
                for (int xx = 0; xx < cur->hw; xx++)
                    for (int yy = 0; yy < cur->hh; yy++) {
                        const int sx = (int)((nx + xx) + lzx), sy = (int)((ny + yy) + lzy);
                        if (!(sx >= 0 && sy >= 0 && sx < width && sy < height) || !blocks(w, sx, sy)) continue;
                        if (w->tiles[sx + (size_t)sy * width].mat->physicsType == SAND_) {
                            kick(w, rkey, tick, sx, sy, 0.0f, -0.5f);
                            cur->vy = (float)(cur->vy * 0.99);
                        } else {
                            collide = true;
                        }
                    }
                if (!collide) {
                    cur->y = ny;
                } else {
                    cur->vy /= 2;
                    cur->ground = 1;
                    break;
                }
            }
        }
        if (std::fabs(cur->vx) >= 1024.0f || std::fabs(cur->vy) >= 1024.0f) {  // 3222-3225
            cur->destroy = 1;
            continue;
        }
        cur->vx = (float)(cur->vx * 0.99);  // 3227-3229
        cur->vy = (float)(cur->vy * 0.99);
    }
}

// "entity fluid displacement & make solid" (game/player.cpp:176-196)
static Cell tiles_object(World* w, int object_mat) {  // Tiles_OBJECT = MaterialInstance(&GENERIC_OBJECT, 0x00ff00) (gds.cpp:318)
    Cell c = w->nothing();
    c.mat = &w->mats[object_mat];
    c.id = (uint32_t)object_mat;
    c.color = 0x00ff00;
    return c;
}
void entities_stamp(World* w, const fse_entity* ents, int n, float lzx, float lzy, int object_mat, uint32_t tick, uint32_t seed,
                    std::vector<int64_t>& object_delete) {
    const uint32_t rkey = rng_key(seed, tick, 8u);
    const int width = w->width, height = w->height;
    for (int e = 0; e < n; e++) {
        const fse_entity& pl = ents[e];
        for (int tx = 0; tx < pl.hw; tx++)
            for (int ty = 0; ty < pl.hh; ty++) {
                const int wx = (int)(tx + pl.x + lzx), wy = (int)(ty + pl.y + lzy);
                if (wx < 0 || wy < 0 || wx >= width || wy >= height) continue;
                const size_t i = wx + (size_t)wy * width;
                const int t = w->tiles[i].mat->physicsType;
                if (t == AIR_) {
                    w->tiles[i] = tiles_object(w, object_mat);
                    object_delete.push_back((int64_t)i);
                } else if (t == SAND_ || t == SOUP_) {
                    const uint32_t cb = rng_cell(rkey, wx, wy);
                    Particle p;
                    p.tile = w->tiles[i];
                    p.x = (float)(wx + (int)(rng_draw(cb, S_STAMP_X) % 3) - 1 - pl.vx);
                    p.y = (float)(wy - std::fabs(pl.vy));
                    p.vx = (float)(-pl.vx / 4 + ((int)(rng_draw(cb, S_STAMP_VX) % 10) - 5) / 5.0f);
                    p.vy = (float)(-pl.vy / 4 + -((int)(rng_draw(cb, S_STAMP_VY) % 5) + 5) / 5.0f);
                    p.ax = 0;
                    p.ay = 0.1f;
                    p.id = entity_particle_id(tick, 1, wx, wy);
                    w->cells.push_back(p);
                    w->tiles[i] = tiles_object(w, object_mat);
                    object_delete.push_back((int64_t)i);
                    w->dirty[i] = 1;
                }
            }
    }
}

// game.cpp:2128-2139: every stamped index becomes Tiles_NOTHING again (dirty is not touched)
void object_delete(World* w, std::vector<int64_t>& object_delete) {
    for (int64_t i : object_delete) w->tiles[(size_t)i] = w->nothing();
    object_delete.clear();
}

}  // namespace fseo

using namespace fseo;
#define OAPI __attribute__((visibility("default")))
static std::vector<int64_t> g_object_delete;  // one list per process is enough for the tests (worlds are used one at a time)

extern "C" {
OAPI int fseo_entities_tick(void* p, fse_entity* ents, int n, float lzx, float lzy, uint32_t tick, uint32_t seed) {
    entities_tick((World*)p, ents, n, lzx, lzy, tick, seed);
    return 0;
}
OAPI int fseo_entities_stamp(void* p, const fse_entity* ents, int n, float lzx, float lzy, int object_mat, uint32_t tick, uint32_t seed) {
    entities_stamp((World*)p, ents, n, lzx, lzy, object_mat, tick, seed, g_object_delete);
    return 0;
}
OAPI int fseo_object_delete(void* p) {
    object_delete((World*)p, g_object_delete);
    return 0;
}
}
