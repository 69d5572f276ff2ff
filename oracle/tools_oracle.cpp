// tools_oracle.cpp — CPU restatement of the grid side of the reference's interactive tools.
// TEST INFRASTRUCTURE ONLY (see fse_oracle.hpp).
//
//   for_line / for_line_cornered   world::forLine, world::forLineCornered        (world.cpp:3250-3313)
//   tool_erase_line                middle-mouse erase brush                       (game.cpp:593-625)
//   tool_pickaxe                   "break with pickaxe"                           (game.cpp:762-808, grid part: 771-790)
//   tool_hammer                    hammer release                                 (game.cpp:843-910, grid part: 856-890)
//   tool_vacuum                    vacuum: aim walk, suck the disc, re-energise   (game.cpp:2456-2585)
//   particles_vacuum_pull          the vacuumCells update                         (game.cpp:2640-2664)
//
// rand() -> the counter RNG: hammer jitter keyed on the segment number, the vacuum's clip radius on (0, 0), per-cell draws on the
// cell, re-energised particles on their id.  Rigid-body pixels (the surfaces the reference edits next to the grid) stay with the
// host's body list; Box2D / audio / UI calls are not part of this path.
#include <cmath>
#include <functional>

#include "fse_oracle.hpp"

namespace fseo {

enum { AIR_ = 0, SOLID_ = 1, SAND_ = 2, SOUP_ = 3, OBJECT_ = 5 };
enum : uint32_t { S_HAMMER_JX = 81, S_HAMMER_JY = 82, S_VAC_CLIP = 83, S_VAC_VX = 84, S_VAC_VY = 85 };

// world.cpp:3250-3276
static void for_line(int width, int x0, int y0, int x1, int y1, const std::function<bool(long)>& fn) {
    const int dx = x1 - x0, dy = y1 - y0;
    int dLong = std::abs(dx), dShort = std::abs(dy);
    long offsetLong = dx > 0 ? 1 : -1, offsetShort = dy > 0 ? width : -width;
    if (dLong < dShort) {
        std::swap(dShort, dLong);
        std::swap(offsetShort, offsetLong);
    }
    int error = dLong / 2;
    long index = (long)y0 * width + x0;
    const long offset[] = {offsetLong, offsetLong + offsetShort};
    const int abs_d[] = {dShort, dShort - dLong};
    for (int i = 0; i <= dLong; ++i) {
        if (fn(index)) return;
        const int errorIsTooBig = error >= dLong;
        index += offset[errorIsTooBig];
        error += abs_d[errorIsTooBig];
    }
}

// world.cpp:3278-3313 (f32 arithmetic, libm cos / sin / atan2 as in the reference; the visited list removes repeats)
static void for_line_cornered(int width, int x0, int y0, int x1, int y1, const std::function<bool(long)>& fn) {
    const float sx = (float)x0, sy = (float)y0, ex = (float)x1, ey = (float)y1;
    float x = std::floor(sx), y = std::floor(sy);
    const float diffX = ex - sx, diffY = ey - sy;
    const float stepX = (diffX > 0) ? 1 : ((diffX < 0) ? -1 : 0);
    const float stepY = (diffY > 0) ? 1 : ((diffY < 0) ? -1 : 0);
    const float xOffset = ex > sx ? (std::ceil(sx) - sx) : (sx - std::floor(sx));
    const float yOffset = ey > sy ? (std::ceil(sy) - sy) : (sy - std::floor(sy));
    const float angle = (float)std::atan2(-diffY, diffX);
    float tMaxX = (float)(xOffset / std::cos(angle));
    float tMaxY = (float)(yOffset / std::sin(angle));
    const float tDeltaX = (float)(1.0 / std::cos(angle));
    const float tDeltaY = (float)(1.0 / std::sin(angle));
    const float manhattanDistance = std::abs(std::floor(ex) - std::floor(sx)) + std::abs(std::floor(ey) - std::floor(sy));
    std::vector<long> visited;
    for (int t = 0; t <= manhattanDistance; ++t) {
        const long idx = (long)(x + y * width);
        bool seen = false;
        for (long v : visited) seen |= v == idx;
        if (!seen && fn(idx)) return;
        visited.push_back(idx);
        if (std::abs(tMaxX) < std::abs(tMaxY) || std::isnan(tMaxY)) {
            tMaxX += tDeltaX;
            x += stepX;
        } else {
            tMaxY += tDeltaY;
            y += stepY;
        }
    }
}

static inline bool in_world(const World* w, long x, long y) { return x >= 0 && y >= 0 && x < w->width && y < w->height; }

// game.cpp:593-625: every non-AIR cell under the brush along the line becomes Tiles_NOTHING through setTile (dirty)
int tool_erase_line(World* w, int x0, int y0, int x1, int y1, int brush) {
    int n = 0;
    for_line(w->width, x0, y0, x1, y1, [&](long index) {
        const int lineX = (int)(index % w->width), lineY = (int)(index / w->width);
        for (int xx = -brush / 2; xx < (int)std::ceil(brush / 2.0); xx++)
            for (int yy = -brush / 2; yy < (int)std::ceil(brush / 2.0); yy++) {
                if (std::abs(xx) + std::abs(yy) == brush) continue;
                if (!in_world(w, lineX + xx, lineY + yy)) continue;  // getTile out of bounds = TEST_SOLID, setTile out of bounds ignored
                const size_t i = (lineX + xx) + (size_t)(lineY + yy) * w->width;
                if (w->tiles[i].mat->physicsType != AIR_) {
                    w->tiles[i] = w->nothing();
                    w->dirty[i] = 1;
                    n++;
                }
            }
        return false;
    });
    return n;
}

// game.cpp:771-790: SOLID cells inside the circle of diameter breakSize leave the grid; their colours fill the ARGB surface of the
// rigid body the host builds from them (pixels row-major xx + yy * size, 0 where nothing was taken)
int tool_pickaxe(World* w, int x, int y, float breakSize, uint32_t* pixels) {
    const int size = (int)breakSize;
    for (int i = 0; i < size * size; i++) pixels[i] = 0;
    int n = 0;
    for (int xx = 0; xx < breakSize; xx++)
        for (int yy = 0; yy < breakSize; yy++) {
            const float cx = (float)((xx / breakSize) - 0.5), cy = (float)((yy / breakSize) - 0.5);
            if (cx * cx + cy * cy > 0.25f) continue;
            if (!in_world(w, x + xx, y + yy)) continue;
            const size_t i = (x + xx) + (size_t)(y + yy) * w->width;
            if (w->tiles[i].mat->physicsType == SOLID_) {
                if (xx < size && yy < size) pixels[xx + yy * size] = w->tiles[i].color;
                w->tiles[i] = w->nothing();
                w->dirty[i] = 1;
                n++;
            }
        }
    return n;
}

static inline uint32_t darken(uint32_t color, float brightness) {  // ME_draw_darken_color (renderer/gpu.cpp:194-201)
    const int a = (color >> 24) & 0xFF;
    const int r = (int)(((color >> 16) & 0xFF) * brightness), g = (int)(((color >> 8) & 0xFF) * brightness), b = (int)((color & 0xFF) * brightness);
    return ((uint32_t)a << 24) | ((uint32_t)r << 16) | ((uint32_t)g << 8) | (uint32_t)b;
}

// game.cpp:843-890: the crack from the hammer point away from the release point, in jittered segments of ~10 cells; SOLID cells on
// it become GENERIC_SAND at half brightness until the crack leaves the solid.  out = {end_x, end_y, n_changed, broke}.
void tool_hammer(World* w, int hammerX, int hammerY, int x, int y, int sand_mat, uint32_t tick, uint32_t seed, int32_t* out) {
    const uint32_t rkey = rng_key(seed, tick, 9u);
    const int dx = hammerX - x, dy = hammerY - y;
    const float len = std::sqrt((float)(dx * dx + dy * dy));
    const int nSegments = (int)(1 + len / 10);
    std::vector<std::pair<int, int>> points;
    for (int i = 0; i < nSegments; i++) {
        int sx = hammerX + (int)((float)(dx / nSegments) * (i + 1));
        int sy = hammerY + (int)((float)(dy / nSegments) * (i + 1));
        const uint32_t cb = rng_cell(rkey, i, 0);
        sx += (int)(rng_draw(cb, S_HAMMER_JX) % 3) - 1;
        sy += (int)(rng_draw(cb, S_HAMMER_JY) % 3) - 1;
        points.push_back({sx, sy});
    }
    long endInd = -1;
    int nTilesChanged = 0;
    bool broke = false;
    for (size_t i = 0; i < points.size(); i++) {
        const int segSx = i == 0 ? hammerX : points[i - 1].first, segSy = i == 0 ? hammerY : points[i - 1].second;
        const int segEx = points[i].first, segEy = points[i].second;
        bool hitSolidYet = false;
        broke = false;
        for_line_cornered(w->width, segSx, segSy, segEx, segEy, [&](long index) {
            if (index < 0 || index >= (long)w->width * w->height) return false;  // the reference indexes unchecked; the oracle skips
            if (w->tiles[index].mat->physicsType != SOLID_) {
                if (hitSolidYet && (std::abs((int)(index % w->width) - segSx) + std::abs((int)(index / w->width) - segSy) > 1)) {
                    broke = true;
                    return true;
                }
                return false;
            }
            hitSolidYet = true;
            Cell c = w->nothing();  // MaterialInstance(&GENERIC_SAND, darken(color, 0.5f)): temperature 0, defaults elsewhere
            c.mat = &w->mats[sand_mat];
            c.id = (uint32_t)sand_mat;
            c.color = darken(w->tiles[index].color, 0.5f);
            w->tiles[index] = c;
            w->dirty[index] = 1;
            endInd = index;
            nTilesChanged++;
            return false;
        });
        if (broke) break;
    }
    out[0] = endInd < 0 ? -1 : (int)(endInd % w->width);
    out[1] = endInd < 0 ? -1 : (int)(endInd / w->width);
    out[2] = nTilesChanged;
    out[3] = broke ? 1 : 0;
}

static inline uint64_t vacuum_particle_id(uint32_t tick, int x, int y) {
    return (3ULL << 62) | (1ULL << 61) | ((uint64_t)(tick & 0x1fffff) << 40) | ((uint64_t)(y & 0xfffff) << 20) | (uint64_t)(x & 0xfffff);
}

// game.cpp:2456-2585.  out = {x, y, cells sucked, particles re-energised}; nothing happens beyond 256 cells (2465)
void tool_vacuum(World* w, int wcx, int wcy, int wmx, int wmy, uint32_t tick, uint32_t seed, int32_t* out) {
    out[0] = out[1] = -1;
    out[2] = out[3] = 0;
    const int mdx = wmx - wcx, mdy = wmy - wcy;
    if (mdx * mdx + mdy * mdy > 256 * 256) return;
    const uint32_t rkey = rng_key(seed, tick, 10u);
    long sind = -1;
    bool inObject = true;
    for_line(w->width, wcx, wcy, wmx, wmy, [&](long ind) {  // 2468-2486
        if (ind < 0 || ind >= (long)w->width * w->height) return false;
        const int t = w->tiles[ind].mat->physicsType;
        if (t == OBJECT_) {
            if (!inObject) {
                sind = ind;
                return true;
            }
        } else {
            inObject = false;
        }
        if (t == SOLID_ || t == SAND_ || t == SOUP_) {
            sind = ind;
            return true;
        }
        return false;
    });
    const int x = sind == -1 ? wmx : (int)(sind % w->width), y = sind == -1 ? wmy : (int)(sind / w->width);
    out[0] = x;
    out[1] = y;
    auto energise = [&](Particle& par, uint32_t cb) {  // 2492-2505 and 2548-2560
        par.vx = ((int)(rng_draw(cb, S_VAC_VX) % 10) - 5) / 5.0f * 1.0f;
        par.vy = ((int)(rng_draw(cb, S_VAC_VY) % 10) - 5) / 5.0f * 1.0f;
        par.ax = -par.vx / 10.0f;
        par.ay = -par.vy / 10.0f;
        if (par.ay == 0 && par.ax == 0) par.ay = 0.01f;
        par.lifetime = 6;
        par.phase = true;
        par.vacuum = true;
    };
    const int rad = 5;
    int clipRadSq = rad * rad;
    clipRadSq += (int)(rng_draw(rng_cell(rkey, 0, 0), S_VAC_CLIP) % (uint32_t)clipRadSq) / 4;  // 2527
    const size_t n_before = w->cells.size();
    for (int xx = -rad; xx <= rad; xx++)
        for (int yy = -rad; yy <= rad; yy++) {
            if (xx * xx + yy * yy > clipRadSq) continue;
            if ((yy == -rad || yy == rad) && (xx == -rad || xx == rad)) continue;
            if (!in_world(w, x + xx, y + yy)) continue;
            const size_t i = (x + xx) + (size_t)(y + yy) * w->width;
            const int t = w->tiles[i].mat->physicsType;
            if (t == SOLID_ || t == SAND_ || t == SOUP_) {
                Particle par;  // CellData(tile, xPos, yPos, 0, 0, 0, 0.01f)
                par.tile = w->tiles[i];
                par.x = (float)(x + xx);
                par.y = (float)(y + yy);
                par.ay = 0.01f;
                energise(par, rng_cell(rkey, x + xx, y + yy));
                par.id = vacuum_particle_id(tick, x + xx, y + yy);
                w->cells.push_back(par);
                w->tiles[i] = w->nothing();
                w->dirty[i] = 1;
                out[2]++;
            }
        }
    for (size_t k = 0; k < n_before; k++) {  // 2543-2583 (the particles just made have phase set and are skipped there anyway)
        Particle& cur = w->cells[k];
        if (!(cur.targetForce == 0 && !cur.phase)) continue;
        bool hit = false;
        for (int xx = -rad; xx <= rad && !hit; xx++)
            for (int yy = -rad; yy <= rad; yy++) {
                if ((yy == -rad || yy == rad) && (xx == -rad || x == rad)) continue;  // sic: `x == rad` (2547)
                if (((int)(cur.x) == (x + xx)) && ((int)(cur.y) == (y + yy))) {
                    hit = true;
                    break;
                }
            }
        if (hit) {
            energise(cur, rng_cell(rkey, (int)(cur.id & 0xffffffffu), (int)(cur.id >> 32)));
            out[3]++;
        }
    }
}

// game.cpp:2640-2664: particles held by the vacuum turn towards the player once their lifetime is up and are collected within
// 10 cells (temporary + lifetime 0: tickCells drops them on its next visit).  Returns the number collected.
int particles_vacuum_pull(World* w, float tx, float ty) {
    int n = 0;
    for (Particle& cur : w->cells) {
        if (!cur.vacuum) continue;
        if (cur.lifetime <= 0) {
            cur.targetForce = 0.45f;
            cur.targetX = tx;
            cur.targetY = ty;
            cur.ax = 0;
            cur.ay = 0.01f;
        }
        const float tdx = cur.targetX - cur.x, tdy = cur.targetY - cur.y;
        if (tdx * tdx + tdy * tdy < 10 * 10) {
            cur.temporary = true;
            cur.lifetime = 0;
            cur.vacuum = false;
            n++;
        }
    }
    return n;
}

}  // namespace fseo

using namespace fseo;
#define OAPI __attribute__((visibility("default")))
extern "C" {
OAPI int fseo_tool_erase_line(void* p, int x0, int y0, int x1, int y1, int brush) { return tool_erase_line((World*)p, x0, y0, x1, y1, brush); }
OAPI int fseo_tool_pickaxe(void* p, int x, int y, float size, uint32_t* pixels) { return tool_pickaxe((World*)p, x, y, size, pixels); }
OAPI int fseo_tool_hammer(void* p, int hx, int hy, int x, int y, int sand_mat, uint32_t tick, uint32_t seed, int32_t* out) {
    tool_hammer((World*)p, hx, hy, x, y, sand_mat, tick, seed, out);
    return 0;
}
OAPI int fseo_tool_vacuum(void* p, int wcx, int wcy, int wmx, int wmy, uint32_t tick, uint32_t seed, int32_t* out) {
    tool_vacuum((World*)p, wcx, wcy, wmx, wmy, tick, seed, out);
    return 0;
}
OAPI int fseo_particles_vacuum_pull(void* p, float tx, float ty) { return particles_vacuum_pull((World*)p, tx, ty); }
}
