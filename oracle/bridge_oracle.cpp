// bridge_oracle.cpp — CPU oracle (TEST INFRASTRUCTURE ONLY) for the rigid-body <-> grid bridge and the fracture outline
// pipeline.  Citations relative to /root/reference/source/engine.
//
//   raster / erase        game.cpp:1711-1815 / 1896-1983 (sequential: bodies in order, pixels tx-major)
//   marching squares      physics/physics_math.cpp:1870-1965 (value / FindPerimeter, saddles 6 and 9 by previous direction)
//   Douglas-Peucker       physics/physics_math.cpp:1766-1843 (simplify / simplify_section / pDistance, tolerance 1)
//   contour discovery     world.cpp:399-509 (updateRigidBodyHitbox) — see outlines() for the one documented deviation
//   flood fill            world.cpp:3330-3429 (physicsCheck + 4-way flood, cap 1000)
//   component labels      no reference counterpart (the reference assigns pixels by nearest triangle centroid,
//                         world.cpp:587-610); north_star prescribes 4-connected labelling, checked against this CCL.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "fse_oracle.hpp"

namespace fseo {

enum { AIR = 0, SOLID = 1, SAND = 2, SOUP = 3 };

static Cell from_pod(World* w, const fse_cell& s) {
    Cell d;
    d.mat = &w->mats[s.mat];
    d.id = s.mat;
    d.color = s.color;
    d.temperature = s.temp;
    d.moved = s.moved != 0;
    d.settleCount = s.settle;
    d.fluidAmount = s.fluid;
    d.fluidAmountDiff = s.fluid_diff;
    return d;
}
static void to_pod(const Cell& s, fse_cell& d) {
    std::memset(&d, 0, sizeof d);
    d.mat = (uint16_t)s.mat->id;
    d.color = s.color;
    d.temp = s.temperature;
    d.moved = s.moved;
    d.settle = s.settleCount;
    d.fluid = s.fluidAmount;
    d.fluid_diff = s.fluidAmountDiff;
}

static const int DIRS[5][2] = {{0, 0}, {1, 0}, {-1, 0}, {0, 1}, {0, -1}};  // game.cpp:1766

// game.cpp:1711-1815.  tiles: per body w*h fse_cell (AIR = empty).  feedback[4*b] = {sand hits, soup hits, placed, 0}.
void bodies_raster(World* W, int n, const int* bw, const int* bh, fse_cell* const* tiles, const fse_xform* xf, uint32_t tick,
                   uint32_t seed, int32_t* feedback) {
    const uint32_t key = rng_key(seed, tick, 7);
    for (int b = 0; b < n; b++) {
        const float x = xf[b].x, y = xf[b].y;
        const float s = std::sin(xf[b].angle), c = std::cos(xf[b].angle);
        int32_t* fb = feedback + 4 * b;
        fb[0] = fb[1] = fb[2] = fb[3] = 0;
        for (int tx = 0; tx < bw[b]; tx++)
            for (int ty = 0; ty < bh[b]; ty++) {
                const fse_cell& rm = tiles[b][tx + ty * bw[b]];
                if ((int)rm.mat == W->ids.air) continue;
                int wx = (int)(tx * c - (ty + 1) * s + x);
                int wy = (int)(tx * s + (ty + 1) * c + y);
                for (auto& d : DIRS) {
                    int wxd = wx + d[0], wyd = wy + d[1];
                    if (wxd < 0 || wyd < 0 || wxd >= W->width || wyd >= W->height) continue;
                    Cell& cell = W->tiles[wxd + wyd * W->width];
                    int ph = cell.mat->physicsType;
                    if (ph == AIR) {
                        cell = from_pod(W, rm);
                        W->dirty[wxd + wyd * W->width] = 1;
                        fb[2]++;
                        break;
                    } else if (ph == SAND || ph == SOUP) {
                        Particle p;  // game.cpp:1791 / 1801: the displaced cell is thrown up as a loose particle
                        p.tile = cell;
                        p.x = (float)wxd;
                        p.y = (float)(wyd - 3);
                        uint32_t cb = rng_cell(key, b, tx + ty * bw[b]);
                        p.vx = (float)(((int)(rng_draw(cb, S_BRIDGE_VX) % 10) - 5) / 10.0f);
                        p.vy = (float)(-(int)(rng_draw(cb, S_BRIDGE_VY) % 5 + 5) / 10.0f);
                        p.ax = 0;
                        p.ay = 0.1f;
                        p.id = (2ULL << 62) | ((uint64_t)(tick & 0xffff) << 40) | ((uint64_t)(b & 0xfffff) << 20) |
                               (uint64_t)((tx + ty * bw[b]) & 0xfffff);
                        W->add_particle(p);
                        cell = from_pod(W, rm);
                        W->dirty[wxd + wyd * W->width] = 1;
                        fb[ph == SAND ? 0 : 1]++;
                        fb[2]++;
                        break;
                    }
                }
            }
    }
}

// game.cpp:1896-1983.  feedback[4*b+3] = pixels destroyed.
void bodies_erase(World* W, int n, const int* bw, const int* bh, fse_cell* const* tiles, const fse_xform* xf, int32_t* feedback) {
    for (int b = 0; b < n; b++) {
        const float x = xf[b].x, y = xf[b].y;
        const float s = std::sin(xf[b].angle), c = std::cos(xf[b].angle);
        int32_t* fb = feedback + 4 * b;
        fb[0] = fb[1] = fb[2] = fb[3] = 0;
        for (int tx = 0; tx < bw[b]; tx++)
            for (int ty = 0; ty < bh[b]; ty++) {
                fse_cell& rm = tiles[b][tx + ty * bw[b]];
                if ((int)rm.mat == W->ids.air) continue;
                int wx = (int)(tx * c - (ty + 1) * s + x);
                int wy = (int)(tx * s + (ty + 1) * c + y);
                bool found = false;
                for (auto& d : DIRS) {
                    int wxd = wx + d[0], wyd = wy + d[1];
                    if (wxd < 0 || wyd < 0 || wxd >= W->width || wyd >= W->height) continue;
                    Cell& cell = W->tiles[wxd + wyd * W->width];
                    if (cell.mat->id == rm.mat) {  // .id == rmat.id: any cell of the same material
                        to_pod(cell, rm);
                        cell = W->nothing();
                        W->dirty[wxd + wyd * W->width] = 1;
                        found = true;
                        fb[2]++;
                        break;
                    }
                }
                if (!found) {  // game.cpp:1959-1965
                    if (wx >= 0 && wy >= 0 && wx < W->width && wy < W->height && (int)W->tiles[wx + wy * W->width].mat->id == W->ids.air) {
                        Cell nn = W->nothing();
                        to_pod(nn, rm);
                        fb[3]++;
                    }
                }
            }
    }
}

// ---- marching squares (physics_math.cpp:1893-1965) ----------------------------------------------------------
static inline bool isSet(int x, int y, int w, int h, const uint8_t* d) { return x <= 0 || x > w || y <= 0 || y > h ? false : d[(y - 1) * w + (x - 1)] != 0; }
static inline int msValue(int x, int y, int w, int h, const uint8_t* d) {
    int sum = 0;
    if (isSet(x, y, w, h, d)) sum |= 1;
    if (isSet(x + 1, y, w, h, d)) sum |= 2;
    if (isSet(x, y + 1, w, h, d)) sum |= 4;
    if (isSet(x + 1, y + 1, w, h, d)) sum |= 8;
    return sum;
}
// direction codes: 0 none, 1 East(1,0), 2 North(0,1), 3 West(-1,0), 4 South(0,-1)
static inline int msDir(int v, int prev) {
    switch (v) {
        case 1: return 2;
        case 2: return 1;
        case 3: return 1;
        case 4: return 3;
        case 5: return 2;
        case 6: return prev == 2 ? 3 : 1;
        case 7: return 1;
        case 8: return 4;
        case 9: return prev == 1 ? 2 : 4;
        case 10: return 4;
        case 11: return 4;
        case 12: return 3;
        case 13: return 2;
        case 14: return 3;
    }
    return 0;
}
static const int DX[5] = {0, 1, 0, -1, 0}, DY[5] = {0, 0, 1, 0, -1};

// physics_math.cpp:1813-1843
static float pDistance(float x, float y, float x1, float y1, float x2, float y2) {
    float A = x - x1, B = y - y1, C = x2 - x1, D = y2 - y1;
    float dot = A * C + B * D;
    float len_sq = C * C + D * D;
    float param = -1;
    if (len_sq != 0) param = dot / len_sq;
    float xx, yy;
    if (param < 0) {
        xx = x1;
        yy = y1;
    } else if (param > 1) {
        xx = x2;
        yy = y2;
    } else {
        xx = x1 + param * C;
        yy = y1 + param * D;
    }
    float dx = x - xx, dy = y - yy;
    return std::sqrt(dx * dx + dy * dy);
}
// physics_math.cpp:1766-1811 (the `omitted` guard only ever sees pts.size(), it is passed by value)
static void simplify_section(const std::vector<float>& px, const std::vector<float>& py, float tol, size_t i, size_t j, std::vector<char>& mark) {
    if (px.size() <= 2) return;
    if (i + 1 == j) return;
    float maxd = -1.0f;
    size_t maxi = i;
    for (size_t k = i + 1; k < j; k++) {
        float d = pDistance(px[k], py[k], px[i], py[i], px[j], py[j]);
        if (d > maxd) {
            maxd = d;
            maxi = k;
        }
    }
    if (maxd <= tol) {
        for (size_t k = i + 1; k < j; k++) mark[k] = 0;
    } else {
        simplify_section(px, py, tol, i, maxi, mark);
        simplify_section(px, py, tol, maxi, j, mark);
    }
}

// Contours of a w*h mask.  Discovery follows world.cpp:412-451: a start candidate is a set pixel whose right / down /
// down-right neighbours are not all set and whose vertex value is not 0 or 15; FindPerimeter runs from it.  The
// reference suppresses re-tracing with an `edgeSeen` pixel map updated while it scans (world.cpp:442, 464-481); here
// a loop is emitted once, from the lowest-index candidate that lies on it (same loops and start points except for
// edgeSeen's index-clamping artefacts at the mask border) — that rule is order-free, so the GPU can apply it in parallel.
// Output: for each contour, simplified points (x0,y0,x1,y1,...), in discovery order.
void outlines(const uint8_t* data, int w, int h, std::vector<std::vector<float>>& out) {
    out.clear();
    const int size = w * h;
    auto is_cand = [&](int i) {
        if (!data[i]) return false;
        int x = i % w, y = i / w, nb = 0;
        if (x + 1 < w) nb += data[i + 1] != 0;
        if (y + 1 < h) nb += data[i + w] != 0;
        if (y + 1 < h && x + 1 < w) nb += data[i + w + 1] != 0;
        if (nb == 3) return false;
        int v = msValue(x, y, w, h, data);
        return v != 0 && v != 15;
    };
    for (int i = 0; i < size; i++) {
        if (!is_cand(i)) continue;
        const int sx = i % w, sy = i / w;
        // trace (FindPerimeter, physics_math.cpp:1893-1965)
        std::vector<int> dirs, lens;
        int x = sx, y = sy, prev = 0;
        bool canonical = true;
        do {
            int d = msDir(msValue(x, y, w, h, data), prev);
            if (!(x == sx && y == sy && prev == 0)) {
                // does a lower-index candidate start this same loop from here?
                if (x >= 0 && y >= 0 && x < w && y < h) {
                    int j = x + y * w;
                    if (j < i && is_cand(j) && msDir(msValue(x, y, w, h, data), 0) == d) {
                        canonical = false;
                        break;
                    }
                }
            }
            if (d == prev) lens.back()++;
            else {
                dirs.push_back(d);
                lens.push_back(1);
                prev = d;
            }
            x += DX[d];
            y -= DY[d];
        } while (x != sx || y != sy);
        if (!canonical) continue;
        std::vector<float> px, py;
        float lx = (float)sx, ly = (float)sy;
        for (size_t k = 0; k < dirs.size(); k++) {  // world.cpp:483-485
            lx += (float)(DX[dirs[k]] * lens[k]);
            ly -= (float)(DY[dirs[k]] * lens[k]);
            px.push_back(lx);
            py.push_back(ly);
        }
        std::vector<char> mark(px.size(), 1);
        if (!px.empty()) simplify_section(px, py, 1.0f, 0, px.size() - 1, mark);  // world.cpp:488
        std::vector<float> poly;
        for (size_t k = 0; k < px.size(); k++)
            if (mark[k]) {
                poly.push_back(px[k]);
                poly.push_back(py[k]);
            }
        if (poly.size() < 6) continue;  // world.cpp:490: fewer than 3 points
        out.push_back(poly);
    }
}

// 4-connected component labels: label = lowest row-major pixel index of the component, -1 for unset pixels.
int ccl(const uint8_t* data, int w, int h, int32_t* labels) {
    const int n = w * h;
    std::vector<int> stack;
    for (int i = 0; i < n; i++) labels[i] = -1;
    int count = 0;
    for (int i = 0; i < n; i++) {
        if (!data[i] || labels[i] >= 0) continue;
        count++;
        stack.assign(1, i);
        labels[i] = i;
        while (!stack.empty()) {
            int p = stack.back();
            stack.pop_back();
            int x = p % w, y = p / w;
            const int nb[4] = {x + 1 < w ? p + 1 : -1, y + 1 < h ? p + w : -1, x > 0 ? p - 1 : -1, y > 0 ? p - w : -1};
            for (int q : nb)
                if (q >= 0 && data[q] && labels[q] < 0) {
                    labels[q] = i;
                    stack.push_back(q);
                }
        }
    }
    return count;
}

// world.cpp:3330-3429: size, bounding box and pixels of the 4-connected SOLID component at (x, y); count = cap+1 when it
// is larger than cap (the reference then does nothing), 0 when the seed is not SOLID.
int flood_component(World* W, int x, int y, int cap, int* bbox, int32_t* pixels) {
    if (x < 0 || y < 0 || x >= W->width || y >= W->height) return 0;
    if (W->tiles[x + y * W->width].mat->physicsType != SOLID) return 0;
    std::vector<int> stack{x + y * W->width}, seen;
    std::vector<char> vis((size_t)W->width * W->height, 0);
    vis[stack[0]] = 1;
    while (!stack.empty()) {
        int p = stack.back();
        stack.pop_back();
        seen.push_back(p);
        if ((int)seen.size() > cap) return cap + 1;
        int px = p % W->width, py = p / W->width;
        const int nb[4][2] = {{px + 1, py}, {px, py + 1}, {px - 1, py}, {px, py - 1}};
        for (auto& q : nb) {
            if (q[0] < 0 || q[1] < 0 || q[0] >= W->width || q[1] >= W->height) continue;
            int qi = q[0] + q[1] * W->width;
            if (!vis[qi] && W->tiles[qi].mat->physicsType == SOLID) {
                vis[qi] = 1;
                stack.push_back(qi);
            }
        }
    }
    std::sort(seen.begin(), seen.end());
    bbox[0] = W->width; bbox[1] = W->height; bbox[2] = 0; bbox[3] = 0;
    for (size_t k = 0; k < seen.size(); k++) {
        int px = seen[k] % W->width, py = seen[k] / W->width;
        bbox[0] = std::min(bbox[0], px); bbox[1] = std::min(bbox[1], py);
        bbox[2] = std::max(bbox[2], px); bbox[3] = std::max(bbox[3], py);
        if (pixels) pixels[k] = seen[k];
    }
    return (int)seen.size();
}

// world::physicsCheck (world.cpp:3330-3411): the 4-connected SOLID component at (x, y), abandoned beyond 1000 cells; 11..1000 cells
// are cut out of the grid (Tiles_NOTHING, dirty) into the tile array of a new rigid body — what makeRigidBody builds from the
// surface of their colours (world.cpp:191-209): OBSIDIAN carrying the cell's colour, AIR elsewhere in the bounding box; 1..10 cells
// are simply deleted.  res = {count (1001 = abandoned, 0 = seed not SOLID), action (0 none, 1 deleted, 2 cut out), min x, min y, w, h};
// tiles (w * h, row-major) is written for action 2 when it holds at least w * h cells, otherwise nothing is changed and -1 returned.
int physics_check(World* W, int x, int y, int32_t* res, fse_cell* tiles, int cap_tiles) {
    for (int q = 0; q < 6; q++) res[q] = 0;
    std::vector<int32_t> px(1001);
    int bbox[4];
    const int count = flood_component(W, x, y, 1000, bbox, px.data());
    res[0] = count;
    if (count <= 0 || count > 1000) return 0;
    res[2] = bbox[0]; res[3] = bbox[1]; res[4] = bbox[2] - bbox[0] + 1; res[5] = bbox[3] - bbox[1] + 1;
    if (count > 10) {
        if ((long long)res[4] * res[5] > cap_tiles) return -1;
        fse_cell air;
        std::memset(&air, 0, sizeof air);
        air.mat = (uint16_t)W->ids.air;
        air.fluid = 2.0f;
        for (int i = 0; i < res[4] * res[5]; i++) tiles[i] = air;
        const Material& ob = W->mats[W->ids.obsidian];
        for (int k = 0; k < count; k++) {
            const int cx = px[k] % W->width, cy = px[k] / W->width;
            fse_cell t = air;
            t.mat = (uint16_t)ob.id;
            t.color = W->tiles[px[k]].color;
            t.temp = ob.createTemp;
            tiles[(cx - res[2]) + (cy - res[3]) * res[4]] = t;
        }
        res[1] = 2;
    } else {
        res[1] = 1;
    }
    for (int k = 0; k < count; k++) {
        W->tiles[px[k]] = W->nothing();
        W->dirty[px[k]] = 1;
    }
    return 0;
}

}  // namespace fseo

using namespace fseo;
extern "C" {
__attribute__((visibility("default"))) int fseo_physics_check(void* p, int x, int y, int32_t* res, fse_cell* tiles, int cap_tiles) {
    return physics_check((World*)p, x, y, res, tiles, cap_tiles);
}
// world::explosion (world.cpp:2294-2332) with rand() replaced by the counter RNG keyed on (seed, tick, x, y) — cells decide
// independently, so the loop order is immaterial.
void explosion(World* w, int cx, int cy, int radius, uint32_t tick, uint32_t seed) {
    const uint32_t rkey = rng_key(seed, tick, 7u);
    const int outer = radius * 2;
    for (int x = cx - outer; x < cx + outer; x++) {
        for (int y = cy - outer; y < cy + outer; y++) {
            if (x < 0 || y < 0 || x >= w->width || y >= w->height) continue;  // getTile OOB = TEST_SOLID, setTile OOB ignored (world.cpp:999-1008)
            Cell tile = w->tiles[x + (size_t)y * w->width];
            if (tile.mat->physicsType == 0) continue;
            const int dx = x - cx, dy = y - cy;
            const uint32_t cb = rng_cell(rkey, x, y);
            bool particle = false, inner = false;
            if (dx * dx + dy * dy < radius * radius) {
                inner = true;
                if (!(tile.mat->physicsType == 1 || rng_draw(cb, S_EXPL_KEEP) % 10 < 6)) particle = true;
            } else if (dx * dx + dy * dy < outer * outer && tile.mat->physicsType != 1) {
                particle = true;
            } else {
                continue;
            }
            if (particle) {
                Particle p;
                if (inner) {
                    const uint32_t r = (tile.color >> 16) & 0xFF, g = (tile.color >> 8) & 0xFF, b = tile.color & 0xFF;
                    tile.color = ((r / 4) << 16) | ((g / 4) << 8) | (b / 4);
                }
                p.tile = tile;
                p.x = (float)x;
                p.y = (float)(inner ? y + 1 : y);
                p.vx = dx / 10.0f + ((int)(rng_draw(cb, S_EXPL_VX) % 10) - 5) / 10.0f;
                p.vy = dy / 6.0f + ((int)(rng_draw(cb, S_EXPL_VY) % 10) - 5) / 10.0f;
                p.ax = 0;
                p.ay = 0.1f;
                p.id = (3ULL << 62) | ((uint64_t)(tick & 0x3fffff) << 40) | ((uint64_t)(y & 0xfffff) << 20) | (uint64_t)(x & 0xfffff);
                w->cells.push_back(p);
            }
            w->tiles[x + (size_t)y * w->width] = w->nothing();
            w->dirty[x + (size_t)y * w->width] = 1;
        }
    }
}

#define OAPI __attribute__((visibility("default")))

OAPI int fseo_bodies_raster(void* p, int n, const int* bw, const int* bh, fse_cell* const* tiles, const fse_xform* xf, uint32_t tick,
                            uint32_t seed, int32_t* feedback) {
    bodies_raster((World*)p, n, bw, bh, tiles, xf, tick, seed, feedback);
    return 0;
}
// The dirty -> texture loop of game::tick (game.cpp:1994-2066).  px_main / px_fire / px_emis / px_flow: width*height RGBA8 texels in the
// reference's byte order (r, g, b, a).  The flow texture and the flowX / flowY reset (2017-2018, 2040-2062) run when the world carries
// the flow accumulators (World::enable_flows) and px_flow is given; had[2] counts the dirty SOUP cells (hadFlow).
void render_dirty(World* w, uint8_t* px_main, uint8_t* px_fire, uint8_t* px_emis, uint8_t* px_flow, int64_t* moving, int64_t* had) {
    had[0] = had[1] = had[2] = 0;
    const bool flows = w->flowX_ != nullptr && px_flow != nullptr;
    for (int i = 0; i < w->n_materials(); i++) moving[i] = 0;
    for (int i = 0; i < w->width * w->height; i++) {
        const unsigned int offset = (unsigned int)i * 4;
        if (!w->dirty[i]) continue;
        had[0]++;
        const Cell& t = w->tiles[i];
        moving[t.mat->id]++;
        if (t.mat->physicsType == AIR) {
            for (int q = 0; q < 4; q++) px_main[offset + q] = px_fire[offset + q] = px_emis[offset + q] = 0;  // ME_ALPHA_TRANSPARENT = 0
            if (flows) w->flowY_[i] = w->flowX_[i] = 0;  // 2017-2018
            continue;
        }
        const uint32_t color = t.color, emit = t.mat->emitColor;
        px_main[offset + 2] = (color >> 0) & 0xff;
        px_main[offset + 1] = (color >> 8) & 0xff;
        px_main[offset + 0] = (color >> 16) & 0xff;
        px_main[offset + 3] = t.mat->alpha;
        px_emis[offset + 2] = (emit >> 0) & 0xff;
        px_emis[offset + 1] = (emit >> 8) & 0xff;
        px_emis[offset + 0] = (emit >> 16) & 0xff;
        px_emis[offset + 3] = (emit >> 24) & 0xff;
        if ((int)t.mat->id == w->ids.fire) {
            px_fire[offset + 2] = (color >> 0) & 0xff;
            px_fire[offset + 1] = (color >> 8) & 0xff;
            px_fire[offset + 0] = (color >> 16) & 0xff;
            px_fire[offset + 3] = t.mat->alpha;
            had[1]++;
        }
        if (!flows) continue;
        if (t.mat->physicsType == SOUP) {  // 2040-2059: the literals 0.25, 0.5, 3.0, 4.0 are doubles in the reference
            float newFlowX = w->prevFlowX[i] + (w->flowX_[i] - w->prevFlowX[i]) * 0.25;
            float newFlowY = w->prevFlowY[i] + (w->flowY_[i] - w->prevFlowY[i]) * 0.25;
            if (newFlowY < 0) newFlowY *= 0.5;
            double a;
            px_flow[offset + 2] = 0;
            a = newFlowY * (3.0 / t.mat->iterations + 0.5) / 4.0 + 0.5;
            px_flow[offset + 1] = std::min(std::max(a, 0.0), 1.0) * 255;
            a = newFlowX * (3.0 / t.mat->iterations + 0.5) / 4.0 + 0.5;
            px_flow[offset + 0] = std::min(std::max(a, 0.0), 1.0) * 255;
            px_flow[offset + 3] = 0xff;
            had[2]++;
            w->prevFlowX[i] = newFlowX;
            w->prevFlowY[i] = newFlowY;
        }
        w->flowY_[i] = 0;
        w->flowX_[i] = 0;
    }
}

// The layer-2 and background loops of the same function (game.cpp:2068-2126): dirty layer-2 cells -> RGBA (AIR: transparent, or the
// grey checker of globaldef.draw_background_grid), dirty background cells -> their ARGB colour; had[0] / had[1] = cells of each kind.
// Clears the two dirty planes afterwards like game.cpp:2154-2155.
void render_layers(World* w, int draw_background_grid, uint8_t* px_layer2, uint8_t* px_bg, int64_t* had) {
    had[0] = had[1] = 0;
    if (w->layer2Id.empty()) return;
    for (int i = 0; i < w->width * w->height; i++) {
        const unsigned int offset = (unsigned int)i * 4;
        if (w->layer2Dirty[i]) {
            had[0]++;
            const Material& m = w->mats[w->layer2Id[i]];
            if (m.physicsType == AIR) {
                if (draw_background_grid) {
                    const uint32_t color = (i % 2) == 0 ? 0x888888 : 0x444444;
                    px_layer2[offset + 2] = (color >> 0) & 0xff;
                    px_layer2[offset + 1] = (color >> 8) & 0xff;
                    px_layer2[offset + 0] = (color >> 16) & 0xff;
                    px_layer2[offset + 3] = 0xff;  // ME_ALPHA_OPAQUE
                } else {
                    for (int q = 0; q < 4; q++) px_layer2[offset + q] = 0;
                }
            } else {
                const uint32_t color = w->layer2Color[i];
                px_layer2[offset + 2] = (color >> 0) & 0xff;
                px_layer2[offset + 1] = (color >> 8) & 0xff;
                px_layer2[offset + 0] = (color >> 16) & 0xff;
                px_layer2[offset + 3] = m.alpha;
            }
        }
        if (w->backgroundDirty[i]) {
            had[1]++;
            const uint32_t color = w->background[i];
            px_bg[offset + 2] = (color >> 0) & 0xff;
            px_bg[offset + 1] = (color >> 8) & 0xff;
            px_bg[offset + 0] = (color >> 16) & 0xff;
            px_bg[offset + 3] = (color >> 24) & 0xff;
        }
    }
    if (had[0]) std::fill(w->layer2Dirty.begin(), w->layer2Dirty.end(), 0);
    if (had[1]) std::fill(w->backgroundDirty.begin(), w->backgroundDirty.end(), 0);
}

// setTileLayer2 over a rectangle / the layer-2 and background part of the chunk merge (world.cpp:1015-1019, 2384-2389)
void layer2_write_rect(World* w, int x0, int y0, int rw, int rh, const fse_cell* src) {
    w->enable_layers();
    for (int y = 0; y < rh; y++)
        for (int x = 0; x < rw; x++) {
            const fse_cell& s = src[x + (size_t)y * rw];
            const size_t i = (size_t)(x0 + x) + (size_t)(y0 + y) * w->width;
            w->layer2Id[i] = s.mat;
            w->layer2Color[i] = s.color;
            w->layer2Temp[i] = s.temp;
            w->layer2Dirty[i] = 1;
        }
}
void layer2_read_rect(World* w, int x0, int y0, int rw, int rh, fse_cell* dst) {
    w->enable_layers();
    for (int y = 0; y < rh; y++)
        for (int x = 0; x < rw; x++) {
            fse_cell& d = dst[x + (size_t)y * rw];
            const size_t i = (size_t)(x0 + x) + (size_t)(y0 + y) * w->width;
            std::memset(&d, 0, sizeof d);
            d.mat = (uint16_t)w->layer2Id[i];
            d.color = w->layer2Color[i];
            d.temp = w->layer2Temp[i];
            d.fluid = 2.0f;  // MaterialInstance default (game_datastruct.hpp:216); chunk files do not keep it
            d.dirty = w->layer2Dirty[i];
        }
}
void background_write_rect(World* w, int x0, int y0, int rw, int rh, const uint32_t* src) {
    w->enable_layers();
    for (int y = 0; y < rh; y++)
        for (int x = 0; x < rw; x++) {
            const size_t i = (size_t)(x0 + x) + (size_t)(y0 + y) * w->width;
            w->background[i] = src[x + (size_t)y * rw];
            w->backgroundDirty[i] = 1;
        }
}
void background_read_rect(World* w, int x0, int y0, int rw, int rh, uint32_t* dst) {
    w->enable_layers();
    for (int y = 0; y < rh; y++)
        for (int x = 0; x < rw; x++) dst[x + (size_t)y * rw] = w->background[(size_t)(x0 + x) + (size_t)(y0 + y) * w->width];
}

// The grid shift of world::tickChunks (world.cpp:2454-2478) and the particle shift (2579-2582), loops as in the reference.
void scroll(World* w, int changeX, int changeY) {
    const int width = w->width, height = w->height;
    if (changeX != 0 || changeY != 0) {
        const bool revX = changeX > 0, revY = changeY > 0;
        for (int y = 0; y < height; y++) {
            const int oldY = revY ? (height - y - 1) : y;
            const int newY = oldY + changeY;
            if (newY < 0 || newY >= height) continue;
            for (int x = 0; x < width; x++) {
                const int oldX = revX ? (width - x - 1) : x;
                const int newX = oldX + changeX;
                if (newX >= 0 && newX < width) {
                    w->tiles[newX + newY * width] = w->tiles[oldX + oldY * width];
                    if (!w->layer2Id.empty()) {  // background and real_layer2 move with the grid (2475-2476); their dirty planes do not
                        w->background[newX + newY * width] = w->background[oldX + oldY * width];
                        w->layer2Id[newX + newY * width] = w->layer2Id[oldX + oldY * width];
                        w->layer2Color[newX + newY * width] = w->layer2Color[oldX + oldY * width];
                        w->layer2Temp[newX + newY * width] = w->layer2Temp[oldX + oldY * width];
                    }
                }
            }
        }
        for (auto& p : w->cells) {
            p.x += changeX;
            p.y += changeY;
        }
    }
}

OAPI int fseo_render_dirty(void* p, uint8_t* px_main, uint8_t* px_fire, uint8_t* px_emis, uint8_t* px_flow, int64_t* moving, int64_t* had) {
    render_dirty((World*)p, px_main, px_fire, px_emis, px_flow, moving, had);
    return 0;
}
OAPI int fseo_render_layers(void* p, int grid, uint8_t* px_layer2, uint8_t* px_bg, int64_t* had) {
    render_layers((World*)p, grid, px_layer2, px_bg, had);
    return 0;
}
OAPI int fseo_layer2_write_rect(void* p, int x, int y, int w, int h, const fse_cell* c) { layer2_write_rect((World*)p, x, y, w, h, c); return 0; }
OAPI int fseo_layer2_read_rect(void* p, int x, int y, int w, int h, fse_cell* c) { layer2_read_rect((World*)p, x, y, w, h, c); return 0; }
OAPI int fseo_background_write_rect(void* p, int x, int y, int w, int h, const uint32_t* c) { background_write_rect((World*)p, x, y, w, h, c); return 0; }
OAPI int fseo_background_read_rect(void* p, int x, int y, int w, int h, uint32_t* c) { background_read_rect((World*)p, x, y, w, h, c); return 0; }
OAPI int fseo_flow_enable(void* p) { ((World*)p)->enable_flows(); return 0; }
// which: 0 flowX, 1 flowY, 2 prevFlowX, 3 prevFlowY; whole planes
OAPI int fseo_flow_read(void* p, int which, float* out) {
    World* w = (World*)p;
    if (!w->flowX_) return -1;
    const std::vector<float>& v = which == 0 ? w->flowX : which == 1 ? w->flowY : which == 2 ? w->prevFlowX : w->prevFlowY;
    std::memcpy(out, v.data(), v.size() * sizeof(float));
    return 0;
}
OAPI int fseo_scroll(void* p, int dx, int dy) {
    scroll((World*)p, dx, dy);
    return 0;
}
OAPI int fseo_explosion(void* p, int cx, int cy, int radius, uint32_t tick, uint32_t seed) {
    explosion((World*)p, cx, cy, radius, tick, seed);
    return 0;
}
OAPI int fseo_bodies_erase(void* p, int n, const int* bw, const int* bh, fse_cell* const* tiles, const fse_xform* xf, int32_t* feedback) {
    bodies_erase((World*)p, n, bw, bh, tiles, xf, feedback);
    return 0;
}
// Flattened contours: pts (x,y pairs) and offsets[n_contours+1] in points; returns n_contours (or -needed if caps too small).
OAPI int fseo_outlines(const uint8_t* data, int w, int h, float* pts, int cap_pts, int32_t* offsets, int cap_contours) {
    std::vector<std::vector<float>> out;
    outlines(data, w, h, out);
    int total = 0;
    for (auto& c : out) total += (int)c.size() / 2;
    if ((int)out.size() > cap_contours || total > cap_pts) return -std::max((int)out.size(), total);
    int o = 0;
    for (size_t k = 0; k < out.size(); k++) {
        offsets[k] = o;
        std::memcpy(pts + 2 * o, out[k].data(), out[k].size() * sizeof(float));
        o += (int)out[k].size() / 2;
    }
    offsets[out.size()] = o;
    return (int)out.size();
}
OAPI int fseo_ccl(const uint8_t* data, int w, int h, int32_t* labels) { return ccl(data, w, h, labels); }
// the two leaf functions of the outline path on their own (pinned against the reference's compiled code in tests/test_ref_pins.py)
OAPI int fseo_ms_value(const uint8_t* data, int w, int h, int x, int y) { return msValue(x, y, w, h, data); }
OAPI float fseo_p_distance(float x, float y, float x1, float y1, float x2, float y2) { return pDistance(x, y, x1, y1, x2, y2); }

// Fracture hand-off restated on the CPU (world::updateRigidBodyHitbox, world.cpp:288-720, with 4-connected membership instead of
// the nearest-centroid assignment of 587-610): per component, in the order of its first pixel, the cropped tile array (305-320),
// the weld flag (620) and the rotated shift of the box corner (350-362).  Returns the number of pieces; -1 on overflow.
OAPI int fseo_body_split(const fse_cell* tiles, int bw, int bh, int air, float angle, int weld_x, int weld_y, fse_body_piece* pieces, int cap_pieces,
                         fse_cell* tiles_out, long cap_tiles) {
    std::vector<uint8_t> mask((size_t)bw * bh);
    for (int i = 0; i < bw * bh; i++) mask[i] = tiles[i].mat != air;
    std::vector<int32_t> labels((size_t)bw * bh);
    ccl(mask.data(), bw, bh, labels.data());
    int n = 0;
    long off = 0;
    const float s = std::sin(angle), c = std::cos(angle);
    for (int root = 0; root < bw * bh; root++) {
        if (labels[root] != root) continue;
        if (n >= cap_pieces) return -1;
        int x0 = bw, y0 = bh, x1 = -1, y1 = -1, cnt = 0, weld = 0;
        for (int i = 0; i < bw * bh; i++) {
            if (labels[i] != root) continue;
            const int x = i % bw, y = i / bw;
            x0 = std::min(x0, x); y0 = std::min(y0, y); x1 = std::max(x1, x); y1 = std::max(y1, y);
            cnt++;
            if (x == weld_x && y == weld_y) weld = 1;
        }
        fse_body_piece& pc = pieces[n++];
        pc.x0 = x0; pc.y0 = y0; pc.w = x1 - x0 + 1; pc.h = y1 - y0 + 1;
        pc.n_pixels = cnt; pc.weld = weld; pc.tile_off = (int32_t)off;
        pc.shift_x = x0 * c - y0 * s;
        pc.shift_y = x0 * s + y0 * c;
        if (off + (long)pc.w * pc.h > cap_tiles) return -1;
        for (int y = 0; y < pc.h; y++)
            for (int x = 0; x < pc.w; x++) {
                const int i = (x0 + x) + (y0 + y) * bw;
                fse_cell t;
                if (labels[i] == root) {
                    t = tiles[i];
                } else {
                    std::memset(&t, 0, sizeof t);
                    t.mat = (uint16_t)air;
                    t.fluid = 2.0f;
                }
                tiles_out[off + x + (long)y * pc.w] = t;
            }
        off += (long)pc.w * pc.h;
    }
    return n;
}
OAPI int fseo_flood_component(void* p, int x, int y, int cap, int* bbox, int32_t* pixels) {
    return flood_component((World*)p, x, y, cap, bbox, pixels);
}
}
