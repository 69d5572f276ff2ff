// fse_tools.cu — grid edits of the reference's interactive tools that run as kernels.
//   fse_explosion             world::explosion                      (world.cpp:2294-2332)
//   fse_tool_erase_line       middle-mouse erase brush              (game.cpp:593-625)
//   fse_tool_pickaxe          "break with pickaxe", grid part       (game.cpp:771-790)
//   fse_tool_hammer           hammer release, grid part             (game.cpp:856-890)
//   fse_tool_vacuum           vacuum: aim walk, disc, re-energise   (game.cpp:2456-2585)
//   fse_particles_vacuum_pull the vacuumCells update                (game.cpp:2640-2664)
// Line rasterisation (world::forLine / forLineCornered, world.cpp:3250-3313) is a pure function of the end points and runs on the
// host with the host's libm, like the reference; the kernels get the list of cells and do everything that touches the grid.
#include <cmath>
#include <cstring>
#include <vector>

#include "fse_device.cuh"
#include "fse_internal.hpp"

namespace fse {

struct ExplArgs {
    Planes p;
    const DevTables* T;
    int W, H;
    int cx, cy, radius;  // centre in global coordinates
    int y_off;           // global row of local row 0 (strip worlds)
    int own_lo, own_hi;  // global rows whose cells this rank turns into particles (every rank clears the rows it holds)
    uint32_t rkey, tick;
    fse_particle* pbuf;
    unsigned int* pcount;
    unsigned int pcap;
};

// one thread per cell of the 4r x 4r square; cells decide independently (the reference's loop carries no state but rand())
__global__ void explosion_kernel(const ExplArgs a) {
    const int outer = a.radius * 2;
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y;
    if (ix >= 2 * outer) return;
    const int x = a.cx - outer + ix, y = a.cy - outer + iy;  // global
    const int yl = y - a.y_off;
    if (x < 0 || yl < 0 || x >= a.W || yl >= a.H) return;  // getTile out of bounds is TEST_SOLID, setTile out of bounds is ignored
    const size_t g = (size_t)yl * a.W + x;
    const uint8_t m = a.p.mat[g];
    const int ph = a.T->phys[m];
    if (ph == P_AIR) return;
    const int dx = x - a.cx, dy = y - a.cy;
    const int d2 = dx * dx + dy * dy;
    const bool inner = d2 < a.radius * a.radius;
    if (!inner && !(d2 < outer * outer && ph != P_SOLID)) return;
    const uint32_t cb = rng_cell(a.rkey, x, y);
    if ((!inner || !(ph == P_SOLID || rng_draw(cb, S_EXPL_KEEP) % 10 < 6)) && y >= a.own_lo && y < a.own_hi) {
        const unsigned int i = atomicAdd(a.pcount, 1u);
        if (i < a.pcap) {
            fse_particle q;
            memset(&q, 0, sizeof q);
            q.tile.mat = m;
            q.tile.moved = (a.p.flg[g] & F_MOVED) ? 1 : 0;
            q.tile.settle = a.p.stl[g];
            q.tile.temp = a.p.tmp[g];
            q.tile.fluid = a.p.fl[g];
            q.tile.fluid_diff = a.p.fd[g];
            uint32_t col = a.p.col[g];
            if (inner) col = ((((col >> 16) & 0xff) / 4) << 16) | ((((col >> 8) & 0xff) / 4) << 8) | ((col & 0xff) / 4);
            q.tile.color = col;
            q.x = (float)x;
            q.y = (float)(inner ? y + 1 : y);
            q.vx = dx / 10.0f + ((int)(rng_draw(cb, S_EXPL_VX) % 10) - 5) / 10.0f;
            q.vy = dy / 6.0f + ((int)(rng_draw(cb, S_EXPL_VY) % 10) - 5) / 10.0f;
            q.ay = 0.1f;
            q.fade_time = 60;
            q.id = (3ULL << 62) | ((uint64_t)(a.tick & 0x3fffff) << 40) | ((uint64_t)(y & 0xfffff) << 20) | (uint64_t)(x & 0xfffff);
            a.pbuf[i] = q;
        }
    }
    // setTile(x, y, Tiles_NOTHING)
    a.p.mat[g] = (uint8_t)a.T->air;
    a.p.flg[g] = F_DIRTY;
    a.p.stl[g] = 0;
    a.p.tmp[g] = 0;
    a.p.col[g] = 0;
    a.p.fl[g] = 2.0f;
    a.p.fd[g] = 0.0f;
}


// ---- tools ---------------------------------------------------------------------------------------------------------------------
enum : uint32_t { S_HAMMER_JX = 81, S_HAMMER_JY = 82, S_VAC_CLIP = 83, S_VAC_VX = 84, S_VAC_VY = 85 };

__device__ __forceinline__ void tool_set_nothing(Planes p, size_t g, int air) {  // setTile(x, y, Tiles_NOTHING): dirty
    p.mat[g] = (uint8_t)air;
    p.flg[g] = F_DIRTY;
    p.stl[g] = 0;
    p.tmp[g] = 0;
    p.col[g] = 0;
    p.fl[g] = 2.0f;
    p.fd[g] = 0.0f;
}

// erase brush: thread (point, brush cell); a cell under several stamps is cleared by each of them with the same result
__global__ void tool_brush_kernel(Planes p, const DevTables* T, int W, int H, int y_off, const long long* pts, int n_pts, int brush) {
    const int lo = -brush / 2, hi = (int)ceil(brush / 2.0), side = hi - lo;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (side <= 0 || i >= (long long)n_pts * side * side) return;
    const int c = (int)(i % (side * side)), pi = (int)(i / (side * side));
    const int xx = lo + c / side, yy = lo + c % side;
    if (abs(xx) + abs(yy) == brush) return;
    const int x = (int)(pts[pi] % W) + xx, y = (int)(pts[pi] / W) + yy - y_off;  // points are global cells, planes hold rows from y_off
    if (x < 0 || y < 0 || x >= W || y >= H) return;
    const size_t g = (size_t)y * W + x;
    if (T->phys[p.mat[g]] != P_AIR) tool_set_nothing(p, g, T->air);
}

// oy0 / oy1: the local rows whose cells this rank reports (strips: its own rows — ghost cells are cleared as well, but reported by their owner)
__global__ void tool_pickaxe_kernel(Planes p, const DevTables* T, int W, int H, int x, int y, float breakSize, uint32_t* pixels, int* n_out, int oy0, int oy1) {
    const int size = (int)breakSize, span = (int)ceilf(breakSize);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= span * span) return;
    const int xx = i / span, yy = i % span;
    const float cx = (float)((xx / breakSize) - 0.5), cy = (float)((yy / breakSize) - 0.5);
    if (cx * cx + cy * cy > 0.25f) return;
    if (x + xx < 0 || y + yy < 0 || x + xx >= W || y + yy >= H) return;
    const size_t g = (size_t)(y + yy) * W + (x + xx);
    if (T->phys[p.mat[g]] != P_SOLID) return;
    const bool report = y + yy >= oy0 && y + yy < oy1;
    if (report && xx < size && yy < size) pixels[xx + yy * size] = p.col[g];
    tool_set_nothing(p, g, T->air);
    if (report) atomicAdd(n_out, 1);
}

// hammer: one warp walks the segments in order, 32 cells of the crack at a time.  A segment's cells are distinct (the host removed
// repeats like the reference's visited list), so they classify independently; what is sequential — "solid seen yet", the break at
// the first non-solid cell more than one step from the segment start — is two ballots.
__global__ void tool_hammer_kernel(Planes p, const DevTables* T, int W, int H, const long long* pts, const int* seg_off, const int* seg_start, int n_seg,
                                   int sand_mat, int* out, int y_off) {
    const int lane = threadIdx.x;
    long long endInd = -1;
    int changed = 0, broke = 0;
    for (int s = 0; s < n_seg && !broke; s++) {
        const int sx = seg_start[2 * s], sy = seg_start[2 * s + 1];
        bool hit = false;
        for (int base = seg_off[s]; base < seg_off[s + 1] && !broke; base += 32) {
            const int j = base + lane;
            const bool valid = j < seg_off[s + 1] && pts[j] >= 0 && pts[j] < (long long)W * H;
            const long long idx = valid ? pts[j] : 0;
            const bool solid = valid && T->phys[p.mat[idx]] == P_SOLID;
            const bool far = valid && (abs((int)(idx % W) - sx) + abs((int)(idx / W) - sy) > 1);
            const unsigned S = __ballot_sync(0xffffffffu, solid), NF = __ballot_sync(0xffffffffu, valid && !solid && far);
            unsigned after = 0xffffffffu;  // lanes at which a solid cell has been seen
            if (!hit) after = S ? ~((2u << (__ffs(S) - 1)) - 1u) : 0u;
            const unsigned B = NF & after;
            const unsigned keep = B ? ((1u << (__ffs(B) - 1)) - 1u) : 0xffffffffu;  // lanes before the break
            const unsigned conv = S & keep;
            if ((conv >> lane) & 1u) {
                const uint32_t col = p.col[idx];  // ME_draw_darken_color(color, 0.5f) (renderer/gpu.cpp:194-201)
                const uint32_t dk = (col & 0xff000000u) | ((uint32_t)(int)(((col >> 16) & 0xff) * 0.5f) << 16) | ((uint32_t)(int)(((col >> 8) & 0xff) * 0.5f) << 8) |
                                    (uint32_t)(int)((col & 0xff) * 0.5f);
                p.mat[idx] = (uint8_t)sand_mat;  // MaterialInstance(&GENERIC_SAND, colour): temperature 0, defaults elsewhere
                p.flg[idx] = F_DIRTY;
                p.stl[idx] = 0;
                p.tmp[idx] = 0;
                p.col[idx] = dk;
                p.fl[idx] = 2.0f;
                p.fd[idx] = 0.0f;
            }
            if (conv) {
                hit = true;
                changed += __popc(conv);
                endInd = __shfl_sync(0xffffffffu, idx, 31 - __clz(conv));
            }
            if (S & keep) hit = true;
            if (B) broke = 1;
        }
    }
    if (lane == 0) {
        out[0] = endInd < 0 ? -1 : (int)(endInd % W);
        out[1] = endInd < 0 ? -1 : (int)(endInd / W) + y_off;  // global row
        out[2] = changed;
        out[3] = broke;
    }
}

struct VacArgs {
    Planes p;
    const DevTables* T;
    int W, H;
    int wcx, wcy, wmx, wmy;
    uint32_t rkey, tick;
    fse_particle* pbuf;
    unsigned int* pcount;
    unsigned int pcap;
    unsigned int n_before;  // particles in the pool before the call
    int* out;               // x, y, cells sucked, particles re-energised
    // strips: global rows (W x H is the whole world, the plane pointers are moved back by the rows above this rank's window); the walk
    // was done by the host from the types along the line (preset: out[0..1] already hold the hit); a rank empties the cells of the
    // square in its window [ylo, yhi) and makes the particles of the rows it owns [own_lo, own_hi)
    int preset, ylo, yhi, own_lo, own_hi;
};
// strips: physics type + 1 of the line cells this rank owns (0 elsewhere; the sum over the ranks is the whole line)
__global__ void vacuum_line_types_kernel(Planes p, const DevTables* T, int W, const long long* cells, int n, int own_lo, int own_hi, unsigned int* types) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long c = cells[i];
    unsigned int t = 0;
    if (c >= 0) {
        const int y = (int)(c / W);
        if (y >= own_lo && y < own_hi) t = 1u + (unsigned int)T->phys[p.mat[c]];
    }
    types[i] = t;
}
__device__ __forceinline__ void vac_energise(fse_particle& q, uint32_t cb) {  // game.cpp:2492-2505 / 2548-2560
    q.vx = ((int)(rng_draw(cb, S_VAC_VX) % 10) - 5) / 5.0f * 1.0f;
    q.vy = ((int)(rng_draw(cb, S_VAC_VY) % 10) - 5) / 5.0f * 1.0f;
    q.ax = -q.vx / 10.0f;
    q.ay = -q.vy / 10.0f;
    if (q.ay == 0 && q.ax == 0) q.ay = 0.01f;
    q.lifetime = 6;
    q.phase = 1;
    q.vacuum = 1;
}
// thread 0 walks forLine from the screen centre towards the mouse until something suckable is hit (game.cpp:2468-2486; the walk
// stops at the first hit, so it is sequential); then one thread per cell of the 11 x 11 square (2525-2541)
__global__ void __launch_bounds__(128) tool_vacuum_kernel(VacArgs a) {
    __shared__ int s_x, s_y;
    if (threadIdx.x == 0 && a.preset) {
        s_x = a.out[0];
        s_y = a.out[1];
    } else if (threadIdx.x == 0) {
        const int dx = a.wmx - a.wcx, dy = a.wmy - a.wcy;
        int dLong = abs(dx), dShort = abs(dy);
        long long offsetLong = dx > 0 ? 1 : -1, offsetShort = dy > 0 ? a.W : -a.W;
        if (dLong < dShort) {
            const int t = dShort; dShort = dLong; dLong = t;
            const long long o = offsetShort; offsetShort = offsetLong; offsetLong = o;
        }
        int error = dLong / 2;
        long long index = (long long)a.wcy * a.W + a.wcx, sind = -1;
        bool inObject = true;
        for (int i = 0; i <= dLong; ++i) {
            if (index >= 0 && index < (long long)a.W * a.H) {
                const int t = a.T->phys[a.p.mat[index]];
                bool stop = false;
                if (t == P_PASSABLE) {  // OBJECT: the player's own stamp is skipped until the walk has left it
                    if (!inObject) stop = true;
                } else {
                    inObject = false;
                }
                if (t == P_SOLID || t == P_SAND || t == P_SOUP) stop = true;
                if (stop) {
                    sind = index;
                    break;
                }
            }
            const int big = error >= dLong;
            index += big ? offsetLong + offsetShort : offsetLong;
            error += big ? dShort - dLong : dShort;
        }
        s_x = sind == -1 ? a.wmx : (int)(sind % a.W);
        s_y = sind == -1 ? a.wmy : (int)(sind / a.W);
        a.out[0] = s_x;
        a.out[1] = s_y;
    }
    __syncthreads();
    const int rad = 5, c = threadIdx.x;
    if (c >= 121) return;
    const int xx = c / 11 - rad, yy = c % 11 - rad;
    int clipRadSq = rad * rad;
    clipRadSq += (int)(rng_draw(rng_cell(a.rkey, 0, 0), S_VAC_CLIP) % (uint32_t)clipRadSq) / 4;
    if (xx * xx + yy * yy > clipRadSq) return;
    if ((yy == -rad || yy == rad) && (xx == -rad || xx == rad)) return;
    const int x = s_x + xx, y = s_y + yy;
    if (x < 0 || y < 0 || x >= a.W || y >= a.H) return;
    if (y < a.ylo || y >= a.yhi) return;
    const size_t g = (size_t)y * a.W + x;
    const int t = a.T->phys[a.p.mat[g]];
    if (!(t == P_SOLID || t == P_SAND || t == P_SOUP)) return;
    if (!(y >= a.own_lo && y < a.own_hi)) {  // a ghost cell: emptied here as well, its owner makes the particle and counts it
        tool_set_nothing(a.p, g, a.T->air);
        return;
    }
    const unsigned int i = atomicAdd(a.pcount, 1u);
    if (i < a.pcap) {
        fse_particle q;
        memset(&q, 0, sizeof q);
        q.tile.mat = a.p.mat[g];
        q.tile.moved = (a.p.flg[g] & F_MOVED) ? 1 : 0;
        q.tile.settle = a.p.stl[g];
        q.tile.color = a.p.col[g];
        q.tile.temp = a.p.tmp[g];
        q.tile.fluid = a.p.fl[g];
        q.tile.fluid_diff = a.p.fd[g];
        q.x = (float)x;
        q.y = (float)y;
        q.fade_time = 60;
        vac_energise(q, rng_cell(a.rkey, x, y));
        q.id = (3ULL << 62) | (1ULL << 61) | ((uint64_t)(a.tick & 0x1fffff) << 40) | ((uint64_t)(y & 0xfffff) << 20) | (uint64_t)(x & 0xfffff);
        a.pbuf[i] = q;
    }
    tool_set_nothing(a.p, g, a.T->air);
    atomicAdd(&a.out[2], 1);
}
// loose particles already flying through the square are caught as well (game.cpp:2543-2583, with its `x == rad` corner test)
__global__ void tool_vacuum_pool_kernel(VacArgs a) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_before) return;
    fse_particle q = a.pbuf[i];
    if (!(q.target_force == 0 && !q.phase)) return;
    const int x = a.out[0], y = a.out[1], rad = 5;
    const int xx = (int)q.x - x, yy = (int)q.y - y;
    if (xx < -rad || xx > rad || yy < -rad || yy > rad) return;
    if ((yy == -rad || yy == rad) && (xx == -rad || x == rad)) return;
    vac_energise(q, rng_cell(a.rkey, (int)(q.id & 0xffffffffu), (int)(q.id >> 32)));
    a.pbuf[i] = q;
    atomicAdd(&a.out[3], 1);
}
__global__ void vacuum_pull_kernel(fse_particle* pbuf, unsigned int n, float tx, float ty, int* n_out) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fse_particle q = pbuf[i];
    if (!q.vacuum) return;
    if (q.lifetime <= 0) {
        q.target_force = 0.45f;
        q.target_x = tx;
        q.target_y = ty;
        q.ax = 0;
        q.ay = 0.01f;
    }
    const float tdx = q.target_x - q.x, tdy = q.target_y - q.y;
    if (tdx * tdx + tdy * tdy < 10 * 10) {
        q.temporary = 1;
        q.lifetime = 0;
        q.vacuum = 0;
        atomicAdd(n_out, 1);
    }
    pbuf[i] = q;
}

// ---- host: world::forLine / forLineCornered (world.cpp:3250-3313) as cell lists ----
static void line_cells(int width, int x0, int y0, int x1, int y1, std::vector<long long>& out) {
    const int dx = x1 - x0, dy = y1 - y0;
    int dLong = std::abs(dx), dShort = std::abs(dy);
    long long stepLong = dx > 0 ? 1 : -1, stepShort = dy > 0 ? width : -width;
    if (dLong < dShort) {
        std::swap(dShort, dLong);
        std::swap(stepShort, stepLong);
    }
    int error = dLong / 2;
    long long index = (long long)y0 * width + x0;
    for (int i = 0; i <= dLong; ++i) {
        out.push_back(index);
        const bool big = error >= dLong;
        index += big ? stepLong + stepShort : stepLong;
        error += big ? dShort - dLong : dShort;
    }
}
static void cornered_cells(int width, int x0, int y0, int x1, int y1, std::vector<long long>& out) {
    const size_t first = out.size();
    const float sx = (float)x0, sy = (float)y0, ex = (float)x1, ey = (float)y1;
    float x = std::floor(sx), y = std::floor(sy);
    const float diffX = ex - sx, diffY = ey - sy;
    const float stepX = (diffX > 0) ? 1.0f : ((diffX < 0) ? -1.0f : 0.0f), stepY = (diffY > 0) ? 1.0f : ((diffY < 0) ? -1.0f : 0.0f);
    const float xOffset = ex > sx ? (std::ceil(sx) - sx) : (sx - std::floor(sx));
    const float yOffset = ey > sy ? (std::ceil(sy) - sy) : (sy - std::floor(sy));
    const float angle = (float)std::atan2(-diffY, diffX);
    float tMaxX = (float)(xOffset / std::cos(angle)), tMaxY = (float)(yOffset / std::sin(angle));
    const float tDeltaX = (float)(1.0 / std::cos(angle)), tDeltaY = (float)(1.0 / std::sin(angle));
    const float manhattan = std::abs(std::floor(ex) - std::floor(sx)) + std::abs(std::floor(ey) - std::floor(sy));
    for (int t = 0; t <= manhattan; ++t) {
        const long long idx = (long long)(x + y * width);
        bool seen = false;
        for (size_t q = first; q < out.size(); q++) seen |= out[q] == idx;
        if (!seen) out.push_back(idx);
        if (std::abs(tMaxX) < std::abs(tMaxY) || std::isnan(tMaxY)) {
            tMaxX += tDeltaX;
            x += stepX;
        } else {
            tMaxY += tDeltaY;
            y += stepY;
        }
    }
}

}  // namespace fse

using namespace fse;

extern "C" FSE_API int fse_explosion(fse_world* w, int32_t cx, int32_t cy, int32_t radius, uint32_t tick, uint32_t seed) {
    if (!w) return fail(FSE_EINVAL, "fse_explosion: null world");
    if (radius <= 0 || radius > 4096) return fail(FSE_EINVAL, "fse_explosion: radius %d out of range (1..4096)", radius);
    cudaError_t e = cudaSetDevice(w->ctx->device);
    if (e != cudaSuccess) return fail(FSE_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    {   // every cell of the blast square may become a loose particle: make room first (the kernels drop what has no slot)
        const long long side_ = 4LL * radius + 2, cells_ = (long long)w->W * w->H;
        if (int r = particles_headroom(w, side_ * side_ < cells_ ? side_ * side_ : cells_, true)) return r;
    }
    ExplArgs a;
    a.p = w->p;
    a.T = w->ctx->d_tabs;
    a.W = w->W;
    a.H = w->H;
    a.cx = cx;
    a.cy = cy;
    a.y_off = w->y_off;
    // Multi-rank strips: every rank makes the call with the same arguments.  Cells decide independently from their own state and the
    // position-keyed RNG, so each rank clears the part of the blast it holds; the particles come from the rank that owns the row.
    a.own_lo = w->strip ? w->own_lo : 0;
    a.own_hi = w->strip ? w->own_hi : w->H;
    a.radius = radius;
    a.rkey = rng_key(seed, tick, 7u);  // the tick's own iterations use 0..cell_iter-1
    a.tick = tick;
    a.pbuf = w->pbuf;
    a.pcount = w->pcount;
    a.pcap = w->pcap;
    const int side = 4 * radius;
    dim3 grid((side + 127) / 128, side);
    explosion_kernel<<<grid, 128, 0, w->stream>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(FSE_ECUDA, "explosion_kernel: %s", cudaGetErrorString(e));
    w->ctx->launches += 1;
    const int cyl = cy - w->y_off;
    const int x0 = cx - 2 * radius < 0 ? 0 : cx - 2 * radius, y0 = cyl - 2 * radius < 0 ? 0 : cyl - 2 * radius;
    const int x1 = cx + 2 * radius > w->W ? w->W : cx + 2 * radius, y1 = cyl + 2 * radius > w->H ? w->H : cyl + 2 * radius;
    if (x1 > x0 && y1 > y0)
        if (int r = fse_wake_rect(w, x0, y0, x1 - x0, y1 - y0)) return r;
    if (w->strip && w->ctx->nranks > 1) return strip_refresh(w, w->stream, 32);  // ghost rows are the owner's again
    return FSE_OK;
}

#define CKT(call)                                                                                 \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return fail(FSE_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

static int tool_scratch(fse_world* w, size_t bytes) {  // device scratch of the tools (cell lists, results)
    if (w->tool_scratch_bytes >= bytes) return FSE_OK;
    CKT(cudaStreamSynchronize(w->stream));
    cudaFree(w->tool_scratch);
    w->tool_scratch = nullptr;
    w->tool_scratch_bytes = 0;
    CKT(cudaMalloc(&w->tool_scratch, bytes + 4096));
    w->tool_scratch_bytes = bytes + 4096;
    return FSE_OK;
}
static int tool_common(fse_world* w, const char* who, bool strips_ok = false) {
    if (!w) return fail(FSE_EINVAL, "%s: null world", who);
    if (w->strip && w->ctx->nranks > 1 && !strips_ok) return fail(FSE_ESTATE, "%s: not available on multi-rank strips", who);
    CKT(cudaSetDevice(w->ctx->device));
    return FSE_OK;
}
static int tool_wake(fse_world* w, int x0, int y0, int x1, int y1) {
    if (x0 > x1) std::swap(x0, x1);
    if (y0 > y1) std::swap(y0, y1);
    x0 = std::max(0, x0); y0 = std::max(0, y0); x1 = std::min(w->W - 1, x1); y1 = std::min(w->H - 1, y1);
    if (x1 < x0 || y1 < y0) return FSE_OK;
    return fse_wake_rect(w, x0, y0, x1 - x0 + 1, y1 - y0 + 1);
}

extern "C" FSE_API int fse_tool_erase_line(fse_world* w, int32_t x0, int32_t y0, int32_t x1, int32_t y1, int32_t brush_size) {
    // multi-rank strips: every rank makes the call; a rank clears the cells of the stroke it holds (the cells decide independently)
    if (int r = tool_common(w, "fse_tool_erase_line", true)) return r;
    if (brush_size < 1 || brush_size > 256) return fail(FSE_EINVAL, "fse_tool_erase_line: brush size %d (1..256)", brush_size);
    const int Hg = w->strip ? w->Hglobal : w->H;
    if (x0 < 0 || y0 < 0 || x1 < 0 || y1 < 0 || x0 >= w->W || x1 >= w->W || y0 >= Hg || y1 >= Hg)
        return fail(FSE_EINVAL, "fse_tool_erase_line: end points outside the world");
    std::vector<long long> pts;
    line_cells(w->W, x0, y0, x1, y1, pts);
    if (int r = tool_scratch(w, pts.size() * sizeof(long long))) return r;
    CKT(cudaMemcpyAsync(w->tool_scratch, pts.data(), pts.size() * sizeof(long long), cudaMemcpyHostToDevice, w->stream));
    const int lo = -brush_size / 2, hi = (int)std::ceil(brush_size / 2.0), side = hi - lo;
    const long long total = (long long)pts.size() * side * side;
    tool_brush_kernel<<<(unsigned)((total + 255) / 256), 256, 0, w->stream>>>(w->p, w->ctx->d_tabs, w->W, w->H, w->y_off, (const long long*)w->tool_scratch, (int)pts.size(), brush_size);
    CKT(cudaGetLastError());
    w->ctx->launches += 1;
    CKT(cudaStreamSynchronize(w->stream));  // pts is a local
    if (int r = tool_wake(w, std::min(x0, x1) - brush_size, std::min(y0, y1) - brush_size - w->y_off, std::max(x0, x1) + brush_size,
                          std::max(y0, y1) + brush_size - w->y_off)) return r;
    if (w->strip && w->ctx->nranks > 1) return strip_refresh(w, w->stream, 32);
    return FSE_OK;
}

extern "C" FSE_API int fse_tool_pickaxe(fse_world* w, int32_t x, int32_t y, float break_size, uint32_t* pixels_out, int32_t* n_out) {
    // multi-rank strips: every rank makes the call; the cells decide independently, so a rank clears the cells of the circle it holds
    // and reports the ones it owns; the reports are summed over the ranks (every pixel has one owner)
    if (int r = tool_common(w, "fse_tool_pickaxe", true)) return r;
    if (!(break_size >= 1.0f) || break_size > 1024.0f || !pixels_out || !n_out) return fail(FSE_EINVAL, "fse_tool_pickaxe: bad argument");
    const bool multi = w->strip && w->ctx->nranks > 1;
    const int size = (int)break_size, span = (int)std::ceil(break_size);
    const size_t pix_bytes = sizeof(uint32_t) * (size_t)size * size, pix_al = (pix_bytes + 15) & ~(size_t)15;
    if (int r = tool_scratch(w, pix_al + 16)) return r;
    uint32_t* d_pix = (uint32_t*)w->tool_scratch;
    int* d_n = (int*)((char*)w->tool_scratch + pix_al);
    CKT(cudaMemsetAsync(w->tool_scratch, 0, pix_al + 16, w->stream));
    if (multi)
        if (int r = strip_refresh(w, w->stream, STRIP_GHOST)) return r;
    tool_pickaxe_kernel<<<(span * span + 127) / 128, 128, 0, w->stream>>>(w->p, w->ctx->d_tabs, w->W, w->H, x, y - w->y_off, break_size, d_pix, d_n,
                                                                           multi ? w->own_lo - w->y_off : 0, multi ? w->own_hi - w->y_off : w->H);
    CKT(cudaGetLastError());
    w->ctx->launches += 1;
    if (multi)
        if (int r = strip_allreduce_u32(w, (unsigned int*)w->tool_scratch, pix_al / 4 + 1, w->stream)) return r;
    CKT(cudaMemcpyAsync(pixels_out, d_pix, pix_bytes, cudaMemcpyDeviceToHost, w->stream));
    CKT(cudaMemcpyAsync(n_out, d_n, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
    CKT(cudaStreamSynchronize(w->stream));
    return tool_wake(w, x - 1, y - 1 - w->y_off, x + span + 1, y + span + 1 - w->y_off);
}

extern "C" FSE_API int fse_tool_hammer(fse_world* w, int32_t hammer_x, int32_t hammer_y, int32_t x, int32_t y, int32_t sand_mat, uint32_t tick, uint32_t seed,
                                       fse_hammer_result* out) {
    // multi-rank strips: the crack is walked in order, so ONE rank runs it — the one that holds the middle row of its box — and the
    // box travels to the neighbours it reaches into afterwards; every rank gets the same result
    if (int r = tool_common(w, "fse_tool_hammer", true)) return r;
    if (!out || sand_mat < 0 || sand_mat >= w->ctx->h_tabs.n) return fail(FSE_EINVAL, "fse_tool_hammer: bad argument");
    const int dx = hammer_x - x, dy = hammer_y - y;
    if (std::abs(dx) > 4096 || std::abs(dy) > 4096) return fail(FSE_EINVAL, "fse_tool_hammer: crack longer than 4096 cells");
    const uint32_t rkey = rng_key(seed, tick, 9u);
    const float len = std::sqrt((float)(dx * dx + dy * dy));
    const int nSegments = (int)(1 + len / 10);  // game.cpp:856
    std::vector<long long> pts;
    std::vector<int> seg_off(1, 0), seg_start;
    int px = hammer_x, py = hammer_y;
    for (int i = 0; i < nSegments; i++) {  // 858-864: segment ends with a +-1 jitter
        int sx = hammer_x + (int)((float)(dx / nSegments) * (i + 1));
        int sy = hammer_y + (int)((float)(dy / nSegments) * (i + 1));
        const uint32_t cb = rng_cell(rkey, i, 0);
        sx += (int)(rng_draw(cb, S_HAMMER_JX) % 3) - 1;
        sy += (int)(rng_draw(cb, S_HAMMER_JY) % 3) - 1;
        seg_start.push_back(px);
        seg_start.push_back(py);
        cornered_cells(w->W, px, py, sx, sy, pts);
        seg_off.push_back((int)pts.size());
        px = sx;
        py = sy;
    }
    const bool multi = w->strip && w->ctx->nranks > 1;
    // the crack runs from the hammer AWAY from (x, y): hammer .. hammer + (dx, dy), segment ends jittered by one cell
    const int bx0 = std::min(hammer_x, hammer_x + dx) - 4, by0 = std::min(hammer_y, hammer_y + dy) - 4, bx1 = std::max(hammer_x, hammer_x + dx) + 4,
              by1 = std::max(hammer_y, hammer_y + dy) + 4;
    int runner = w->ctx->rank;
    if (multi) {
        if (int r = strip_runner_of_rows(w, by0, by1, "fse_tool_hammer", &runner)) return r;
        if (int r = strip_refresh(w, w->stream, STRIP_GHOST)) return r;
        const long long shift = (long long)w->y_off * w->W;  // cells and segment starts in the runner's local rows
        for (long long& c : pts) c = c >= 0 ? c - shift : c;
        for (size_t i = 1; i < seg_start.size(); i += 2) seg_start[i] -= w->y_off;
    }
    const size_t b0 = pts.size() * sizeof(long long), b1 = seg_off.size() * sizeof(int), b2 = seg_start.size() * sizeof(int);
    const size_t o1 = (b0 + 15) & ~(size_t)15, o2 = o1 + ((b1 + 15) & ~(size_t)15), o3 = o2 + ((b2 + 15) & ~(size_t)15);
    if (int r = tool_scratch(w, o3 + 16)) return r;
    char* base = (char*)w->tool_scratch;
    CKT(cudaMemsetAsync(base + o3, 0, 16, w->stream));
    if (runner == w->ctx->rank) {
        if (b0) CKT(cudaMemcpyAsync(base, pts.data(), b0, cudaMemcpyHostToDevice, w->stream));
        CKT(cudaMemcpyAsync(base + o1, seg_off.data(), b1, cudaMemcpyHostToDevice, w->stream));
        CKT(cudaMemcpyAsync(base + o2, seg_start.data(), b2, cudaMemcpyHostToDevice, w->stream));
        tool_hammer_kernel<<<1, 32, 0, w->stream>>>(w->p, w->ctx->d_tabs, w->W, w->H, (const long long*)base, (const int*)(base + o1), (const int*)(base + o2), nSegments,
                                                   sand_mat, (int*)(base + o3), w->y_off);
        CKT(cudaGetLastError());
        w->ctx->launches += 1;
    }
    if (multi) {
        std::vector<int4> rect[4];
        strip_rects_of_box(w, runner, bx0, by0, bx1, by1, rect);
        if (int r = strip_push_rects(w, rect, w->stream)) return r;
        if (int r = strip_allreduce_u32(w, (unsigned int*)(base + o3), 4, w->stream)) return r;  // the others contribute zeros
    }
    int res[4];
    CKT(cudaMemcpyAsync(res, base + o3, sizeof res, cudaMemcpyDeviceToHost, w->stream));
    CKT(cudaStreamSynchronize(w->stream));
    out->end_x = res[0];
    out->end_y = res[1];
    out->n_changed = res[2];
    out->broke = res[3];
    return tool_wake(w, bx0, by0 - w->y_off, bx1, by1 - w->y_off);
}

extern "C" FSE_API int fse_tool_vacuum(fse_world* w, int32_t wcx, int32_t wcy, int32_t wmx, int32_t wmy, uint32_t tick, uint32_t seed, fse_vacuum_result* out) {
    // multi-rank strips: every rank makes the call.  The walk from the screen centre towards the mouse only reads: every rank reports the
    // physics types of the line cells it owns, the sum over the ranks is the whole line, and the host of every rank finds the same hit
    // on it; the cells of the 11 x 11 square are independent (a rank empties the ones in its window and makes the particles of the
    // rows it owns), the loose particles are caught by the rank whose pool they are in; the counts are summed.
    if (int r = tool_common(w, "fse_tool_vacuum", true)) return r;
    if (!out) return fail(FSE_EINVAL, "fse_tool_vacuum: null result");
    out->x = out->y = -1;
    out->n_sucked = out->n_caught = 0;
    const bool multi = w->strip && w->ctx->nranks > 1;
    const int Hg = w->strip ? w->Hglobal : w->H;
    const int mdx = wmx - wcx, mdy = wmy - wcy;
    if (mdx * mdx + mdy * mdy > 256 * 256) return FSE_OK;  // game.cpp:2465: out of reach
    if (wcx < 0 || wcy < 0 || wcx >= w->W || wcy >= Hg) return fail(FSE_EINVAL, "fse_tool_vacuum: centre outside the world");
    int64_t n_before = 0;
    if (int r = particles_headroom(w, 121, true)) return r;
    if (int r = fse_particles_count(w, &n_before)) return r;
    VacArgs a;
    memset(&a, 0, sizeof a);
    a.p = w->p; a.T = w->ctx->d_tabs; a.W = w->W; a.H = w->H;
    a.wcx = wcx; a.wcy = wcy; a.wmx = wmx; a.wmy = wmy;
    a.rkey = rng_key(seed, tick, 10u); a.tick = tick;
    a.pbuf = w->pbuf; a.pcount = w->pcount; a.pcap = w->pcap;
    a.n_before = (unsigned int)n_before;
    a.ylo = 0; a.yhi = w->H; a.own_lo = 0; a.own_hi = w->H;
    if (!multi) {
        if (int r = tool_scratch(w, 64)) return r;
        CKT(cudaMemsetAsync(w->tool_scratch, 0, 16, w->stream));
        a.out = (int*)w->tool_scratch;
    } else {
        if (int r = strip_refresh(w, w->stream, STRIP_GHOST)) return r;
        const size_t back = (size_t)w->y_off * w->W;  // global rows from here on
        a.p.mat -= back; a.p.flg -= back; a.p.stl -= back; a.p.tmp -= back; a.p.col -= back; a.p.fl -= back; a.p.fd -= back;
        a.H = Hg;
        a.ylo = w->y_off; a.yhi = w->y_off + w->H; a.own_lo = w->own_lo; a.own_hi = w->own_hi;
        // the cells of the walk (game.cpp:2468-2486: forLine on linear cell indices), then their types from the ranks that own them
        int dLong = std::abs(mdx), dShort = std::abs(mdy);
        long long offsetLong = mdx > 0 ? 1 : -1, offsetShort = mdy > 0 ? w->W : -w->W;
        if (dLong < dShort) {
            std::swap(dLong, dShort);
            std::swap(offsetLong, offsetShort);
        }
        std::vector<long long> cells((size_t)dLong + 1);
        int error = dLong / 2;
        long long index = (long long)wcy * w->W + wcx;
        for (int i = 0; i <= dLong; ++i) {
            cells[i] = (index >= 0 && index < (long long)w->W * Hg) ? index : -1;
            const int big = error >= dLong;
            index += big ? offsetLong + offsetShort : offsetLong;
            error += big ? dShort - dLong : dShort;
        }
        const size_t nb = cells.size() * sizeof(long long), o_types = (nb + 15) & ~(size_t)15, o_out = o_types + ((cells.size() * 4 + 15) & ~(size_t)15);
        if (int r = tool_scratch(w, o_out + 64)) return r;
        char* base = (char*)w->tool_scratch;
        unsigned int* d_types = (unsigned int*)(base + o_types);
        a.out = (int*)(base + o_out);
        CKT(cudaMemcpyAsync(base, cells.data(), nb, cudaMemcpyHostToDevice, w->stream));
        vacuum_line_types_kernel<<<(unsigned)((cells.size() + 127) / 128), 128, 0, w->stream>>>(a.p, a.T, w->W, (const long long*)base, (int)cells.size(), w->own_lo, w->own_hi, d_types);
        CKT(cudaGetLastError());
        w->ctx->launches += 1;
        if (int r = strip_allreduce_u32(w, d_types, cells.size(), w->stream)) return r;
        std::vector<unsigned int> types(cells.size());
        CKT(cudaMemcpyAsync(types.data(), d_types, cells.size() * 4, cudaMemcpyDeviceToHost, w->stream));
        CKT(cudaStreamSynchronize(w->stream));
        long long sind = -1;
        bool inObject = true;
        for (size_t i = 0; i < cells.size(); i++) {
            if (cells[i] < 0) continue;
            const int t = (int)types[i] - 1;
            bool stop = false;
            if (t == P_PASSABLE) {  // OBJECT: the player's own stamp is skipped until the walk has left it
                if (!inObject) stop = true;
            } else {
                inObject = false;
            }
            if (t == P_SOLID || t == P_SAND || t == P_SOUP) stop = true;
            if (stop) {
                sind = cells[i];
                break;
            }
        }
        int hit[4] = {sind == -1 ? wmx : (int)(sind % w->W), sind == -1 ? wmy : (int)(sind / w->W), 0, 0};
        CKT(cudaMemcpyAsync(a.out, hit, sizeof hit, cudaMemcpyHostToDevice, w->stream));
        CKT(cudaStreamSynchronize(w->stream));  // hit is a local
        a.preset = 1;
    }
    tool_vacuum_kernel<<<1, 128, 0, w->stream>>>(a);
    CKT(cudaGetLastError());
    w->ctx->launches += 1;
    if (n_before) {
        tool_vacuum_pool_kernel<<<(unsigned)((n_before + 127) / 128), 128, 0, w->stream>>>(a);
        CKT(cudaGetLastError());
        w->ctx->launches += 1;
    }
    if (multi)
        if (int r = strip_allreduce_u32(w, (unsigned int*)a.out + 2, 2, w->stream)) return r;  // cells sucked, particles caught
    int res[4];
    CKT(cudaMemcpyAsync(res, a.out, sizeof res, cudaMemcpyDeviceToHost, w->stream));
    CKT(cudaStreamSynchronize(w->stream));
    out->x = res[0]; out->y = res[1]; out->n_sucked = res[2]; out->n_caught = res[3];
    return tool_wake(w, res[0] - 7, res[1] - 7 - w->y_off, res[0] + 7, res[1] + 7 - w->y_off);
}

extern "C" FSE_API int fse_particles_vacuum_pull(fse_world* w, float target_x, float target_y, int32_t* n_collected) {
    // multi-rank strips: every rank pulls the particles of its own pool, the collected ones are summed
    if (int r = tool_common(w, "fse_particles_vacuum_pull", true)) return r;
    const bool multi = w->strip && w->ctx->nranks > 1;
    int64_t n = 0;
    if (int r = fse_particles_count(w, &n)) return r;
    if (n_collected) *n_collected = 0;
    if (n == 0 && !multi) return FSE_OK;
    if (int r = tool_scratch(w, 64)) return r;
    CKT(cudaMemsetAsync(w->tool_scratch, 0, 16, w->stream));
    if (n) {
        vacuum_pull_kernel<<<(unsigned)((n + 127) / 128), 128, 0, w->stream>>>(w->pbuf, (unsigned int)n, target_x, target_y, (int*)w->tool_scratch);
        CKT(cudaGetLastError());
        w->ctx->launches += 1;
    }
    if (multi)  // every rank takes part, also one whose pool is empty
        if (int r = strip_allreduce_u32(w, (unsigned int*)w->tool_scratch, 1, w->stream)) return r;
    if (n_collected) {
        CKT(cudaMemcpyAsync(n_collected, w->tool_scratch, sizeof(int), cudaMemcpyDeviceToHost, w->stream));
        CKT(cudaStreamSynchronize(w->stream));
    }
    return FSE_OK;
}
