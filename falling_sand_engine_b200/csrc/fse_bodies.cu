// fse_bodies.cu — rigid-body <-> grid bridge: the raster and erase loops of game::tick
// (reference: source/engine/game.cpp:1711-1815 and 1896-1983) as kernels.  Box2D itself stays on the host.
//
// The reference visits bodies in order and, inside a body, pixels tx-major; each pixel tries the five offsets
// (0,0),(1,0),(-1,0),(0,1),(0,-1) against the grid as the pixels before it left it.  That order is kept EXACTLY, in
// parallel: every pixel has a rank (its position in the reference's visiting order) and a footprint (its five candidate
// cells).  In a round, every pending pixel publishes its rank on its footprint cells with atomicMin; a pixel acts only if
// it holds the minimum on all five, i.e. when no earlier pixel that could still change one of its cells is pending.  Pixels
// that act in the same round have disjoint footprints, so the outcome equals the sequential loop.
//
// One launch per call: a CTA takes the next body (atomic ticket, so bodies start in index order), waits for the lower-index
// bodies whose footprint box overlaps its own to finish — bodies that do not overlap commute, overlapping ones keep the
// reference's body order — and then runs the rounds of its own pixels between block barriers, with the claim map of its
// footprint box in shared memory (boxes up to 64x64 cells; larger bodies use a claim plane in global memory).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fse_internal.hpp"

namespace fse {

struct BodyArgs {
    Planes p;
    const DevTables* T;
    int W, H;
    int n_bodies;
    const int* off;     // [n+1] first pixel rank of each body
    const int* bw;
    const int* bh;
    fse_cell* tiles;    // all bodies, body b at off[b], index tx + ty*w
    const float4* xf;   // x, y, sin, cos
    uint8_t* pending;   // per pixel rank
    uint32_t* claim;    // W*H, 0xffffffff = free
    unsigned int* counters;  // [4] pending count
    int4* feedback;
    fse_particle* pbuf;
    unsigned int* pcount;
    unsigned int pcap;
    uint32_t rkey;
    uint32_t tick;
    int air;
    uint8_t* awake;
    int acols, arows;
    int n_pixels;
    int4* aabb;              // per body: footprint box (x0, y0, x1, y1), inclusive
    unsigned int* done;      // per body: 1 once all its pixels have acted
    unsigned int* ticket;    // [0] next body, [1] most rounds any body needed
    // strip worlds: the planes hold global rows [y_off, y_off + H); transforms stay global (the float -> int truncation must see the same
    // sums as on one world) and the row offset is taken off the integer result.  exec (null on one world): the rank that runs body b
    int y_off;
    const int* exec;
    int rank;
};

__device__ __forceinline__ void world_pos(const BodyArgs& a, int b, int tx, int ty, int& wx, int& wy) {
    const float4 t = a.xf[b];  // x, y, s, c
    wx = (int)(tx * t.w - (ty + 1) * t.z + t.x);  // game.cpp:1763
    wy = (int)(tx * t.z + (ty + 1) * t.w + t.y) - a.y_off;  // game.cpp:1764
}
__constant__ int c_dirs[5][2] = {{0, 0}, {1, 0}, {-1, 0}, {0, 1}, {0, -1}};  // game.cpp:1766

// footprint box of a body: every candidate cell of every pixel, with one cell of slack for the float -> int truncation
__global__ void bodies_aabb_kernel(BodyArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) a.ticket[0] = a.ticket[1] = a.ticket[2] = a.ticket[3] = a.ticket[4] = a.ticket[5] = 0;
    if (b >= a.n_bodies) return;
    const float4 t = a.xf[b];
    float x0 = 3.0e38f, y0 = 3.0e38f, x1 = -3.0e38f, y1 = -3.0e38f;
    for (int q = 0; q < 4; q++) {
        const float tx = (q & 1) ? (float)(a.bw[b] - 1) : 0.0f, ty1 = (q & 2) ? (float)a.bh[b] : 1.0f;
        const float fx = tx * t.w - ty1 * t.z + t.x, fy = tx * t.z + ty1 * t.w + t.y;
        x0 = fminf(x0, fx); x1 = fmaxf(x1, fx);
        y0 = fminf(y0, fy); y1 = fmaxf(y1, fy);
    }
    const float lim = 1.0e9f;
    x0 = fmaxf(fminf(x0, lim), -lim); x1 = fmaxf(fminf(x1, lim), -lim);
    y0 = fmaxf(fminf(y0, lim), -lim); y1 = fmaxf(fminf(y1, lim), -lim);
    a.aabb[b] = make_int4((int)floorf(x0) - 3, (int)floorf(y0) - 3 - a.y_off, (int)ceilf(x1) + 3, (int)ceilf(y1) + 3 - a.y_off);
    a.done[b] = 0;
}

__device__ __forceinline__ void wake3x3(const BodyArgs& a, int x, int y) {
    if (!a.awake) return;
    const int ci = x / CHUNK, cj = (y + a.y_off) / CHUNK;
    for (int dj = -1; dj <= 1; dj++)
        for (int di = -1; di <= 1; di++) {
            const int ni = ci + di, nj = cj + dj;
            if (ni >= 0 && nj >= 0 && ni < a.acols && nj < a.arows) a.awake[nj * a.acols + ni] = 1;
        }
}

__device__ __forceinline__ void write_cell(const BodyArgs& a, size_t g, const fse_cell& t) {
    a.p.mat[g] = (uint8_t)t.mat;
    a.p.flg[g] = (uint8_t)((t.moved ? F_MOVED : 0) | F_DIRTY);
    a.p.stl[g] = t.settle;
    a.p.tmp[g] = t.temp;
    a.p.col[g] = t.color;
    a.p.fl[g] = t.fluid;
    a.p.fd[g] = t.fluid_diff;
}
__device__ __forceinline__ fse_cell read_cell(const BodyArgs& a, size_t g) {
    fse_cell c;
    const uint8_t f = __ldcg(a.p.flg + g);
    c.mat = __ldcg(a.p.mat + g);
    c.moved = (f & F_MOVED) ? 1 : 0;
    c.settle = __ldcg(a.p.stl + g);
    c.color = __ldcg(a.p.col + g);
    c.temp = __ldcg(a.p.tmp + g);
    c.dirty = 0;
    c._pad = 0;
    c.fluid = __ldcg(a.p.fl + g);
    c.fluid_diff = __ldcg(a.p.fd + g);
    return c;
}

__device__ __forceinline__ unsigned int bodies_ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// one pixel acts (game.cpp:1768-1811 raster, 1918-1965 erase); grid cells are read at L2 (another SM may have written them)
template <bool ERASE>
__device__ __noinline__ void body_pixel_act(const BodyArgs& a, int b, int tx, int ty, int wx, int wy) {
    fse_cell* tile = &a.tiles[a.off[b] + tx + ty * a.bw[b]];
    const fse_cell rm = *tile;
    int4* fb = &a.feedback[b];
    if (!ERASE) {
        for (int d = 0; d < 5; d++) {
            const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
            if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
            const size_t g = (size_t)y * a.W + x;
            const int ph = a.T->phys[__ldcg(a.p.mat + g)];
            if (ph == P_AIR) {
                write_cell(a, g, rm);
                atomicAdd(&fb->z, 1);
                wake3x3(a, x, y);
                break;
            } else if (ph == P_SAND || ph == P_SOUP) {
                const unsigned int i = atomicAdd(a.pcount, 1u);
                if (i < a.pcap) {  // the displaced cell is thrown up as a loose particle (game.cpp:1791 / 1801)
                    fse_particle p;
                    memset(&p, 0, sizeof p);
                    p.tile = read_cell(a, g);
                    p.x = (float)x;
                    p.y = (float)(y + a.y_off - 3);  // particles carry global coordinates
                    const int pix = tx + ty * a.bw[b];
                    const uint32_t cb = rng_cell(a.rkey, b, pix);
                    p.vx = (float)(((int)(rng_draw(cb, S_BRIDGE_VX) % 10) - 5) / 10.0f);
                    p.vy = (float)(-(int)(rng_draw(cb, S_BRIDGE_VY) % 5 + 5) / 10.0f);
                    p.ay = 0.1f;
                    p.fade_time = 60;
                    p.id = (2ULL << 62) | ((uint64_t)(a.tick & 0xffff) << 40) | ((uint64_t)(b & 0xfffff) << 20) | (uint64_t)(pix & 0xfffff);
                    a.pbuf[i] = p;
                }
                write_cell(a, g, rm);
                atomicAdd(ph == P_SAND ? &fb->x : &fb->y, 1);
                atomicAdd(&fb->z, 1);
                wake3x3(a, x, y);
                break;
            }
        }
    } else {
        bool found = false;
        for (int d = 0; d < 5; d++) {
            const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
            if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
            const size_t g = (size_t)y * a.W + x;
            if (__ldcg(a.p.mat + g) == rm.mat) {  // .id == rmat.id (any cell of the same material)
                *tile = read_cell(a, g);
                fse_cell nothing;
                memset(&nothing, 0, sizeof nothing);
                nothing.mat = (uint16_t)a.air;
                nothing.fluid = 2.0f;
                write_cell(a, g, nothing);
                atomicAdd(&fb->z, 1);
                wake3x3(a, x, y);
                found = true;
                break;
            }
        }
        if (!found && wx >= 0 && wy >= 0 && wx < a.W && wy < a.H && __ldcg(a.p.mat + (size_t)wy * a.W + wx) == a.air) {  // 1959-1965
            fse_cell nothing;
            memset(&nothing, 0, sizeof nothing);
            nothing.mat = (uint16_t)a.air;
            nothing.fluid = 2.0f;
            *tile = nothing;
            atomicAdd(&fb->w, 1);
        }
    }
}

// release the higher-rank neighbours of body pixel (tx, ty) whose footprint meets its own (bit order: see the fast path below)
__device__ __forceinline__ void body_release(int* s_cnt, uint32_t dep, int tx, int ty, int bhei) {
    int bit = 0;
    for (int dx = 0; dx <= 3; dx++)
        for (int dy = (dx == 0 ? 1 : -3); dy <= 3; dy++, bit++)
            if ((dep >> bit) & 1u) atomicSub(&s_cnt[(tx + dx) * bhei + ty + dy], 1);
}

// body_pixel_act for the fast path: the material plane of the footprint box is staged in shared memory (s_mat), so a pixel
// decides without a trip to HBM and its neighbours are released as soon as its stores are issued; the full cell is only read
// from the grid where the reference copies it (displaced sand / liquid, erased pixels).
template <bool ERASE>
__device__ __noinline__ void body_pixel_act_fast(const BodyArgs& a, int b, int tx, int ty, int lx, int ly, int4 box, int aw, uint8_t* s_mat,
                                                 int* s_cnt, uint32_t dep, int bhei) {
    fse_cell* tile = &a.tiles[a.off[b] + tx + ty * a.bw[b]];
    const fse_cell rm = *tile;
    int4* fb = &a.feedback[b];
    const int wx = lx + box.x, wy = ly + box.y;
    if (!ERASE) {
        for (int d = 0; d < 5; d++) {
            const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
            if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
            uint8_t* sm = &s_mat[(ly + c_dirs[d][1]) * aw + lx + c_dirs[d][0]];
            const size_t g = (size_t)y * a.W + x;
            const int ph = a.T->phys[*sm];
            if (ph == P_AIR) {
                write_cell(a, g, rm);
                *sm = (uint8_t)rm.mat;
                atomicAdd(&fb->z, 1);
                wake3x3(a, x, y);
                break;
            } else if (ph == P_SAND || ph == P_SOUP) {
                const unsigned int i = atomicAdd(a.pcount, 1u);
                if (i < a.pcap) {  // the displaced cell is thrown up as a loose particle (game.cpp:1791 / 1801)
                    fse_particle p;
                    memset(&p, 0, sizeof p);
                    p.tile = read_cell(a, g);
                    p.x = (float)x;
                    p.y = (float)(y + a.y_off - 3);  // particles carry global coordinates
                    const int pix = tx + ty * a.bw[b];
                    const uint32_t cb = rng_cell(a.rkey, b, pix);
                    p.vx = (float)(((int)(rng_draw(cb, S_BRIDGE_VX) % 10) - 5) / 10.0f);
                    p.vy = (float)(-(int)(rng_draw(cb, S_BRIDGE_VY) % 5 + 5) / 10.0f);
                    p.ay = 0.1f;
                    p.fade_time = 60;
                    p.id = (2ULL << 62) | ((uint64_t)(a.tick & 0xffff) << 40) | ((uint64_t)(b & 0xfffff) << 20) | (uint64_t)(pix & 0xfffff);
                    a.pbuf[i] = p;
                }
                write_cell(a, g, rm);
                *sm = (uint8_t)rm.mat;
                atomicAdd(ph == P_SAND ? &fb->x : &fb->y, 1);
                atomicAdd(&fb->z, 1);
                wake3x3(a, x, y);
                break;
            }
        }
        __threadfence_block();
        body_release(s_cnt, dep, tx, ty, bhei);
    } else {
        bool found = false;
        fse_cell got;
        for (int d = 0; d < 5; d++) {
            const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
            if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
            uint8_t* sm = &s_mat[(ly + c_dirs[d][1]) * aw + lx + c_dirs[d][0]];
            if (*sm == rm.mat) {  // .id == rmat.id (any cell of the same material)
                const size_t g = (size_t)y * a.W + x;
                got = read_cell(a, g);
                fse_cell nothing;
                memset(&nothing, 0, sizeof nothing);
                nothing.mat = (uint16_t)a.air;
                nothing.fluid = 2.0f;
                write_cell(a, g, nothing);
                *sm = (uint8_t)a.air;
                atomicAdd(&fb->z, 1);
                wake3x3(a, x, y);
                found = true;
                break;
            }
        }
        const bool lost = !found && wx >= 0 && wy >= 0 && wx < a.W && wy < a.H && s_mat[ly * aw + lx] == a.air;  // 1959-1965
        __threadfence_block();
        body_release(s_cnt, dep, tx, ty, bhei);
        if (found) {
            *tile = got;  // waits for the cell's loads; the neighbours are already on their way
        } else if (lost) {
            fse_cell nothing;
            memset(&nothing, 0, sizeof nothing);
            nothing.mat = (uint16_t)a.air;
            nothing.fluid = 2.0f;
            *tile = nothing;
            atomicAdd(&fb->w, 1);
        }
    }
}

// ---- fracture hand-off (world::updateRigidBodyHitbox, world.cpp:288-720, device part) ---------------------------------------------
// A body whose pixels were carved apart becomes one body per piece.  One CTA per body: the solid mask (tile != AIR, the alpha test of
// world.cpp:294-303) is labelled in shared memory (4-connected components by membership, north_star; the reference assigns pixels to
// the nearest triangle centroid instead, world.cpp:587-610), every piece gets its bounding box (the crop of world.cpp:305-320 applied
// to the piece, as the recursion at world.cpp:700-706 ends up doing), its pixel count and the weld flag (world.cpp:620), and its
// tiles are gathered into a w x h array of their own (AIR where the box covers another piece).  Pieces are numbered by their first
// pixel in row-major order.  Triangulation (TPPL) and the b2Body calls stay on the host.
constexpr int SPLIT_MAX_PIXELS = 128 * 128;
constexpr int SPLIT_MAX_PIECES = 1024;
struct SplitArgs {
    const fse_cell* tiles;  // body tiles, bw x bh
    int bw, bh, air;
    int weld_x, weld_y;
    fse_body_piece* pieces;  // out
    int cap_pieces;
    fse_cell* tiles_out;
    long long cap_tiles;
    int* result;  // [0] pieces found, [1] tiles used, [2] overflow flags (1 pieces, 2 tiles)
};
__global__ void __launch_bounds__(256) bodies_split_kernel(SplitArgs a) {
    extern __shared__ int split_smem[];
    int* L = split_smem;                          // labels: root = lowest pixel index of the component, -1 = empty
    int* box = split_smem + a.bw * a.bh;          // per piece: x0, y0, x1, y1, count, weld, root, tile offset (8 ints)
    __shared__ int n_roots, s_scan[256];
    const int tid = threadIdx.x, n = a.bw * a.bh, w = a.bw, h = a.bh;
    for (int i = tid; i < n; i += 256) L[i] = a.tiles[i].mat != a.air ? i : -1;
    __syncthreads();
    L = ccl_relax(L, split_smem + a.bw * a.bh + 8 * a.cap_pieces, n, w, h);  // second label buffer behind the piece boxes
    // number the roots in row-major order: per-thread counts over contiguous ranges, scanned across the CTA
    const int per = (n + 255) / 256, lo = tid * per, hi = min(n, lo + per);
    int mine = 0;
    for (int i = lo; i < hi; i++) mine += L[i] == i;
    s_scan[tid] = mine;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int q = 0; q < 256; q++) {
            const int v = s_scan[q];
            s_scan[q] = acc;
            acc += v;
        }
        n_roots = acc;
    }
    __syncthreads();
    const int np = n_roots;
    if (np > a.cap_pieces || np > SPLIT_MAX_PIECES) {
        if (tid == 0) {
            a.result[0] = np;
            a.result[2] = 1;
        }
        return;
    }
    {
        int k = s_scan[tid];
        for (int i = lo; i < hi; i++)
            if (L[i] == i) {
                int* b = box + 8 * k;
                b[0] = w; b[1] = h; b[2] = -1; b[3] = -1; b[4] = 0; b[5] = 0; b[6] = i; b[7] = 0;
                k++;
            }
    }
    __syncthreads();
    // root -> piece number: the root pixel's own label slot is free to carry it (stored as -2 - piece; members still point at the root)
    for (int k = tid; k < np; k += 256) L[box[8 * k + 6]] = -2 - k;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        int l = L[i];
        if (l == -1) continue;
        const int k = l <= -2 ? -2 - l : -2 - L[l];
        int* b = box + 8 * k;
        const int x = i % w, y = i / w;
        atomicMin(&b[0], x); atomicMin(&b[1], y); atomicMax(&b[2], x); atomicMax(&b[3], y);
        atomicAdd(&b[4], 1);
        if (x == a.weld_x && y == a.weld_y) b[5] = 1;
    }
    __syncthreads();
    if (tid == 0) {
        long long off = 0;
        for (int k = 0; k < np; k++) {
            int* b = box + 8 * k;
            b[7] = (int)off;
            off += (long long)(b[2] - b[0] + 1) * (b[3] - b[1] + 1);
        }
        a.result[0] = np;
        a.result[1] = (int)(off > 0x7fffffff ? 0x7fffffff : off);
        a.result[2] = off > a.cap_tiles ? 2 : 0;
    }
    __syncthreads();
    for (int k = tid; k < np; k += 256) {
        const int* b = box + 8 * k;
        fse_body_piece pc;
        pc.x0 = b[0]; pc.y0 = b[1]; pc.w = b[2] - b[0] + 1; pc.h = b[3] - b[1] + 1;
        pc.n_pixels = b[4]; pc.weld = b[5]; pc.tile_off = b[7];
        pc.shift_x = pc.shift_y = 0.0f;  // filled in by the host (sin / cos of the body angle on the host's libm, world.cpp:350-358)
        a.pieces[k] = pc;
    }
    if (a.result[2]) return;
    for (int k = 0; k < np; k++) {
        const int* b = box + 8 * k;
        const int pw = b[2] - b[0] + 1, ph = b[3] - b[1] + 1;
        for (int c = tid; c < pw * ph; c += 256) {
            const int x = b[0] + c % pw, y = b[1] + c / pw, i = x + y * w;
            const int l = L[i];
            const bool member = l != -1 && (l <= -2 ? -2 - l : -2 - L[l]) == k;
            fse_cell t;
            if (member) {
                t = a.tiles[i];
            } else {
                memset(&t, 0, sizeof t);
                t.mat = (uint16_t)a.air;
                t.fluid = 2.0f;
            }
            a.tiles_out[b[7] + c] = t;
        }
    }
}

constexpr int BODY_TB = 128;
constexpr int BODY_PPT = 8;   // pixels per thread kept in registers (bodies up to BODY_TB * BODY_PPT pixels)
constexpr int BODY_SCLAIM = 64 * 64;  // cells of a footprint box whose claim map fits shared memory

template <bool ERASE>
__global__ void __launch_bounds__(BODY_TB, 8) bodies_body_kernel(BodyArgs a) {
    __shared__ uint32_t s_claim[BODY_SCLAIM];
    __shared__ int s_b;
    const int tid = threadIdx.x;
    if (tid == 0) s_b = (int)atomicAdd(&a.ticket[0], 1u);
    __syncthreads();
    const int b = s_b;
    if (b >= a.n_bodies) return;
    if (a.exec && a.exec[b] != a.rank) {  // strips: another rank runs this body (and every body it could share cells with)
        if (tid == 0) atomicExch(a.done + b, 1u);
        return;
    }

    const int4 box = a.aabb[b];
    const int bwid = a.bw[b], bhei = a.bh[b], npix = bwid * bhei, off = a.off[b];
    // the reference's body order matters only between bodies that can touch the same cells
    for (int j = tid; j < b; j += BODY_TB) {
        const int4 o = a.aabb[j];
        if (o.x <= box.z && box.x <= o.z && o.y <= box.w && box.y <= o.w)
            while (bodies_ld_acquire(a.done + j) == 0u) __nanosleep(100);
    }
    __threadfence();
#ifdef FSE_BODIES_CYCLES  // profiling aid: cycles of the setup and of the rounds, summed over bodies (FSE_BODIES_DEBUG=1 prints them)
    __syncthreads();
    long long c1 = clock64();
#endif
    const int aw = box.z - box.x + 1, ah = box.w - box.y + 1;
    const bool smem = (long long)aw * ah <= BODY_SCLAIM;
    const bool fast = smem && npix <= BODY_TB * BODY_PPT && aw <= 255 && ah <= 255;  // positions inside the box are kept in 8 + 8 bits
    if (smem && !fast)
        for (int i = tid; i < aw * ah; i += BODY_TB) s_claim[i] = 0xffffffffu;
    for (int o = tid; o < npix; o += BODY_TB) {
        const int tx = o / bhei, ty = o % bhei;
        a.pending[off + o] = a.tiles[off + tx + ty * bwid].mat != a.air ? 1 : 0;
    }
    __syncthreads();
    // claim word of world cell (x, y); cells outside the world are never claimed
    auto claim_at = [&](int x, int y) -> uint32_t* {
        return smem ? &s_claim[(y - box.y) * aw + (x - box.x)] : &a.claim[(size_t)y * a.W + x];
    };
    int rounds = 0;
    if (fast) {
        // common case (bodies up to 1024 pixels, box up to 64x64).  Two footprints share a cell exactly when their centres are at
        // Manhattan distance <= 2, and two body pixels can only land that close when they are less than 4 apart in the body
        // (rotation keeps the distance, truncation moves a coordinate by < 1): every pixel counts the live lower-rank pixels of
        // its 7x7 body neighbourhood whose footprint meets its own, acts when that count is 0 — all earlier pixels that could
        // still change one of its cells are done, the same condition as holding the minimum on all five cells — and then
        // releases its higher-rank neighbours.  No claim map, one counter read per pending pixel and round.
        uint32_t* s_pos = s_claim;                                  // x | y << 8 inside the box, bit 16 live
        int* s_cnt = reinterpret_cast<int*>(s_claim + BODY_TB * BODY_PPT);
        uint8_t* s_mat = reinterpret_cast<uint8_t*>(s_claim + 2 * BODY_TB * BODY_PPT);  // material plane of the box (<= 64 x 64 bytes)
        for (int i = tid; i < aw * ah; i += BODY_TB) {
            const int x = box.x + i % aw, y = box.y + i / aw;
            s_mat[i] = (x >= 0 && y >= 0 && x < a.W && y < a.H) ? __ldcg(a.p.mat + (size_t)y * a.W + x) : (uint8_t)0;
        }
        const float4 t = a.xf[b];
        uint32_t px[BODY_PPT], dep[BODY_PPT];  // px: position | state << 16 (1 pending, 0 done or dead); dep: higher-rank neighbours to release
#pragma unroll
        for (int k = 0; k < BODY_PPT; k++) {
            const int o = tid + k * BODY_TB;
            px[k] = dep[k] = 0;
            if (o < npix) {
                const int tx = o / bhei, ty = o % bhei;
                const int wx = (int)(tx * t.w - (ty + 1) * t.z + t.x);  // game.cpp:1763
                const int wy = (int)(tx * t.z + (ty + 1) * t.w + t.y) - a.y_off;  // game.cpp:1764
                px[k] = (uint32_t)(wx - box.x) | ((uint32_t)(wy - box.y) << 8) | ((uint32_t)a.pending[off + o] << 16);
                s_pos[o] = px[k];
                s_cnt[o] = 0;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BODY_PPT; k++) {
            if (!(px[k] >> 16)) continue;
            const int o = tid + k * BODY_TB;
            const int tx = o / bhei, ty = o % bhei;
            const int lx = px[k] & 0xff, ly = (px[k] >> 8) & 0xff;
            int lower = 0, bit = 0;
            for (int dx = -3; dx <= 3; dx++)
                for (int dy = -3; dy <= 3; dy++) {
                    if (dx == 0 && dy == 0) continue;
                    const bool higher = dx > 0 || (dx == 0 && dy > 0);  // rank = tx * h + ty
                    const int ux = tx + dx, uy = ty + dy;
                    if (ux >= 0 && uy >= 0 && ux < bwid && uy < bhei) {
                        const uint32_t q = s_pos[ux * bhei + uy];
                        if ((q >> 16) && abs((int)(q & 0xff) - lx) + abs((int)((q >> 8) & 0xff) - ly) <= 2) {
                            if (higher) dep[k] |= 1u << bit;
                            else lower++;
                        }
                    }
                    if (higher) bit++;
                }
            s_cnt[o] = lower;
        }
        __syncthreads();
#ifdef FSE_BODIES_CYCLES
        if (tid == 0) atomicAdd(&a.ticket[3], (unsigned int)((clock64() - c1) >> 4));
        c1 = clock64();
#endif
        for (;;) {
            bool any = false;
#pragma unroll
            for (int k = 0; k < BODY_PPT; k++) {
                if (!(px[k] >> 16)) continue;
                any = true;
                const int o = tid + k * BODY_TB;
                if (*reinterpret_cast<volatile int*>(&s_cnt[o]) != 0) continue;
                __threadfence_block();  // acquire side of body_release: the releasing pixels' s_mat / grid writes are visible from here on
                const int tx = o / bhei, ty = o % bhei;
                body_pixel_act_fast<ERASE>(a, b, tx, ty, (int)(px[k] & 0xff), (int)((px[k] >> 8) & 0xff), box, aw, s_mat, s_cnt, dep[k], bhei);
                px[k] &= 0xffffu;
            }
            if (!__syncthreads_or(any)) break;
            rounds++;
        }
    } else
    for (;;) {
        bool any = false;
        for (int o = tid; o < npix; o += BODY_TB) {  // pending pixels publish their rank on their footprint
            if (!a.pending[off + o]) continue;
            any = true;
            int wx, wy;
            world_pos(a, b, o / bhei, o % bhei, wx, wy);
#pragma unroll
            for (int d = 0; d < 5; d++) {
                const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
                if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
                atomicMin(claim_at(x, y), (uint32_t)o);
            }
        }
        if (!__syncthreads_or(any)) break;
        rounds++;
        for (int o = tid; o < npix; o += BODY_TB) {  // a pixel that holds all five cells acts
            if (!a.pending[off + o]) continue;
            const int tx = o / bhei, ty = o % bhei;
            int wx, wy;
            world_pos(a, b, tx, ty, wx, wy);
            bool mine = true;
#pragma unroll
            for (int d = 0; d < 5; d++) {
                const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
                if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
                mine &= (smem ? *claim_at(x, y) : __ldcg(claim_at(x, y))) == (uint32_t)o;
            }
            if (!mine) continue;
            body_pixel_act<ERASE>(a, b, tx, ty, wx, wy);
            a.pending[off + o] = 2;
        }
        __syncthreads();
        for (int o = tid; o < npix; o += BODY_TB) {  // footprints are released; pixels that acted retire
            const uint8_t st = a.pending[off + o];
            if (!st) continue;
            if (st == 2) a.pending[off + o] = 0;
            int wx, wy;
            world_pos(a, b, o / bhei, o % bhei, wx, wy);
#pragma unroll
            for (int d = 0; d < 5; d++) {
                const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
                if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
                *claim_at(x, y) = 0xffffffffu;
            }
        }
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
#ifdef FSE_BODIES_CYCLES
    if (tid == 0) {
        atomicAdd(&a.ticket[4], (unsigned int)((clock64() - c1) >> 4));
        atomicAdd(&a.ticket[5], (unsigned int)rounds);
    }
#endif
    if (tid == 0) {
        atomicMax(&a.ticket[1], (unsigned int)rounds);
        __threadfence();
        atomicExch(a.done + b, 1u);
    }
}


// ---- multi-rank strips ---------------------------------------------------------------------------------------------------------------
// Every rank makes the call with the same transforms.  A body is run by ONE rank — the owner of the middle row of its footprint box —
// together with every body whose box overlaps its own (transitively: overlapping bodies keep the reference's body order, so they must
// meet on one rank; bodies that do not overlap commute).  The executing rank holds the whole box (own rows + GHOST rows, refreshed
// before the call); afterwards the part of the box that lies in a neighbour's rows travels there as a rectangle of cells, so both
// ranks agree on every row they share again.  Feedback is summed over the ranks (the others contribute zeros), and after an erase the
// tile arrays — replicated on every rank — are brought back in step the same way.
// x[i] = word i of the tiles if this rank ran the body the word belongs to, 0 otherwise (the sum over the ranks is the runner's copy)
__global__ void bodies_tiles_select_kernel(const uint32_t* tiles, uint32_t* x, const int* off, const int* exec, int n_bodies, int rank, int n_pixels) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    constexpr int CW = (int)(sizeof(fse_cell) / 4);  // words per body pixel
    static_assert(sizeof(fse_cell) % 4 == 0, "body tiles are exchanged as 32-bit words");
    if (i >= (size_t)n_pixels * CW) return;
    const int pix = (int)(i / CW);
    int lo = 0, hi = n_bodies;  // body of the pixel: off[lo] <= pix < off[lo + 1]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= pix) lo = mid;
        else hi = mid;
    }
    x[i] = exec[lo] == rank ? tiles[i] : 0u;
}

}  // namespace fse

using namespace fse;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return fail(FSE_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

struct fse_bodies {
    int n = 0, n_pixels = 0;
    std::vector<int> off, bw, bh;
    int *d_off = nullptr, *d_bw = nullptr, *d_bh = nullptr;
    fse_cell* d_tiles = nullptr;
    float4* d_xf = nullptr;
    uint8_t* d_pending = nullptr;
    uint32_t* d_claim = nullptr;
    int4* d_feedback = nullptr;
    int4* d_aabb = nullptr;
    unsigned int* d_done = nullptr;  // [n] done flags, then the ticket and the round maximum
    bool big = false;                // some body's footprint box may exceed the shared-memory claim map: claim plane in global memory
    // multi-rank strips
    int* d_exec = nullptr;           // [n] rank that runs each body in this call
    uint32_t* d_tiles_x = nullptr;   // n_pixels fse_cell (as words): the executed bodies' tiles, summed over the ranks after an erase
    void release() {
        cudaFree(d_off); cudaFree(d_bw); cudaFree(d_bh); cudaFree(d_tiles); cudaFree(d_xf); cudaFree(d_pending); cudaFree(d_claim);
        cudaFree(d_feedback); cudaFree(d_aabb); cudaFree(d_done);
        cudaFree(d_exec); cudaFree(d_tiles_x);
        d_exec = nullptr; d_tiles_x = nullptr;
        d_off = d_bw = d_bh = nullptr; d_tiles = nullptr; d_xf = nullptr; d_pending = nullptr; d_claim = nullptr; d_feedback = nullptr;
        d_aabb = nullptr; d_done = nullptr;
    }
};

void fse_bodies_free(fse_world* w) {
    if (w->bodies) {
        w->bodies->release();
        delete w->bodies;
        w->bodies = nullptr;
    }
}

extern "C" FSE_API int fse_bodies_upload(fse_world* w, const fse_body_desc* bodies, int32_t n) {
    if (!w || (!bodies && n > 0) || n < 0) return fail(FSE_EINVAL, "fse_bodies_upload: bad argument");
    // multi-rank strips: every rank uploads every body (the tile arrays are replicated; fse_bodies_erase brings them back in step)
    CK(cudaSetDevice(w->ctx->device));
    if (!w->bodies) w->bodies = new fse_bodies();
    fse_bodies* B = w->bodies;
    uint32_t* keep_claim = B->d_claim;
    B->d_claim = nullptr;
    B->release();
    B->d_claim = keep_claim;
    B->n = n;
    B->big = false;
    B->off.assign(n + 1, 0);
    B->bw.resize(n);
    B->bh.resize(n);
    for (int i = 0; i < n; i++) {
        if (bodies[i].w < 1 || bodies[i].h < 1 || bodies[i].w > 1024 || bodies[i].h > 1024 || !bodies[i].tiles)
            return fail(FSE_EINVAL, "fse_bodies_upload: body %d is %dx%d", i, bodies[i].w, bodies[i].h);
        B->bw[i] = bodies[i].w;
        B->bh[i] = bodies[i].h;
        B->off[i + 1] = B->off[i] + bodies[i].w * bodies[i].h;
        const double side = std::ceil(std::hypot((double)bodies[i].w, (double)bodies[i].h + 1.0)) + 8.0;  // bound of bodies_aabb_kernel's box
        if (side * side > (double)BODY_SCLAIM) B->big = true;
    }
    B->n_pixels = B->off[n];
    if (n == 0) return FSE_OK;
    std::vector<fse_cell> all((size_t)B->n_pixels);
    for (int i = 0; i < n; i++) memcpy(&all[B->off[i]], bodies[i].tiles, sizeof(fse_cell) * (size_t)bodies[i].w * bodies[i].h);
    CK(cudaMalloc(&B->d_off, sizeof(int) * (n + 1)));
    CK(cudaMalloc(&B->d_bw, sizeof(int) * n));
    CK(cudaMalloc(&B->d_bh, sizeof(int) * n));
    CK(cudaMalloc(&B->d_tiles, sizeof(fse_cell) * (size_t)B->n_pixels));
    CK(cudaMalloc(&B->d_xf, sizeof(float4) * n));
    CK(cudaMalloc(&B->d_pending, (size_t)B->n_pixels));
    CK(cudaMalloc(&B->d_feedback, sizeof(int4) * n));
    CK(cudaMalloc(&B->d_aabb, sizeof(int4) * n));
    CK(cudaMalloc(&B->d_done, sizeof(unsigned int) * (n + 8)));
    if (B->big && !B->d_claim) {
        CK(cudaMalloc(&B->d_claim, sizeof(uint32_t) * (size_t)w->W * w->H));
        CK(cudaMemsetAsync(B->d_claim, 0xff, sizeof(uint32_t) * (size_t)w->W * w->H, w->stream));
    }
    CK(cudaMemcpyAsync(B->d_off, B->off.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, w->stream));
    CK(cudaMemcpyAsync(B->d_bw, B->bw.data(), sizeof(int) * n, cudaMemcpyHostToDevice, w->stream));
    CK(cudaMemcpyAsync(B->d_bh, B->bh.data(), sizeof(int) * n, cudaMemcpyHostToDevice, w->stream));
    CK(cudaMemcpyAsync(B->d_tiles, all.data(), sizeof(fse_cell) * (size_t)B->n_pixels, cudaMemcpyHostToDevice, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return FSE_OK;
}


// ---- multi-rank strips: who runs which body, and which rectangles travel afterwards (host; every rank computes the same plan) -------
struct StripPlan {
    std::vector<int> exec;
    std::vector<int4> rect[4];  // up send, up recv, down send, down recv (x0, y0 in local rows, w, h)
};
static int plan_strip_bodies(fse_world* w, const fse_bodies* B, const fse_xform* xf, int n, StripPlan& P, const char* who) {
    std::vector<int4> box(n);
    for (int b = 0; b < n; b++) {  // the box of bodies_aabb_kernel, in global rows
        const float sn = std::sin(xf[b].angle), cs = std::cos(xf[b].angle);
        float x0 = 3.0e38f, y0 = 3.0e38f, x1 = -3.0e38f, y1 = -3.0e38f;
        for (int q = 0; q < 4; q++) {
            const float tx = (q & 1) ? (float)(B->bw[b] - 1) : 0.0f, ty1 = (q & 2) ? (float)B->bh[b] : 1.0f;
            const float fx = tx * cs - ty1 * sn + xf[b].x, fy = tx * sn + ty1 * cs + xf[b].y;
            x0 = std::fmin(x0, fx); x1 = std::fmax(x1, fx);
            y0 = std::fmin(y0, fy); y1 = std::fmax(y1, fy);
        }
        const float lim = 1.0e9f;
        x0 = std::fmax(std::fmin(x0, lim), -lim); x1 = std::fmax(std::fmin(x1, lim), -lim);
        y0 = std::fmax(std::fmin(y0, lim), -lim); y1 = std::fmax(std::fmin(y1, lim), -lim);
        box[b] = make_int4((int)std::floor(x0) - 3, (int)std::floor(y0) - 3, (int)std::ceil(x1) + 3, (int)std::ceil(y1) + 3);
    }
    if (int r = strip_group_runners(w, box, who, "body", P.exec)) return r;
    for (int q = 0; q < 4; q++) P.rect[q].clear();
    for (int b = 0; b < n; b++) strip_rects_of_box(w, P.exec[b], box[b].x, box[b].y, box[b].z, box[b].w, P.rect);
    return FSE_OK;
}

template <bool ERASE>
static int run_bridge(fse_world* w, const fse_xform* xf, int32_t n, uint32_t tick, uint32_t seed, fse_body_feedback* out) {
    fse_bodies* B = w->bodies;
    if (!B || B->n != n) return fail(FSE_ESTATE, "bodies: %d transforms for %d uploaded bodies (fse_bodies_upload first)", n, B ? B->n : 0);
    if (n == 0) return FSE_OK;
    CK(cudaSetDevice(w->ctx->device));
    if (!ERASE)  // every body pixel may displace a cell into the particle pool
        if (int r = particles_headroom(w, B->n_pixels, false)) return r;
    const bool multi = w->strip && w->ctx->nranks > 1;
    StripPlan plan;
    if (multi) {
        // the runner of a body works on its ghost rows as well: they must be the owner's rows first
        if (int r = strip_refresh(w, w->stream, STRIP_GHOST)) return r;
        if (int r = plan_strip_bodies(w, B, xf, n, plan, ERASE ? "fse_bodies_erase" : "fse_bodies_raster")) return r;
        if (!B->d_exec) CK(cudaMalloc(&B->d_exec, sizeof(int) * n));
        CK(cudaMemcpyAsync(B->d_exec, plan.exec.data(), sizeof(int) * n, cudaMemcpyHostToDevice, w->stream));
    }
    std::vector<float4> h(n);
    for (int i = 0; i < n; i++) h[i] = make_float4(xf[i].x, xf[i].y, std::sin(xf[i].angle), std::cos(xf[i].angle));  // game.cpp:1763-1764 on the host's libm
    CK(cudaMemcpyAsync(B->d_xf, h.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, w->stream));
    CK(cudaMemsetAsync(B->d_feedback, 0, sizeof(int4) * n, w->stream));
    BodyArgs a;
    a.p = w->p; a.T = w->ctx->d_tabs; a.W = w->W; a.H = w->H;
    a.n_bodies = n; a.off = B->d_off; a.bw = B->d_bw; a.bh = B->d_bh; a.tiles = B->d_tiles; a.xf = B->d_xf;
    a.pending = B->d_pending; a.claim = B->d_claim; a.counters = nullptr; a.feedback = B->d_feedback;
    a.pbuf = w->pbuf; a.pcount = w->pcount; a.pcap = w->pcap;
    a.rkey = rng_key(seed, tick, 7); a.tick = tick; a.air = w->ctx->h_tabs.air;
    a.awake = w->active_on ? w->d_awake : nullptr; a.acols = w->acols; a.arows = w->arows;
    a.n_pixels = B->n_pixels;
    a.aabb = B->d_aabb; a.done = B->d_done; a.ticket = B->d_done + n;
    a.y_off = w->y_off; a.exec = multi ? B->d_exec : nullptr; a.rank = w->ctx->rank;
    bodies_aabb_kernel<<<(n + 127) / 128, 128, 0, w->stream>>>(a);
    bodies_body_kernel<ERASE><<<n, BODY_TB, 0, w->stream>>>(a);
    CK(cudaGetLastError());
    w->ctx->launches += 2;
    if (multi) {
        // the boxes of the bodies this rank ran, as far as they lie in a neighbour's rows, travel there (and the neighbours' come here)
        if (int r = strip_push_rects(w, plan.rect, w->stream)) return r;
        // feedback: the runner's numbers on every rank
        if (int r = strip_allreduce_u32(w, (unsigned int*)B->d_feedback, (size_t)4 * n, w->stream)) return r;
        if (ERASE) {  // the erase rewrote the tiles of the bodies this rank ran: every rank gets every runner's copy
            if (!B->d_tiles_x) CK(cudaMalloc(&B->d_tiles_x, sizeof(fse_cell) * (size_t)B->n_pixels));
            const size_t words = (size_t)B->n_pixels * (sizeof(fse_cell) / 4);
            bodies_tiles_select_kernel<<<(unsigned)((words + 255) / 256), 256, 0, w->stream>>>((const uint32_t*)B->d_tiles, B->d_tiles_x, B->d_off, B->d_exec, n,
                                                                                                w->ctx->rank, B->n_pixels);
            CK(cudaGetLastError());
            w->ctx->launches += 1;
            if (int r = strip_allreduce_u32(w, B->d_tiles_x, words, w->stream)) return r;
            CK(cudaMemcpyAsync(B->d_tiles, B->d_tiles_x, sizeof(fse_cell) * (size_t)B->n_pixels, cudaMemcpyDeviceToDevice, w->stream));
        }
    }
    if (out) {  // the feedback read below joins the stream anyway: report the rounds of the slowest body with it
        unsigned int rounds = 0;
        CK(cudaMemcpyAsync(&rounds, B->d_done + n + 1, sizeof rounds, cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
        w->last_bridge_rounds = (int)rounds;
        if (getenv("FSE_BODIES_DEBUG")) {
            unsigned int t[6];
            cudaMemcpy(t, B->d_done + n, sizeof t, cudaMemcpyDeviceToHost);
            fprintf(stderr, "bodies %s: %d bodies, slowest body took %u rounds; per body: setup %.0f cycles, rounds %.0f cycles, %.1f rounds\n",
                    ERASE ? "erase" : "raster", n, rounds, 16.0 * t[3] / n, 16.0 * t[4] / n, (double)t[5] / n);
        }
    }
    if (out) {
        std::vector<int4> fb(n);
        CK(cudaMemcpy(fb.data(), B->d_feedback, sizeof(int4) * n, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) {
            out[i].sand_hits = fb[i].x;
            out[i].soup_hits = fb[i].y;
            out[i].placed = fb[i].z;
            out[i].destroyed = fb[i].w;
        }
    }
    return FSE_OK;
}

extern "C" FSE_API int fse_bodies_raster(fse_world* w, const fse_xform* xf, int32_t n, uint32_t tick, uint32_t seed, fse_body_feedback* out) {
    if (!w || (!xf && n > 0)) return fail(FSE_EINVAL, "fse_bodies_raster: null argument");
    return run_bridge<false>(w, xf, n, tick, seed, out);
}

extern "C" FSE_API int fse_bodies_erase(fse_world* w, const fse_xform* xf, int32_t n, fse_body_feedback* out, uint8_t* needs_update) {
    if (!w || (!xf && n > 0)) return fail(FSE_EINVAL, "fse_bodies_erase: null argument");
    if (int r = run_bridge<true>(w, xf, n, 0, 0, out)) return r;
    if (needs_update)
        for (int i = 0; i < n; i++) needs_update[i] = 1;  // game.cpp:1982
    return FSE_OK;
}

extern "C" FSE_API int fse_bodies_read(fse_world* w, int32_t body, fse_cell* tiles_out) {
    if (!w || !tiles_out || !w->bodies || body < 0 || body >= w->bodies->n) return fail(FSE_EINVAL, "fse_bodies_read: bad argument");
    CK(cudaSetDevice(w->ctx->device));
    fse_bodies* B = w->bodies;
    CK(cudaMemcpyAsync(tiles_out, B->d_tiles + B->off[body], sizeof(fse_cell) * (size_t)B->bw[body] * B->bh[body], cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return FSE_OK;
}

extern "C" FSE_API int fse_bodies_split(fse_world* w, int32_t body, float angle, int32_t weld_x, int32_t weld_y, fse_body_piece* pieces, int32_t cap_pieces,
                                        int32_t* n_pieces, fse_cell* tiles_out, int64_t cap_tiles) {
    if (!w || !w->bodies || body < 0 || body >= w->bodies->n || !pieces || !n_pieces || !tiles_out || cap_pieces < 1 || cap_tiles < 1)
        return fail(FSE_EINVAL, "fse_bodies_split: bad argument");
    CK(cudaSetDevice(w->ctx->device));
    fse_bodies* B = w->bodies;
    const int bw = B->bw[body], bh = B->bh[body];
    if ((long long)bw * bh > SPLIT_MAX_PIXELS) return fail(FSE_EINVAL, "fse_bodies_split: body of %d x %d pixels (at most %d)", bw, bh, SPLIT_MAX_PIXELS);
    if (cap_pieces > SPLIT_MAX_PIECES) cap_pieces = SPLIT_MAX_PIECES;
    const size_t o_tiles = ((sizeof(fse_body_piece) * (size_t)cap_pieces + 255) / 256) * 256, o_res = o_tiles + ((sizeof(fse_cell) * (size_t)cap_tiles + 255) / 256) * 256;
    if (w->outline_scratch_bytes < o_res + 64) {
        CK(cudaStreamSynchronize(w->stream));
        cudaFree(w->outline_scratch);
        w->outline_scratch = nullptr;
        w->outline_scratch_bytes = 0;
        CK(cudaMalloc(&w->outline_scratch, o_res + 64));
        w->outline_scratch_bytes = o_res + 64;
    }
    char* base = (char*)w->outline_scratch;
    SplitArgs a;
    a.tiles = B->d_tiles + B->off[body];
    a.bw = bw; a.bh = bh; a.air = w->ctx->h_tabs.air;
    a.weld_x = weld_x; a.weld_y = weld_y;
    a.pieces = (fse_body_piece*)base;
    a.cap_pieces = cap_pieces;
    a.tiles_out = (fse_cell*)(base + o_tiles);
    a.cap_tiles = cap_tiles;
    a.result = (int*)(base + o_res);
    CK(cudaMemsetAsync(a.result, 0, 16, w->stream));
    const size_t smem = sizeof(int) * (2 * (size_t)bw * bh + 8 * (size_t)cap_pieces);  // labels, piece boxes, second label buffer (ccl_relax)
    static bool configured = false;
    if (!configured) {
        CK(cudaFuncSetAttribute(bodies_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(int) * (2 * SPLIT_MAX_PIXELS + 8 * SPLIT_MAX_PIECES))));
        configured = true;
    }
    bodies_split_kernel<<<1, 256, smem, w->stream>>>(a);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    int res[4];
    CK(cudaMemcpyAsync(res, a.result, 16, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    *n_pieces = res[0];
    if (res[2] == 1) return fail(FSE_ENOMEM, "fse_bodies_split: %d pieces exceed the caller's capacity (%d)", res[0], cap_pieces);
    if (res[2] == 2) return fail(FSE_ENOMEM, "fse_bodies_split: the pieces need %d tiles, the caller gave %lld", res[1], (long long)cap_tiles);
    if (res[0] > 0) {
        CK(cudaMemcpy(pieces, a.pieces, sizeof(fse_body_piece) * (size_t)res[0], cudaMemcpyDeviceToHost));
        if (res[1] > 0) CK(cudaMemcpy(tiles_out, a.tiles_out, sizeof(fse_cell) * (size_t)res[1], cudaMemcpyDeviceToHost));
    }
    const float s = std::sin(angle), c = std::cos(angle);  // world.cpp:350-358: the crop moves the body by the rotated corner
    for (int k = 0; k < res[0]; k++) {
        pieces[k].shift_x = pieces[k].x0 * c - pieces[k].y0 * s;
        pieces[k].shift_y = pieces[k].x0 * s + pieces[k].y0 * c;
    }
    return FSE_OK;
}
