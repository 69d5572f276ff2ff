// fse_outline.cu — fracture detection and collider outlines on the GPU.
//
// Reference: world::updateRigidBodyHitbox / updateChunkMesh (source/engine/world.cpp:288-720, 722-959) trace every
// perimeter of an alpha / SOLID mask with MarchingSquares::FindPerimeter and simplify it with Douglas-Peucker
// (source/engine/physics/physics_math.cpp:1766-1965); world::physicsCheck (world.cpp:3330-3429) flood-fills a SOLID blob.
// Here, per mask:
//   * ccl_kernel            4-connected component labels (label = lowest pixel index of the component) by min-propagation
//                           with pointer jumping in one CTA per mask;
//   * contour_kernel        one thread per start candidate (world.cpp:412-451): it walks its marching-squares loop and
//                           gives up as soon as it meets a lower-index candidate of the same loop, so exactly one thread
//                           per loop survives — the order-free replacement of the reference's edgeSeen scan; the survivor
//                           run-length-compresses the walk (physics_math.cpp:1951-1957), runs Douglas-Peucker with an
//                           explicit stack (tolerance 1, first/last kept) and appends a contour record;
//   * flood_kernel          bounded BFS (cap 1000) from a seed with a shared-memory hash set.
// TPPL hole removal / ear clipping and b2Body creation stay on the host (they feed Box2D, which stays on the host).
#include <algorithm>
#include <vector>

#include "fse_internal.hpp"

namespace fse {

// ---- marching squares (physics_math.cpp:1870-1949) -------------------------------------------------------------
__device__ __forceinline__ bool ms_set(const uint8_t* d, int x, int y, int w, int h) {
    return x <= 0 || x > w || y <= 0 || y > h ? false : d[(y - 1) * w + (x - 1)] != 0;
}
__device__ __forceinline__ int ms_value(const uint8_t* d, int x, int y, int w, int h) {
    return (ms_set(d, x, y, w, h) ? 1 : 0) | (ms_set(d, x + 1, y, w, h) ? 2 : 0) | (ms_set(d, x, y + 1, w, h) ? 4 : 0) |
           (ms_set(d, x + 1, y + 1, w, h) ? 8 : 0);
}
// direction codes: 1 East(1,0) 2 North(0,1) 3 West(-1,0) 4 South(0,-1); prev selects the branch at saddles 6 / 9
__device__ __forceinline__ int ms_dir(int v, int prev) {
    // nibble v of the constant = direction of case v: {0,2,1,1,3,2,-,1,4,-,4,4,3,2,3,0}
    if (v == 6) return prev == 2 ? 3 : 1;
    if (v == 9) return prev == 1 ? 2 : 4;
    return (int)((0x0323440410231120ULL >> (4 * v)) & 0xF);
}
__device__ __forceinline__ int dir_dx(int d) { return d == 1 ? 1 : (d == 3 ? -1 : 0); }
__device__ __forceinline__ int dir_dy(int d) { return d == 2 ? 1 : (d == 4 ? -1 : 0); }

__device__ __forceinline__ bool is_candidate(const uint8_t* d, int i, int w, int h) {  // world.cpp:413-451
    if (!d[i]) return false;
    const int x = i % w, y = i / w;
    int nb = 0;
    if (x + 1 < w) nb += d[i + 1] != 0;
    if (y + 1 < h) nb += d[i + w] != 0;
    if (y + 1 < h && x + 1 < w) nb += d[i + w + 1] != 0;
    if (nb == 3) return false;
    const int v = ms_value(d, x, y, w, h);
    return v != 0 && v != 15;
}

// physics_math.cpp:1813-1843
__device__ __forceinline__ float p_distance(float x, float y, float x1, float y1, float x2, float y2) {
    const float A = x - x1, B = y - y1, C = x2 - x1, D = y2 - y1;
    const float dot = A * C + B * D;
    const float len_sq = C * C + D * D;
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
    const float dx = x - xx, dy = y - yy;
    return sqrtf(dx * dx + dy * dy);
}

struct OutlineArgs {
    const uint8_t* masks;
    int n, w, h;
    int32_t* labels;   // n*w*h
    int* ncomp;        // n
    float* pool;
    unsigned int* pool_used;
    unsigned int pool_cap;  // floats
    // Per cell of every mask (index mask * w * h + cell): the contour that starts at this candidate cell, if it survived.  The contours
    // come out ordered by (mask, candidate index) — the reference's discovery order — through two scans on the device, so the host
    // neither sorts records nor copies scratch: cell_kept -> rank inside the mask and point offset inside the mask (contour_rank_kernel)
    // -> offsets of the masks (mask_scan_kernel) -> compacted points and offsets (contour_scatter_kernel).
    int* cell_kept;          // points of the simplified contour, 0 = no contour starts here (zeroed before the launch)
    unsigned int* cell_off;  // pool offset of its points
    int* cell_rank;          // contours of the same mask with a lower candidate index
    int* cell_ptoff;         // points of those contours
    int* mask_cnt;           // [n] contours per mask, [n] points per mask behind them
    int* mask_off;           // [n + 1] first contour of each mask, [n] first point of each mask behind them
    int* pt_off;             // [contours + 1] first point of each contour (output)
    float* out_pts;          // compacted points (output)
    unsigned int out_cap;    // floats
    unsigned int* n_recs;    // [0] contours, [1] points (set by mask_scan_kernel)
    unsigned int rec_cap;
    int* overflow;
    int want_labels;   // copy the labels out (callers that only need contours and component counts skip 4 bytes per cell)
};

// 4-connected component labels of one mask per CTA (label = lowest pixel index of the component, -1 = unset).  Masks of up to
// CCL_SMEM_CELLS cells (every rigid body of config 4 and every 128 x 96 chunk crop) are labelled in shared memory and only the
// finished labels go to HBM — and only when the caller asked for them; larger masks iterate on the global array.
constexpr int CCL_SMEM_CELLS = 12288;
__global__ void __launch_bounds__(256) ccl_kernel(OutlineArgs a) {
    extern __shared__ int ccl_smem[];
    const int m = blockIdx.x;
    const uint8_t* d = a.masks + (size_t)m * a.w * a.h;
    const int n = a.w * a.h, w = a.w, h = a.h;
    const bool in_smem = n <= CCL_SMEM_CELLS;
    int32_t* G = a.labels + (size_t)m * a.w * a.h;
    int* L = in_smem ? reinterpret_cast<int*>(ccl_smem) : reinterpret_cast<int*>(G);
    for (int i = threadIdx.x; i < n; i += blockDim.x) L[i] = d[i] ? i : -1;
    __syncthreads();
    if (in_smem) {
        L = ccl_relax(L, ccl_smem + n, n, w, h);  // two label buffers in shared memory, one read and one written per sweep
    } else {
        __shared__ int changed;
        for (;;) {  // labels in global memory: in-place sweeps (monotone: every order of updates ends at the same labels)
            if (threadIdx.x == 0) changed = 0;
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int l = L[i];
                if (l < 0) continue;
                const int x = i % w, y = i / w;
                int best = l;
                if (x + 1 < w && L[i + 1] >= 0) best = min(best, L[i + 1]);
                if (x > 0 && L[i - 1] >= 0) best = min(best, L[i - 1]);
                if (y + 1 < h && L[i + w] >= 0) best = min(best, L[i + w]);
                if (y > 0 && L[i - w] >= 0) best = min(best, L[i - w]);
                best = min(best, L[best]);  // pointer jumping
                if (best < l) {
                    L[i] = best;
                    atomicMin(&L[l], best);  // pull the old representative along
                    changed = 1;
                }
            }
            __syncthreads();
            if (!changed) break;
            __syncthreads();
        }
    }
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int mine = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        mine += L[i] == i;
        if (in_smem && a.want_labels) G[i] = L[i];
    }
    atomicAdd(&cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) a.ncomp[m] = cnt;
}

// One thread per mask cell.  A start candidate walks its marching-squares loop and gives up as soon as it meets a lower-index
// candidate of the same loop, so exactly one thread per loop survives; the survivor walks the loop once more to lay down one vertex
// per direction run.  Douglas-Peucker then runs one loop at a time with the WHOLE WARP: the farthest-vertex search of a section —
// the O(n) part of every split — is spread over the lanes (ties go to the lowest index, like the sequential scan), so the long outer
// loop of a body no longer makes the kernel's time on one thread.
__global__ void contour_kernel(OutlineArgs a) {
    const int gi = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int per = a.w * a.h, w = a.w, h = a.h;
    const bool in_range = gi < a.n * per;
    const int m = in_range ? gi / per : 0, i = in_range ? gi % per : 0;
    const uint8_t* d = a.masks + (size_t)m * per;
    bool alive = in_range && is_candidate(d, i, w, h);
    const int sx = i % w, sy = i / w;
    int runs = 0;
    unsigned int off = 0;
    if (alive) {  // first walk: canonical? how many direction runs?
        int x = sx, y = sy, prev = 0;
        do {
            const int v = ms_value(d, x, y, w, h);
            const int dir = ms_dir(v, prev);
            if (!(x == sx && y == sy && prev == 0) && x >= 0 && y >= 0 && x < w && y < h) {
                const int j = x + y * w;
                if (j < i && is_candidate(d, j, w, h) && ms_dir(v, 0) == dir) {  // a lower candidate owns this loop
                    alive = false;
                    break;
                }
            }
            if (dir != prev) runs++;
            prev = dir;
            x += dir_dx(dir);
            y -= dir_dy(dir);
        } while (x != sx || y != sy);
    }
    if (alive) {  // scratch: px[runs] py[runs] mark[runs] stack[2*runs] out[2*runs]  (floats / ints of the same size)
        const unsigned int need = 7u * (unsigned int)runs;
        off = atomicAdd(a.pool_used, need);
        if (off + need > a.pool_cap) {
            atomicExch(a.overflow, 1);
            alive = false;
        }
    }
    if (alive) {  // second walk: one vertex at the end of every run (world.cpp:483-485)
        float* px = a.pool + off;
        float* py = px + runs;
        int x = sx, y = sy, prev = 0, k = -1;
        do {
            const int dir = ms_dir(ms_value(d, x, y, w, h), prev);
            if (dir != prev) k++;
            prev = dir;
            x += dir_dx(dir);
            y -= dir_dy(dir);
            px[k] = (float)x;
            py[k] = (float)y;
        } while (x != sx || y != sy);
    }
    __syncwarp();
    // simplify(worldMesh, 1) (physics_math.cpp:1766-1811), explicit stack instead of recursion, one loop of this warp at a time
    unsigned todo = __ballot_sync(0xffffffffu, alive);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int np = __shfl_sync(0xffffffffu, runs, src);
        const unsigned int o = __shfl_sync(0xffffffffu, off, src);
        float* px = a.pool + o;
        float* py = px + np;
        int* mark = reinterpret_cast<int*>(py + np);
        int* stack = mark + np;
        float* out = reinterpret_cast<float*>(stack + 2 * np);
        for (int q = lane; q < np; q += 32) mark[q] = 1;
        __syncwarp();
        if (np > 2) {
            int sp = 0;
            if (lane == 0) {
                stack[0] = 0;
                stack[1] = np - 1;
            }
            sp = 2;
            __syncwarp();
            while (sp > 0) {
                const int jj = stack[sp - 1], ii = stack[sp - 2];
                sp -= 2;
                __syncwarp();  // every lane has read the section before a push overwrites its slots
                if (ii + 1 == jj) continue;
                const float x1 = px[ii], y1 = py[ii], x2 = px[jj], y2 = py[jj];
                float maxd = -1.0f;
                int maxi = ii;
                for (int q = ii + 1 + lane; q < jj; q += 32) {
                    const float dist = p_distance(px[q], py[q], x1, y1, x2, y2);
                    if (dist > maxd) {
                        maxd = dist;
                        maxi = q;
                    }
                }
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) {  // the largest distance; of equal ones the lowest index, like the sequential scan
                    const float od = __shfl_xor_sync(0xffffffffu, maxd, sh);
                    const int oi = __shfl_xor_sync(0xffffffffu, maxi, sh);
                    if (od > maxd || (od == maxd && oi < maxi)) {
                        maxd = od;
                        maxi = oi;
                    }
                }
                if (maxd <= 1.0f) {
                    for (int q = ii + 1 + lane; q < jj; q += 32) mark[q] = 0;
                } else {
                    if (lane == 0) {
                        stack[sp] = ii; stack[sp + 1] = maxi;
                        stack[sp + 2] = maxi; stack[sp + 3] = jj;
                    }
                    sp += 4;
                }
                __syncwarp();
            }
        }
        __syncwarp();
        // kept vertices, in order: 32 at a time, positions from a ballot
        int kept = 0;
        for (int base = 0; base < np; base += 32) {
            const int q = base + lane;
            const bool keep = q < np && mark[q];
            const unsigned kb = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int pos = kept + __popc(kb & ((1u << lane) - 1u));
                out[2 * pos] = px[q];
                out[2 * pos + 1] = py[q];
            }
            kept += __popc(kb);
        }
        if (lane == src && kept >= 3) {  // world.cpp:490
            a.cell_kept[gi] = kept;
            a.cell_off[gi] = o + 5u * (unsigned int)np;
        }
        __syncwarp();
    }
}

// exclusive scan of (flag, value) pairs over the 256 threads of a CTA; returns this thread's offsets, *tot_f / *tot_v get the CTA totals
__device__ __forceinline__ void block_scan2(int f, int v, int& ex_f, int& ex_v, int& tot_f, int& tot_v, int* ws /* 16 ints of shared memory */) {
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    int inf = f, inv = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int tf = __shfl_up_sync(0xffffffffu, inf, d), tv = __shfl_up_sync(0xffffffffu, inv, d);
        if (lane >= d) { inf += tf; inv += tv; }
    }
    if (lane == 31) { ws[wp] = inf; ws[8 + wp] = inv; }
    __syncthreads();
    int of = 0, ov = 0, sf = 0, sv = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        if (q < wp) { of += ws[q]; ov += ws[8 + q]; }
        sf += ws[q]; sv += ws[8 + q];
    }
    ex_f = of + inf - f;
    ex_v = ov + inv - v;
    tot_f = sf;
    tot_v = sv;
    __syncthreads();
}

// one CTA per mask: rank and point offset of every contour inside its mask, contour and point totals of the mask
__global__ void __launch_bounds__(256) contour_rank_kernel(OutlineArgs a) {
    __shared__ int ws[16];
    const int m = blockIdx.x, per = a.w * a.h;
    const size_t base = (size_t)m * per;
    int carry_c = 0, carry_p = 0;
    for (int tile = 0; tile < per; tile += 256) {
        const int i = tile + threadIdx.x;
        const int k = i < per ? a.cell_kept[base + i] : 0;
        int ec, ep, tc, tp;
        block_scan2(k > 0, k, ec, ep, tc, tp, ws);
        if (k > 0) {
            a.cell_rank[base + i] = carry_c + ec;
            a.cell_ptoff[base + i] = carry_p + ep;
        }
        carry_c += tc;
        carry_p += tp;
    }
    if (threadIdx.x == 0) {
        a.mask_cnt[m] = carry_c;
        a.mask_cnt[a.n + m] = carry_p;
    }
}

// one CTA: offsets of the masks (exclusive scans of the per-mask totals), the grand totals
__global__ void __launch_bounds__(256) mask_scan_kernel(OutlineArgs a) {
    __shared__ int ws[16];
    int carry_c = 0, carry_p = 0;
    for (int tile = 0; tile < a.n; tile += 256) {
        const int m = tile + threadIdx.x;
        const int c = m < a.n ? a.mask_cnt[m] : 0, p = m < a.n ? a.mask_cnt[a.n + m] : 0;
        int ec, ep, tc, tp;
        block_scan2(c, p, ec, ep, tc, tp, ws);
        if (m < a.n) {
            a.mask_off[m] = carry_c + ec;
            a.mask_off[a.n + 1 + m] = carry_p + ep;
        }
        carry_c += tc;
        carry_p += tp;
    }
    if (threadIdx.x == 0) {
        a.mask_off[a.n] = carry_c;
        a.n_recs[0] = (unsigned int)carry_c;
        a.n_recs[1] = (unsigned int)carry_p;
        if ((unsigned int)carry_c > a.rec_cap || 2u * (unsigned int)carry_p > a.out_cap) atomicExch(a.overflow, 1);
        else a.pt_off[carry_c] = carry_p;
    }
}

// one thread per candidate cell with a contour: its points go to their final place
__global__ void contour_scatter_kernel(OutlineArgs a) {
    const int gi = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = a.w * a.h;
    if (gi >= a.n * per || *a.overflow) return;
    const int kept = a.cell_kept[gi];
    if (kept <= 0) return;
    const int m = gi / per;
    const int c = a.mask_off[m] + a.cell_rank[gi], o = a.mask_off[a.n + 1 + m] + a.cell_ptoff[gi];
    a.pt_off[c] = o;
    const float* src = a.pool + a.cell_off[gi];
    float* dst = a.out_pts + 2 * (size_t)o;
    for (int q = 0; q < 2 * kept; q++) dst[q] = src[q];
}

__global__ void solid_mask_kernel(const uint8_t* mat, const DevTables* T, int W, int x0, int y0, int rw, int rh, uint8_t* out) {
    const size_t n = (size_t)rw * rh;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = T->phys[mat[(size_t)(y0 + i / rw) * W + (x0 + i % rw)]] == P_SOLID;
}

// physicsCheck_flood (world.cpp:3413-3429): wavefront BFS of the 4-connected SOLID component, one CTA.
constexpr int FLOOD_HASH = 8192;
__global__ void flood_kernel(const uint8_t* mat, const DevTables* T, int W, int H, int sx, int sy, int cap, int* out_pixels, int* out_count) {
    __shared__ int table[FLOOD_HASH];
    __shared__ int q_head, q_tail, q_next;
    for (int i = threadIdx.x; i < FLOOD_HASH; i += blockDim.x) table[i] = -1;
    if (threadIdx.x == 0) {
        q_head = 0;
        q_tail = 0;
        *out_count = 0;
    }
    __syncthreads();
    if (T->phys[mat[(size_t)sy * W + sx]] != P_SOLID) return;
    if (threadIdx.x == 0) {
        const int p = sx + sy * W;
        table[(unsigned)(p * 2654435761u) % FLOOD_HASH] = p;
        out_pixels[0] = p;
        q_tail = 1;
    }
    __syncthreads();
    for (;;) {
        const int head = q_head, tail = q_tail;
        if (head >= tail || tail > cap) break;
        if (threadIdx.x == 0) q_next = tail;
        __syncthreads();
        for (int f = head + threadIdx.x / 4; f < tail; f += blockDim.x / 4) {
            const int p = out_pixels[f];
            const int px = p % W, py = p / W, d = threadIdx.x & 3;
            const int nx = px + (d == 0) - (d == 2), ny = py + (d == 1) - (d == 3);
            if (nx < 0 || ny < 0 || nx >= W || ny >= H) continue;
            if (T->phys[mat[(size_t)ny * W + nx]] != P_SOLID) continue;
            const int q = nx + ny * W;
            unsigned hsh = (unsigned)(q * 2654435761u) % FLOOD_HASH;
            bool fresh = false;
            for (int probe = 0; probe < FLOOD_HASH; probe++) {
                const int old = atomicCAS(&table[hsh], -1, q);
                if (old == -1) { fresh = true; break; }
                if (old == q) break;
                hsh = (hsh + 1) % FLOOD_HASH;
            }
            if (fresh) {
                const int slot = atomicAdd(&q_next, 1);
                if (slot <= cap) out_pixels[slot] = q;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            q_head = tail;
            q_tail = q_next;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_count = q_tail > cap ? cap + 1 : q_tail;
}

// world::physicsCheck, the cut (world.cpp:3352-3395): every cell of the component becomes Tiles_NOTHING (dirty); with `tiles` the cell's
// colour goes into the w x h tile array of the new body as OBSIDIAN (what makeRigidBody builds from the colour surface, world.cpp:191-209).
__global__ void physcheck_cut_kernel(Planes p, const DevTables* T, int W, const int* pixels, int n, int minx, int miny, int bw, fse_cell* tiles) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = pixels[i];
    if (tiles) {
        fse_cell t;
        memset(&t, 0, sizeof t);
        t.mat = (uint16_t)T->obsidian;
        t.color = p.col[g];
        t.temp = T->ctemp[T->obsidian];
        t.fluid = 2.0f;
        tiles[(g % W - minx) + (g / W - miny) * bw] = t;
    }
    p.mat[g] = (uint8_t)T->air;
    p.flg[g] = F_DIRTY;
    p.stl[g] = 0;
    p.tmp[g] = 0;
    p.col[g] = 0;
    p.fl[g] = 2.0f;
    p.fd[g] = 0.0f;
}
__global__ void fill_air_tiles_kernel(fse_cell* tiles, int n, uint16_t air) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fse_cell t;
    memset(&t, 0, sizeof t);
    t.mat = air;
    t.fluid = 2.0f;
    tiles[i] = t;
}

static cudaError_t grow_scratch(fse_world* w, size_t need) {
    if (w->outline_scratch_bytes >= need) return cudaSuccess;
    cudaFree(w->outline_scratch);
    w->outline_scratch = nullptr;
    w->outline_scratch_bytes = 0;
    cudaError_t e = cudaMalloc(&w->outline_scratch, need);
    if (e == cudaSuccess) w->outline_scratch_bytes = need;
    return e;
}

}  // namespace fse

using namespace fse;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return fail(FSE_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

extern "C" FSE_API int fse_mask_outline(fse_world* w, const uint8_t* masks, int32_t n_masks, int32_t mw, int32_t mh, int32_t* labels,
                                        int32_t* n_components, float* pts, int32_t cap_pts, int32_t* pt_off, int32_t cap_contours,
                                        int32_t* mask_off) {
    if (!w || !masks || n_masks < 1 || mw < 1 || mh < 1 || mw > 2048 || mh > 2048 || !pts || !pt_off || !mask_off)
        return fail(FSE_EINVAL, "fse_mask_outline: bad argument");
    CK(cudaSetDevice(w->ctx->device));
    const size_t per = (size_t)mw * mh, total = per * n_masks;
    if (total > ((size_t)1 << 30)) return fail(FSE_EINVAL, "fse_mask_outline: %zu mask cells in one call (at most 2^30)", total);
    const unsigned int pool_cap = (unsigned int)std::min<size_t>((size_t)1 << 30, 16 * total + 4096);
    const unsigned int rec_cap = (unsigned int)std::min<size_t>((size_t)1 << 26, total / 2 + 1024);
    const unsigned int out_cap = pool_cap / 7 * 2 + 64;  // a contour's scratch is 7 x its runs, its output at most 2 x its runs
    // scratch layout: masks | labels | ncomp | counters | cell arrays | mask totals / offsets | contour offsets | output points | pool
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t o_masks = 0, o_labels = up(total), o_ncomp = up(o_labels + total * 4), o_cnt = up(o_ncomp + (size_t)n_masks * 4),
                 o_kept = o_cnt + 256, o_coff = up(o_kept + total * 4), o_rank = up(o_coff + total * 4), o_ptoff = up(o_rank + total * 4),
                 o_mcnt = up(o_ptoff + total * 4), o_moff = up(o_mcnt + (size_t)n_masks * 8), o_poff = up(o_moff + ((size_t)n_masks * 2 + 1) * 4),
                 o_out = up(o_poff + ((size_t)rec_cap + 1) * 4), o_pool = up(o_out + (size_t)out_cap * 4), bytes = o_pool + (size_t)pool_cap * 4;
    CK(grow_scratch(w, bytes));
    char* base = (char*)w->outline_scratch;
    OutlineArgs a;
    a.masks = (const uint8_t*)(base + o_masks);
    a.n = n_masks; a.w = mw; a.h = mh;
    a.labels = (int32_t*)(base + o_labels);
    a.ncomp = (int*)(base + o_ncomp);
    a.pool_used = (unsigned int*)(base + o_cnt);
    a.n_recs = a.pool_used + 1;                // [1] contours, [2] points
    a.overflow = (int*)(a.pool_used + 3);
    a.cell_kept = (int*)(base + o_kept);
    a.cell_off = (unsigned int*)(base + o_coff);
    a.cell_rank = (int*)(base + o_rank);
    a.cell_ptoff = (int*)(base + o_ptoff);
    a.mask_cnt = (int*)(base + o_mcnt);
    a.mask_off = (int*)(base + o_moff);
    a.pt_off = (int*)(base + o_poff);
    a.out_pts = (float*)(base + o_out);
    a.out_cap = out_cap;
    a.rec_cap = rec_cap;
    a.pool = (float*)(base + o_pool);
    a.pool_cap = pool_cap;
    // host traffic goes through pinned staging buffers (masks in; counters, labels, component counts and mask offsets out in one batch;
    // contour offsets and points — already in their final order and place — once the counts are known)
    const size_t st_labels = up(total), st_ncomp = up(st_labels + total * 4), st_moff = up(st_ncomp + (size_t)n_masks * 4),
                 st_cnt = up(st_moff + ((size_t)n_masks + 1) * 4), st_fixed = st_cnt + 256;
    if (w->outline_pinned_bytes < st_fixed) {
        if (w->outline_pinned) cudaFreeHost(w->outline_pinned);
        w->outline_pinned = nullptr;
        w->outline_pinned_bytes = 0;
        CK(cudaMallocHost(&w->outline_pinned, st_fixed));
        w->outline_pinned_bytes = st_fixed;
    }
    char* pin = (char*)w->outline_pinned;
    memcpy(pin, masks, total);
    CK(cudaMemcpyAsync(base + o_masks, pin, total, cudaMemcpyHostToDevice, w->stream));
    CK(cudaMemsetAsync(base + o_cnt, 0, 256, w->stream));
    CK(cudaMemsetAsync(a.cell_kept, 0, total * 4, w->stream));
    a.want_labels = labels != nullptr;
    const size_t ccl_smem_bytes = per <= (size_t)CCL_SMEM_CELLS ? 2 * per * sizeof(int32_t) : 0;  // two label buffers (ccl_relax)
    static bool ccl_configured = false;
    if (!ccl_configured) {
        CK(cudaFuncSetAttribute(ccl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * CCL_SMEM_CELLS * (int)sizeof(int32_t)));
        ccl_configured = true;
    }
    ccl_kernel<<<n_masks, 256, ccl_smem_bytes, w->stream>>>(a);
    CK(cudaGetLastError());
    contour_kernel<<<(unsigned int)((total + 127) / 128), 128, 0, w->stream>>>(a);
    CK(cudaGetLastError());
    contour_rank_kernel<<<n_masks, 256, 0, w->stream>>>(a);
    CK(cudaGetLastError());
    mask_scan_kernel<<<1, 256, 0, w->stream>>>(a);
    CK(cudaGetLastError());
    contour_scatter_kernel<<<(unsigned int)((total + 127) / 128), 128, 0, w->stream>>>(a);
    CK(cudaGetLastError());
    w->ctx->launches += 5;
    CK(cudaMemcpyAsync(pin + st_cnt, base + o_cnt, 16, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaMemcpyAsync(pin + st_moff, a.mask_off, ((size_t)n_masks + 1) * 4, cudaMemcpyDeviceToHost, w->stream));
    if (labels) CK(cudaMemcpyAsync(pin + st_labels, a.labels, total * 4, cudaMemcpyDeviceToHost, w->stream));
    if (n_components) CK(cudaMemcpyAsync(pin + st_ncomp, a.ncomp, (size_t)n_masks * 4, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    unsigned int hc[4];
    memcpy(hc, pin + st_cnt, sizeof hc);
    if (hc[3]) return fail(FSE_ENOMEM, "fse_mask_outline: contour scratch overflow (%u scratch floats, %u contours, %u points)", hc[0], hc[1], hc[2]);
    const unsigned int nrec = hc[1], npts = hc[2];
    if ((int64_t)nrec > cap_contours || (int64_t)npts > cap_pts)
        return fail(FSE_ENOMEM, "fse_mask_outline: %u contours / %u points exceed the caller's capacity (%d / %d)", nrec, npts, cap_contours, cap_pts);
    const size_t v_off = up(sizeof(float) * 2 * (size_t)npts), var_bytes = v_off + sizeof(int) * ((size_t)nrec + 1);
    if (w->outline_pinned2_bytes < var_bytes) {
        if (w->outline_pinned2) cudaFreeHost(w->outline_pinned2);
        w->outline_pinned2 = nullptr;
        w->outline_pinned2_bytes = 0;
        CK(cudaMallocHost(&w->outline_pinned2, var_bytes + var_bytes / 2 + 4096));
        w->outline_pinned2_bytes = var_bytes + var_bytes / 2 + 4096;
    }
    char* pin2 = (char*)w->outline_pinned2;
    if (npts) CK(cudaMemcpyAsync(pin2, a.out_pts, sizeof(float) * 2 * (size_t)npts, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaMemcpyAsync(pin2 + v_off, a.pt_off, sizeof(int) * ((size_t)nrec + 1), cudaMemcpyDeviceToHost, w->stream));
    if (labels) memcpy(labels, pin + st_labels, total * 4);  // overlaps the two copies above
    if (n_components) memcpy(n_components, pin + st_ncomp, (size_t)n_masks * 4);
    memcpy(mask_off, pin + st_moff, ((size_t)n_masks + 1) * 4);
    CK(cudaStreamSynchronize(w->stream));
    if (npts) memcpy(pts, pin2, sizeof(float) * 2 * (size_t)npts);
    memcpy(pt_off, pin2 + v_off, sizeof(int) * ((size_t)nrec + 1));
    return FSE_OK;
}

extern "C" FSE_API int fse_solid_mask(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, uint8_t* mask) {
    if (!w || !mask || rw <= 0 || rh <= 0 || x < 0 || y < w->y_off || x + rw > w->W || y + rh > w->y_off + w->H)
        return fail(FSE_EINVAL, "fse_solid_mask: bad rect");
    CK(cudaSetDevice(w->ctx->device));
    CK(grow_scratch(w, (size_t)rw * rh));
    solid_mask_kernel<<<148 * 4, 256, 0, w->stream>>>(w->p.mat, w->ctx->d_tabs, w->W, x, y - w->y_off, rw, rh, (uint8_t*)w->outline_scratch);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    CK(cudaMemcpyAsync(mask, w->outline_scratch, (size_t)rw * rh, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return FSE_OK;
}

extern "C" FSE_API int fse_flood_component(fse_world* w, int32_t x, int32_t y, int32_t cap, int32_t* count, int32_t* bbox, int32_t* pixels) {
    if (!w || !count || cap < 1 || cap > FLOOD_HASH / 4) return fail(FSE_EINVAL, "fse_flood_component: bad argument (cap <= %d)", FLOOD_HASH / 4);
    *count = 0;
    if (x < 0 || y < w->y_off || x >= w->W || y >= w->y_off + w->H) return FSE_OK;  // getTile(x,y) out of range: TEST_SOLID but no flood
    CK(cudaSetDevice(w->ctx->device));
    CK(grow_scratch(w, sizeof(int) * (size_t)(cap + 8)));
    int* d_pix = (int*)w->outline_scratch;
    int* d_cnt = d_pix + cap + 4;
    flood_kernel<<<1, 256, 0, w->stream>>>(w->p.mat, w->ctx->d_tabs, w->W, w->H, x, y - w->y_off, cap, d_pix, d_cnt);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    int n = 0;
    CK(cudaMemcpyAsync(&n, d_cnt, sizeof n, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    *count = n;
    if (n == 0 || n > cap) return FSE_OK;
    std::vector<int> px(n);
    CK(cudaMemcpy(px.data(), d_pix, sizeof(int) * n, cudaMemcpyDeviceToHost));
    std::sort(px.begin(), px.end());
    int bb[4] = {w->W, w->Hglobal, 0, 0};
    for (int i = 0; i < n; i++) {
        const int cx = px[i] % w->W, cy = px[i] / w->W + w->y_off;
        bb[0] = std::min(bb[0], cx); bb[1] = std::min(bb[1], cy); bb[2] = std::max(bb[2], cx); bb[3] = std::max(bb[3], cy);
        if (pixels) pixels[i] = cx + cy * w->W;
    }
    if (bbox) memcpy(bbox, bb, sizeof bb);
    return FSE_OK;
}

// world::physicsCheck(x, y) (world.cpp:3330-3411) as one call: flood, then delete (1..10 cells) or cut out into a body (11..1000).
// d_tiles_out: where the cut-out tiles lie on the device afterwards (action 2).  strip_edges: fail when a component of <= 1000 cells touches a
// row edge of this rank's window that is not an edge of the world (the flood could not follow it there)
static int physics_check_local(fse_world* w, int32_t x, int32_t y, fse_physcheck_result* out, fse_cell* tiles_out, int32_t cap_tiles, fse_cell** d_tiles_out,
                               bool strip_edges) {
    memset(out, 0, sizeof *out);
    *d_tiles_out = nullptr;
    if (x < 0 || y < w->y_off || x >= w->W || y >= w->y_off + w->H) return FSE_OK;
    CK(cudaSetDevice(w->ctx->device));
    const int cap = 1000;
    // scratch: [pixels cap + 4][count 4] ... tiles behind them
    CK(grow_scratch(w, sizeof(int) * (size_t)(cap + 8)));
    int* d_pix = (int*)w->outline_scratch;
    int* d_cnt = d_pix + cap + 4;
    flood_kernel<<<1, 256, 0, w->stream>>>(w->p.mat, w->ctx->d_tabs, w->W, w->H, x, y - w->y_off, cap, d_pix, d_cnt);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    int n = 0;
    CK(cudaMemcpyAsync(&n, d_cnt, sizeof n, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    out->count = n;
    if (n == 0 || n > cap) return FSE_OK;
    std::vector<int> px(n);
    CK(cudaMemcpy(px.data(), d_pix, sizeof(int) * n, cudaMemcpyDeviceToHost));
    int bb[4] = {w->W, w->H, 0, 0};
    for (int i = 0; i < n; i++) {
        const int cx = px[i] % w->W, cy = px[i] / w->W;
        bb[0] = std::min(bb[0], cx); bb[1] = std::min(bb[1], cy); bb[2] = std::max(bb[2], cx); bb[3] = std::max(bb[3], cy);
    }
    out->x = bb[0]; out->y = bb[1] + w->y_off; out->w = bb[2] - bb[0] + 1; out->h = bb[3] - bb[1] + 1;
    if (strip_edges && ((w->y_off > 0 && bb[1] == 0) || (w->y_off + w->H < w->Hglobal && bb[3] == w->H - 1)))
        return fail(FSE_ESTATE, "fse_physics_check: the component at (%d, %d) reaches the edge of the rows rank %d holds (rows %d..%d); nothing was changed", x, y,
                    w->ctx->rank, w->y_off, w->y_off + w->H - 1);
    const bool body = n > 10;
    fse_cell* d_tiles = nullptr;
    const size_t area = (size_t)out->w * out->h;
    if (body) {
        if (!tiles_out || (int64_t)area > (int64_t)cap_tiles)
            return fail(FSE_EINVAL, "fse_physics_check: the component's %d x %d box needs %zu tiles (cap_tiles = %d); nothing was changed", out->w, out->h,
                        area, cap_tiles);
        // the pixel list must survive growing the scratch: put list and tiles into one allocation
        const size_t off = (sizeof(int) * (size_t)(cap + 8) + 63) / 64 * 64;
        if (w->outline_scratch_bytes < off + area * sizeof(fse_cell)) {
            CK(grow_scratch(w, off + area * sizeof(fse_cell)));
            d_pix = (int*)w->outline_scratch;
            CK(cudaMemcpyAsync(d_pix, px.data(), sizeof(int) * n, cudaMemcpyHostToDevice, w->stream));
        }
        d_tiles = (fse_cell*)((char*)w->outline_scratch + off);
        fill_air_tiles_kernel<<<(int)((area + 255) / 256), 256, 0, w->stream>>>(d_tiles, (int)area, (uint16_t)w->ctx->h_tabs.air);
        CK(cudaGetLastError());
        w->ctx->launches += 1;
    }
    physcheck_cut_kernel<<<(n + 255) / 256, 256, 0, w->stream>>>(w->p, w->ctx->d_tabs, w->W, d_pix, n, bb[0], bb[1], out->w, d_tiles);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    if (w->active_on)
        if (int r = fse_wake_rect(w, bb[0], bb[1], out->w, out->h)) return r;
    if (body) {
        CK(cudaMemcpyAsync(tiles_out, d_tiles, area * sizeof(fse_cell), cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
    }
    out->action = body ? 2 : 1;
    *d_tiles_out = d_tiles;
    return FSE_OK;
}

extern "C" FSE_API int fse_physics_check(fse_world* w, int32_t x, int32_t y, fse_physcheck_result* out, fse_cell* tiles_out, int32_t cap_tiles) {
    if (!w || !out) return fail(FSE_EINVAL, "fse_physics_check: null argument");
    fse_cell* d_tiles = nullptr;
    if (!(w->strip && w->ctx->nranks > 1)) return physics_check_local(w, x, y, out, tiles_out, cap_tiles, &d_tiles, false);
    // multi-rank strips: every rank makes the call; the owner of row y floods, deletes or cuts on its rows + ghost rows (refreshed first), the
    // result and the cut-out tiles are summed over the ranks (the others contribute zeros), the component's box travels to the neighbours
    memset(out, 0, sizeof *out);
    if (x < 0 || y < 0 || x >= w->W || y >= w->Hglobal) return FSE_OK;
    CK(cudaSetDevice(w->ctx->device));
    int runner = 0;
    if (int r = strip_runner_of_rows(w, y, y, nullptr, &runner)) return r;
    if (int r = strip_refresh(w, w->stream, STRIP_GHOST)) return r;
    int share[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (runner == w->ctx->rank) {
        const int rc = physics_check_local(w, x, y, out, tiles_out, cap_tiles, &d_tiles, true);
        share[0] = out->count; share[1] = out->action; share[2] = out->x; share[3] = out->y; share[4] = out->w; share[5] = out->h; share[6] = rc;
    }
    unsigned int* d_share = nullptr;  // (a small allocation of its own: the scratch holds the runner's tiles)
    CK(cudaMalloc((void**)&d_share, sizeof share));
    cudaError_t ce = cudaMemcpyAsync(d_share, share, sizeof share, cudaMemcpyHostToDevice, w->stream);
    int rr = ce == cudaSuccess ? strip_allreduce_u32(w, d_share, 8, w->stream) : fail(FSE_ECUDA, "fse_physics_check: %s", cudaGetErrorString(ce));
    if (!rr) {
        ce = cudaMemcpyAsync(share, d_share, sizeof share, cudaMemcpyDeviceToHost, w->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(w->stream);
        if (ce != cudaSuccess) rr = fail(FSE_ECUDA, "fse_physics_check: %s", cudaGetErrorString(ce));
    }
    cudaFree(d_share);
    if (rr) return rr;
    if (share[6]) {  // the runner refused (tile buffer too small, component at the edge of its rows): every rank reports it, nothing was changed
        out->count = share[0]; out->x = share[2]; out->y = share[3]; out->w = share[4]; out->h = share[5];
        if (runner == w->ctx->rank) return share[6];  // the thread-local message is the runner's own
        return fail(share[6], "fse_physics_check: rank %d, which holds row %d, refused the call (component of %d cells, box %d x %d)", runner, y, share[0], share[4], share[5]);
    }
    out->count = share[0]; out->action = share[1]; out->x = share[2]; out->y = share[3]; out->w = share[4]; out->h = share[5];
    if (out->action == 0) return FSE_OK;
    if (out->action == 2) {  // the new body's tiles on every rank
        const size_t area = (size_t)out->w * out->h, words = area * (sizeof(fse_cell) / 4);
        if (runner != w->ctx->rank) {
            if (!tiles_out || (int64_t)area > (int64_t)cap_tiles) return fail(FSE_EINVAL, "fse_physics_check: cap_tiles differs between the ranks");
            CK(grow_scratch(w, area * sizeof(fse_cell)));
            d_tiles = (fse_cell*)w->outline_scratch;
            CK(cudaMemsetAsync(d_tiles, 0, area * sizeof(fse_cell), w->stream));
        }
        if (int r = strip_allreduce_u32(w, (unsigned int*)d_tiles, words, w->stream)) return r;
        if (runner != w->ctx->rank) {
            CK(cudaMemcpyAsync(tiles_out, d_tiles, area * sizeof(fse_cell), cudaMemcpyDeviceToHost, w->stream));
            CK(cudaStreamSynchronize(w->stream));
        }
    }
    std::vector<int4> rect[4];
    strip_rects_of_box(w, runner, out->x, out->y, out->x + out->w - 1, out->y + out->h - 1, rect);
    return strip_push_rects(w, rect, w->stream);
}
