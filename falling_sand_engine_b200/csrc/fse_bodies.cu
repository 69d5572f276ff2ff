// fse_bodies.cu — rigid-body <-> grid bridge: the raster and erase loops of game::tick
// (reference: source/engine/game.cpp:1711-1815 and 1896-1983) as kernels.  Box2D itself stays on the host.
//
// The reference visits bodies in order and, inside a body, pixels tx-major; each pixel tries the five offsets
// (0,0),(1,0),(-1,0),(0,1),(0,-1) against the grid as the pixels before it left it.  That order is kept EXACTLY, in
// parallel: every pixel has a rank (its position in the reference's visiting order) and a footprint (its five candidate
// cells).  In a round, every pending pixel publishes its rank on its footprint cells with atomicMin; a pixel acts only if
// it holds the minimum on all five, i.e. when no earlier pixel that could still change one of its cells is pending.  Pixels
// that act in the same round have disjoint footprints, so the outcome equals the sequential loop.
#include <cmath>
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
};

// rank -> (body, tx, ty) in the reference's visiting order (tx outer, ty inner)
__device__ __forceinline__ bool locate(const BodyArgs& a, int r, int& b, int& tx, int& ty) {
    int lo = 0, hi = a.n_bodies - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (a.off[mid] <= r) lo = mid;
        else hi = mid - 1;
    }
    b = lo;
    const int o = r - a.off[b];
    tx = o / a.bh[b];
    ty = o % a.bh[b];
    return true;
}
__device__ __forceinline__ void world_pos(const BodyArgs& a, int b, int tx, int ty, int& wx, int& wy) {
    const float4 t = a.xf[b];  // x, y, s, c
    wx = (int)(tx * t.w - (ty + 1) * t.z + t.x);  // game.cpp:1763
    wy = (int)(tx * t.z + (ty + 1) * t.w + t.y);  // game.cpp:1764
}
__constant__ int c_dirs[5][2] = {{0, 0}, {1, 0}, {-1, 0}, {0, 1}, {0, -1}};  // game.cpp:1766

__global__ void bodies_init_kernel(BodyArgs a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_pixels) return;
    int b, tx, ty;
    locate(a, r, b, tx, ty);
    const bool live = a.tiles[a.off[b] + tx + ty * a.bw[b]].mat != a.air;
    a.pending[r] = live ? 1 : 0;
    if (live) atomicAdd(&a.counters[4], 1u);
}

__global__ void bodies_mark_kernel(BodyArgs a, int reset) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_pixels || !a.pending[r]) return;
    if (reset && a.pending[r] == 2) {  // acted in the round that just ended: free its footprint, then retire
        a.pending[r] = 0;
    }
    int b, tx, ty, wx, wy;
    locate(a, r, b, tx, ty);
    world_pos(a, b, tx, ty, wx, wy);
#pragma unroll
    for (int d = 0; d < 5; d++) {
        const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
        if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
        if (reset) a.claim[(size_t)y * a.W + x] = 0xffffffffu;
        else atomicMin(&a.claim[(size_t)y * a.W + x], (uint32_t)r);
    }
}

__device__ __forceinline__ void wake3x3(const BodyArgs& a, int x, int y) {
    if (!a.awake) return;
    const int ci = x / CHUNK, cj = y / CHUNK;
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
    const uint8_t f = a.p.flg[g];
    c.mat = a.p.mat[g];
    c.moved = (f & F_MOVED) ? 1 : 0;
    c.settle = a.p.stl[g];
    c.color = a.p.col[g];
    c.temp = a.p.tmp[g];
    c.dirty = 0;
    c._pad = 0;
    c.fluid = a.p.fl[g];
    c.fluid_diff = a.p.fd[g];
    return c;
}

template <bool ERASE>
__global__ void bodies_act_kernel(BodyArgs a) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_pixels || a.pending[r] != 1) return;
    int b, tx, ty, wx, wy;
    locate(a, r, b, tx, ty);
    world_pos(a, b, tx, ty, wx, wy);
    bool mine = true;
#pragma unroll
    for (int d = 0; d < 5; d++) {
        const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
        if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
        mine &= a.claim[(size_t)y * a.W + x] == (uint32_t)r;
    }
    if (!mine) {
        atomicAdd(&a.counters[5], 1u);  // still pending after this round
        return;
    }
    fse_cell* tile = &a.tiles[a.off[b] + tx + ty * a.bw[b]];
    const fse_cell rm = *tile;
    int4* fb = &a.feedback[b];
    if (!ERASE) {
        for (int d = 0; d < 5; d++) {  // game.cpp:1768-1811
            const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
            if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
            const size_t g = (size_t)y * a.W + x;
            const int ph = a.T->phys[a.p.mat[g]];
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
                    p.y = (float)(y - 3);
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
        for (int d = 0; d < 5; d++) {  // game.cpp:1918-1957
            const int x = wx + c_dirs[d][0], y = wy + c_dirs[d][1];
            if (x < 0 || y < 0 || x >= a.W || y >= a.H) continue;
            const size_t g = (size_t)y * a.W + x;
            if (a.p.mat[g] == rm.mat) {  // .id == rmat.id (any cell of the same material)
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
        if (!found && wx >= 0 && wy >= 0 && wx < a.W && wy < a.H && a.p.mat[(size_t)wy * a.W + wx] == a.air) {  // 1959-1965
            fse_cell nothing;
            memset(&nothing, 0, sizeof nothing);
            nothing.mat = (uint16_t)a.air;
            nothing.fluid = 2.0f;
            *tile = nothing;
            atomicAdd(&fb->w, 1);
        }
    }
    a.pending[r] = 2;  // acted; its footprint is released by the reset pass
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
    void release() {
        cudaFree(d_off); cudaFree(d_bw); cudaFree(d_bh); cudaFree(d_tiles); cudaFree(d_xf); cudaFree(d_pending); cudaFree(d_claim);
        cudaFree(d_feedback);
        d_off = d_bw = d_bh = nullptr; d_tiles = nullptr; d_xf = nullptr; d_pending = nullptr; d_claim = nullptr; d_feedback = nullptr;
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
    if (w->strip && w->ctx->nranks > 1) return fail(FSE_ESTATE, "fse_bodies_upload: bodies on multi-rank strip worlds are not implemented");
    CK(cudaSetDevice(w->ctx->device));
    if (!w->bodies) w->bodies = new fse_bodies();
    fse_bodies* B = w->bodies;
    uint32_t* keep_claim = B->d_claim;
    B->d_claim = nullptr;
    B->release();
    B->d_claim = keep_claim;
    B->n = n;
    B->off.assign(n + 1, 0);
    B->bw.resize(n);
    B->bh.resize(n);
    for (int i = 0; i < n; i++) {
        if (bodies[i].w < 1 || bodies[i].h < 1 || bodies[i].w > 1024 || bodies[i].h > 1024 || !bodies[i].tiles)
            return fail(FSE_EINVAL, "fse_bodies_upload: body %d is %dx%d", i, bodies[i].w, bodies[i].h);
        B->bw[i] = bodies[i].w;
        B->bh[i] = bodies[i].h;
        B->off[i + 1] = B->off[i] + bodies[i].w * bodies[i].h;
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
    if (!B->d_claim) {
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

template <bool ERASE>
static int run_bridge(fse_world* w, const fse_xform* xf, int32_t n, uint32_t tick, uint32_t seed, fse_body_feedback* out) {
    fse_bodies* B = w->bodies;
    if (!B || B->n != n) return fail(FSE_ESTATE, "bodies: %d transforms for %d uploaded bodies (fse_bodies_upload first)", n, B ? B->n : 0);
    if (n == 0) return FSE_OK;
    CK(cudaSetDevice(w->ctx->device));
    std::vector<float4> h(n);
    for (int i = 0; i < n; i++) h[i] = make_float4(xf[i].x, xf[i].y, std::sin(xf[i].angle), std::cos(xf[i].angle));  // game.cpp:1763-1764 on the host's libm
    CK(cudaMemcpyAsync(B->d_xf, h.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, w->stream));
    CK(cudaMemsetAsync(B->d_feedback, 0, sizeof(int4) * n, w->stream));
    CK(cudaMemsetAsync(w->pcount + 4, 0, 2 * sizeof(unsigned int), w->stream));
    BodyArgs a;
    a.p = w->p; a.T = w->ctx->d_tabs; a.W = w->W; a.H = w->H;
    a.n_bodies = n; a.off = B->d_off; a.bw = B->d_bw; a.bh = B->d_bh; a.tiles = B->d_tiles; a.xf = B->d_xf;
    a.pending = B->d_pending; a.claim = B->d_claim; a.counters = w->pcount; a.feedback = B->d_feedback;
    a.pbuf = w->pbuf; a.pcount = w->pcount; a.pcap = w->pcap;
    a.rkey = rng_key(seed, tick, 7); a.tick = tick; a.air = w->ctx->h_tabs.air;
    a.awake = w->active_on ? w->d_awake : nullptr; a.acols = w->acols; a.arows = w->arows;
    a.n_pixels = B->n_pixels;
    const int TB = 256, G = (B->n_pixels + TB - 1) / TB;
    bodies_init_kernel<<<G, TB, 0, w->stream>>>(a);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    unsigned int pending = 0;
    CK(cudaMemcpyAsync(&pending, w->pcount + 4, sizeof pending, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    int rounds = 0;
    while (pending > 0) {
        if (++rounds > 100000) return fail(FSE_ESTATE, "bodies: dependency rounds did not converge");
        CK(cudaMemsetAsync(w->pcount + 5, 0, sizeof(unsigned int), w->stream));
        bodies_mark_kernel<<<G, TB, 0, w->stream>>>(a, 0);
        bodies_act_kernel<ERASE><<<G, TB, 0, w->stream>>>(a);
        bodies_mark_kernel<<<G, TB, 0, w->stream>>>(a, 1);
        CK(cudaGetLastError());
        w->ctx->launches += 3;
        CK(cudaMemcpyAsync(&pending, w->pcount + 5, sizeof pending, cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
    }
    w->last_bridge_rounds = rounds;
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
