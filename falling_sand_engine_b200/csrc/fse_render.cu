// fse_render.cu — the streaming passes next to the tick; all plain HBM-bound kernels.
//   fse_render_dirty  : dirty cells -> RGBA texels of the main / fire / emission / flow planes + movingTiles (game.cpp:1994-2066)
//   fse_render_layers : dirty layer-2 / background cells -> their RGBA planes (game.cpp:2068-2126, 2154-2155)
//   fse_flow_*, fse_layer2_*, fse_background_* : the planes those loops read (world.hpp:112-119)
//   fse_scroll        : the grid shift of world::tickChunks when the camera moves (world.cpp:2454-2478, 2579-2582)
#include <cstring>
#include <utility>

#include "fse_device.cuh"
#include "fse_internal.hpp"

namespace fse {

struct RenderStatsDev {
    unsigned long long dirty, fire;
    unsigned long long moving[FSE_MAX_MATERIALS];
    unsigned long long flow;
    unsigned long long layer2, background;  // fse_render_layers
};

// byte order of the reference's texture arrays: [0] = r = color >> 16, [1] = g, [2] = b, [3] = alpha (game.cpp:2022-2025)
__device__ __forceinline__ uint32_t texel(uint32_t color, uint32_t alpha) {
    return ((color >> 16) & 0xffu) | (color & 0xff00u) | ((color & 0xffu) << 16) | (alpha << 24);
}

// One thread per 4 cells: the flag word decides whether anything has to be touched at all (1 B / cell for a clean world);
// a dirty cell costs mat + colour in and two or three texels out.
// flow (optional): flowX | flowY | prevFlowX | prevFlowY planes of n floats and the flow texture (game.cpp:2017-2018, 2040-2062)
__global__ void __launch_bounds__(256) render_dirty_kernel(Planes p, const DevTables* __restrict__ T, size_t n, uint32_t* __restrict__ px_main,
                                                           uint32_t* __restrict__ px_fire, uint32_t* __restrict__ px_emis, RenderStatsDev* st,
                                                           float* __restrict__ flow, uint32_t* __restrict__ px_flow) {
    __shared__ unsigned int hist[FSE_MAX_MATERIALS];
    __shared__ unsigned int s_dirty, s_fire, s_flow;
    for (int i = threadIdx.x; i < FSE_MAX_MATERIALS; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) s_dirty = s_fire = s_flow = 0;
    __syncthreads();
    const int fire_id = T->fire;
    const size_t n4 = n / 4;
    const uint32_t* flg32 = reinterpret_cast<const uint32_t*>(p.flg);
    const uint32_t* mat32 = reinterpret_cast<const uint32_t*>(p.mat);
    unsigned int my_dirty = 0, my_fire = 0, my_flow = 0;
    auto one = [&](size_t g, uint32_t m) {
        my_dirty++;
        atomicAdd(&hist[m], 1u);
        const int ph = T->phys[m];
        if (flow) {
            if (ph == P_SOUP) {  // the reference computes these in double (0.25, 0.5, 3.0, 4.0 are double literals) and stores floats
                const float fx = flow[g], fy = flow[n + g], px = flow[2 * n + g], py = flow[3 * n + g];
                const float nx = (float)((double)px + (double)__fsub_rn(fx, px) * 0.25);
                float ny = (float)((double)py + (double)__fsub_rn(fy, py) * 0.25);
                if (ny < 0) ny = (float)((double)ny * 0.5);
                const double k = 3.0 / (double)T->lut.iters[m] + 0.5;
                const double ay = fmin(fmax((double)ny * k / 4.0 + 0.5, 0.0), 1.0) * 255.0;
                const double ax = fmin(fmax((double)nx * k / 4.0 + 0.5, 0.0), 1.0) * 255.0;
                // bytes r = x, g = y, b = 0, a = 0xff; double -> u8 truncates
                px_flow[g] = (uint32_t)(uint8_t)(int)ax | ((uint32_t)(uint8_t)(int)ay << 8) | 0xff000000u;
                flow[2 * n + g] = nx;
                flow[3 * n + g] = ny;
                my_flow++;
            }
            flow[g] = 0.0f;
            flow[n + g] = 0.0f;
        }
        if (ph == P_AIR) {  // game.cpp:2000-2017: transparent black in all three planes
            px_main[g] = 0;
            px_fire[g] = 0;
            px_emis[g] = 0;
            return;
        }
        const uint32_t color = p.col[g], emit = T->emit_color[m], alpha = T->alpha[m];
        px_main[g] = texel(color, alpha);
        px_emis[g] = texel(emit, emit >> 24);
        if ((int)m == fire_id) {
            px_fire[g] = texel(color, alpha);
            my_fire++;
        }
    };
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t fw = flg32[i];
        if (!(fw & (0x01010101U * F_DIRTY))) continue;
        const uint32_t mw = mat32[i];
#pragma unroll
        for (int q = 0; q < 4; q++)
            if ((fw >> (8 * q)) & F_DIRTY) one(4 * i + q, (mw >> (8 * q)) & 0xffu);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) {  // cells beyond the last whole word
        const size_t g = 4 * n4 + threadIdx.x;
        if (p.flg[g] & F_DIRTY) one(g, p.mat[g]);
    }
    atomicAdd(&s_dirty, my_dirty);
    atomicAdd(&s_fire, my_fire);
    if (my_flow) atomicAdd(&s_flow, my_flow);
    __syncthreads();
    for (int i = threadIdx.x; i < FSE_MAX_MATERIALS; i += blockDim.x)
        if (hist[i]) atomicAdd(&st->moving[i], (unsigned long long)hist[i]);
    if (threadIdx.x == 0) {
        if (s_dirty) atomicAdd(&st->dirty, (unsigned long long)s_dirty);
        if (s_fire) atomicAdd(&st->fire, (unsigned long long)s_fire);
        if (s_flow) atomicAdd(&st->flow, (unsigned long long)s_flow);
    }
}

// game.cpp:2068-2126: dirty layer-2 cells -> RGBA (AIR: transparent, or the grey checker of globaldef.draw_background_grid), dirty
// background cells -> their ARGB colour; both dirty bits are cleared (game.cpp:2154-2155).  One thread per 4 cells, the dirty word decides.
__global__ void __launch_bounds__(256) render_layers_kernel(const uint8_t* __restrict__ l2_mat, const uint32_t* __restrict__ l2_col,
                                                            const uint32_t* __restrict__ bg, uint8_t* __restrict__ dirty,
                                                            const DevTables* __restrict__ T, size_t n, int grid, uint32_t* __restrict__ px_l2,
                                                            uint32_t* __restrict__ px_bg, RenderStatsDev* st) {
    unsigned int my_l2 = 0, my_bg = 0;
    auto one = [&](size_t g, uint32_t d) {
        if (d & 1u) {
            my_l2++;
            const uint32_t m = l2_mat[g];
            if (T->phys[m] == P_AIR) px_l2[g] = grid ? texel((g % 2) == 0 ? 0x888888u : 0x444444u, 0xffu) : 0u;
            else px_l2[g] = texel(l2_col[g], T->alpha[m]);
        }
        if (d & 2u) {
            my_bg++;
            const uint32_t c = bg[g];
            px_bg[g] = texel(c, c >> 24);
        }
    };
    const size_t n4 = n / 4;
    uint32_t* d32 = reinterpret_cast<uint32_t*>(dirty);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t dw = d32[i];
        if (!dw) continue;
#pragma unroll
        for (int q = 0; q < 4; q++)
            if ((dw >> (8 * q)) & 3u) one(4 * i + q, (dw >> (8 * q)) & 3u);
        d32[i] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) {
        const size_t g = 4 * n4 + threadIdx.x;
        if (dirty[g]) one(g, dirty[g]);
        dirty[g] = 0;
    }
    my_l2 = __reduce_add_sync(0xffffffffu, my_l2);
    my_bg = __reduce_add_sync(0xffffffffu, my_bg);
    if ((threadIdx.x & 31) == 0) {
        if (my_l2) atomicAdd(&st->layer2, (unsigned long long)my_l2);
        if (my_bg) atomicAdd(&st->background, (unsigned long long)my_bg);
    }
}

// layer-2 cells / background colours of a rect (setTileLayer2, the chunk merge world.cpp:2384-2389): AoS staging <-> planes
__global__ void layer2_write_kernel(uint8_t* l2_mat, uint32_t* l2_col, int16_t* l2_tmp, uint8_t* dirty, int W, int x0, int y0, int rw, int rh,
                                    const fse_cell* __restrict__ src) {
    const size_t n = (size_t)rw * rh;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t g = (size_t)(y0 + i / rw) * W + (x0 + i % rw);
        const fse_cell c = src[i];
        l2_mat[g] = (uint8_t)c.mat;
        l2_col[g] = c.color;
        l2_tmp[g] = c.temp;
        dirty[g] |= 1u;
    }
}
__global__ void layer2_read_kernel(const uint8_t* l2_mat, const uint32_t* l2_col, const int16_t* l2_tmp, const uint8_t* dirty, int W, int x0,
                                   int y0, int rw, int rh, fse_cell* __restrict__ dst) {
    const size_t n = (size_t)rw * rh;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t g = (size_t)(y0 + i / rw) * W + (x0 + i % rw);
        fse_cell c;
        memset(&c, 0, sizeof c);
        c.mat = l2_mat[g];
        c.color = l2_col[g];
        c.temp = l2_tmp[g];
        c.fluid = 2.0f;  // MaterialInstance default (game_datastruct.hpp:216); a chunk file keeps id / colour / temperature only
        c.dirty = dirty[g] & 1u;
        dst[i] = c;
    }
}
__global__ void background_write_kernel(uint32_t* bg, uint8_t* dirty, int W, int x0, int y0, int rw, int rh, const uint32_t* __restrict__ src) {
    const size_t n = (size_t)rw * rh;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t g = (size_t)(y0 + i / rw) * W + (x0 + i % rw);
        bg[g] = src[i];
        dirty[g] |= 2u;
    }
}
__global__ void fill_u8_kernel(uint8_t* p, size_t n, uint8_t v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// dst[Y][X] = src[Y - dy][X - dx] where that source cell exists, src[Y][X] otherwise (the reference's in-place loop leaves such cells
// alone); the dirty bit of the flag plane stays where it is (world::dirty is not shifted).  All seven planes in one pass from one set of
// planes into the other; fse_scroll swaps the sets afterwards, so the scroll moves 2 x 17 B per cell instead of 4 x.
__global__ void __launch_bounds__(256) scroll_planes_kernel(Planes dst, Planes src, int W, int H, int dx, int dy) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x;
    if (X >= W) return;
    const int sx = X - dx;
    const bool okx = sx >= 0 && sx < W;
    for (int Y = blockIdx.y; Y < H; Y += gridDim.y) {
        const int sy = Y - dy;
        const size_t g = (size_t)Y * W + X;
        const size_t f = (okx && sy >= 0 && sy < H) ? (size_t)sy * W + sx : g;
        dst.mat[g] = src.mat[f];
        dst.flg[g] = (uint8_t)((src.flg[f] & ~F_DIRTY) | (src.flg[g] & F_DIRTY));
        dst.stl[g] = src.stl[f];
        dst.tmp[g] = src.tmp[f];
        dst.col[g] = src.col[f];
        dst.fl[g] = src.fl[f];
        dst.fd[g] = src.fd[f];
    }
}
// The same for shifts by a multiple of 4 columns on a world whose width is one (every chunk-aligned camera move): a thread moves 4
// cells of every plane with one 4 / 8 / 16-byte access each — a group of 4 cells has its source entirely inside or outside the world.
__global__ void __launch_bounds__(256) scroll_planes4_kernel(Planes dst, Planes src, int W, int H, int dx, int dy) {
    const int X = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (X >= W) return;
    const int sx = X - dx;
    const bool okx = sx >= 0 && sx + 4 <= W;
    for (int Y = blockIdx.y; Y < H; Y += gridDim.y) {
        const int sy = Y - dy;
        const size_t g = ((size_t)Y * W + X) >> 2;
        const size_t f = (okx && sy >= 0 && sy < H) ? ((size_t)sy * W + sx) >> 2 : g;
        reinterpret_cast<uint32_t*>(dst.mat)[g] = reinterpret_cast<const uint32_t*>(src.mat)[f];
        const uint32_t dirty = 0x01010101u * F_DIRTY;
        reinterpret_cast<uint32_t*>(dst.flg)[g] = (reinterpret_cast<const uint32_t*>(src.flg)[f] & ~dirty) | (reinterpret_cast<const uint32_t*>(src.flg)[g] & dirty);
        reinterpret_cast<uint32_t*>(dst.stl)[g] = reinterpret_cast<const uint32_t*>(src.stl)[f];
        reinterpret_cast<uint2*>(dst.tmp)[g] = reinterpret_cast<const uint2*>(src.tmp)[f];
        reinterpret_cast<uint4*>(dst.col)[g] = reinterpret_cast<const uint4*>(src.col)[f];
        reinterpret_cast<uint4*>(dst.fl)[g] = reinterpret_cast<const uint4*>(src.fl)[f];
        reinterpret_cast<uint4*>(dst.fd)[g] = reinterpret_cast<const uint4*>(src.fd)[f];
    }
}
// Vertical shifts on multi-rank strips: the source row of a local row may belong to a neighbour.  Those rows arrive packed plane after
// plane (strip_shift_rows): ext.pl[q] = plane q of the global rows [ext.lo, ext.lo + ext.rows).  "Inside the world" is decided on
// GLOBAL rows, so every rank shifts exactly the cells the single world would.
struct ExtRows {
    const unsigned char* pl[7];
    int lo, rows;
};
__global__ void __launch_bounds__(256) scroll_planes_strip_kernel(Planes dst, Planes src, ExtRows ext, int W, int H, int dx, int dy, int y_off, int Hg) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x;
    if (X >= W) return;
    const int sx = X - dx;
    const bool okx = sx >= 0 && sx < W;
    for (int Y = blockIdx.y; Y < H; Y += gridDim.y) {
        const int sg = Y + y_off - dy;  // global source row
        const size_t g = (size_t)Y * W + X;
        const int sl = sg - y_off, se = sg - ext.lo;
        const uint8_t own_flg = src.flg[g];
        if (okx && sg >= 0 && sg < Hg && sl >= 0 && sl < H) {
            const size_t f = (size_t)sl * W + sx;
            dst.mat[g] = src.mat[f];
            dst.flg[g] = (uint8_t)((src.flg[f] & ~F_DIRTY) | (own_flg & F_DIRTY));
            dst.stl[g] = src.stl[f];
            dst.tmp[g] = src.tmp[f];
            dst.col[g] = src.col[f];
            dst.fl[g] = src.fl[f];
            dst.fd[g] = src.fd[f];
        } else if (okx && sg >= 0 && sg < Hg && se >= 0 && se < ext.rows) {
            const size_t f = (size_t)se * W + sx;
            dst.mat[g] = ext.pl[0][f];
            dst.flg[g] = (uint8_t)((ext.pl[1][f] & ~F_DIRTY) | (own_flg & F_DIRTY));
            dst.stl[g] = ext.pl[2][f];
            dst.tmp[g] = reinterpret_cast<const int16_t*>(ext.pl[3])[f];
            dst.col[g] = reinterpret_cast<const uint32_t*>(ext.pl[4])[f];
            dst.fl[g] = reinterpret_cast<const float*>(ext.pl[5])[f];
            dst.fd[g] = reinterpret_cast<const float*>(ext.pl[6])[f];
        } else {  // no source inside the world: the cell stays (the reference's in-place loop leaves it alone)
            dst.mat[g] = src.mat[g];
            dst.flg[g] = own_flg;
            dst.stl[g] = src.stl[g];
            dst.tmp[g] = src.tmp[g];
            dst.col[g] = src.col[g];
            dst.fl[g] = src.fl[g];
            dst.fd[g] = src.fd[g];
        }
    }
}
// background and real_layer2 move with the grid (world.cpp:2475-2476); layer2Dirty / backgroundDirty do not
__global__ void __launch_bounds__(256) scroll_layers_kernel(uint8_t* d_mat, int16_t* d_tmp, uint32_t* d_col, uint32_t* d_bg, const uint8_t* s_mat,
                                                            const int16_t* s_tmp, const uint32_t* s_col, const uint32_t* s_bg, int W, int H, int dx,
                                                            int dy) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x;
    if (X >= W) return;
    const int sx = X - dx;
    const bool okx = sx >= 0 && sx < W;
    for (int Y = blockIdx.y; Y < H; Y += gridDim.y) {
        const int sy = Y - dy;
        const size_t g = (size_t)Y * W + X;
        const size_t f = (okx && sy >= 0 && sy < H) ? (size_t)sy * W + sx : g;
        d_mat[g] = s_mat[f];
        d_tmp[g] = s_tmp[f];
        d_col[g] = s_col[f];
        d_bg[g] = s_bg[f];
    }
}

__global__ void scroll_particles_kernel(fse_particle* pbuf, const unsigned int* pcount, unsigned int pcap, float dx, float dy) {
    const unsigned int n = *pcount < pcap ? *pcount : pcap;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        pbuf[i].x += dx;  // world.cpp:2579-2582
        pbuf[i].y += dy;
    }
}

}  // namespace fse

using namespace fse;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return fail(FSE_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

static inline int grid_sm(fse_world* w) { return w->ctx->sm_count > 0 ? w->ctx->sm_count * 8 : 148 * 8; }  // 8 CTAs of 256 threads per SM, grid-stride

namespace fse {
void render_free(fse_world* w) {
    cudaFree(w->d_pixels); cudaFree(w->d_render_stats);
    cudaFree(w->p_shadow.mat); cudaFree(w->p_shadow.flg); cudaFree(w->p_shadow.stl); cudaFree(w->p_shadow.tmp); cudaFree(w->p_shadow.col);
    cudaFree(w->p_shadow.fl); cudaFree(w->p_shadow.fd);
    cudaFree(w->d_flow); cudaFree(w->d_pixels_flow);
    cudaFree(w->l2_mat); cudaFree(w->l2_tmp); cudaFree(w->l2_col); cudaFree(w->bg_col); cudaFree(w->layer_dirty); cudaFree(w->d_pixels_layers);
    cudaFree(w->l2_mat_s); cudaFree(w->l2_tmp_s); cudaFree(w->l2_col_s); cudaFree(w->bg_col_s);
}
}  // namespace fse

static int ensure_render_stats(fse_world* w) {
    if (!w->d_render_stats) CK(cudaMalloc(&w->d_render_stats, sizeof(RenderStatsDev)));
    return FSE_OK;
}

extern "C" FSE_API int fse_pixels_enable(fse_world* w, int enable) {
    if (!w) return fail(FSE_EINVAL, "fse_pixels_enable: null world");
    CK(cudaSetDevice(w->ctx->device));
    const size_t n = (size_t)w->W * w->H;
    if (enable && !w->d_pixels) {
        CK(cudaMalloc((void**)&w->d_pixels, 3 * n * sizeof(uint32_t)));
        CK(cudaMemsetAsync(w->d_pixels, 0, 3 * n * sizeof(uint32_t), w->stream));
        if (int r = ensure_render_stats(w)) return r;
    } else if (!enable && w->d_pixels) {
        CK(cudaStreamSynchronize(w->stream));
        cudaFree(w->d_pixels);
        w->d_pixels = nullptr;
    }
    return FSE_OK;
}

// world::flowX / flowY / prevFlowX / prevFlowY (world.hpp:116-119) and the flow texture.  While enabled, pass 1 of the tick adds every
// liquid flow it decides to the source cell's accumulators (world.cpp:1334, 1374, 1402, 1432) and fse_render_dirty consumes them.
extern "C" FSE_API int fse_flow_enable(fse_world* w, int enable) {
    if (!w) return fail(FSE_EINVAL, "fse_flow_enable: null world");
    CK(cudaSetDevice(w->ctx->device));
    const size_t n = (size_t)w->W * w->H;
    if (enable && !w->d_flow) {
        float* fl = nullptr;
        uint32_t* px = nullptr;
        if (cudaMalloc((void**)&fl, 4 * n * sizeof(float)) != cudaSuccess || cudaMalloc((void**)&px, n * sizeof(uint32_t)) != cudaSuccess) {
            cudaFree(fl);  // all or nothing: a world with accumulators but no texture would make fse_render_dirty write through null
            cudaGetLastError();
            return fail(FSE_ENOMEM, "fse_flow_enable: no memory for %zu cells of flow planes", n);
        }
        w->d_flow = fl;
        w->d_pixels_flow = px;
        CK(cudaMemsetAsync(w->d_flow, 0, 4 * n * sizeof(float), w->stream));
        CK(cudaMemsetAsync(w->d_pixels_flow, 0, n * sizeof(uint32_t), w->stream));
    } else if (!enable && w->d_flow) {
        CK(cudaStreamSynchronize(w->stream));
        cudaFree(w->d_flow);
        cudaFree(w->d_pixels_flow);
        w->d_flow = nullptr;
        w->d_pixels_flow = nullptr;
    }
    return FSE_OK;
}

extern "C" FSE_API int fse_flow_read(fse_world* w, int which, int32_t x, int32_t y, int32_t rw, int32_t rh, float* out) {
    if (!w || !out || which < 0 || which > 3) return fail(FSE_EINVAL, "fse_flow_read: bad argument");
    if (!w->d_flow) return fail(FSE_ESTATE, "fse_flow_read: fse_flow_enable first");
    if (int r = check_rect(w, x, y, rw, rh, "fse_flow_read")) return r;
    CK(cudaSetDevice(w->ctx->device));
    const float* src = w->d_flow + (size_t)which * w->W * w->H + (size_t)(y - w->y_off) * w->W + x;
    CK(cudaMemcpy2DAsync(out, (size_t)rw * 4, src, (size_t)w->W * 4, (size_t)rw * 4, rh, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return FSE_OK;
}

static uint32_t* pixel_plane(fse_world* w, int which) {
    const size_t n = (size_t)w->W * w->H;
    if (which >= 0 && which <= 2) return w->d_pixels ? w->d_pixels + (size_t)which * n : nullptr;
    if (which == FSE_PIXELS_FLOW) return w->d_pixels_flow;
    if (which == FSE_PIXELS_LAYER2 || which == FSE_PIXELS_BACKGROUND) return w->d_pixels_layers ? w->d_pixels_layers + (size_t)(which - FSE_PIXELS_LAYER2) * n : nullptr;
    return nullptr;
}

extern "C" FSE_API void* fse_pixels_device(fse_world* w, int which) {
    if (!w) return nullptr;
    return pixel_plane(w, which);
}

extern "C" FSE_API int fse_render_dirty(fse_world* w, fse_render_stats* out) {
    if (!w) return fail(FSE_EINVAL, "fse_render_dirty: null world");
    if (!w->d_pixels) return fail(FSE_ESTATE, "fse_render_dirty: fse_pixels_enable first");
    CK(cudaSetDevice(w->ctx->device));
    const size_t n = (size_t)w->W * w->H;
    CK(cudaMemsetAsync(w->d_render_stats, 0, sizeof(RenderStatsDev), w->stream));
    render_dirty_kernel<<<grid_sm(w), 256, 0, w->stream>>>(w->p, w->ctx->d_tabs, n, w->d_pixels, w->d_pixels + n, w->d_pixels + 2 * n,
                                                           (RenderStatsDev*)w->d_render_stats, w->d_flow, w->d_pixels_flow);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    if (out) {
        RenderStatsDev h;
        CK(cudaMemcpyAsync(&h, w->d_render_stats, sizeof h, cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
        out->dirty = (int64_t)h.dirty;
        out->fire = (int64_t)h.fire;
        for (int i = 0; i < FSE_MAX_MATERIALS; i++) out->moving[i] = (int64_t)h.moving[i];
        out->flow = (int64_t)h.flow;
    }
    return FSE_OK;
}

extern "C" FSE_API int fse_pixels_read(fse_world* w, int which, int32_t x, int32_t y, int32_t rw, int32_t rh, uint8_t* rgba) {
    if (!w || !rgba) return fail(FSE_EINVAL, "fse_pixels_read: bad argument");
    const uint32_t* plane = pixel_plane(w, which);
    if (!plane) return fail(which < 0 || which > FSE_PIXELS_BACKGROUND ? FSE_EINVAL : FSE_ESTATE, "fse_pixels_read: plane %d is not there (fse_pixels_enable / fse_flow_enable / fse_layer2_write_rect first)", which);
    const int yl = y - w->y_off;
    if (rw <= 0 || rh <= 0 || x < 0 || yl < 0 || x + rw > w->W || yl + rh > w->H) return fail(FSE_EINVAL, "fse_pixels_read: rect outside the world");
    CK(cudaSetDevice(w->ctx->device));
    const uint32_t* src = plane + (size_t)yl * w->W + x;
    CK(cudaMemcpy2DAsync(rgba, (size_t)rw * 4, src, (size_t)w->W * 4, (size_t)rw * 4, rh, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return FSE_OK;
}

// ---- layer 2 and background (world.hpp:112-113; setTileLayer2 world.cpp:1015-1019; chunk merge 2384-2389) ---------------------------------
static int ensure_layers(fse_world* w) {
    if (w->l2_mat) return FSE_OK;
    if (!w->ctx->has_materials) return fail(FSE_ESTATE, "layer planes: fse_materials_set first");
    const size_t n = (size_t)w->W * w->H;
    {   // all or nothing: l2_mat != null is what "the layer planes exist" means everywhere else
        void* q[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        const size_t bytes[6] = {n, n * 2, n * 4, n * 4, n, 2 * n * 4};
        for (int i = 0; i < 6; i++)
            if (cudaMalloc(&q[i], bytes[i]) != cudaSuccess) {
                for (int j = 0; j < i; j++) cudaFree(q[j]);
                cudaGetLastError();
                return fail(FSE_ENOMEM, "layer planes: no memory for %zu cells", n);
            }
        w->l2_mat = (uint8_t*)q[0]; w->l2_tmp = (int16_t*)q[1]; w->l2_col = (uint32_t*)q[2]; w->bg_col = (uint32_t*)q[3];
        w->layer_dirty = (uint8_t*)q[4]; w->d_pixels_layers = (uint32_t*)q[5];
    }
    fill_u8_kernel<<<grid_sm(w), 256, 0, w->stream>>>(w->l2_mat, n, (uint8_t)w->ctx->h_tabs.air);  // Tiles_NOTHING (world.cpp:117-119)
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(w->l2_tmp, 0, n * 2, w->stream));
    CK(cudaMemsetAsync(w->l2_col, 0, n * 4, w->stream));
    CK(cudaMemsetAsync(w->bg_col, 0, n * 4, w->stream));
    CK(cudaMemsetAsync(w->layer_dirty, 0, n, w->stream));
    CK(cudaMemsetAsync(w->d_pixels_layers, 0, 2 * n * 4, w->stream));
    w->ctx->launches += 1;
    return ensure_render_stats(w);
}

static const size_t LAYER_STAGE_CELLS = (size_t)16 << 20;

extern "C" FSE_API int fse_layer2_write_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, const fse_cell* cells) {
    if (!w || !cells) return fail(FSE_EINVAL, "fse_layer2_write_rect: null argument");
    if (int r = check_rect(w, x, y, rw, rh, "fse_layer2_write_rect")) return r;
    CK(cudaSetDevice(w->ctx->device));
    if (int r = ensure_layers(w)) return r;
    y -= w->y_off;
    const int nmat = w->ctx->h_tabs.n;
    int band = (int)(LAYER_STAGE_CELLS / (size_t)rw);
    band = band < 1 ? 1 : (band > rh ? rh : band);
    if (int r = ensure_stage(w, (size_t)band * rw)) return r;
    for (int yy = 0; yy < rh; yy += band) {
        const int hb = rh - yy < band ? rh - yy : band;
        const fse_cell* src = cells + (size_t)yy * rw;
        unsigned int worst = 0;
        for (size_t i = 0; i < (size_t)hb * rw; i++) worst = src[i].mat > worst ? src[i].mat : worst;
        if ((int)worst >= nmat) return fail(FSE_EINVAL, "fse_layer2_write_rect: cell material %u >= %d (material table size)", worst, nmat);
        CK(cudaMemcpyAsync(w->d_stage, src, (size_t)hb * rw * sizeof(fse_cell), cudaMemcpyHostToDevice, w->stream));
        layer2_write_kernel<<<grid_sm(w), 256, 0, w->stream>>>(w->l2_mat, w->l2_col, w->l2_tmp, w->layer_dirty, w->W, x, y + yy, rw, hb, w->d_stage);
        CK(cudaGetLastError());
        w->ctx->launches += 1;
        if (yy + band < rh) CK(cudaStreamSynchronize(w->stream));
    }
    return FSE_OK;
}

extern "C" FSE_API int fse_layer2_read_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, fse_cell* cells) {
    if (!w || !cells) return fail(FSE_EINVAL, "fse_layer2_read_rect: null argument");
    if (int r = check_rect(w, x, y, rw, rh, "fse_layer2_read_rect")) return r;
    CK(cudaSetDevice(w->ctx->device));
    if (int r = ensure_layers(w)) return r;
    y -= w->y_off;
    int band = (int)(LAYER_STAGE_CELLS / (size_t)rw);
    band = band < 1 ? 1 : (band > rh ? rh : band);
    if (int r = ensure_stage(w, (size_t)band * rw)) return r;
    for (int yy = 0; yy < rh; yy += band) {
        const int hb = rh - yy < band ? rh - yy : band;
        layer2_read_kernel<<<grid_sm(w), 256, 0, w->stream>>>(w->l2_mat, w->l2_col, w->l2_tmp, w->layer_dirty, w->W, x, y + yy, rw, hb, w->d_stage);
        CK(cudaGetLastError());
        w->ctx->launches += 1;
        CK(cudaMemcpyAsync(cells + (size_t)yy * rw, w->d_stage, (size_t)hb * rw * sizeof(fse_cell), cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
    }
    return FSE_OK;
}

extern "C" FSE_API int fse_background_write_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, const uint32_t* colors) {
    if (!w || !colors) return fail(FSE_EINVAL, "fse_background_write_rect: null argument");
    if (int r = check_rect(w, x, y, rw, rh, "fse_background_write_rect")) return r;
    CK(cudaSetDevice(w->ctx->device));
    if (int r = ensure_layers(w)) return r;
    y -= w->y_off;
    const size_t per = sizeof(fse_cell) / sizeof(uint32_t);  // colours that fit one staged cell
    int band = (int)(LAYER_STAGE_CELLS * per / (size_t)rw);
    band = band < 1 ? 1 : (band > rh ? rh : band);
    if (int r = ensure_stage(w, ((size_t)band * rw + per - 1) / per)) return r;
    for (int yy = 0; yy < rh; yy += band) {
        const int hb = rh - yy < band ? rh - yy : band;
        CK(cudaMemcpyAsync(w->d_stage, colors + (size_t)yy * rw, (size_t)hb * rw * 4, cudaMemcpyHostToDevice, w->stream));
        background_write_kernel<<<grid_sm(w), 256, 0, w->stream>>>(w->bg_col, w->layer_dirty, w->W, x, y + yy, rw, hb, (const uint32_t*)w->d_stage);
        CK(cudaGetLastError());
        w->ctx->launches += 1;
        if (yy + band < rh) CK(cudaStreamSynchronize(w->stream));
    }
    return FSE_OK;
}

extern "C" FSE_API int fse_background_read_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, uint32_t* colors) {
    if (!w || !colors) return fail(FSE_EINVAL, "fse_background_read_rect: null argument");
    if (int r = check_rect(w, x, y, rw, rh, "fse_background_read_rect")) return r;
    CK(cudaSetDevice(w->ctx->device));
    if (int r = ensure_layers(w)) return r;
    const uint32_t* src = w->bg_col + (size_t)(y - w->y_off) * w->W + x;
    CK(cudaMemcpy2DAsync(colors, (size_t)rw * 4, src, (size_t)w->W * 4, (size_t)rw * 4, rh, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return FSE_OK;
}

extern "C" FSE_API int fse_render_layers(fse_world* w, int draw_background_grid, int64_t* n_layer2, int64_t* n_background) {
    if (!w) return fail(FSE_EINVAL, "fse_render_layers: null world");
    if (n_layer2) *n_layer2 = 0;
    if (n_background) *n_background = 0;
    if (!w->l2_mat) return FSE_OK;  // nothing was ever written to either layer: nothing is dirty
    CK(cudaSetDevice(w->ctx->device));
    const size_t n = (size_t)w->W * w->H;
    RenderStatsDev* st = (RenderStatsDev*)w->d_render_stats;
    CK(cudaMemsetAsync(&st->layer2, 0, 2 * sizeof(unsigned long long), w->stream));
    render_layers_kernel<<<grid_sm(w), 256, 0, w->stream>>>(w->l2_mat, w->l2_col, w->bg_col, w->layer_dirty, w->ctx->d_tabs, n,
                                                            draw_background_grid ? 1 : 0, w->d_pixels_layers, w->d_pixels_layers + n, st);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    if (n_layer2 || n_background) {
        unsigned long long h[2];
        CK(cudaMemcpyAsync(h, &st->layer2, sizeof h, cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
        if (n_layer2) *n_layer2 = (int64_t)h[0];
        if (n_background) *n_background = (int64_t)h[1];
    }
    return FSE_OK;
}

extern "C" FSE_API int fse_scroll(fse_world* w, int32_t dx, int32_t dy) {
    if (!w) return fail(FSE_EINVAL, "fse_scroll: null world");
    // multi-rank strips: a horizontal shift never leaves a rank's rows (every rank makes the call and shifts what it holds, ghost rows
    // included); a vertical one would move rows between ranks and is left to the host (save, shift, reload)
    const bool strip_dy = w->strip && w->ctx->nranks > 1 && dy != 0;
    if (dx == 0 && dy == 0) return FSE_OK;
    CK(cudaSetDevice(w->ctx->device));
    const size_t n = (size_t)w->W * w->H;
    unsigned char* ext_rows = nullptr;
    int ext_lo = 0, ext_n = 0;
    if (strip_dy) {
        // A vertical shift moves rows between ranks: every rank sends the |dy| rows its neighbour's window slides onto and receives as
        // many from the other side (one message each way), then shifts what it holds, ghost rows included, reading the rows beyond its
        // own window from the message.  The same limit on every rank (so that all of them make the same NCCL calls): |dy| must stay
        // inside the shortest strip.
        if (w->l2_mat) return fail(FSE_ESTATE, "fse_scroll: vertical shifts of the layer-2 / background planes on multi-rank strips are not available");
        const int nr = w->ctx->nranks, nz = (w->Hglobal - 2 * CHUNK) / CHUNK;
        int min_own = w->Hglobal;
        for (int r = 0; r < nr; r++) {
            const int j0 = (int)((int64_t)nz * r / nr), j1 = (int)((int64_t)nz * (r + 1) / nr);
            const int lo = r == 0 ? 0 : CHUNK + CHUNK * j0, hi = r == nr - 1 ? w->Hglobal : CHUNK + CHUNK * j1;
            min_own = hi - lo < min_own ? hi - lo : min_own;
        }
        const int ghost = w->own_lo - w->y_off > 0 ? w->own_lo - w->y_off : (w->y_off + w->H) - w->own_hi;  // one of the two exists on a multi-rank strip
        const int nrows = dy > 0 ? dy : -dy;
        if (nrows > min_own - ghost)
            return fail(FSE_ESTATE, "fse_scroll: a vertical shift of %d rows on strips whose shortest one owns %d rows (at most %d)", dy, min_own, min_own - ghost);
        if (int r = strip_refresh(w, w->stream, ghost)) return r;  // every local row is the owner's before it becomes somebody's source
        int lo;
        if (dy > 0) {  // content moves down: the neighbour below needs the rows above its window
            lo = (w->own_hi - ghost - nrows) - w->y_off;
            ext_lo = w->y_off - nrows;
        } else {       // content moves up: the neighbour above needs the rows below its window
            lo = (w->own_lo + ghost) - w->y_off;
            ext_lo = w->y_off + w->H;
        }
        if (int r = strip_shift_rows(w, lo, lo + nrows, dy > 0, &ext_rows, w->stream)) return r;
        ext_n = ext_rows ? nrows : 0;
    }
    if (dx > -w->W && dx < w->W && (strip_dy || (dy > -w->H && dy < w->H))) {  // otherwise no cell has a source inside the world
        if (!w->p_shadow.mat) {  // all or nothing
            void* q[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            const size_t bytes[7] = {n, n, n, n * 2, n * 4, n * 4, n * 4};
            for (int i = 0; i < 7; i++)
                if (cudaMalloc(&q[i], bytes[i]) != cudaSuccess) {
                    for (int j = 0; j < i; j++) cudaFree(q[j]);
                    cudaGetLastError();
                    return fail(FSE_ENOMEM, "fse_scroll: no memory for the second plane set (%zu cells)", n);
                }
            w->p_shadow.mat = (uint8_t*)q[0]; w->p_shadow.flg = (uint8_t*)q[1]; w->p_shadow.stl = (uint8_t*)q[2]; w->p_shadow.tmp = (int16_t*)q[3];
            w->p_shadow.col = (uint32_t*)q[4]; w->p_shadow.fl = (float*)q[5]; w->p_shadow.fd = (float*)q[6];
        }
        dim3 grid((w->W + 255) / 256, w->H < 2048 ? w->H : 2048);
        if (strip_dy) {
            ExtRows ext;
            size_t so = 0;
            const size_t es[7] = {1, 1, 1, 2, 4, 4, 4};
            for (int q = 0; q < 7; q++) {
                ext.pl[q] = ext_rows ? ext_rows + so : nullptr;
                so += (size_t)ext_n * w->W * es[q];
            }
            ext.lo = ext_lo;
            ext.rows = ext_n;
            scroll_planes_strip_kernel<<<grid, 256, 0, w->stream>>>(w->p_shadow, w->p, ext, w->W, w->H, dx, dy, w->y_off, w->Hglobal);
        } else if (dx % 4 == 0 && w->W % 4 == 0) {
            dim3 grid4((w->W / 4 + 255) / 256, w->H < 4096 ? w->H : 4096);
            scroll_planes4_kernel<<<grid4, 256, 0, w->stream>>>(w->p_shadow, w->p, w->W, w->H, dx, dy);
        } else {
            scroll_planes_kernel<<<grid, 256, 0, w->stream>>>(w->p_shadow, w->p, w->W, w->H, dx, dy);
        }
        CK(cudaGetLastError());
        std::swap(w->p, w->p_shadow);  // kernels take the planes from the world at launch time; the stream orders them after the shift
        w->ctx->launches += 1;
        if (w->l2_mat) {
            if (!w->l2_mat_s) {
                void* q[4] = {nullptr, nullptr, nullptr, nullptr};
                const size_t bytes[4] = {n, n * 2, n * 4, n * 4};
                for (int i = 0; i < 4; i++)
                    if (cudaMalloc(&q[i], bytes[i]) != cudaSuccess) {
                        for (int j = 0; j < i; j++) cudaFree(q[j]);
                        cudaGetLastError();
                        return fail(FSE_ENOMEM, "fse_scroll: no memory for the second set of layer planes (%zu cells); the grid was shifted, the layers were not", n);
                    }
                w->l2_mat_s = (uint8_t*)q[0]; w->l2_tmp_s = (int16_t*)q[1]; w->l2_col_s = (uint32_t*)q[2]; w->bg_col_s = (uint32_t*)q[3];
            }
            scroll_layers_kernel<<<grid, 256, 0, w->stream>>>(w->l2_mat_s, w->l2_tmp_s, w->l2_col_s, w->bg_col_s, w->l2_mat, w->l2_tmp, w->l2_col,
                                                              w->bg_col, w->W, w->H, dx, dy);
            CK(cudaGetLastError());
            std::swap(w->l2_mat, w->l2_mat_s); std::swap(w->l2_tmp, w->l2_tmp_s); std::swap(w->l2_col, w->l2_col_s); std::swap(w->bg_col, w->bg_col_s);
            w->ctx->launches += 1;
        }
    }
    scroll_particles_kernel<<<256, 256, 0, w->stream>>>(w->pbuf, w->pcount, w->pcap, (float)dx, (float)dy);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    if (w->active_on) return fse_wake_rect(w, 0, 0, w->W, w->H);
    return FSE_OK;
}
