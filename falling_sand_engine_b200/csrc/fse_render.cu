// fse_render.cu — the two streaming passes next to the tick; both are plain HBM-bound kernels.
//   fse_render_dirty : dirty cells -> RGBA texels of the main / fire / emission planes + movingTiles (game.cpp:1994-2060)
//   fse_scroll       : the grid shift of world::tickChunks when the camera moves (world.cpp:2454-2478, 2579-2582)
#include <cstring>

#include "fse_device.cuh"
#include "fse_internal.hpp"

namespace fse {

struct RenderStatsDev {
    unsigned long long dirty, fire;
    unsigned long long moving[FSE_MAX_MATERIALS];
};

// byte order of the reference's texture arrays: [0] = r = color >> 16, [1] = g, [2] = b, [3] = alpha (game.cpp:2022-2025)
__device__ __forceinline__ uint32_t texel(uint32_t color, uint32_t alpha) {
    return ((color >> 16) & 0xffu) | (color & 0xff00u) | ((color & 0xffu) << 16) | (alpha << 24);
}

// One thread per 4 cells: the flag word decides whether anything has to be touched at all (1 B / cell for a clean world);
// a dirty cell costs mat + colour in and two or three texels out.
__global__ void __launch_bounds__(256) render_dirty_kernel(Planes p, const DevTables* __restrict__ T, size_t n, uint32_t* __restrict__ px_main,
                                                           uint32_t* __restrict__ px_fire, uint32_t* __restrict__ px_emis, RenderStatsDev* st) {
    __shared__ unsigned int hist[FSE_MAX_MATERIALS];
    __shared__ unsigned int s_dirty, s_fire;
    for (int i = threadIdx.x; i < FSE_MAX_MATERIALS; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) s_dirty = s_fire = 0;
    __syncthreads();
    const int fire_id = T->fire;
    const size_t n4 = n / 4;
    const uint32_t* flg32 = reinterpret_cast<const uint32_t*>(p.flg);
    const uint32_t* mat32 = reinterpret_cast<const uint32_t*>(p.mat);
    unsigned int my_dirty = 0, my_fire = 0;
    auto one = [&](size_t g, uint32_t m) {
        my_dirty++;
        atomicAdd(&hist[m], 1u);
        if (T->phys[m] == P_AIR) {  // game.cpp:2000-2017: transparent black in all three planes
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
    __syncthreads();
    for (int i = threadIdx.x; i < FSE_MAX_MATERIALS; i += blockDim.x)
        if (hist[i]) atomicAdd(&st->moving[i], (unsigned long long)hist[i]);
    if (threadIdx.x == 0) {
        if (s_dirty) atomicAdd(&st->dirty, (unsigned long long)s_dirty);
        if (s_fire) atomicAdd(&st->fire, (unsigned long long)s_fire);
    }
}

// plane[Y][X] = old[Y - dy][X - dx] where the source exists; other cells keep their content.  `old` is a copy of the plane.
template <typename E, bool FLAGS>
__global__ void __launch_bounds__(256) scroll_plane_kernel(E* __restrict__ plane, const E* __restrict__ old, int W, int H, int dx, int dy) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x;
    if (X >= W) return;
    const int sx = X - dx;
    if (sx < 0 || sx >= W) return;
    for (int Y = blockIdx.y; Y < H; Y += gridDim.y) {
        const int sy = Y - dy;
        if (sy < 0 || sy >= H) continue;
        E v = old[(size_t)sy * W + sx];
        if (FLAGS) v = (E)((v & ~(E)F_DIRTY) | (old[(size_t)Y * W + X] & (E)F_DIRTY));  // world::dirty is not shifted
        plane[(size_t)Y * W + X] = v;
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

extern "C" FSE_API int fse_pixels_enable(fse_world* w, int enable) {
    if (!w) return fail(FSE_EINVAL, "fse_pixels_enable: null world");
    CK(cudaSetDevice(w->ctx->device));
    const size_t n = (size_t)w->W * w->H;
    if (enable && !w->d_pixels) {
        CK(cudaMalloc((void**)&w->d_pixels, 3 * n * sizeof(uint32_t)));
        CK(cudaMemsetAsync(w->d_pixels, 0, 3 * n * sizeof(uint32_t), w->stream));
        if (!w->d_render_stats) CK(cudaMalloc(&w->d_render_stats, sizeof(RenderStatsDev)));
    } else if (!enable && w->d_pixels) {
        CK(cudaStreamSynchronize(w->stream));
        cudaFree(w->d_pixels);
        w->d_pixels = nullptr;
    }
    return FSE_OK;
}

extern "C" FSE_API void* fse_pixels_device(fse_world* w, int which) {
    if (!w || !w->d_pixels || which < 0 || which > 2) return nullptr;
    return w->d_pixels + (size_t)which * w->W * w->H;
}

extern "C" FSE_API int fse_render_dirty(fse_world* w, fse_render_stats* out) {
    if (!w) return fail(FSE_EINVAL, "fse_render_dirty: null world");
    if (!w->d_pixels) return fail(FSE_ESTATE, "fse_render_dirty: fse_pixels_enable first");
    CK(cudaSetDevice(w->ctx->device));
    const size_t n = (size_t)w->W * w->H;
    CK(cudaMemsetAsync(w->d_render_stats, 0, sizeof(RenderStatsDev), w->stream));
    const int grid = w->ctx->sm_count > 0 ? w->ctx->sm_count * 8 : 148 * 8;  // 8 resident CTAs of 256 threads per SM, grid-stride
    render_dirty_kernel<<<grid, 256, 0, w->stream>>>(w->p, w->ctx->d_tabs, n, w->d_pixels, w->d_pixels + n, w->d_pixels + 2 * n,
                                                     (RenderStatsDev*)w->d_render_stats);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    if (out) {
        RenderStatsDev h;
        CK(cudaMemcpyAsync(&h, w->d_render_stats, sizeof h, cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
        out->dirty = (int64_t)h.dirty;
        out->fire = (int64_t)h.fire;
        for (int i = 0; i < FSE_MAX_MATERIALS; i++) out->moving[i] = (int64_t)h.moving[i];
    }
    return FSE_OK;
}

extern "C" FSE_API int fse_pixels_read(fse_world* w, int which, int32_t x, int32_t y, int32_t rw, int32_t rh, uint8_t* rgba) {
    if (!w || !rgba || which < 0 || which > 2) return fail(FSE_EINVAL, "fse_pixels_read: bad argument");
    if (!w->d_pixels) return fail(FSE_ESTATE, "fse_pixels_read: fse_pixels_enable first");
    const int yl = y - w->y_off;
    if (rw <= 0 || rh <= 0 || x < 0 || yl < 0 || x + rw > w->W || yl + rh > w->H) return fail(FSE_EINVAL, "fse_pixels_read: rect outside the world");
    CK(cudaSetDevice(w->ctx->device));
    const uint32_t* src = w->d_pixels + (size_t)which * w->W * w->H + (size_t)yl * w->W + x;
    CK(cudaMemcpy2DAsync(rgba, (size_t)rw * 4, src, (size_t)w->W * 4, (size_t)rw * 4, rh, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return FSE_OK;
}

template <typename E, bool FLAGS>
static cudaError_t scroll_plane(fse_world* w, E* plane, int dx, int dy) {
    const size_t bytes = (size_t)w->W * w->H * sizeof(E);
    cudaError_t e = cudaMemcpyAsync(w->scroll_scratch, plane, bytes, cudaMemcpyDeviceToDevice, w->stream);
    if (e != cudaSuccess) return e;
    dim3 grid((w->W + 255) / 256, w->H < 2048 ? w->H : 2048);
    scroll_plane_kernel<E, FLAGS><<<grid, 256, 0, w->stream>>>(plane, (const E*)w->scroll_scratch, w->W, w->H, dx, dy);
    return cudaGetLastError();
}

extern "C" FSE_API int fse_scroll(fse_world* w, int32_t dx, int32_t dy) {
    if (!w) return fail(FSE_EINVAL, "fse_scroll: null world");
    if (w->strip && w->ctx->nranks > 1) return fail(FSE_ESTATE, "fse_scroll: not available on multi-rank strips");
    if (dx == 0 && dy == 0) return FSE_OK;
    CK(cudaSetDevice(w->ctx->device));
    const size_t need = (size_t)w->W * w->H * sizeof(uint32_t);
    if (w->scroll_scratch_bytes < need) {
        cudaFree(w->scroll_scratch);
        w->scroll_scratch = nullptr;
        w->scroll_scratch_bytes = 0;
        CK(cudaMalloc(&w->scroll_scratch, need));
        w->scroll_scratch_bytes = need;
    }
    if (dx > -w->W && dx < w->W && dy > -w->H && dy < w->H) {  // otherwise no cell has a source inside the world
        CK((scroll_plane<uint8_t, false>(w, w->p.mat, dx, dy)));
        CK((scroll_plane<uint8_t, true>(w, w->p.flg, dx, dy)));
        CK((scroll_plane<uint8_t, false>(w, w->p.stl, dx, dy)));
        CK((scroll_plane<uint16_t, false>(w, reinterpret_cast<uint16_t*>(w->p.tmp), dx, dy)));  // bit copies
        CK((scroll_plane<uint32_t, false>(w, w->p.col, dx, dy)));
        CK((scroll_plane<uint32_t, false>(w, reinterpret_cast<uint32_t*>(w->p.fl), dx, dy)));
        CK((scroll_plane<uint32_t, false>(w, reinterpret_cast<uint32_t*>(w->p.fd), dx, dy)));
        w->ctx->launches += 7;
    }
    scroll_particles_kernel<<<256, 256, 0, w->stream>>>(w->pbuf, w->pcount, w->pcap, (float)dx, (float)dy);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    if (w->active_on) return fse_wake_rect(w, 0, 0, w->W, w->H);
    return FSE_OK;
}
