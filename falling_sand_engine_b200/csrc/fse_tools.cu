// fse_tools.cu — grid edits of the reference's interactive tools that run as kernels.
//   fse_explosion : world::explosion (world.cpp:2294-2332)
#include <cstring>

#include "fse_device.cuh"
#include "fse_internal.hpp"

namespace fse {

struct ExplArgs {
    Planes p;
    const DevTables* T;
    int W, H;
    int cx, cy, radius;
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
    const int x = a.cx - outer + ix, y = a.cy - outer + iy;
    if (x < 0 || y < 0 || x >= a.W || y >= a.H) return;  // getTile out of bounds is TEST_SOLID, setTile out of bounds is ignored
    const size_t g = (size_t)y * a.W + x;
    const uint8_t m = a.p.mat[g];
    const int ph = a.T->phys[m];
    if (ph == P_AIR) return;
    const int dx = x - a.cx, dy = y - a.cy;
    const int d2 = dx * dx + dy * dy;
    const bool inner = d2 < a.radius * a.radius;
    if (!inner && !(d2 < outer * outer && ph != P_SOLID)) return;
    const uint32_t cb = rng_cell(a.rkey, x, y);
    if (!inner || !(ph == P_SOLID || rng_draw(cb, S_EXPL_KEEP) % 10 < 6)) {
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

}  // namespace fse

using namespace fse;

extern "C" FSE_API int fse_explosion(fse_world* w, int32_t cx, int32_t cy, int32_t radius, uint32_t tick, uint32_t seed) {
    if (!w) return fail(FSE_EINVAL, "fse_explosion: null world");
    if (radius <= 0 || radius > 4096) return fail(FSE_EINVAL, "fse_explosion: radius %d out of range (1..4096)", radius);
    if (w->strip && w->ctx->nranks > 1) return fail(FSE_ESTATE, "fse_explosion: not available on multi-rank strips");
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
    a.cy = cy - w->y_off;
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
    const int x0 = cx - 2 * radius < 0 ? 0 : cx - 2 * radius, y0 = a.cy - 2 * radius < 0 ? 0 : a.cy - 2 * radius;
    const int x1 = cx + 2 * radius > w->W ? w->W : cx + 2 * radius, y1 = a.cy + 2 * radius > w->H ? w->H : a.cy + 2 * radius;
    if (x1 > x0 && y1 > y0) return fse_wake_rect(w, x0, y0, x1 - x0, y1 - y0);
    return FSE_OK;
}
