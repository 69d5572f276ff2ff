// fse_aux.cu — small kernels around the tick: AoS<->SoA rect transfer (world::getTile/setTile, frame()
// merge and chunkSaveCache, world.cpp:999-1008, 2374-2391, 2780-2792), world statistics, dirty clear,
// and world::tickTemperature() (world.cpp:1950-2004).
#include "fse_device.cuh"

namespace fse {

// ---- AoS rect -> planes ------------------------------------------------------------------------------
__global__ void write_rect_kernel(Planes p, int W, int x0, int y0, int rw, int rh, const fse_cell* __restrict__ src) {
    const size_t n = (size_t)rw * rh;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % rw), y = (int)(i / rw);
        size_t g = (size_t)(y0 + y) * W + (x0 + x);
        fse_cell c = src[i];
        p.mat[g] = (uint8_t)c.mat;
        p.flg[g] = (uint8_t)((c.moved ? F_MOVED : 0) | (c.dirty ? F_DIRTY : 0));
        p.stl[g] = c.settle;
        p.tmp[g] = c.temp;
        p.col[g] = c.color;
        p.fl[g] = c.fluid;
        p.fd[g] = c.fluid_diff;
    }
}

__global__ void read_rect_kernel(Planes p, int W, int x0, int y0, int rw, int rh, fse_cell* __restrict__ dst) {
    const size_t n = (size_t)rw * rh;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % rw), y = (int)(i / rw);
        size_t g = (size_t)(y0 + y) * W + (x0 + x);
        fse_cell c;
        uint8_t f = p.flg[g];
        c.mat = p.mat[g];
        c.moved = (f & F_MOVED) ? 1 : 0;
        c.settle = p.stl[g];
        c.color = p.col[g];
        c.temp = p.tmp[g];
        c.dirty = (f & F_DIRTY) ? 1 : 0;
        c._pad = 0;
        c.fluid = p.fl[g];
        c.fluid_diff = p.fd[g];
        dst[i] = c;
    }
}

// fresh world: every cell is a default MaterialInstance() = Tiles_NOTHING (gds.cpp:310-312)
__global__ void fill_air_kernel(Planes p, size_t n, uint8_t air) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        p.mat[i] = air;
        p.flg[i] = 0;
        p.stl[i] = 0;
        p.tmp[i] = 0;
        p.col[i] = 0;
        p.fl[i] = 2.0f;
        p.fd[i] = 0.0f;
    }
}

// memset(dirty, false, W*H) (game.cpp:2153): clear bit1 of the flag plane, 16 bytes per thread
__global__ void clear_dirty_kernel(uint4* flg, size_t n16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = flg[i];
        const uint32_t m = ~(0x01010101U * F_DIRTY);
        v.x &= m; v.y &= m; v.z &= m; v.w &= m;
        flg[i] = v;
    }
}

// ---- statistics (movingTiles-style histogram, game.cpp:1991-2000, + parity hash) ----------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

struct DevStats {
    unsigned long long hash;
    unsigned long long count[FSE_MAX_MATERIALS];
    double mass[FSE_MAX_MATERIALS];
    unsigned long long n_dirty, n_moved;
};

__global__ void stats_kernel(Planes p, int W, int x0, int y0, int rw, int rh, int yoff, const DevTables* T, DevStats* out) {
    __shared__ unsigned int cnt[FSE_MAX_MATERIALS];
    __shared__ double mass[FSE_MAX_MATERIALS];
    __shared__ unsigned long long sh_hash, sh_dirty, sh_moved;
    for (int i = threadIdx.x; i < FSE_MAX_MATERIALS; i += blockDim.x) {
        cnt[i] = 0;
        mass[i] = 0.0;
    }
    if (threadIdx.x == 0) sh_hash = sh_dirty = sh_moved = 0;
    __syncthreads();
    unsigned long long h = 0, nd = 0, nm = 0;
    const size_t n = (size_t)rw * rh;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int x = x0 + (int)(i % rw), y = y0 + (int)(i / rw);
        size_t g = (size_t)y * W + x;
        uint8_t m = p.mat[g], f = p.flg[g], st = p.stl[g];
        uint32_t col = p.col[g];
        int16_t tmp = p.tmp[g];
        float fl = p.fl[g], fd = p.fd[g];
        uint64_t a = ((uint64_t)(uint32_t)x << 32) | (uint32_t)(y + yoff);  // hash on global coordinates
        uint64_t b = ((uint64_t)m << 48) | ((uint64_t)(f & F_MOVED) << 40) | ((uint64_t)st << 32) | col;
        uint64_t d = ((uint64_t)(uint16_t)tmp << 32) | __float_as_uint(fl);
        uint64_t hh = mix64(a + 0x9E3779B97F4A7C15ULL);
        hh = mix64(hh ^ b);
        hh = mix64(hh ^ d);
        hh = mix64(hh ^ (uint64_t)__float_as_uint(fd));
        h += hh;
        atomicAdd(&cnt[m], 1u);
        if (T->phys[m] == P_SOUP) atomicAdd(&mass[m], (double)fl + (double)fd);
        nd += (f & F_DIRTY) ? 1 : 0;
        nm += (f & F_MOVED) ? 1 : 0;
    }
    atomicAdd(&sh_hash, h);
    atomicAdd(&sh_dirty, nd);
    atomicAdd(&sh_moved, nm);
    __syncthreads();
    for (int i = threadIdx.x; i < FSE_MAX_MATERIALS; i += blockDim.x) {
        if (cnt[i]) atomicAdd(&out->count[i], (unsigned long long)cnt[i]);
        if (mass[i] != 0.0) atomicAdd(&out->mass[i], mass[i]);
    }
    if (threadIdx.x == 0) {
        atomicAdd(&out->hash, sh_hash);
        atomicAdd(&out->n_dirty, sh_dirty);
        atomicAdd(&out->n_moved, sh_moved);
    }
}

// ---- world::tickTemperature() (world.cpp:1950-2004) -----------------------------------------------------
// Jacobi 3x3 stencil on the i16 temperature plane: order-independent, so bit-exact against the reference
// order.  One thread per cell of a 32x8 tile staged (with a 1-cell halo) in shared memory; the new
// temperatures go to a second plane that then becomes the temperature plane (the reference's newTemps[] two-loop form; only the
// cells around the zone are copied).
constexpr int TT_W = 128, TT_H = 16, TT_P = TT_W + 4;  // outputs per block; row pitch of the shared-memory tile (130 used)

// Each loaded cell's contribution is computed once — factor = float(abs(t) / 64) * conductionOther(mat), t * factor — and every
// output cell then adds the nine (t * factor, factor) pairs of its neighbourhood in the reference's order (xa outer, ya inner,
// world.cpp:1976-1984).  A neighbour with t == 0, which the reference skips, contributes (+-0, 0): adding it leaves v and n
// unchanged bit for bit, so no branch is needed.  A thread walks 8 rows of one column with the 3 x 3 window in registers.
__global__ void __launch_bounds__(256) temperature_kernel(const uint8_t* __restrict__ mat, const int16_t* __restrict__ tmp,
                                                          int16_t* __restrict__ out, int W, int H, int zx, int zy, int zw, int zh,
                                                          const DevTables* __restrict__ T, uint8_t* awake, int acols, int yoff) {
    __shared__ float2 sc[TT_H + 2][TT_P];   // (t * factor, factor)
    __shared__ int16_t st[TT_H + 2][TT_P];
    __shared__ uint8_t sm[TT_H + 2][TT_P];
    __shared__ float condO[FSE_MAX_MATERIALS];
    __shared__ float condS[FSE_MAX_MATERIALS];
    __shared__ uint32_t addT[FSE_MAX_MATERIALS];
    const int tid = threadIdx.x;
    for (int i = tid; i < FSE_MAX_MATERIALS; i += blockDim.x) {
        condO[i] = T->cond_other[i];
        condS[i] = T->cond_self[i];
        addT[i] = T->add_temp[i];
    }
    __syncthreads();
    const int bx = zx + blockIdx.x * TT_W, by = zy + blockIdx.y * TT_H;
    auto load = [&](int ly, int lx) {  // tile cell (ly, lx) <- world cell (by + ly - 1, bx + lx - 1)
        const int gx = bx + lx - 1, gy = by + ly - 1;
        int t = 0;
        uint8_t m = 0;
        if (gx < W && gy < H) {  // tiles may hang over the zone; those cells feed no output inside it
            const size_t g = (size_t)gy * W + gx;
            t = tmp[g];
            m = mat[g];
        }
        const float factor = __fmul_rn((float)(abs(t) / 64), condO[m]);
        st[ly][lx] = (int16_t)t;
        sm[ly][lx] = m;
        sc[ly][lx] = make_float2(__fmul_rn((float)t, factor), factor);
    };
    for (int r = tid >> 7; r < TT_H + 2; r += 2) load(r, 1 + (tid & 127));      // the 128 middle columns, two rows per sweep
    if (tid < 2 * (TT_H + 2)) load(tid >> 1, (tid & 1) ? TT_W + 1 : 0);          // the two halo columns
    __syncthreads();
    const int lx = tid & 127, ly0 = (tid >> 7) * (TT_H / 2);
    const int x = bx + lx;
    if (x >= zx + zw) return;
    float2 w0[3], w1[3], w2[3];  // window rows y-1, y, y+1; index = xa + 1
#pragma unroll
    for (int q = 0; q < 3; q++) {
        w0[q] = sc[ly0][lx + q];
        w1[q] = sc[ly0 + 1][lx + q];
    }
#pragma unroll
    for (int k = 0; k < TT_H / 2; k++) {
        const int ly = ly0 + k, y = by + ly;
#pragma unroll
        for (int q = 0; q < 3; q++) w2[q] = sc[ly + 2][lx + q];
        if (y < zy + zh) {
            float n = 0.01f;
            float v = 0.0f;
#pragma unroll
            for (int q = 0; q < 3; q++) {  // xa = q - 1; ya = -1, 0, 1
                v = __fadd_rn(v, w0[q].x); n = __fadd_rn(n, w0[q].y);
                v = __fadd_rn(v, w1[q].x); n = __fadd_rn(n, w1[q].y);
                v = __fadd_rn(v, w2[q].x); n = __fadd_rn(n, w2[q].y);
            }
            const int t0 = st[ly + 1][lx + 1];
            const uint8_t m0 = sm[ly + 1][lx + 1];
            int nt;
            if (v != 0.0f) {
                float cs = condS[m0];
                float a = __fmul_rn(__fdiv_rn(v, n), cs);
                float b = __fmul_rn((float)t0, __fsub_rn(1.0f, cs));
                float r = __fadd_rn(__fadd_rn((float)addT[m0], a), b);
                nt = (int)r;  // i32 newTemps[] (world.hpp), truncation toward zero
            } else {
                nt = (int)(addT[m0] + (uint32_t)t0);  // unsigned wrap, as u32 + i16 in the reference
            }
            out[(size_t)y * W + x] = (int16_t)nt;  // real_tiles[].temperature = newTemps[] (i16 wrap)
            // active-region tracking: a temperature change can arm a reaction (world.cpp:1181-1204) in a sleeping chunk
            if (awake && (int16_t)nt != (int16_t)t0 && (T->lut.mflags[m0] & MF_REACT)) awake[((y + yoff) / CHUNK) * acols + x / CHUNK] = 1;
        }
#pragma unroll
        for (int q = 0; q < 3; q++) {
            w0[q] = w1[q];
            w1[q] = w2[q];
        }
    }
}

// The cells OUTSIDE the zone (top and bottom bands, left and right margins) copied from the old plane into the new one, so the two
// planes can simply trade places afterwards: the zone itself is never copied back.
__global__ void copy_outside_zone_i16_kernel(const int16_t* __restrict__ src, int16_t* __restrict__ dst, int W, int H, int zx, int zy, int zw,
                                             int zh) {
    const size_t n_top = (size_t)zy * W, n_bot = (size_t)(H - zy - zh) * W, n_left = (size_t)zh * zx, n_right = (size_t)zh * (W - zx - zw);
    const size_t n = n_top + n_bot + n_left + n_right;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t g;
        if (i < n_top) g = i;
        else if (i < n_top + n_bot) g = (size_t)(zy + zh) * W + (i - n_top);
        else if (i < n_top + n_bot + n_left) {
            const size_t k = i - n_top - n_bot;
            g = (size_t)(zy + k / zx) * W + k % zx;
        } else {
            const size_t k = i - n_top - n_bot - n_left;
            const int mr = W - zx - zw;
            g = (size_t)(zy + k / mr) * W + zx + zw + k % mr;
        }
        dst[g] = src[g];
    }
}

// ---- host launchers --------------------------------------------------------------------------------
static inline int grid_for(size_t n, int block) {
    size_t g = (n + block - 1) / block;
    if (g > 148 * 16) g = 148 * 16;  // grid-stride: a multiple of the SM count
    if (g < 1) g = 1;
    return (int)g;
}

cudaError_t launch_write_rect(Planes p, int W, int x0, int y0, int rw, int rh, const fse_cell* src, cudaStream_t s) {
    write_rect_kernel<<<grid_for((size_t)rw * rh, 256), 256, 0, s>>>(p, W, x0, y0, rw, rh, src);
    return cudaGetLastError();
}
cudaError_t launch_read_rect(Planes p, int W, int x0, int y0, int rw, int rh, fse_cell* dst, cudaStream_t s) {
    read_rect_kernel<<<grid_for((size_t)rw * rh, 256), 256, 0, s>>>(p, W, x0, y0, rw, rh, dst);
    return cudaGetLastError();
}
cudaError_t launch_fill_air(Planes p, size_t n, uint8_t air, cudaStream_t s) {
    fill_air_kernel<<<grid_for(n, 256), 256, 0, s>>>(p, n, air);
    return cudaGetLastError();
}
cudaError_t launch_clear_dirty(Planes p, size_t n, cudaStream_t s) {
    clear_dirty_kernel<<<grid_for(n / 16, 256), 256, 0, s>>>(reinterpret_cast<uint4*>(p.flg), n / 16);
    return cudaGetLastError();
}
cudaError_t launch_stats(Planes p, int W, int x0, int y0, int rw, int rh, int yoff, const DevTables* T, void* out, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(DevStats), s);
    if (e != cudaSuccess) return e;
    stats_kernel<<<grid_for((size_t)rw * rh, 256), 256, 0, s>>>(p, W, x0, y0, rw, rh, yoff, T, (DevStats*)out);
    return cudaGetLastError();
}
size_t dev_stats_bytes() { return sizeof(DevStats); }

// new temperatures of the zone -> `scratch`, the cells around the zone copied over; the caller swaps p.tmp and scratch afterwards
cudaError_t launch_temperature(Planes p, int16_t* scratch, int W, int H, int zx, int zy, int zw, int zh, const DevTables* T, uint8_t* awake,
                               int acols, int yoff, cudaStream_t s) {
    dim3 grid((zw + TT_W - 1) / TT_W, (zh + TT_H - 1) / TT_H);
    temperature_kernel<<<grid, 256, 0, s>>>(p.mat, p.tmp, scratch, W, H, zx, zy, zw, zh, T, awake, acols, yoff);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t outside = (size_t)W * H - (size_t)zw * zh;
    if (outside) copy_outside_zone_i16_kernel<<<grid_for(outside, 256), 256, 0, s>>>(p.tmp, scratch, W, H, zx, zy, zw, zh);
    return cudaGetLastError();
}

}  // namespace fse
