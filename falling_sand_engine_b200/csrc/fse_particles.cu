// fse_particles.cu — loose particles: world::tickCells() (reference: source/engine/world.cpp:2030-2195) on the GPU.
//
// Schedule.  The reference walks its std::vector<CellData*> in order and every deposit is seen by
// the particles after it.  Here (1) every particle is integrated — target attraction, v += a, sub-stepped collision,
// object pass-through state machine — against the grid as it was when the call started, one thread per particle;
// (2) particles that hit something resolve their deposit in rounds: each proposes its start cell or the first free
// cell of the reference's 32x32 square spiral (or a same-material liquid cell to merge into), an open-addressing
// claim table keeps the LOWEST particle id per contested cell, winners write the cell, losers look again next round;
// (3) survivors are compacted into a fresh buffer.  Everything is keyed on particle ids, so the result does not
// depend on thread scheduling or on the order of the pool.
#include "fse_internal.hpp"

namespace fse {

struct PState {
    fse_particle adv;   // state after integration
    long long cand;     // proposed cell (x + y*W) or -1
    int lx, ly;         // start cell
    short sx, sy, sdx, sdy;
    int sj;
    unsigned char status;  // 0 alive, 1 dead, 2 wants start cell, 3 spiral
    unsigned char merge;
};

struct PArgs {
    Planes p;
    const DevTables* T;
    int W, H;
    int zx, zy, zw, zh;
    fse_particle* pbuf;
    PState* st;
    unsigned int n;
    unsigned int* counters;  // [0] live, [1] pending, [2] compacted
    long long* keys;
    unsigned long long* vals;
    unsigned int tmask;
    uint8_t* awake;  // active-region flags (null = off)
    int acols, arows;
    unsigned int* list;    // particles that hit something (indices into st), appended by the integrate kernel
    unsigned int n_list;   // entries of list the deposit rounds run over
    const unsigned int* prev_pending;  // particles still pending after the previous round (null in round 0): 0 = nothing left to do
    unsigned int* my_pending;          // this round's count
};

__device__ __forceinline__ int phys_at(const PArgs& a, int x, int y) { return a.T->phys[a.p.mat[(size_t)y * a.W + x]]; }

__global__ void particles_integrate_kernel(PArgs a) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    fse_particle cur = a.pbuf[i];
    PState s;
    s.cand = -1;
    s.sx = 0; s.sy = 0; s.sdx = 0; s.sdy = -1; s.sj = 0;
    s.merge = 0;
    s.status = 0;
    const int W = a.W, H = a.H;
    do {
        if (cur.temporary && cur.lifetime <= 0) { s.status = 1; break; }  // 2033
        if (cur.target_force != 0) {  // 2039-2051
            float tdx = cur.target_x - cur.x;
            float tdy = cur.target_y - cur.y;
            float normFac = sqrtf(tdx * tdx + tdy * tdy);
            cur.vx += tdx / normFac * cur.target_force;
            cur.vy += tdy / normFac * cur.target_force;
            if (normFac < 100) {
                cur.vx *= 0.95f;
                cur.vy *= 0.95f;
            }
        }
        const int lx = (int)cur.x, ly = (int)cur.y;
        s.lx = lx;
        s.ly = ly;
        if (cur.x < 0 || (int)(cur.x) >= W || cur.y < 0 || (int)(cur.y) >= H) { s.status = 1; break; }  // 2056
        if (!(lx >= a.zx && ly >= a.zy && lx < a.zx + a.zw && ly < a.zy + a.zh)) break;                // 2061
        cur.vx += cur.ax;
        cur.vy += cur.ay;
        const int div = (int)((fabsf(cur.vx) + fabsf(cur.vy)) + 1);  // 2066
        const float dvx = cur.vx / div;
        const float dvy = cur.vy / div;
        bool done = false;
        for (int k = 0; k < div; k++) {
            cur.x += dvx;
            cur.y += dvy;
            if (cur.x < 0 || (int)(cur.x) >= W || cur.y < 0 || (int)(cur.y) >= H) { s.status = 1; done = true; break; }  // 2075
            const int ph = phys_at(a, (int)cur.x, (int)cur.y);
            if (!cur.phase && ph != P_AIR) {  // 2080
                const bool isObject = ph == P_PASSABLE;  // PhysicsType::OBJECT == PASSABLE == 5
                if (cur.in_object_state == 0) cur.in_object_state = isObject ? 1 : 2;
                else if (cur.in_object_state == 1 && !isObject) cur.in_object_state = 2;
                if (!isObject || cur.in_object_state == 2) {
                    if (cur.temporary) { s.status = 1; done = true; break; }  // 2098
                    s.status = phys_at(a, lx, ly) != P_AIR ? 3 : 2;            // 2104 / 2159
                    done = true;
                    break;
                }
            }
        }
        if (done) break;
        if (cur.lifetime > 0) cur.lifetime--;  // 2170
    } while (false);
    s.adv = cur;
    a.st[i] = s;
    if (s.status >= 2) a.list[atomicAdd(&a.counters[1], 1u)] = i;  // the deposit rounds only visit these (order is irrelevant: ids decide)
}

__device__ __forceinline__ unsigned int hash_cell(long long c) {
    unsigned long long z = (unsigned long long)c * 0x9E3779B97F4A7C15ULL;
    return (unsigned int)(z >> 32);
}

__global__ void particles_propose_kernel(PArgs a) {
    const unsigned int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= a.n_list || (a.prev_pending && *a.prev_pending == 0)) return;
    PState* sp = &a.st[a.list[li]];
    int status = sp->status;
    if (status < 2) return;
    const int W = a.W, H = a.H;
    const float cx = sp->adv.x, cy = sp->adv.y;
    long long cand = -1;
    int merge = 0;
    if (status == 2) {
        if (phys_at(a, sp->lx, sp->ly) == P_AIR) cand = sp->lx + (long long)sp->ly * W;
        else status = 3;
    }
    if (status == 3) {
        int sx = sp->sx, sy = sp->sy, sdx = sp->sdx, sdy = sp->sdy, sj = sp->sj;
        const int myMat = sp->adv.tile.mat;
        const bool amSoup = a.T->phys[myMat] == P_SOUP;
        while (sj < 32 * 32) {  // 2116-2147
            if (-16 <= sx && sx <= 16 && -16 <= sy && sy <= 16) {
                const int px = (int)(cx + sx), py = (int)(cy + sy);
                if (px >= 0 && py >= 0 && px < W && py < H) {
                    const int m = a.p.mat[(size_t)py * W + px];
                    if (a.T->phys[m] == P_AIR) { cand = px + (long long)py * W; break; }
                    if (amSoup && m == myMat) { cand = px + (long long)py * W; merge = 1; break; }
                }
            }
            if ((sx == sy) || ((sx < 0) && (sx == -sy)) || ((sx > 0) && (sx == 1 - sy))) {
                int t = sdx;
                sdx = -sdy;
                sdy = t;
            }
            sx += sdx;
            sy += sdy;
            sj++;
        }
        sp->sx = (short)sx; sp->sy = (short)sy; sp->sdx = (short)sdx; sp->sdy = (short)sdy; sp->sj = sj;
        if (cand < 0) {  // 2154-2157: bounce
            sp->adv.vy = -4.0f;
            sp->adv.y -= 16.0f;
            sp->status = 0;
            sp->cand = -1;
            return;
        }
    }
    sp->status = (unsigned char)status;
    sp->cand = cand;
    sp->merge = (unsigned char)merge;
    // claim: lowest id per cell
    unsigned int h = hash_cell(cand) & a.tmask;
    for (;;) {
        long long prev = (long long)atomicCAS((unsigned long long*)&a.keys[h], (unsigned long long)-1LL, (unsigned long long)cand);
        if (prev == -1LL || prev == cand) {
            atomicMin(&a.vals[h], (unsigned long long)sp->adv.id);
            break;
        }
        h = (h + 1) & a.tmask;
    }
}

__global__ void particles_commit_kernel(PArgs a) {
    const unsigned int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= a.n_list || (a.prev_pending && *a.prev_pending == 0)) return;
    PState* sp = &a.st[a.list[li]];
    if (sp->status < 2) return;
    const long long cand = sp->cand;
    unsigned int h = hash_cell(cand) & a.tmask;
    while (a.keys[h] != cand) h = (h + 1) & a.tmask;
    if (a.vals[h] != (unsigned long long)sp->adv.id) {
        atomicAdd(a.my_pending, 1u);  // still pending
        return;
    }
    const size_t g = (size_t)cand;
    const fse_cell t = sp->adv.tile;
    if (sp->merge) {
        a.p.fl[g] += t.fluid;  // 2133
        a.p.flg[g] |= F_DIRTY;
    } else {  // real_tiles[...] = cur->tile (2127 / 2160)
        a.p.mat[g] = (uint8_t)t.mat;
        a.p.flg[g] = (uint8_t)((t.moved ? F_MOVED : 0) | F_DIRTY);
        a.p.stl[g] = t.settle;
        a.p.tmp[g] = t.temp;
        a.p.col[g] = t.color;
        a.p.fl[g] = t.fluid;
        a.p.fd[g] = t.fluid_diff;
    }
    if (a.awake) {  // a deposit can un-settle the cells around it: wake the 3x3 chunks
        const int ci = (int)(cand % a.W) / CHUNK, cj = (int)(cand / a.W) / CHUNK;
        for (int dj = -1; dj <= 1; dj++)
            for (int di = -1; di <= 1; di++) {
                const int ni = ci + di, nj = cj + dj;
                if (ni >= 0 && nj >= 0 && ni < a.acols && nj < a.arows) a.awake[nj * a.acols + ni] = 1;
            }
    }
    sp->status = 1;
}

__global__ void particles_compact_kernel(PArgs a, fse_particle* out) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const PState* sp = &a.st[i];
    if (sp->status == 1) return;
    const fse_particle p = sp->status == 0 ? sp->adv : a.pbuf[i];  // still pending: retried next tick from its old state
    if (p.y > (float)a.H) return;                                  // 2190
    const unsigned int o = atomicAdd(&a.counters[2], 1u);
    out[o] = p;
}

static cudaError_t grow(void** p, size_t* have, size_t need) {
    if (*have >= need) return cudaSuccess;
    cudaFree(*p);
    *p = nullptr;
    *have = 0;
    size_t n = need + need / 2 + (1 << 20);  // head room: a growing particle pool should not reallocate on every call
    cudaError_t e = cudaMalloc(p, n);
    if (e == cudaSuccess) *have = n;
    return e;
}

}  // namespace fse

using namespace fse;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return fail(FSE_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

extern "C" FSE_API int fse_particles_tick(fse_world* w, const fse_rect* z) {
    if (!w || !z) return fail(FSE_EINVAL, "fse_particles_tick: null argument");
    if (w->strip && w->ctx->nranks > 1)
        return fail(FSE_ESTATE, "fse_particles_tick: particle migration between strips is not implemented yet (single-GPU worlds only)");
    CK(cudaSetDevice(w->ctx->device));
    unsigned int n = 0;
    {
        int64_t n64 = 0;
        if (int r = fse_particles_count(w, &n64)) return r;  // clamps an overflowed counter and records the drops
        n = (unsigned int)n64;
        w->particles_promised = 0;
    }
    if (n == 0) return FSE_OK;
    CK(grow(&w->part_scratch, &w->part_scratch_bytes, sizeof(PState) * (size_t)n));
    CK(grow((void**)&w->pbuf2, &w->pbuf2_bytes, sizeof(fse_particle) * (size_t)w->pcap));
    CK(grow((void**)&w->part_list, &w->part_list_bytes, sizeof(unsigned int) * ((size_t)n + 32)));  // + one pending counter per round
    PArgs a;
    a.p = w->p;
    a.T = w->ctx->d_tabs;
    a.W = w->W; a.H = w->H;
    a.zx = z->x; a.zy = z->y; a.zw = z->w; a.zh = z->h;
    a.pbuf = w->pbuf;
    a.st = (PState*)w->part_scratch;
    a.n = n;
    a.counters = w->pcount;
    a.keys = nullptr; a.vals = nullptr; a.tmask = 0;
    a.awake = w->active_on ? w->d_awake : nullptr;
    a.acols = w->acols; a.arows = w->arows;
    a.list = (unsigned int*)w->part_list;
    a.n_list = 0;
    const int B = 128;
    const int G = (int)((n + B - 1) / B);
    CK(cudaMemsetAsync(w->pcount + 1, 0, 2 * sizeof(unsigned int), w->stream));
    particles_integrate_kernel<<<G, B, 0, w->stream>>>(a);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    unsigned int pending = 0;
    CK(cudaMemcpyAsync(&pending, w->pcount + 1, sizeof pending, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    a.n_list = pending;  // every particle that hit something; the ones that are done drop out by their status byte
    const int GL = (int)((pending + B - 1) / B);
    if (pending > 0) {
        // All rounds are enqueued back to back: a round whose predecessor left nothing pending returns at once, so the host does not
        // have to read a counter between rounds.  The claim table is sized for the first round, the fullest one.
        size_t tsz = 1024;
        while (tsz < (size_t)pending * 2) tsz <<= 1;
        CK(grow((void**)&w->claim_keys, &w->claim_keys_bytes, tsz * sizeof(long long)));
        CK(grow((void**)&w->claim_vals, &w->claim_vals_bytes, tsz * sizeof(unsigned long long)));
        unsigned int* round_cnt = (unsigned int*)w->part_list + n;
        CK(cudaMemsetAsync(round_cnt, 0, 32 * sizeof(unsigned int), w->stream));
        a.keys = (long long*)w->claim_keys;
        a.vals = (unsigned long long*)w->claim_vals;
        a.tmask = (unsigned int)(tsz - 1);
        for (int round = 0; round < FSE_PARTICLE_ROUNDS; round++) {
            CK(cudaMemsetAsync(w->claim_keys, 0xff, tsz * sizeof(long long), w->stream));
            CK(cudaMemsetAsync(w->claim_vals, 0xff, tsz * sizeof(unsigned long long), w->stream));
            a.prev_pending = round ? round_cnt + round - 1 : nullptr;
            a.my_pending = round_cnt + round;
            particles_propose_kernel<<<GL, B, 0, w->stream>>>(a);
            CK(cudaGetLastError());
            particles_commit_kernel<<<GL, B, 0, w->stream>>>(a);
            CK(cudaGetLastError());
            w->ctx->launches += 2;
        }
    }
    particles_compact_kernel<<<G, B, 0, w->stream>>>(a, w->pbuf2);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    // live count <- compacted count; swap pools
    CK(cudaMemcpyAsync(w->pcount, w->pcount + 2, sizeof(unsigned int), cudaMemcpyDeviceToDevice, w->stream));
    fse_particle* t = w->pbuf;
    w->pbuf = w->pbuf2;
    w->pbuf2 = t;
    size_t tb = sizeof(fse_particle) * (size_t)w->pcap;
    w->pbuf2_bytes = tb;
    return FSE_OK;
}
