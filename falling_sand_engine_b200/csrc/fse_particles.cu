// fse_particles.cu — loose particles: world::tickCells() (reference: source/engine/world.cpp:2030-2195) on the GPU.
//
// Schedule.  The reference walks its std::vector<CellData*> in order and every deposit is seen by
// the particles after it.  Here (1) every particle is integrated — target attraction, v += a, sub-stepped collision,
// object pass-through state machine — against the grid as it was when the call started, one thread per particle; a
// particle that simply flew on goes straight into the next tick's pool; (2) particles that hit something resolve their
// deposit in rounds: each proposes its start cell or the first free cell of the reference's 32x32 square spiral (or a
// same-material liquid cell to merge into; the whole warp searches one particle's spiral), an open-addressing claim
// table keeps the LOWEST particle id per contested cell, winners write the cell, losers look again next round;
// (3) what is left of them (bounced, or still pending: retried next tick from the old state) joins the new pool.
// Everything is keyed on particle ids, so the result does not depend on thread scheduling or on the order of the pool.
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
    unsigned int src;      // index of the particle in the pool the call started from (its state before integration)
};

// proposal of a particle for a cell in the band around a strip cut, as it travels to the neighbour rank
struct PProp {
    long long cell;          // x + y * W in global coordinates
    unsigned long long id;
    fse_cell tile;
    int merge;
};
static_assert(sizeof(PProp) == 40, "PProp travels over NCCL as raw bytes");

struct PArgs {
    Planes p;
    const DevTables* T;
    int W, H;        // H = height of the whole world (global rows)
    int y_off, Hl;   // the planes hold global rows [y_off, y_off + Hl) (strip worlds; plain worlds: 0, H)
    unsigned int* oob;  // counts grid probes outside the held rows (a particle faster than the ghost band; strips only)
    int zx, zy, zw, zh;
    fse_particle* pbuf;
    PState* st;
    unsigned int n;
    unsigned int* counters;  // [0] live, [1] pending, [2] compacted
    long long* keys;
    unsigned long long* vals;
    unsigned int tmask;
    int round;       // deposit round: part of every claim key, so the table is cleared once per call, not once per round
    uint8_t* awake;  // active-region flags (null = off)
    int acols, arows;
    fse_particle* out;     // the pool of the next tick: survivors are appended as soon as their fate is known (counters[2])
    unsigned int n_list;   // particles that hit something: st[0 .. n_list), appended by the integrate kernel (counters[1])
    const unsigned int* prev_pending;  // particles still pending after the previous round (null in round 0): 0 = nothing left to do
    unsigned int* my_pending;          // this round's count
    // strips (null / 0 on plain worlds)
    unsigned int* n_props;             // proposals made this round
    PProp *band_up, *band_down;        // proposals for the band of `ghost` rows on either side of the cut above / below
    unsigned int* band_cnt;            // [0] up, [1] down
    int own_lo, own_hi, ghost;         // owned global rows
};

// strips: every particle to the rank that owns its row (world.cpp has one list; here ownership follows int(y))
__global__ void particles_partition_kernel(const fse_particle* in, unsigned int n, int own_lo, int own_hi, int has_up, int has_down,
                                           fse_particle* keep, fse_particle* up, fse_particle* down, unsigned int* cnt) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fse_particle p = in[i];
    const long long row = (long long)floorf(p.y);
    if (has_up && row < own_lo) up[atomicAdd(&cnt[1], 1u)] = p;
    else if (has_down && row >= own_hi) down[atomicAdd(&cnt[2], 1u)] = p;
    else keep[atomicAdd(&cnt[0], 1u)] = p;
}

// cell index in the local planes of global cell (x, y), or -1 when this rank does not hold the row
__device__ __forceinline__ long long local_cell(const PArgs& a, int x, int y) {
    const int r = y - a.y_off;
    if (r < 0 || r >= a.Hl) {
        if (a.oob) atomicAdd(a.oob, 1u);
        return -1;
    }
    return (long long)r * a.W + x;
}
__device__ __forceinline__ int phys_at(const PArgs& a, int x, int y) {
    const long long g = local_cell(a, x, y);
    return g < 0 ? (int)P_SOLID : (int)a.T->phys[a.p.mat[g]];
}

// slot for one more record behind *counter for every thread that wants one: one atomic per warp.  All 32 lanes must call.
__device__ __forceinline__ unsigned int warp_append(unsigned int* counter, bool want) {
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (!m) return 0;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned int)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + (unsigned int)__popc(m & ((1u << lane) - 1u));
}

// One thread per particle.  A particle that simply flew on (the bulk of a pool of millions) goes straight into the next tick's pool:
// 80 bytes in, 80 bytes out.  Only the ones that hit something get a PState record for the deposit rounds.
// The 80-byte AoS records move through shared memory: the CTA's 128 records are loaded as 640 consecutive 16-byte words (a thread
// reading its own record straight from global memory touches 32 different sectors per warp instruction: the kernel was bound by L1
// wavefronts, not by DRAM), a thread takes its record from there (stride 80 B = 20 banks: conflict-free for 128-bit accesses), and a
// warp's survivors are packed back into the warp's slice and leave as one contiguous run to the slots its single atomic reserved.
constexpr int PINT_THREADS = 128;
constexpr int PREC_WORDS = (int)(sizeof(fse_particle) / 16);
static_assert(sizeof(fse_particle) % 16 == 0, "particle records move as 16-byte words");
__global__ void __launch_bounds__(PINT_THREADS) particles_integrate_kernel(PArgs a) {
    __shared__ uint4 stage[PINT_THREADS * PREC_WORDS];
    const unsigned int base = blockIdx.x * PINT_THREADS;
    const unsigned int i = base + threadIdx.x;
    const bool valid = i < a.n;
    {
        const unsigned int nrec = a.n - base < (unsigned int)PINT_THREADS ? a.n - base : (unsigned int)PINT_THREADS;
        const uint4* src = reinterpret_cast<const uint4*>(a.pbuf + base);
        for (unsigned int q = threadIdx.x; q < nrec * PREC_WORDS; q += PINT_THREADS) stage[q] = src[q];
    }
    __syncthreads();
    fse_particle cur;
    if (valid) cur = *reinterpret_cast<const fse_particle*>(&stage[threadIdx.x * PREC_WORDS]);
    PState s;
    s.status = 1;
    if (valid) {
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
    }
    const bool alive = valid && s.status == 0 && !(cur.y > (float)a.H);  // 2190: particles below the world are dropped
    const bool pend = valid && s.status >= 2;
    {   // survivors of this warp: packed into the warp's slice of the stage, then out as one contiguous run
        const int lane = threadIdx.x & 31;
        const unsigned am = __ballot_sync(0xffffffffu, alive);
        const int k = __popc(am);
        if (k) {
            uint4* slice = &stage[(threadIdx.x & ~31) * PREC_WORDS];
            __syncwarp();  // every lane has taken its record out of the slice
            if (alive) *reinterpret_cast<fse_particle*>(&slice[__popc(am & ((1u << lane) - 1u)) * PREC_WORDS]) = cur;
            unsigned int o = 0;
            if (lane == 0) o = atomicAdd(&a.counters[2], (unsigned int)k);
            o = __shfl_sync(0xffffffffu, o, 0);
            __syncwarp();
            uint4* dst = reinterpret_cast<uint4*>(a.out + o);
            for (int q = lane; q < k * PREC_WORDS; q += 32) dst[q] = slice[q];
        }
    }
    const unsigned int q = warp_append(&a.counters[1], pend);  // the deposit rounds only visit these (order is irrelevant: ids decide)
    if (pend) {
        s.adv = cur;
        s.src = i;
        a.st[q] = s;
    }
}

__device__ __forceinline__ unsigned int hash_cell(long long c) {
    unsigned long long z = (unsigned long long)c * 0x9E3779B97F4A7C15ULL;
    return (unsigned int)(z >> 32);
}

// claim: lowest id per cell (open addressing).  Keys are (global cell index, round): entries of earlier rounds never match and only
// take up room — every proposed cell has exactly one winner, so all rounds together insert at most as many keys as particles were
// pending at the start, which is what the table is sized for.
__device__ __forceinline__ long long claim_key(const PArgs& a, long long cell) { return cell * 32 + a.round; }
__device__ __forceinline__ unsigned int claim_slot(const PArgs& a, long long cell) {  // slot of a key that is in the table
    const long long key = claim_key(a, cell);
    unsigned int h = hash_cell(key) & a.tmask;
    while (a.keys[h] != key) h = (h + 1) & a.tmask;
    return h;
}
__device__ __forceinline__ void claim_cell(const PArgs& a, long long cell, unsigned long long id) {
    const long long cand = claim_key(a, cell);
    unsigned int h = hash_cell(cand) & a.tmask;
    for (;;) {
        long long prev = (long long)atomicCAS((unsigned long long*)&a.keys[h], (unsigned long long)-1LL, (unsigned long long)cand);
        if (prev == -1LL || prev == cand) {
            atomicMin(&a.vals[h], id);
            break;
        }
        h = (h + 1) & a.tmask;
    }
}
__device__ __forceinline__ void deposit_cell(const PArgs& a, size_t g, const fse_cell& t, int merge) {
    if (merge) {
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
}

// The reference's 32 x 32 square spiral (world.cpp:2116-2147) as a table: SPIRAL[j] = the (sx, sy) offset tested at step j.  Filled on
// the host by the very loop of the reference (spiral_table_init) so that a warp can test 32 consecutive steps at once.
__constant__ signed char SPIRAL[1024][2];

__device__ __forceinline__ void propose_one(const PArgs& a, unsigned int li) {
    const int lane = threadIdx.x & 31;
    const bool live = li < a.n_list;
    PState* sp = live ? &a.st[li] : nullptr;
    int status = live ? sp->status : 0;
    if (status < 2) status = 0;
    const int W = a.W, H = a.H;
    float cx = 0.0f, cy = 0.0f;
    long long cand = -1;
    int merge = 0, sj = 0, myMat = 0;
    if (status) {
        cx = sp->adv.x;
        cy = sp->adv.y;
        myMat = sp->adv.tile.mat;
        sj = sp->sj;
    }
    if (status == 2) {
        if (phys_at(a, sp->lx, sp->ly) == P_AIR) cand = sp->lx + (long long)sp->ly * W;
        else status = 3;
    }
    // Spiral searches, one particle at a time by the whole warp: lane l tests step sj + l of the spiral, the lowest step that finds AIR
    // (or the particle's own liquid to merge into) wins — exactly the cell the reference's sequential loop stops at, in 1/32 of the
    // dependent loads.  (A single thread walking up to 1024 steps of two dependent global loads each made this kernel's time.)
    unsigned need = __ballot_sync(0xffffffffu, status == 3);
    while (need) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        const float bcx = __shfl_sync(0xffffffffu, cx, src), bcy = __shfl_sync(0xffffffffu, cy, src);
        const int bmat = __shfl_sync(0xffffffffu, myMat, src);
        int base = __shfl_sync(0xffffffffu, sj, src);
        const bool amSoup = a.T->phys[bmat] == P_SOUP;
        int found = 32 * 32, fmerge = 0;
        while (base < 32 * 32) {
            const int j = base + lane;
            int hit = 0;  // 1 AIR, 2 same liquid
            if (j < 32 * 32) {
                const int sx = SPIRAL[j][0], sy = SPIRAL[j][1];
                if (-16 <= sx && sx <= 16 && -16 <= sy && sy <= 16) {
                    const int px = (int)(bcx + sx), py = (int)(bcy + sy);
                    if (px >= 0 && py >= 0 && px < W && py < H) {
                        const long long g = local_cell(a, px, py);
                        if (g >= 0) {
                            const int m = a.p.mat[g];
                            if (a.T->phys[m] == P_AIR) hit = 1;
                            else if (amSoup && m == bmat) hit = 2;
                        }
                    }
                }
            }
            const unsigned hb = __ballot_sync(0xffffffffu, hit != 0);
            if (hb) {
                const int l = __ffs(hb) - 1;
                found = base + l;
                fmerge = __shfl_sync(0xffffffffu, hit, l) == 2;
                break;
            }
            base += 32;
        }
        if (lane == src) {
            sj = found;
            if (found < 32 * 32) {
                const int sx = SPIRAL[found][0], sy = SPIRAL[found][1];
                cand = (int)(cx + sx) + (long long)(int)(cy + sy) * W;
                merge = fmerge;
            }
        }
    }
    if (!status) return;
    if (status == 3) {
        sp->sj = sj;
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
    claim_cell(a, cand, sp->adv.id);
    if (a.n_props) atomicAdd(a.n_props, 1u);
    // strips: a proposal for the band around a cut also goes to the neighbour (both ranks pick the same winner)
    if (a.band_up || a.band_down) {
        const int row = (int)(cand / a.W);
        PProp pr;
        pr.cell = cand;
        pr.id = sp->adv.id;
        pr.tile = sp->adv.tile;
        pr.merge = merge;
        if (a.band_up && row >= a.own_lo - a.ghost && row < a.own_lo + a.ghost) a.band_up[atomicAdd(&a.band_cnt[0], 1u)] = pr;
        if (a.band_down && row >= a.own_hi - a.ghost && row < a.own_hi + a.ghost) a.band_down[atomicAdd(&a.band_cnt[1], 1u)] = pr;
    }
}
// Grid-stride over the pending list in whole warps (the spiral search is warp-cooperative).  Rounds after the first are launched with a
// small grid: a round whose predecessor left nothing pending returns at once, and the few losers of a contested cell do not need a
// thread per particle of the first round.
__global__ void particles_propose_kernel(PArgs a) {
    if (a.prev_pending && *a.prev_pending == 0) return;
    const unsigned int up = (a.n_list + 31u) & ~31u;
    for (unsigned int base = blockIdx.x * blockDim.x; base < up; base += gridDim.x * blockDim.x) propose_one(a, base + threadIdx.x);
}

// strips: the neighbours' band proposals join the claim table ...
__global__ void particles_ext_claim_kernel(PArgs a, const PProp* ext, unsigned int n) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) claim_cell(a, ext[i].cell, ext[i].id);
}
// ... and the ones that won are written into the rows this rank holds (the owner of the particle writes its own copy)
__global__ void particles_ext_commit_kernel(PArgs a, const PProp* ext, unsigned int n) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PProp pr = ext[i];
    const unsigned int h = claim_slot(a, pr.cell);
    if (a.vals[h] != pr.id) return;
    const int r = (int)(pr.cell / a.W) - a.y_off;
    if (r < 0 || r >= a.Hl) return;
    deposit_cell(a, (size_t)r * a.W + (size_t)(pr.cell % a.W), pr.tile, pr.merge);
}

__device__ __forceinline__ void commit_one(const PArgs& a, unsigned int li) {
    if (li >= a.n_list) return;
    PState* sp = &a.st[li];
    if (sp->status < 2) return;
    const long long cand = sp->cand;
    const unsigned int h = claim_slot(a, cand);
    if (a.vals[h] != (unsigned long long)sp->adv.id) {
        atomicAdd(a.my_pending, 1u);  // still pending
        return;
    }
    {
        const int r = (int)(cand / a.W) - a.y_off;  // a winner outside the held rows is written by the neighbour that holds the cell
        if (r >= 0 && r < a.Hl) deposit_cell(a, (size_t)r * a.W + (size_t)(cand % a.W), sp->adv.tile, sp->merge);
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
__global__ void particles_commit_kernel(PArgs a) {
    if (a.prev_pending && *a.prev_pending == 0) return;
    for (unsigned int li = blockIdx.x * blockDim.x + threadIdx.x; li < a.n_list; li += gridDim.x * blockDim.x) commit_one(a, li);
}

// After the rounds: a particle that deposited is gone; one that bounced flies on; one that is still pending is retried next tick from
// the state it had before this call.
__global__ void particles_finish_kernel(PArgs a) {
    const unsigned int li = blockIdx.x * blockDim.x + threadIdx.x;
    const PState* sp = li < a.n_list ? &a.st[li] : nullptr;
    const bool keep = sp && sp->status != 1;
    fse_particle p;
    if (keep) p = sp->status == 0 ? sp->adv : a.pbuf[sp->src];
    const bool alive = keep && !(p.y > (float)a.H);  // 2190
    const unsigned int o = warp_append(&a.counters[2], alive);
    if (alive) a.out[o] = p;
}

// SPIRAL <- the offsets the loop of world.cpp:2116-2147 visits, step by step (once per device)
static cudaError_t spiral_table_init(fse_ctx* c) {
    if (c->spiral_ready) return cudaSuccess;
    signed char t[1024][2];
    int sx = 0, sy = 0, sdx = 0, sdy = -1;
    for (int j = 0; j < 32 * 32; j++) {
        // offsets beyond +-16 are never used by the reference (its range check); clamp so they stay out of range in 8 bits
        t[j][0] = (signed char)(sx < -17 ? -17 : (sx > 17 ? 17 : sx));
        t[j][1] = (signed char)(sy < -17 ? -17 : (sy > 17 ? 17 : sy));
        if ((sx == sy) || ((sx < 0) && (sx == -sy)) || ((sx > 0) && (sx == 1 - sy))) {
            int q = sdx;
            sdx = -sdy;
            sdy = q;
        }
        sx += sdx;
        sy += sdy;
    }
    cudaError_t e = cudaMemcpyToSymbol(SPIRAL, t, sizeof t);
    if (e == cudaSuccess) c->spiral_ready = true;
    return e;
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

// ---- strips: the protocol of DESIGN.md §5 (validated on the host by tests/test_strips_cpu.py with the oracle's staged schedule) ----
// 1. ghost rows fresh (32 rows either side of a cut), every particle to the rank that owns its row;
// 2. each rank integrates its particles against its rows + ghost rows;
// 3. per deposit round: proposals for the band of 32 rows on either side of a cut travel to that neighbour, both ranks put
//    them into their claim tables (the lowest id wins whoever owns the particle) and both write the winners' cells they hold;
// 4. rounds stop when no rank proposed anything (one all-reduce per round).
// Variable-length arrays travel as two grouped exchanges: the counts, then the payload.
struct StripBufs {
    unsigned int* d_cnt = nullptr;   // [0..2] partition keep/up/down, [3..4] band up/down, [5..6] received up/down, [7] proposals, [8] oob
    unsigned int* h_cnt = nullptr;   // pinned mirror
    void* send[2] = {nullptr, nullptr};
    void* recv[2] = {nullptr, nullptr};
    size_t send_bytes[2] = {0, 0}, recv_bytes[2] = {0, 0};
};
static StripBufs* strip_bufs(fse_world* w) {
    if (!w->particle_strip) {
        StripBufs* b = new StripBufs();
        if (cudaMalloc((void**)&b->d_cnt, 16 * sizeof(unsigned int)) != cudaSuccess || cudaMallocHost((void**)&b->h_cnt, 16 * sizeof(unsigned int)) != cudaSuccess) {
            delete b;
            return nullptr;
        }
        w->particle_strip = b;
    }
    return (StripBufs*)w->particle_strip;
}
void fse::particles_strip_free(fse_world* w) {
    StripBufs* b = (StripBufs*)w->particle_strip;
    if (!b) return;
    cudaFree(b->d_cnt);
    cudaFreeHost(b->h_cnt);
    for (int q = 0; q < 2; q++) {
        cudaFree(b->send[q]);
        cudaFree(b->recv[q]);
    }
    delete b;
    w->particle_strip = nullptr;
}

// send n_up / n_down records of `es` bytes to the neighbours, receive theirs; counts first.  *got_up / *got_down = records received
// (in b->recv[0] / b->recv[1]).  One host sync for the counts.
static int exchange_records(fse_world* w, StripBufs* b, const void* up, unsigned int n_up, const void* down, unsigned int n_down, size_t es,
                            unsigned int* got_up, unsigned int* got_down) {
    const bool has_up = w->ctx->rank > 0, has_down = w->ctx->rank + 1 < w->ctx->nranks;
    unsigned int* c = b->d_cnt;
    b->h_cnt[9] = n_up;
    b->h_cnt[10] = n_down;
    CK(cudaMemcpyAsync(c + 9, b->h_cnt + 9, 2 * sizeof(unsigned int), cudaMemcpyHostToDevice, w->stream));
    CK(cudaMemsetAsync(c + 5, 0, 2 * sizeof(unsigned int), w->stream));
    if (int r = strip_sendrecv(w, c + 9, has_up ? 4 : 0, c + 5, has_up ? 4 : 0, c + 10, has_down ? 4 : 0, c + 6, has_down ? 4 : 0, w->stream)) return r;
    CK(cudaMemcpyAsync(b->h_cnt + 5, c + 5, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    const unsigned int in_up = has_up ? b->h_cnt[5] : 0, in_down = has_down ? b->h_cnt[6] : 0;
    CK(grow(&b->recv[0], &b->recv_bytes[0], (size_t)in_up * es + 64));
    CK(grow(&b->recv[1], &b->recv_bytes[1], (size_t)in_down * es + 64));
    if (int r = strip_sendrecv(w, up, has_up ? (size_t)n_up * es : 0, b->recv[0], (size_t)in_up * es, down, has_down ? (size_t)n_down * es : 0, b->recv[1],
                               (size_t)in_down * es, w->stream))
        return r;
    *got_up = in_up;
    *got_down = in_down;
    return FSE_OK;
}

static int particles_tick_strips(fse_world* w, const fse_rect* z) {
    StripBufs* b = strip_bufs(w);
    if (!b) return fail(FSE_ENOMEM, "fse_particles_tick: strip buffers");
    const bool has_up = w->ctx->rank > 0, has_down = w->ctx->rank + 1 < w->ctx->nranks;
    const int GH = 32;  // ghost rows a strip holds beyond the rows it owns (strips.GHOST)
    const int B = 128;
    // 1a. ghost rows: the owner's rows next to each cut
    if (int r = strip_refresh(w, w->stream, GH)) return r;
    // 1b. every particle to the rank that owns its row
    int64_t n64 = 0;
    if (int r = fse_particles_count(w, &n64)) return r;
    w->particles_promised = 0;
    unsigned int n = (unsigned int)n64;
    CK(grow((void**)&w->pbuf2, &w->pbuf2_bytes, sizeof(fse_particle) * (size_t)w->pcap));
    CK(grow(&b->send[0], &b->send_bytes[0], sizeof(fse_particle) * (size_t)n + 64));
    CK(grow(&b->send[1], &b->send_bytes[1], sizeof(fse_particle) * (size_t)n + 64));
    CK(cudaMemsetAsync(b->d_cnt, 0, 16 * sizeof(unsigned int), w->stream));
    if (n) {
        particles_partition_kernel<<<(n + B - 1) / B, B, 0, w->stream>>>(w->pbuf, n, w->own_lo, w->own_hi, has_up, has_down, w->pbuf2, (fse_particle*)b->send[0],
                                                                     (fse_particle*)b->send[1], b->d_cnt);
        CK(cudaGetLastError());
        w->ctx->launches += 1;
    }
    CK(cudaMemcpyAsync(b->h_cnt, b->d_cnt, 3 * sizeof(unsigned int), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    const unsigned int n_keep = b->h_cnt[0], go_up = b->h_cnt[1], go_down = b->h_cnt[2];
    unsigned int in_up = 0, in_down = 0;
    if (int r = exchange_records(w, b, b->send[0], go_up, b->send[1], go_down, sizeof(fse_particle), &in_up, &in_down)) return r;
    const size_t total = (size_t)n_keep + in_up + in_down;
    if (total > w->pcap) {  // arrivals do not fit: grow both pools (pbuf2 holds the kept particles, pbuf is free to go)
        CK(cudaStreamSynchronize(w->stream));
        size_t cap = (size_t)w->pcap * 2;
        while (cap < total) cap *= 2;
        fse_particle* nb = nullptr;
        CK(cudaMalloc(&nb, sizeof(fse_particle) * cap));
        if (n_keep) CK(cudaMemcpy(nb, w->pbuf2, sizeof(fse_particle) * (size_t)n_keep, cudaMemcpyDeviceToDevice));
        cudaFree(w->pbuf);
        cudaFree(w->pbuf2);
        w->pbuf = nullptr;
        w->pbuf2 = nb;
        w->pbuf2_bytes = sizeof(fse_particle) * cap;
        CK(cudaMalloc(&w->pbuf, sizeof(fse_particle) * cap));
        w->pcap = (unsigned int)cap;
    }
    if (in_up) CK(cudaMemcpyAsync(w->pbuf2 + n_keep, b->recv[0], sizeof(fse_particle) * (size_t)in_up, cudaMemcpyDeviceToDevice, w->stream));
    if (in_down) CK(cudaMemcpyAsync(w->pbuf2 + n_keep + in_up, b->recv[1], sizeof(fse_particle) * (size_t)in_down, cudaMemcpyDeviceToDevice, w->stream));
    {   // the partitioned pool becomes the live one
        fse_particle* t = w->pbuf;
        w->pbuf = w->pbuf2;
        w->pbuf2 = t;
        w->pbuf2_bytes = sizeof(fse_particle) * (size_t)w->pcap;
    }
    n = (unsigned int)total;
    // 2. integrate (global coordinates; the planes hold rows [y_off, y_off + H))
    CK(grow(&w->part_scratch, &w->part_scratch_bytes, sizeof(PState) * ((size_t)n + 1)));
    CK(grow((void**)&w->part_list, &w->part_list_bytes, sizeof(unsigned int) * ((size_t)n + 32)));
    PArgs a;
    memset(&a, 0, sizeof a);
    a.p = w->p;
    a.T = w->ctx->d_tabs;
    a.W = w->W; a.H = w->Hglobal;
    a.y_off = w->y_off; a.Hl = w->H;
    a.oob = b->d_cnt + 8;
    a.zx = z->x; a.zy = z->y; a.zw = z->w; a.zh = z->h;
    a.pbuf = w->pbuf;
    a.st = (PState*)w->part_scratch;
    a.n = n;
    a.counters = w->pcount;
    a.out = w->pbuf2;
    a.own_lo = w->own_lo; a.own_hi = w->own_hi; a.ghost = GH;
    const int G = (int)((n + B - 1) / B);
    CK(cudaMemsetAsync(w->pcount + 1, 0, 2 * sizeof(unsigned int), w->stream));
    if (n) {
        particles_integrate_kernel<<<G, B, 0, w->stream>>>(a);
        CK(cudaGetLastError());
        w->ctx->launches += 1;
    }
    unsigned int pending = 0;
    CK(cudaMemcpyAsync(&pending, w->pcount + 1, sizeof pending, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    a.n_list = pending;
    const int GL = (int)((pending + B - 1) / B);
    // band buffers: at most every pending particle proposes into a band
    CK(grow(&b->send[0], &b->send_bytes[0], sizeof(PProp) * (size_t)pending + 64));
    CK(grow(&b->send[1], &b->send_bytes[1], sizeof(PProp) * (size_t)pending + 64));
    a.band_up = has_up ? (PProp*)b->send[0] : nullptr;
    a.band_down = has_down ? (PProp*)b->send[1] : nullptr;
    a.band_cnt = b->d_cnt + 3;
    a.n_props = b->d_cnt + 7;
    unsigned int* round_cnt = (unsigned int*)w->part_list + n;
    CK(cudaMemsetAsync(round_cnt, 0, 32 * sizeof(unsigned int), w->stream));
    // 3. deposit rounds
    for (int round = 0; round < FSE_PARTICLE_ROUNDS; round++) {
        size_t tsz = 1024;
        while (tsz < ((size_t)pending + 4096) * 2) tsz <<= 1;
        CK(grow((void**)&w->claim_keys, &w->claim_keys_bytes, tsz * sizeof(long long)));
        CK(grow((void**)&w->claim_vals, &w->claim_vals_bytes, tsz * sizeof(unsigned long long)));
        a.keys = (long long*)w->claim_keys;
        a.vals = (unsigned long long*)w->claim_vals;
        a.tmask = (unsigned int)(tsz - 1);
        CK(cudaMemsetAsync(w->claim_keys, 0xff, tsz * sizeof(long long), w->stream));
        CK(cudaMemsetAsync(w->claim_vals, 0xff, tsz * sizeof(unsigned long long), w->stream));
        CK(cudaMemsetAsync(b->d_cnt + 3, 0, 2 * sizeof(unsigned int), w->stream));
        CK(cudaMemsetAsync(b->d_cnt + 7, 0, sizeof(unsigned int), w->stream));
        a.round = round;
        a.prev_pending = round ? round_cnt + round - 1 : nullptr;
        a.my_pending = round_cnt + round;
        if (pending) {
            particles_propose_kernel<<<GL, B, 0, w->stream>>>(a);
            CK(cudaGetLastError());
            w->ctx->launches += 1;
        }
        // anything proposed anywhere?  (also tells every rank when to stop)
        CK(cudaMemcpyAsync(b->d_cnt + 11, b->d_cnt + 7, sizeof(unsigned int), cudaMemcpyDeviceToDevice, w->stream));
        if (int r = strip_allreduce_u32(w, b->d_cnt + 11, 1, w->stream)) return r;
        CK(cudaMemcpyAsync(b->h_cnt + 3, b->d_cnt + 3, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, w->stream));
        CK(cudaMemcpyAsync(b->h_cnt + 11, b->d_cnt + 11, sizeof(unsigned int), cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
        if (b->h_cnt[11] == 0) break;
        unsigned int ext_up = 0, ext_down = 0;
        if (int r = exchange_records(w, b, b->send[0], has_up ? b->h_cnt[3] : 0, b->send[1], has_down ? b->h_cnt[4] : 0, sizeof(PProp), &ext_up, &ext_down)) return r;
        if ((size_t)pending + ext_up + ext_down > tsz / 2)
            return fail(FSE_ESTATE, "fse_particles_tick: %u + %u band proposals from the neighbours overflow the claim table", ext_up, ext_down);
        for (int side = 0; side < 2; side++) {
            const unsigned int ne = side ? ext_down : ext_up;
            if (!ne) continue;
            particles_ext_claim_kernel<<<(ne + B - 1) / B, B, 0, w->stream>>>(a, (const PProp*)b->recv[side], ne);
            CK(cudaGetLastError());
            w->ctx->launches += 1;
        }
        if (pending) {
            particles_commit_kernel<<<GL, B, 0, w->stream>>>(a);
            CK(cudaGetLastError());
            w->ctx->launches += 1;
        }
        for (int side = 0; side < 2; side++) {
            const unsigned int ne = side ? ext_down : ext_up;
            if (!ne) continue;
            particles_ext_commit_kernel<<<(ne + B - 1) / B, B, 0, w->stream>>>(a, (const PProp*)b->recv[side], ne);
            CK(cudaGetLastError());
            w->ctx->launches += 1;
        }
    }
    // 4. survivors of the rounds (they change owner at the start of the next call if their row now belongs to a neighbour)
    if (pending) {
        particles_finish_kernel<<<GL, B, 0, w->stream>>>(a);
        CK(cudaGetLastError());
        w->ctx->launches += 1;
    }
    CK(cudaMemcpyAsync(w->pcount, w->pcount + 2, sizeof(unsigned int), cudaMemcpyDeviceToDevice, w->stream));
    CK(cudaMemcpyAsync(b->h_cnt + 8, b->d_cnt + 8, sizeof(unsigned int), cudaMemcpyDeviceToHost, w->stream));
    fse_particle* t = w->pbuf;
    w->pbuf = w->pbuf2;
    w->pbuf2 = t;
    w->pbuf2_bytes = sizeof(fse_particle) * (size_t)w->pcap;
    CK(cudaStreamSynchronize(w->stream));
    w->particles_seen = n;
    if (b->h_cnt[8]) return fail(FSE_ESTATE, "fse_particles_tick: %u grid probes left the %d ghost rows of the strip (a particle faster than the halo band)", b->h_cnt[8], GH);
    return FSE_OK;
}

extern "C" FSE_API int fse_particles_tick(fse_world* w, const fse_rect* z) {
    if (!w || !z) return fail(FSE_EINVAL, "fse_particles_tick: null argument");
    CK(cudaSetDevice(w->ctx->device));
    CK(spiral_table_init(w->ctx));
    if (w->strip && w->ctx->nranks > 1) return particles_tick_strips(w, z);
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
    memset(&a, 0, sizeof a);
    a.p = w->p;
    a.T = w->ctx->d_tabs;
    a.W = w->W; a.H = w->Hglobal ? w->Hglobal : w->H;
    a.y_off = w->y_off; a.Hl = w->H;
    a.zx = z->x; a.zy = z->y; a.zw = z->w; a.zh = z->h;
    a.pbuf = w->pbuf;
    a.st = (PState*)w->part_scratch;
    a.n = n;
    a.counters = w->pcount;
    a.keys = nullptr; a.vals = nullptr; a.tmask = 0;
    a.awake = w->active_on ? w->d_awake : nullptr;
    a.acols = w->acols; a.arows = w->arows;
    a.out = w->pbuf2;
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
        CK(cudaMemsetAsync(w->claim_keys, 0xff, tsz * sizeof(long long), w->stream));  // once: the keys carry the round
        CK(cudaMemsetAsync(w->claim_vals, 0xff, tsz * sizeof(unsigned long long), w->stream));
        const int GS = GL < 592 ? GL : 592;  // rounds after the first: a few losers, or nothing at all
        for (int round = 0; round < FSE_PARTICLE_ROUNDS; round++) {
            a.round = round;
            a.prev_pending = round ? round_cnt + round - 1 : nullptr;
            a.my_pending = round_cnt + round;
            particles_propose_kernel<<<round ? GS : GL, B, 0, w->stream>>>(a);
            CK(cudaGetLastError());
            particles_commit_kernel<<<round ? GS : GL, B, 0, w->stream>>>(a);
            CK(cudaGetLastError());
            w->ctx->launches += 2;
        }
    }
    if (pending > 0) {
        particles_finish_kernel<<<GL, B, 0, w->stream>>>(a);
        CK(cudaGetLastError());
        w->ctx->launches += 1;
    }
    // live count <- compacted count; swap pools
    CK(cudaMemcpyAsync(w->pcount, w->pcount + 2, sizeof(unsigned int), cudaMemcpyDeviceToDevice, w->stream));
    fse_particle* t = w->pbuf;
    w->pbuf = w->pbuf2;
    w->pbuf2 = t;
    size_t tb = sizeof(fse_particle) * (size_t)w->pcap;
    w->pbuf2_bytes = tb;
    return FSE_OK;
}
