// fse_tick.cu — the per-chunk cell update of world::tick() (reference: source/engine/world.cpp:1036-1948)
// as hand-written sm_100a kernels: shared-memory layout, PTX helpers (mbarrier, cp.async.bulk = TMA bulk copies, SASS UBLKCP), the
// cell load / store / create helpers and the row IO shared by the kernels of fse_tick_rows.cuh (included at the end), the
// active-chunk compaction and the launch logic of a colour phase.
//
// One CTA = one 128x128 chunk of one colour phase (world.cpp:1057-1077), streamed bottom-up through a shared-memory window of rows:
// rows enter by cp.async.bulk completing on per-slot mbarriers and leave by cp.async.bulk shared->global once the pass is done with
// them.  The round-1 in-place "classes" schedule (4 interleaved column classes per warp, tick_chunk_kernel) was removed in round 2:
// it measured slower than the rows schedule at every size, and every rule change had to be made twice.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (strict FP like the reference, xmake.lua:38).
#include <cstdlib>
#include "fse_device.cuh"

namespace fse {

// ---- geometry of the shared-memory ring ----------------------------------------------------------
constexpr int RING = 28;        // rows resident
constexpr int PF = 2;           // rows prefetched ahead of need
constexpr int HALO_DN = 10;     // rows below the chunk that are read (sand pillar probe, world.cpp:1620)
constexpr int HALO_UP = 5;      // rows above the chunk that can be read/written (interaction reach)
constexpr int HALO_WR = 5;      // rows below the chunk that can be written
constexpr int L12 = 7;          // pass-2 row lags pass-1 row by 7
constexpr int L23 = 11;         // pass-3 row lags pass-2 row by 11
constexpr int LAST_ROW = CHUNK - 1 + HALO_UP;             // 132
constexpr int STORE_LAG = L12 + L23 + 1;                  // a row is stored the step after pass 3 left it
constexpr int N_STEPS = LAST_ROW + STORE_LAG + 1;         // 152

constexpr int HX8 = 16;   // halo columns loaded for the u8 planes (16-byte granules)
constexpr int HXW = 8;    // halo columns loaded for the 16/32-bit planes
constexpr int P8 = CHUNK + 2 * HX8;   // 160
constexpr int PW = CHUNK + 2 * HXW;   // 144
constexpr int OFF_MAT = 0;
constexpr int OFF_FLG = OFF_MAT + P8;
constexpr int OFF_STL = OFF_FLG + P8;
constexpr int OFF_TMP = OFF_STL + P8;           // 480
constexpr int OFF_COL = OFF_TMP + PW * 2;       // 768
constexpr int OFF_FL = OFF_COL + PW * 4;        // 1344
constexpr int OFF_FD = OFF_FL + PW * 4;         // 1920
constexpr int ROW_BYTES = OFF_FD + PW * 4;      // 2496
static_assert(ROW_BYTES % 16 == 0 && OFF_TMP % 16 == 0 && OFF_COL % 16 == 0, "bulk copies need 16-byte alignment");

// core/const.h:23-34
constexpr float FLUID_MaxValue = 0.5f;
constexpr float FLUID_MinValue = 0.0005f;
constexpr float FLUID_MaxCompression = 0.1f;
constexpr float FLUID_MinFlow = 0.05f;
constexpr float FLUID_MaxFlow = 8.0f;
constexpr float FLUID_FlowSpeed = 1.0f;

// Every tick kernel starts its dynamic shared memory with the same head, so the rule code reaches the material LUT, the row
// flags and the row window at compile-time offsets (plain LDS/STS, no pointer loads).
struct SmemHead {
    Lut lut;
    unsigned char rowmod[32];  // row must be stored back (any plane or the dirty bit changed)
    unsigned char rowchg[32];  // cell state other than the dirty bit changed (active-region tracking)
    unsigned char rowvis[32];  // row got tickVisited marks only (per-pass kernels persist them through HBM)
    unsigned char rowlazy[32]; // pass 1 skipped the row: its tickVisited marks are implicit (iter >= iterations of the cell's material)
};
static_assert(sizeof(SmemHead) % 128 == 0, "the row window behind the head must stay 128-byte aligned");
extern __shared__ __align__(128) unsigned char fse_smem[];
#define LUTP (reinterpret_cast<const Lut*>(fse_smem))
#define ROWMOD (fse_smem + offsetof(SmemHead, rowmod))
#define ROWCHG (fse_smem + offsetof(SmemHead, rowchg))
#define ROWVIS (fse_smem + offsetof(SmemHead, rowvis))
#define ROWLAZY (fse_smem + offsetof(SmemHead, rowlazy))
#define RINGP (fse_smem + sizeof(SmemHead))

struct CellR {
    uint8_t mat;
    uint8_t moved;
    uint8_t stl;
    int16_t tmp;
    uint32_t col;
    float fl;
    float fd;
};

struct Ctx {
    const DevTables* T;
    fse_particle* pbuf;
    unsigned int* pcount;
    unsigned int pcap;
    uint32_t rkey;
    uint32_t tick;
    int iter;
    int nmat;
    int yoff;  // global y of local row 0 (strip worlds), 0 otherwise
    int ringn, ringmask, koff;  // ring geometry of the running kernel (rows kernels): slot(k) = (k + koff) mod ringn
    int air, fire, water, lava, steam, obsidian;
    float *flowx, *flowy;  // world::flowX / flowY (render-only accumulators), null when the world does not keep them
    int W;                 // cells per world row (index of a cell in the flow planes)
};

// ---- PTX helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    uint32_t a = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- ring accessors: s = ring slot of a row, j = column index (x - cx + HX8) -----------------------
__device__ __forceinline__ int slot_of_row(int k) { return (k + HALO_DN) % RING; }  // k = rows above the chunk's bottom row
#define MAT(s, j) (RINGP[(s) * ROW_BYTES + OFF_MAT + (j)])
#define FLG(s, j) (RINGP[(s) * ROW_BYTES + OFF_FLG + (j)])
#define STL(s, j) (RINGP[(s) * ROW_BYTES + OFF_STL + (j)])
#define TMP(s, j) (*reinterpret_cast<int16_t*>(RINGP + (s) * ROW_BYTES + OFF_TMP + ((j) - (HX8 - HXW)) * 2))
#define COL(s, j) (*reinterpret_cast<uint32_t*>(RINGP + (s) * ROW_BYTES + OFF_COL + ((j) - (HX8 - HXW)) * 4))
#define FL(s, j) (*reinterpret_cast<float*>(RINGP + (s) * ROW_BYTES + OFF_FL + ((j) - (HX8 - HXW)) * 4))
#define FD(s, j) (*reinterpret_cast<float*>(RINGP + (s) * ROW_BYTES + OFF_FD + ((j) - (HX8 - HXW)) * 4))
#define PHYS(s, j) (LUTP->phys[MAT(s, j)])

__device__ __forceinline__ CellR ldc(const Ctx& c, int s, int j) {
    CellR r;
    r.mat = MAT(s, j);
    r.moved = FLG(s, j) & F_MOVED;
    r.stl = STL(s, j);
    r.tmp = TMP(s, j);
    r.col = COL(s, j);
    r.fl = FL(s, j);
    r.fd = FD(s, j);
    return r;
}
// real_tiles[i] = cell (+ dirty[i] / tickVisited[i] = true as requested by `setbits`).  p0 = mat | stl << 8 | tmp << 16,
// p1 = moved | setbits << 8.  Inlined at ~20 sites: out-of-line copies (tried in both rounds) make the kernel LARGER — 4832 / 4904 SASS
// instructions instead of 4744 — because every call site then saves and restores registers around the call.
__device__ __forceinline__ void stc_raw(const Ctx* cp, int s, int j, uint32_t p0, uint32_t p1, uint32_t col, float fl, float fd) {
    (void)cp;
    MAT(s, j) = (uint8_t)p0;
    const uint8_t f = FLG(s, j);
    FLG(s, j) = (uint8_t)((f & (F_DIRTY | F_VISITED)) | (p1 & 0xff) | ((p1 >> 8) & 0xff));
    STL(s, j) = (uint8_t)(p0 >> 8);
    TMP(s, j) = (int16_t)(p0 >> 16);
    COL(s, j) = col;
    FL(s, j) = fl;
    FD(s, j) = fd;
    ROWMOD[s] = 1;
    ROWCHG[s] = 1;
}
__device__ __forceinline__ void stc(const Ctx& c, int s, int j, const CellR& r, uint8_t setbits) {
    stc_raw(&c, s, j, (uint32_t)r.mat | ((uint32_t)r.stl << 8) | ((uint32_t)(uint16_t)r.tmp << 16), (r.moved ? 1u : 0u) | ((uint32_t)setbits << 8),
            r.col, r.fl, r.fd);
}
__device__ __forceinline__ void set_moved(const Ctx& c, int s, int j, bool v) {
    uint8_t f = FLG(s, j);
    uint8_t g = v ? (f | F_MOVED) : (f & ~F_MOVED);
    if (g != f) {
        FLG(s, j) = g;
        ROWMOD[s] = 1;
        ROWCHG[s] = 1;
    }
}
// Tiles_NOTHING (game_datastruct.cpp:312)
__device__ __forceinline__ CellR nothing(const Ctx& c) {
    CellR r;
    r.mat = (uint8_t)c.air;
    r.moved = 0;
    r.stl = 0;
    r.tmp = 0;
    r.col = 0;
    r.fl = 2.0f;
    r.fd = 0.0f;
    return r;
}
// TilesCreate(id, x, y) (game_datastruct.cpp:485-574) with the per-material policy of fse_material.
// The table walk is a rare path: kept out of line, result returned in registers (colour | temperature << 32).
__device__ __noinline__ uint64_t create_color_temp(const DevTables* T, uint32_t rkey, int m, int x, int y) {
    uint32_t col = T->color[m];
    int kind = T->ckind[m];
    if (kind == FSE_COLOR_JITTER) {
        uint32_t rr = rng_draw(rng_cell(rkey, x, y), S_CREATE_COLOR);
        uint32_t jr = T->jrange[m];
        col = col + ((rr % (jr ? jr : 1u)) << T->jshift[m]);
    } else if (kind == FSE_COLOR_POSITIONAL) {
        col = col ^ (pos_hash(x, y) & 0x0f0f0fU);
    }
    return (uint64_t)col | ((uint64_t)(uint16_t)T->ctemp[m] << 32);
}
__device__ __forceinline__ CellR create(const Ctx& c, int m, int x, int y) {
    const uint64_t ct = create_color_temp(c.T, c.rkey, m, x, y);
    CellR r;
    r.mat = (uint8_t)m;
    r.moved = 0;
    r.stl = 0;
    r.tmp = (int16_t)(uint16_t)(ct >> 32);
    r.col = (uint32_t)ct;
    r.fl = 2.0f;
    r.fd = 0.0f;
    return r;
}
// MaterialInstance(tile.mat, tile.color, tile.temperature) with fluidAmount = 0 (world.cpp:1328-1329)
__device__ __forceinline__ CellR fresh_fluid(const CellR& t) {
    CellR r;
    r.mat = t.mat;
    r.moved = 0;
    r.stl = 0;
    r.tmp = t.tmp;
    r.col = t.col;
    r.fl = 0.0f;
    r.fd = 0.0f;
    return r;
}

__device__ __forceinline__ uint64_t particle_id(const Ctx& c, int x, int y, int k) {
    return ((uint64_t)(c.tick & 0xfffff) << 42) | ((uint64_t)(c.iter & 3) << 40) | ((uint64_t)(y & 0x3ffff) << 22) |
           ((uint64_t)(x & 0x3ffff) << 4) | (uint64_t)(k & 15);
}

// cells.push_back(new CellData(tile, x, y, vx, vy, 0, ay)) (world.cpp:1110,1222,1298).  Out of line, scalar
// arguments only (no stack frame): packed = mat | moved << 8 | settle << 16, tmp in the high half.
__device__ __noinline__ void emit_particle_raw(fse_particle* pbuf, unsigned int* pcount, unsigned int pcap, uint32_t packed, uint32_t col,
                                               float fl, float fd, float px, float py, float vx, float vy, float ay, int lifetime,
                                               int fade, uint64_t id) {
    unsigned int i = atomicAdd(pcount, 1u);
    if (i >= pcap) return;  // counted; the host reports overflow
    fse_particle p;
    p.tile.mat = (uint16_t)(packed & 0xff);
    p.tile.moved = (uint8_t)((packed >> 8) & 1);
    p.tile.settle = (uint8_t)((packed >> 16) & 0xff);
    p.tile.color = col;
    p.tile.temp = (int16_t)(fade >> 16);
    p.tile.dirty = 0;
    p.tile._pad = 0;
    p.tile.fluid = fl;
    p.tile.fluid_diff = fd;
    p.x = px; p.y = py; p.vx = vx; p.vy = vy; p.ax = 0.0f; p.ay = ay;
    p.target_x = 0.0f; p.target_y = 0.0f; p.target_force = 0.0f;
    p.lifetime = lifetime;
    p.fade_time = fade & 0xffff;
    p.phase = 0;
    p.temporary = (uint8_t)((packed >> 24) & 1);
    p.in_object_state = 0;
    p.vacuum = 0;
    p._pad2 = 0;
    p.id = id;
    uint4* dst = reinterpret_cast<uint4*>(pbuf + i);
    const uint4* src = reinterpret_cast<const uint4*>(&p);
#pragma unroll
    for (int q = 0; q < (int)(sizeof(fse_particle) / 16); q++) dst[q] = src[q];
}
__device__ __forceinline__ void emit_particle(const Ctx& c, const CellR& t, float px, float py, float vx, float vy, float ay,
                                              bool temporary, int lifetime, int fade, uint64_t id) {
    uint32_t packed = (uint32_t)t.mat | ((t.moved ? 1u : 0u) << 8) | ((uint32_t)t.stl << 16) | ((temporary ? 1u : 0u) << 24);
    emit_particle_raw(c.pbuf, c.pcount, c.pcap, packed, t.col, t.fl, t.fd, px, py, vx, vy, ay, lifetime,
                      (fade & 0xffff) | ((int)(uint16_t)t.tmp << 16), id);
}

// world.cpp:1021-1034
__device__ __forceinline__ float vertical_flow(float remaining, float dest) {
    float sum = remaining + dest;
    float value;
    if (sum <= FLUID_MaxValue) {
        value = FLUID_MaxValue;
    } else if (sum < 2 * FLUID_MaxValue + FLUID_MaxCompression) {
        value = (FLUID_MaxValue * FLUID_MaxValue + sum * FLUID_MaxCompression) / (FLUID_MaxValue + FLUID_MaxCompression);
    } else {
        value = (sum + FLUID_MaxCompression) / 2.0f;
    }
    return value;
}

// canMoveBelow* (world.cpp:1206,1214-1215,1609-1610)
__device__ __forceinline__ bool can_sink(const Ctx& c, int s, int j, float myDensity) {
    uint8_t m = MAT(s, j);
    int t = LUTP->phys[m];
    return t == P_AIR || (t != P_SOLID && LUTP->dens[m] < myDensity);
}

// one liquid outflow into a neighbour (world.cpp:1324-1333 and its three siblings)
__device__ __forceinline__ void pour(const Ctx& c, int s, int j, int nbPhys, const CellR& tile, float flow) {
    if (nbPhys == P_AIR) {
        CellR n = fresh_fluid(tile);
        n.fd = flow;  // 0.0f + flow
        stc(c, s, j, n, 0);
    } else {
        FD(s, j) = FD(s, j) + flow;
        ROWMOD[s] = 1;
        ROWCHG[s] = 1;
    }
}

// tile.mat->interact && nInteractions[below.id] > 0 (world.cpp:1153), from the shared-memory partner bitmap
__device__ __forceinline__ bool has_interaction(const Ctx& c, uint8_t m, uint8_t mb) {
    const uint8_t mf = LUTP->mflags[m];
    if (!(mf & MF_INTERACT)) return false;
    if (mf & MF_INTERACT_SLOW) return c.T->inter_off[m * c.nmat + mb + 1] > c.T->inter_off[m * c.nmat + mb];
    const int r = LUTP->irow[m];
    return r != 0 && ((LUTP->ibits[r - 1][mb >> 5] >> (mb & 31)) & 1u);
}

// ---- IO: one row in, one row out ------------------------------------------------------------------------
__device__ __forceinline__ void issue_row_load_any(const TickParams& P, unsigned char* ring, unsigned long long* bars, unsigned char* rowmod,
                                                   unsigned char* rowchg, int k, int cx, int cy) {
    const int q = slot_of_row(k);
    const size_t y = (size_t)(cy + CHUNK - 1 - k);
    unsigned char* row = ring + q * ROW_BYTES;
    unsigned long long* bar = &bars[q];
    const size_t o8 = y * P.W + (cx - HX8);
    const size_t ow = y * P.W + (cx - HXW);
    rowmod[q] = 0;
    rowchg[q] = 0;
    if (k < -HALO_WR) {  // probe-only rows: pass 2 reads material types down to y+10, nothing else
        mbar_expect_tx(bar, P8);
        bulk_g2s(row + OFF_MAT, P.p.mat + o8, P8, bar);
        return;
    }
    mbar_expect_tx(bar, ROW_BYTES);
    bulk_g2s(row + OFF_MAT, P.p.mat + o8, P8, bar);
    bulk_g2s(row + OFF_FLG, P.p.flg + o8, P8, bar);
    bulk_g2s(row + OFF_STL, P.p.stl + o8, P8, bar);
    bulk_g2s(row + OFF_TMP, P.p.tmp + ow, PW * 2, bar);
    bulk_g2s(row + OFF_COL, P.p.col + ow, PW * 4, bar);
    bulk_g2s(row + OFF_FL, P.p.fl + ow, PW * 4, bar);
    bulk_g2s(row + OFF_FD, P.p.fd + ow, PW * 4, bar);
}

__device__ __forceinline__ void issue_row_store_any(const TickParams& P, unsigned char* ring, int k, int cx, int cy) {
    const int q = slot_of_row(k);
    const size_t y = (size_t)(cy + CHUNK - 1 - k);
    unsigned char* row = ring + q * ROW_BYTES;
    const size_t o8 = y * P.W + (cx - HX8);
    const size_t ow = y * P.W + (cx - HXW);
    bulk_s2g(P.p.mat + o8, row + OFF_MAT, P8);
    bulk_s2g(P.p.flg + o8, row + OFF_FLG, P8);
    bulk_s2g(P.p.stl + o8, row + OFF_STL, P8);
    bulk_s2g(P.p.tmp + ow, row + OFF_TMP, PW * 2);
    bulk_s2g(P.p.col + ow, row + OFF_COL, PW * 4);
    bulk_s2g(P.p.fl + ow, row + OFF_FL, PW * 4);
    bulk_s2g(P.p.fd + ow, row + OFF_FD, PW * 4);
}

// ---- active-region tracking (new behaviour whose contract is "identical to a full sweep") --------------
// A chunk may be skipped only if processing it again cannot change anything for ANY random draw.  When a finished row
// leaves the ring the IO warp checks that each of its 128 cells is inert:
//   AIR / SOLID / PASSABLE-but-not-FIRE : no rule applies;
//   SAND : cannot sink below nor into either lower diagonal (world.cpp:1206, 1609-1610), moved == false (so pass 2 leaves
//          it alone, 1647-1654), no pair interaction armed, its temperature reaction (if any) not firing;
//   SOUP : settled (moved, 1307), fluidAmountDiff == 0, amount >= FLUID_MinValue, not over AIR (1283);
//   GAS, FIRE : never inert (they roll dice every visit).
// lane handles columns 4*lane .. 4*lane+3 of row slot q (qb = the row below).
__device__ bool row_is_inert(const Ctx& c, int q, int qb, int lane) {
    bool inert = true;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const int j = HX8 + 4 * lane + b;
        const uint8_t m = MAT(q, j);
        const int ph = LUTP->phys[m];
        if (ph == P_AIR || ph == P_SOLID) continue;
        if (ph == P_PASSABLE) {
            inert &= (int)m != c.fire;
        } else if (ph == P_SAND) {
            const float d = LUTP->dens[m];
            const uint8_t mf = LUTP->mflags[m];
            bool ok = !(FLG(q, j) & F_MOVED) && !can_sink(c, qb, j, d) && !can_sink(c, qb, j - 1, d) && !can_sink(c, qb, j + 1, d);
            if (mf & MF_INTERACT) ok = ok && !has_interaction(c, m, MAT(qb, j));
            if (mf & MF_REACT) {
                if (mf & MF_REACT_MULTI) {
                    ok = false;
                } else {
                    const Lut::Rx rx = LUTP->rx[m];
                    const int16_t t = TMP(q, j);
                    ok = ok && !((rx.type == FSE_REACT_TEMPERATURE_BELOW && t < rx.thr) || (rx.type == FSE_REACT_TEMPERATURE_ABOVE && t > rx.thr));
                }
            }
            inert &= ok;
        } else if (ph == P_SOUP) {
            inert &= (FLG(q, j) & F_MOVED) && FD(q, j) == 0.0f && FL(q, j) >= FLUID_MinValue && PHYS(qb, j) != P_AIR;
        } else {
            inert = false;
        }
    }
    return inert;
}

// ---- active-chunk compaction: awake chunks of one colour -> dense list + count (warp-aggregated append) --------------
__global__ void compact_active_kernel(const uint8_t* __restrict__ awake, int acols, int ci0, int cj0, int ncx, int ncy, int* __restrict__ list,
                                      int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = ncx * ncy;
    bool on = false;
    int v = 0;
    if (i < n) {
        const int cxi = i % ncx, cyi = i / ncx;
        on = awake[(cj0 + 2 * cyi) * acols + (ci0 + 2 * cxi)] != 0;
        v = cxi | (cyi << 16);
    }
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (on) list[base + __popc(m & ((1u << lane) - 1))] = v;
}

cudaError_t launch_compact_active(const uint8_t* awake, int acols, int ci0, int cj0, int ncx, int ncy, int* list, int* count, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    const int n = ncx * ncy;
    compact_active_kernel<<<(n + 255) / 256, 256, 0, s>>>(awake, acols, ci0, cj0, ncx, ncy, list, count);
    return cudaGetLastError();
}

}  // namespace fse

#include "fse_tick_rows.cuh"

namespace fse {

// ---- host launcher ----------------------------------------------------------------------------------
size_t tick_smem_bytes() { return sizeof(SmemRows); }

cudaError_t launch_tick_phase(const TickParams& P, int n_chunks, cudaStream_t stream, int* launched, const TickFork* fork) {
    *launched = 0;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tick_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemRows));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tick_pass_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tick_pass_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tick_pass_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tick_pass_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (n_chunks <= 0) return cudaSuccess;
    // The fused kernel keeps 2 chunks per SM in flight and its chain is the longest pass, not the sum of the three: a phase that
    // fits one wave of it (small worlds, the cut-adjacent chunk rows of a strip) is faster there.  Results are identical.
    const bool small = !P.awake && n_chunks <= P.fused_max_chunks;
    if (P.schedule == FSE_SCHEDULE_ROWS && !P.fused && !small && (!P.awake || P.chunk_state)) {  // one kernel per pass
        // The three passes of a chunk only depend on each other, so the phase is cut into parts that run on their own streams:
        // while the last pass-1 CTAs of one part drain, another part's pass 2 already fills the SMs (the kernels are latency
        // bound and a phase is only ~1.3 waves of CTAs).
        static const int pad = getenv("FSE_PASS_SMEM_PAD") ? atoi(getenv("FSE_PASS_SMEM_PAD")) : 0;  // occupancy experiments
        const int parts = (fork && n_chunks >= fork->min_chunks) ? fork->parts : 1;
        TickParams Q = P;
        if (parts > 1) {
            cudaEventRecord(fork->ev_fork, stream);
            for (int q = 1; q < parts; q++) cudaStreamWaitEvent(fork->aux[q - 1], fork->ev_fork, 0);
        }
        for (int q = 0; q < parts; q++) {
            cudaStream_t st = q ? fork->aux[q - 1] : stream;
            // slices of the chunk list; with a dealt-out longest-first list (lpt_build_kernel) they are the parts it was dealt into
            const bool dealt = P.lpt_parts == parts && parts > 1;
            const int lo = dealt ? part_lo(n_chunks, parts, q) : (int)((long long)n_chunks * q / parts);
            const int hi = dealt ? part_lo(n_chunks, parts, q + 1) : (int)((long long)n_chunks * (q + 1) / parts);
            if (hi <= lo) continue;  // fewer chunks than parts
            Q.chunk_base = P.chunk_base + lo;
            if (Q.rowmask && !Q.split) {
                classify_rows_kernel<<<(hi - lo) * 4, 1024, sizeof(Lut), st>>>(Q);
                *launched += 1;
            }
            if (Q.rowmask && Q.split) {
                // no settled-row skipping in pass 1, but pass 1 sorts the rows for pass 2: liquid-only rows are applied by a row-parallel
                // kernel, and pass 2 steps only the rows that still hold powder or gas (its row-skipping instantiation, fed by pass 1)
                tick_pass_kernel<1, false><<<hi - lo, PassGeom<1>::THREADS, sizeof(SmemPass<1>) + pad, st>>>(Q);
                tick_pass2_apply_kernel<<<(hi - lo) * (CHUNK / 4), 128, 0, st>>>(Q);
                tick_pass_kernel<2, true><<<hi - lo, PassGeom<2>::THREADS, sizeof(SmemPass<2>) + pad, st>>>(Q);
                *launched += 1;
            } else if (Q.rowmask) {
                tick_pass_kernel<1, true><<<hi - lo, PassGeom<1>::THREADS, sizeof(SmemPass<1>) + pad, st>>>(Q);
                tick_pass_kernel<2, true><<<hi - lo, PassGeom<2>::THREADS, sizeof(SmemPass<2>) + pad, st>>>(Q);
            } else {
                tick_pass_kernel<1, false><<<hi - lo, PassGeom<1>::THREADS, sizeof(SmemPass<1>) + pad, st>>>(Q);
                tick_pass_kernel<2, false><<<hi - lo, PassGeom<2>::THREADS, sizeof(SmemPass<2>) + pad, st>>>(Q);
            }
            tick_pass3_kernel<<<(hi - lo) * (CHUNK / 4), 128, 0, st>>>(Q);
            *launched += 3;
        }
        for (int q = 1; q < parts; q++) {
            cudaEventRecord(fork->ev_join[q - 1], fork->aux[q - 1]);
            cudaStreamWaitEvent(stream, fork->ev_join[q - 1], 0);
        }
        if (P.awake && P.chunk_state && P.chunk_list) {  // a sleeping chunk can only be woken by a neighbour's record, so the
            apply_chunk_state_kernel<<<(n_chunks + 127) / 128, 128, 0, stream>>>(P, n_chunks);  // flags change after all passes
            *launched += 1;
        }
    } else {
        tick_rows_kernel<<<n_chunks, ROWS_THREADS, sizeof(SmemRows), stream>>>(P);
        *launched = 1;
    }
    return cudaGetLastError();
}

cudaError_t launch_lpt_build(const unsigned int* cost, int n, int ncx, int* list, cudaStream_t stream, const int* members, int parts) {
    lpt_build_kernel<<<1, 1024, 0, stream>>>(cost, n, ncx, list, members, parts);
    return cudaGetLastError();
}

}  // namespace fse
