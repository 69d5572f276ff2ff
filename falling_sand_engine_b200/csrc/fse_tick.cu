// fse_tick.cu — the per-chunk cell update of world::tick() (reference: source/engine/world.cpp:1036-1948)
// as a hand-written sm_100a kernel.
//
// One CTA = one 128x128 chunk of one colour phase (world.cpp:1057-1077).  The chunk is streamed
// bottom-up through a 28-row shared-memory ring: rows enter by cp.async.bulk (TMA bulk copy engine,
// SASS UBLKCP) completing on per-slot mbarriers, and leave by cp.async.bulk shared->global once the
// last pass is done with them.  Warps 0/1/2 run the reference's pass 1/2/3 software-pipelined 7 and 11
// rows apart (exactly equivalent to running the passes one after another); warp 3 is the
// IO warp.  Inside a row a warp visits the 128 columns as 4 interleaved classes x = 4*lane + c, so the
// 32 cells processed together are >= 4 columns apart and every +-1-column rule commutes.
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

struct __align__(128) Smem {
    SmemHead h;
    unsigned char ring[RING * ROW_BYTES];
    unsigned long long bar[RING];
};

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
__device__ __forceinline__ int rs(int s, int dy) {  // slot of the row dy below (+) / above (-) the row in slot s
    int v = s - dy;
    if (v < 0) v += RING;
    if (v >= RING) v -= RING;
    return v;
}
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
// real_tiles[i] = cell (+ dirty[i] / tickVisited[i] = true as requested by `setbits`).  Out of line on purpose: the rule
// code stores cells at ~40 places and the kernel is instruction-cache bound (profiles/r1_tick_ncu.md), so one copy of the
// store sequence beats 40 inlined ones.  p0 = mat | stl << 8 | tmp << 16, p1 = moved | setbits << 8.
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
__device__ __forceinline__ void set_bits(const Ctx& c, int s, int j, uint8_t bits) {  // dirty and/or visited
    uint8_t f = FLG(s, j);
    FLG(s, j) = f | bits;
    if ((bits & F_DIRTY) && !(f & F_DIRTY)) ROWMOD[s] = 1;
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

// ---- FIRE (world.cpp:1101-1146), one fire cell handled by the whole warp: lane i < 25 owns neighbour
// (xx, yy) = (i / 5 - 2, i % 5 - 2), the reference's loop order, so the RNG slots are identical.  Must be
// called convergently; (s, jf, xf) are warp-uniform.
__device__ void fire_coop(const Ctx& c, int s, int jf, int xf, int y, int lane) {
    const uint8_t f0 = FLG(s, jf);
    if (f0 & F_VISITED) return;  // 1091
    if (c.iter >= (int)LUTP->iters[c.fire]) {  // 1093-1096
        if (lane == 0) FLG(s, jf) = f0 | F_VISITED;
        return;
    }
    const uint32_t cb = rng_cell(c.rkey, xf, y);
    // 1102-1107 edits a local copy that is never stored
    if (lane == 0 && rng_draw(cb, S_FIRE_EMBER) % 10 == 0) {  // 1109-1119
        CellR tile = ldc(c, s, jf);
        float vx = ((int)(rng_draw(cb, S_FIRE_EMBER_VX) % 10) - 5) / 20.0f;
        float vy = -((int)(rng_draw(cb, S_FIRE_EMBER_VY) % 10) / 10.0f) / 3.0f + -0.5f;
        emit_particle(c, tile, (float)xf, (float)(y - 1), vx, vy, 0.01f, true, 30, 10, particle_id(c, xf, y, 15));
    }
    if (rng_draw(cb, S_FIRE_DIE) % 150 == 0) {  // 1121-1125
        if (lane == 0) stc(c, s, jf, nothing(c), F_DIRTY | F_VISITED);
        return;
    }
    bool solid = false;  // 1127-1144
    int xx = 0, yy = 0, s2 = s;
    if (lane < 25) {
        xx = lane / 5 - 2;
        yy = lane % 5 - 2;
        s2 = rs(s, yy);
        solid = PHYS(s2, jf + xx) == P_SOLID;
    }
    const bool foundAny = __any_sync(0xffffffffu, solid);
    if (solid && rng_draw(cb, S_FIRE_IGNITE0 + lane) % 500 == 0)
        stc(c, s2, jf + xx, create(c, c.fire, xf + xx, y + yy), F_DIRTY | F_VISITED);
    if (!foundAny && lane == 0 && rng_draw(cb, S_FIRE_DIE_ALONE) % 120 == 0) stc(c, s, jf, nothing(c), F_DIRTY | F_VISITED);
}

// ---- pass 1: one cell (world.cpp:1089-1586) ----------------------------------------------------------
__device__ void visit1(const Ctx& c, int s, int j, int x, int y) {
    uint8_t f0 = FLG(s, j);
    if (f0 & F_VISITED) return;  // 1091
    const uint8_t m = MAT(s, j);
    if (c.iter >= (int)LUTP->iters[m]) {  // 1093-1096
        FLG(s, j) = f0 | F_VISITED;
        return;
    }
    const int type = LUTP->phys[m];
    if (type == P_AIR || type == P_SOLID) return;  // no rule matches (FIRE is PASSABLE)
    const uint32_t cb = rng_cell(c.rkey, x, y);
    const int sb = rs(s, 1);  // row below

    if (m == c.fire) return;  // FIRE cells are handled warp-cooperatively by fire_coop() (pass1_row)

    if (type == P_SAND) {  // 1148-1267
        const uint8_t mb = MAT(sb, j);
        const int below = LUTP->phys[mb];
        const uint8_t mf = LUTP->mflags[m];

        if ((mf & MF_INTERACT) && has_interaction(c, m, mb)) {  // 1153-1179
            const int n = c.nmat;
            int lo = c.T->inter_off[m * n + mb], hi = c.T->inter_off[m * n + mb + 1];
            {
                for (int i = lo; i < hi; i++) {
                    fse_interaction in = c.T->inter[i];
                    int rad = (int)in.data2;
                    if (in.type == FSE_INTERACT_TRANSFORM_MATERIAL) {
                        for (int xx = in.ofs_x - rad; xx <= in.ofs_x + rad; xx++)
                            for (int yy = in.ofs_y - rad; yy <= in.ofs_y + rad; yy++) {
                                int s2 = rs(s, yy);
                                if (MAT(s2, j + xx) == mb) stc(c, s2, j + xx, create(c, in.data1, x + xx, y + yy), F_DIRTY | F_VISITED);
                            }
                    } else if (in.type == FSE_INTERACT_SPAWN_MATERIAL) {
                        for (int xx = in.ofs_x - rad; xx <= in.ofs_x + rad; xx++)
                            for (int yy = in.ofs_y - rad; yy <= in.ofs_y + rad; yy++) {
                                int s2 = rs(s, yy);
                                if ((xx == 0 && yy == 0) || MAT(s2, j + xx) == c.air)
                                    stc(c, s2, j + xx, create(c, in.data1, x + xx, y + yy), F_DIRTY | F_VISITED);
                            }
                    }
                }
                return;  // 1178
            }
        }
        if (mf & MF_REACT) {  // 1181-1204
            bool react = false;
            const int16_t temp = TMP(s, j);
            if (!(mf & MF_REACT_MULTI)) {
                const Lut::Rx rx = LUTP->rx[m];
                bool hit = (rx.type == FSE_REACT_TEMPERATURE_BELOW && temp < rx.thr) || (rx.type == FSE_REACT_TEMPERATURE_ABOVE && temp > rx.thr);
                if (hit) {
                    CellR n = create(c, rx.prod, x, y);
                    n.tmp = temp;
                    stc(c, s, j, n, F_DIRTY | F_VISITED);
                    react = true;
                }
            } else {
                for (int i = c.T->react_off[m]; i < c.T->react_off[m + 1]; i++) {
                    fse_interaction in = c.T->react[i];
                    bool hit = (in.type == FSE_REACT_TEMPERATURE_BELOW && temp < in.data1) ||
                               (in.type == FSE_REACT_TEMPERATURE_ABOVE && temp > in.data1);
                    if (hit) {
                        CellR n = create(c, (int)in.data2, x, y);
                        n.tmp = temp;
                        stc(c, s, j, n, F_DIRTY | F_VISITED);
                        react = true;
                    }
                }
            }
            if (react) return;
        }
        const float myDens = LUTP->dens[m];
        bool canMoveBelow = (below == P_AIR || (below != P_SOLID && LUTP->dens[mb] < myDens));  // 1206
        if (!canMoveBelow) return;
        bool canL = can_sink(c, sb, j - 1, myDens);
        bool canR = can_sink(c, sb, j + 1, myDens);
        if ((canL || canR) && rng_draw(cb, S_SAND_HESITATE) % 20 == 0) return;  // 1217
        CellR tile = ldc(c, s, j);
        CellR belowTile = ldc(c, sb, j);
        if (below == P_AIR && PHYS(rs(s, 2), j) == P_AIR && PHYS(rs(s, 3), j) == P_AIR && PHYS(rs(s, 4), j) == P_AIR) {
            // 1218-1225: free fall -> loose particle; the cell takes a copy of the air below
            stc(c, s, j, belowTile, F_DIRTY);
            float vx = ((int)(rng_draw(cb, S_SAND_PART_VX) % 10) - 5) / 20.0f;
            float vy = -((int)(rng_draw(cb, S_SAND_PART_VY) % 2) + 3) / 10.0f + 1.5f;
            emit_particle(c, tile, (float)x, (float)(y + 1), vx, vy, 0.1f, false, 0, 60, particle_id(c, x, y, 14));
        } else {  // 1227-1239: swap with the cell below
            stc(c, s, j, belowTile, F_DIRTY);
            if (rng_draw(cb, S_SAND_MOVED) % 2 == 0) tile.moved = 1;
            stc(c, sb, j, tile, F_DIRTY | F_VISITED);
        }
        if (rng_draw(cb, S_SAND_TX_SELF) % 2 == 0) {  // 1242-1266
            if (x > 0 && PHYS(sb, j - 1) == P_SAND) {
                if (rng_draw(cb, S_SAND_TX_L) % 2 == 0) set_moved(c, sb, j - 1, true);
            }
            if (PHYS(sb, j + 1) == P_SAND) {
                if (rng_draw(cb, S_SAND_TX_R) % 2 == 0) set_moved(c, sb, j + 1, true);
            }
        }
    } else if (type == P_SOUP) {  // 1269-1568
        CellR tile = ldc(c, s, j);
        if (tile.fl == 0.0f) return;  // 1275
        if (tile.fl < FLUID_MinValue) {  // 1277-1281
            FL(s, j) = 0.0f;
            ROWMOD[s] = 1;
            ROWCHG[s] = 1;
            return;
        }
        const uint8_t mb0 = MAT(sb, j);
        const int bottomPhys = LUTP->phys[mb0];
        if ((double)tile.fl > 0.005 && bottomPhys == P_AIR && PHYS(rs(s, 2), j) == P_AIR && PHYS(rs(s, 3), j) == P_AIR &&
            PHYS(rs(s, 4), j) == P_AIR) {  // 1283-1305
            stc(c, s, j, nothing(c), F_DIRTY);
            int n = (int)(tile.fl / 4);
            if (n < 1) n = 1;
            for (int i = 0; i < n; i++) {
                CellR nt = fresh_fluid(tile);
                nt.fl = tile.fl / n;
                float vx = ((int)(rng_draw(cb, S_SOUP_PART0 + 2 * (i & 7)) % 10) - 5) / 30.0f;
                float vy = -((int)(rng_draw(cb, S_SOUP_PART0 + 2 * (i & 7) + 1) % 2) + 3) / 10.0f + 1.0f;
                emit_particle(c, nt, (float)x, (float)(y + 1), vx, vy, 0.1f, false, 0, 60, particle_id(c, x, y, i & 7));
            }
            return;
        }
        if (tile.moved) return;  // 1307: settled

        const float startValue = tile.fl;
        float remainingValue = tile.fl;
        const float bottomFl = FL(sb, j);
        const bool airBelow = bottomPhys == P_AIR;

        if ((airBelow && c.iter <= 2) || mb0 == m) {  // 1315-1334
            float dstFl = bottomPhys == P_SOUP ? bottomFl : 0.0f;
            float flow = vertical_flow(startValue, dstFl) - dstFl;
            if (bottomFl > 0 && flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = fmaxf(flow, 0.0f);
            if (flow > fminf(FLUID_MaxFlow, startValue)) flow = fminf(FLUID_MaxFlow, startValue);
            if (flow != 0) {
                remainingValue -= flow;
                tile.fd -= flow;
                pour(c, sb, j, bottomPhys, tile, flow);
            }
        } else if (c.iter == 0 && bottomPhys == P_SOUP && mb0 != m) {  // 1335-1341
            if (rng_draw(cb, S_SOUP_SWAP_DOWN) % 10 == 0) {
                CellR bottom = ldc(c, sb, j);
                stc(c, s, j, bottom, 0);
                stc(c, sb, j, tile, 0);
                return;
            }
        }
        if (remainingValue < FLUID_MinValue) {  // 1343-1347
            tile.fd -= remainingValue;
            stc(c, s, j, tile, 0);
            return;
        }
        const uint8_t ml = MAT(s, j - 1), mr = MAT(s, j + 1);
        const int leftPhys = LUTP->phys[ml], rightPhys = LUTP->phys[mr];
        const float leftFl = FL(s, j - 1), rightFl = FL(s, j + 1);
        const bool canMoveLeft = (leftPhys == P_AIR || ml == m) && !airBelow;    // 1350
        const bool canMoveRight = (rightPhys == P_AIR || mr == m) && !airBelow;  // 1353
        if (canMoveLeft) {  // 1355-1375
            float dstFl = leftPhys == P_SOUP ? leftFl : 0.0f;
            float flow = (remainingValue - dstFl) / (canMoveRight ? 3.0f : 2.0f);
            if (flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = fmaxf(flow, 0.0f);
            if (flow > fminf(FLUID_MaxFlow, remainingValue)) flow = fminf(FLUID_MaxFlow, remainingValue);
            if (flow != 0) {
                remainingValue -= flow;
                tile.fd -= flow;
                pour(c, s, j - 1, leftPhys, tile, flow);
            }
        }
        if (remainingValue < FLUID_MinValue) {  // 1377-1381
            tile.fd -= remainingValue;
            stc(c, s, j, tile, 0);
            return;
        }
        if (canMoveRight) {  // 1383-1403 (divisor is 2.0f in both arms)
            float dstFl = rightPhys == P_SOUP ? rightFl : 0.0f;
            float flow = (remainingValue - dstFl) / 2.0f;
            if (flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = fmaxf(flow, 0.0f);
            if (flow > fminf(FLUID_MaxFlow, remainingValue)) flow = fminf(FLUID_MaxFlow, remainingValue);
            if (flow != 0) {
                remainingValue -= flow;
                tile.fd -= flow;
                pour(c, s, j + 1, rightPhys, tile, flow);
            }
        }
        if (remainingValue < FLUID_MinValue) {  // 1405-1409
            tile.fd -= remainingValue;
            stc(c, s, j, tile, 0);
            return;
        }
        const int st = rs(s, -1);  // row above
        const uint8_t mt = MAT(st, j);
        const int topPhys = LUTP->phys[mt];
        if (topPhys == P_AIR || mt == m) {  // 1413-1432
            float dstFl = topPhys == P_SOUP ? FL(st, j) : 0.0f;
            float flow = remainingValue - vertical_flow(remainingValue, dstFl);
            if (flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = fmaxf(flow, 0.0f);
            if (flow > fminf(FLUID_MaxFlow, remainingValue)) flow = fminf(FLUID_MaxFlow, remainingValue);
            if (flow != 0) {
                remainingValue -= flow;
                tile.fd -= flow;
                pour(c, st, j, topPhys, tile, flow);
            }
        } else if (c.iter == 0 && topPhys == P_SOUP && mt != m) {  // 1433-1439
            if (rng_draw(cb, S_SOUP_SWAP_UP) % 10 == 0) {
                CellR top = ldc(c, st, j);
                stc(c, s, j, top, 0);
                stc(c, st, j, tile, 0);
                return;
            }
        }
        if (remainingValue < FLUID_MinValue) {  // 1441-1445
            tile.fd -= remainingValue;
            stc(c, s, j, tile, 0);
            return;
        }
        uint8_t bits = 0;
        if (startValue == remainingValue) {  // 1447-1451
            tile.stl = (uint8_t)(tile.stl + 1);
            if (tile.stl >= 10) tile.moved = 1;
        } else {  // 1452-1458: un-settle the liquid neighbours (types as read before the flows)
            bits = F_DIRTY;
            if (topPhys == P_SOUP) set_moved(c, st, j, false);
            if (bottomPhys == P_SOUP) set_moved(c, sb, j, false);
            if (leftPhys == P_SOUP) set_moved(c, s, j - 1, false);
            if (rightPhys == P_SOUP) set_moved(c, s, j + 1, false);
        }
        stc(c, s, j, tile, bits);  // 1460

        if (m == c.water && MAT(sb, j) == c.lava) {  // 1519-1537
            stc(c, s, j, create(c, c.steam, x, y), F_DIRTY);
            stc(c, sb, j, create(c, c.obsidian, x, y + 1), F_DIRTY | F_VISITED);
            for (int xx = -1; xx <= 1; xx++)
                for (int yy = 0; yy <= 2; yy++) {
                    int s2 = rs(s, yy);
                    if (MAT(s2, j + xx) == c.lava) stc(c, s2, j + xx, create(c, c.obsidian, x + xx, y + yy), F_DIRTY | F_VISITED);
                }
        }
    } else if (type == P_GAS) {  // 1569-1585
        const int st = rs(s, -1);
        int above = PHYS(st, j), aboveL = PHYS(st, j - 1), aboveR = PHYS(st, j + 1);
        if (above == P_AIR && !((aboveL == P_AIR || aboveR == P_AIR) && rng_draw(cb, S_GAS1) % 2 == 0)) {
            CellR tile = ldc(c, s, j);
            CellR up = ldc(c, st, j);
            stc(c, s, j, up, F_DIRTY);
            stc(c, st, j, tile, F_DIRTY | F_VISITED);
        }
    }
}

// ---- pass 2: one cell (world.cpp:1594-1820) ----------------------------------------------------------
__device__ void visit2(const Ctx& c, int s, int j, int x, int y) {
    const uint8_t f0 = FLG(s, j);
    if (f0 & F_VISITED) return;  // 1596
    const uint8_t m = MAT(s, j);
    const int type = LUTP->phys[m];
    if (type == P_SAND) {  // 1602-1727
        const uint32_t cb = rng_cell(c.rkey, x, y);
        const int sb = rs(s, 1);
        const float myDens = LUTP->dens[m];
        const bool canL = can_sink(c, sb, j - 1, myDens);
        const bool canR = can_sink(c, sb, j + 1, myDens);
        bool stoppedByFriction = !(f0 & F_MOVED);  // 1612
        const int slip = LUTP->slip[m];
        bool nowMoved = (f0 & F_MOVED) != 0;  // real_tiles[idx].moved (the local `tile` copy keeps the old flag)
        if (!(canL || canR)) {
            // 1647-1654 fires whatever the pillar probe decides (an un-stick at 1637 is overwritten at 1648)
            set_moved(c, s, j, false);
            return;
        }
        if (stoppedByFriction) {  // 1617-1645
            int drop = 0;
            for (int pil = 0; pil < 10; pil++) {
                int sp = rs(s, 1 + pil);
                if (PHYS(sp, j - 1) == P_AIR || PHYS(sp, j + 1) == P_AIR) drop++;
            }
            int d = drop + 1 - (int)LUTP->maxstab[m];
            if (d > 0) {
                int chance = 1000 / d;
                if (chance < 1000 && rng_draw(cb, S_SAND2_UNSTICK) % chance == 0) {
                    stoppedByFriction = false;
                    nowMoved = true;
                }
            }
        }
        if (stoppedByFriction || !(canL || canR)) {  // 1647-1654
            set_moved(c, s, j, false);
            return;
        }
        if (nowMoved != ((f0 & F_MOVED) != 0)) set_moved(c, s, j, nowMoved);
        const bool shouldMove = rng_draw(cb, S_SAND2_SHOULD) % (2 * slip) != 0;  // 1656
        if (shouldMove && (canL || canR)) {  // 1658-1673
            if (rng_draw(cb, S_SAND2_TX_SELF) % 2 == 0) {
                if (PHYS(sb, j) == P_SAND) {
                    if (rng_draw(cb, S_SAND2_TX_OTHER) % 2 == 0) set_moved(c, sb, j, true);
                }
            }
        }
        const bool goL = shouldMove && canL && (!canR || rng_draw(cb, S_SAND2_LR) % 2 == 0);  // 1675
        const bool goR = !goL && shouldMove && canR;                                          // 1698
        if (goL || goR) {
            const int jd = goL ? j - 1 : j + 1;
            CellR tile = ldc(c, s, j);
            tile.moved = (f0 & F_MOVED) ? 1 : 0;  // the by-value copy taken at 1598
            CellR diag = ldc(c, sb, jd);
            if (PHYS(s, jd) == P_AIR) {
                // the displaced diagonal cell rises beside us; left slide marks it visited, right slide does not (1679 vs 1700-1704)
                stc(c, s, jd, diag, goL ? (uint8_t)(F_DIRTY | F_VISITED) : F_DIRTY);
                stc(c, s, j, nothing(c), F_DIRTY);
            } else {
                stc(c, s, j, diag, F_DIRTY | F_VISITED);
            }
            if (rng_draw(cb, S_SAND2_RESTICK) % (20 * slip) == 0) tile.moved = 0;  // 1688 / 1711
            stc(c, sb, jd, tile, F_DIRTY | F_VISITED);
        } else {
            set_moved(c, s, j, false);  // 1721-1727
        }
    } else if (type == P_SOUP) {  // 1728-1745
        float a = FL(s, j) + FD(s, j);
        if (a < FLUID_MinValue) {
            stc(c, s, j, nothing(c), F_DIRTY | F_VISITED);
        } else {
            const float fd = FD(s, j);
            FL(s, j) = a;
            FD(s, j) = 0.0f;
            FLG(s, j) = f0 | F_DIRTY | F_VISITED;
            ROWMOD[s] = 1;
            if (fd != 0.0f) ROWCHG[s] = 1;  // amount += 0 leaves the cell as it was (only dirty[] is re-set)
        }
    } else if (type == P_GAS) {  // 1799-1819
        const int st = rs(s, -1);
        int aboveL = PHYS(st, j - 1), aboveR = PHYS(st, j + 1);
        int jd = 0;
        if (aboveL == P_AIR && !(aboveR == P_AIR && rng_draw(rng_cell(c.rkey, x, y), S_GAS2) % 2 == 0))
            jd = j - 1;
        else if (aboveR == P_AIR)
            jd = j + 1;
        if (jd) {
            CellR tile = ldc(c, s, j);
            CellR other = ldc(c, st, jd);
            stc(c, s, j, other, F_DIRTY);
            stc(c, st, jd, tile, F_DIRTY | F_VISITED);
        }
    }
}

// ---- pass 3: one cell (world.cpp:1828-1891) ----------------------------------------------------------
__device__ void visit3(const Ctx& c, int s, int j, int x, int y) {
    if (FLG(s, j) & F_VISITED) return;  // 1830
    const uint8_t m = MAT(s, j);
    if (LUTP->phys[m] != P_GAS) return;
    int l = PHYS(s, j - 1), r = PHYS(s, j + 1);
    const uint32_t cb = rng_cell(c.rkey, x, y);
    int jd = 0;
    if (l == P_AIR && !(r == P_AIR && rng_draw(cb, S_GAS3) % 2 == 0))
        jd = j - 1;
    else if (r == P_AIR)
        jd = j + 1;
    if (jd) {
        CellR tile = ldc(c, s, j);
        CellR other = ldc(c, s, jd);
        stc(c, s, j, other, F_DIRTY);
        stc(c, s, jd, tile, F_DIRTY | F_VISITED);
    } else if (m == c.steam) {  // 1883-1888
        if (rng_draw(cb, S_STEAM) % 10 == 0) stc(c, s, j, create(c, c.water, x, y), F_DIRTY);
    }
}

// ---- one row of each pass ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_word(const unsigned char* p) { return *reinterpret_cast<const uint32_t*>(p); }

__device__ void pass1_row(const Ctx& c, int k, int cx, int cy, int lane) {
    const int s = slot_of_row(k);
    const int y = cy + c.yoff + CHUNK - 1 - k;  // global row (RNG keys, particle positions)
    const int jw = HX8 + 4 * lane;
    // row-level vote: does any cell of this row act in pass 1, and does the row hold FIRE / interacting powders?
    uint32_t mw = ld_word(&MAT(s, jw));
    uint32_t fw = ld_word(&FLG(s, jw));
    uint32_t gate = 0;
    bool act = false, spec = false;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        uint32_t m = (mw >> (8 * b)) & 0xff;
        bool vis = (fw >> (8 * b)) & F_VISITED;
        bool gated = c.iter >= (int)LUTP->iters[m];
        int ph = LUTP->phys[m];
        if (!vis && gated) gate |= (uint32_t)F_VISITED << (8 * b);
        const bool fire = (int)m == c.fire;
        if (!vis && !gated && (ph == P_SAND || ph == P_SOUP || ph == P_GAS || fire)) act = true;
        if (fire || (LUTP->mflags[m] & MF_INTERACT)) spec = true;
    }
    if (!__any_sync(0xffffffffu, act)) {
        // inert row: the only effect of pass 1 is tickVisited = true on cells past their iteration count (1093-1096)
        if (gate) *reinterpret_cast<uint32_t*>(&FLG(s, jw)) = fw | gate;
        return;
    }
    const bool special_row = __any_sync(0xffffffffu, spec);
    const int sb = rs(s, 1);
    for (int cc = 0; cc < 4; cc++) {
        const int j = jw + cc;
        const int x = cx + 4 * lane + cc;
        if (!special_row) {  // no FIRE and no interacting powder can exist in this row during this step
            visit1(c, s, j, x, y);
            __syncwarp();
            continue;
        }
        // classify at the start of the sub-step
        int phase = 0;
        const uint8_t m = MAT(s, j);
        if ((int)m == c.fire) {
            phase = 1 + (lane & 1);
        } else if (LUTP->phys[m] == P_SAND && has_interaction(c, m, MAT(sb, j))) {
            phase = 3 + (lane % 3);
        }
        const unsigned special = __ballot_sync(0xffffffffu, phase != 0);
        if (phase == 0) visit1(c, s, j, x, y);
        __syncwarp();
        if (special) {
            for (int ph = 1; ph <= 2; ph++) {  // FIRE cells: even lanes, then odd lanes (footprints +-2 columns, 8 apart)
                unsigned fm = __ballot_sync(0xffffffffu, phase == ph);
                while (fm) {
                    const int src = __ffs(fm) - 1;
                    fm &= fm - 1;
                    fire_coop(c, s, jw - 4 * lane + 4 * src + cc, cx + 4 * src + cc, y, lane);
                    __syncwarp();
                }
            }
            for (int ph = 3; ph < 6; ph++) {  // interacting powders: lanes l % 3 (footprints +-5 columns, 12 apart)
                if (__ballot_sync(0xffffffffu, phase == ph)) {
                    if (phase == ph) visit1(c, s, j, x, y);
                    __syncwarp();
                }
            }
        }
    }
}

__device__ void pass2_row(const Ctx& c, int k, int cx, int cy, int lane) {
    const int s = slot_of_row(k);
    const int y = cy + c.yoff + CHUNK - 1 - k;  // global row (RNG keys, particle positions)
    const int jw = HX8 + 4 * lane;
    uint32_t mw = ld_word(&MAT(s, jw));
    uint32_t fw = ld_word(&FLG(s, jw));
    bool act = false;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        int ph = LUTP->phys[(mw >> (8 * b)) & 0xff];
        bool vis = (fw >> (8 * b)) & F_VISITED;
        if (!vis && (ph == P_SAND || ph == P_SOUP || ph == P_GAS)) act = true;
    }
    if (!__any_sync(0xffffffffu, act)) return;
    for (int cc = 0; cc < 4; cc++) {
        visit2(c, s, jw + cc, cx + 4 * lane + cc, y);
        __syncwarp();
    }
}

__device__ void pass3_row(const Ctx& c, int k, int cx, int cy, int lane) {
    const int s = slot_of_row(k);
    const int y = cy + c.yoff + CHUNK - 1 - k;  // global row (RNG keys, particle positions)
    const int jw = HX8 + 4 * lane;
    uint32_t mw = ld_word(&MAT(s, jw));
    uint32_t fw = ld_word(&FLG(s, jw));
    bool act = false;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        int ph = LUTP->phys[(mw >> (8 * b)) & 0xff];
        bool vis = (fw >> (8 * b)) & F_VISITED;
        if (!vis && ph == P_GAS) act = true;
    }
    if (!__any_sync(0xffffffffu, act)) return;
    for (int cc = 0; cc < 4; cc++) {
        visit3(c, s, jw + cc, cx + 4 * lane + cc, y);
        __syncwarp();
    }
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

__device__ __forceinline__ void issue_row_load(const TickParams& P, Smem& S, int k, int cx, int cy) {
    issue_row_load_any(P, S.ring, S.bar, S.h.rowmod, S.h.rowchg, k, cx, cy);
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

__device__ __forceinline__ void issue_row_store(const TickParams& P, Smem& S, int k, int cx, int cy) { issue_row_store_any(P, S.ring, k, cx, cy); }

__global__ void __launch_bounds__(128, 3) tick_chunk_kernel(const __grid_constant__ TickParams P) {
    unsigned char* const smem_raw = fse_smem;
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    int cxi, cyi;
    if (P.list_count && (int)blockIdx.x >= *P.list_count) return;  // over-provisioned grid of the active-chunk pass
    if (P.chunk_list) {
        int v = P.chunk_list[blockIdx.x];
        cxi = v & 0xffff;
        cyi = v >> 16;
    } else {
        cxi = blockIdx.x % P.ncx;
        cyi = blockIdx.x / P.ncx;
    }
    const int cx = P.x0 + cxi * 2 * CHUNK;
    const int cy = P.y0 + cyi * 2 * CHUNK;

    const DevTables* T = P.tabs;
    {
        const uint4* src = reinterpret_cast<const uint4*>(&T->lut);
        uint4* dst = reinterpret_cast<uint4*>(&S.h.lut);
        for (int i = tid; i < (int)(sizeof(Lut) / 16); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
        for (int q = 0; q < RING; q++) mbar_init(&S.bar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    Ctx c;
    c.T = T;
    c.pbuf = P.pbuf;
    c.pcount = P.pcount;
    c.pcap = P.pcap;
    c.rkey = P.rkey;
    c.tick = P.tick;
    c.iter = P.iter;
    c.nmat = T->n;
    c.yoff = P.y_off;
    c.air = T->air; c.fire = T->fire; c.water = T->water; c.lava = T->lava; c.steam = T->steam; c.obsidian = T->obsidian;
    c.flowx = c.flowy = nullptr;  // the classes schedule keeps no flow accumulators (fse_flow_enable refuses it)
    c.W = P.W;

    // prologue: rows -HALO_DN .. HALO_UP+PF-1
    if (warp == 3 && lane == 0) {
        for (int k = -HALO_DN; k < HALO_UP + PF; k++) issue_row_load(P, S, k, cx, cy);
    }
    for (int k = -HALO_DN; k < HALO_UP; k++) mbar_wait(&S.bar[slot_of_row(k)], 0);

    bool io_modified = false, io_inert = true;  // IO warp only (active-region tracking)
    for (int t = 0; t < N_STEPS; t++) {
        const int kw = t + HALO_UP;  // newest row this step may touch
        if (kw <= LAST_ROW) mbar_wait(&S.bar[slot_of_row(kw)], (uint32_t)(((kw + HALO_DN) / RING) & 1));
        fence_proxy_async();  // order this thread's shared-memory writes before the IO warp's bulk stores
        __syncthreads();
        if (warp == 0) {
            if (t < CHUNK) pass1_row(c, t, cx, cy, lane);
        } else if (warp == 1) {
            const int k = t - L12;
            if (k >= 0 && k < CHUNK) pass2_row(c, k, cx, cy, lane);
        } else if (warp == 2) {
            const int k = t - L12 - L23;
            if (k >= 0 && k < CHUNK) pass3_row(c, k, cx, cy, lane);
        } else {
            // store the row that pass 3 left in the previous step (rows below/above the chunk ride the same schedule)
            const int ks = t - STORE_LAG - 0;
            if (ks >= -HALO_WR && ks <= LAST_ROW) {
                const int q = slot_of_row(ks);
                if (P.awake && ks >= 0 && ks < CHUNK) io_inert &= row_is_inert(c, q, rs(q, 1), lane);
                io_modified |= S.h.rowchg[q] != 0;
                if (S.h.rowmod[q]) {
                    uint32_t* fw = reinterpret_cast<uint32_t*>(S.ring + q * ROW_BYTES + OFF_FLG);
                    for (int w = lane; w < P8 / 4; w += 32) fw[w] &= 0x7f7f7f7fU;  // tickVisited never reaches HBM
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) issue_row_store(P, S, ks, cx, cy);
                }
                if (lane == 0) bulk_commit();
            }
            const int kl = t + HALO_UP + PF;
            if (kl <= LAST_ROW && lane == 0) {
                bulk_wait_read<1>();  // the slot's previous row (stored two steps ago) has left shared memory
                issue_row_load(P, S, kl, cx, cy);
            }
        }
    }
    if (warp == 3) {
        if (P.awake) {
            // wake the 3x3 neighbourhood if anything changed; go to sleep if nothing changed and every cell is inert
            const bool inert = __all_sync(0xffffffffu, io_inert);
            const int ci = cx / CHUNK, cj = (cy + P.y_off) / CHUNK;
            if (io_modified) {
                if (lane < 9) {
                    const int ni = ci + lane % 3 - 1, nj = cj + lane / 3 - 1;
                    if (ni >= 0 && nj >= 0 && ni < P.acols && nj < P.arows) P.awake[nj * P.acols + ni] = 1;
                }
            } else if (inert && lane == 0 && !P.never_sleep) {
                P.awake[cj * P.acols + ci] = 0;
            }
        }
        if (lane == 0) bulk_wait_all();
    }
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
        cudaError_t e = cudaFuncSetAttribute(tick_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(tick_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemRows));
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
            if (Q.rowmask) {
                classify_rows_kernel<<<(hi - lo) * 4, 1024, sizeof(Lut), st>>>(Q);
                *launched += 1;
            }
            if (Q.rowmask) {
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
    } else if (P.schedule == FSE_SCHEDULE_ROWS) {
        tick_rows_kernel<<<n_chunks, ROWS_THREADS, sizeof(SmemRows), stream>>>(P);
        *launched = 1;
    } else {
        tick_chunk_kernel<<<n_chunks, 128, sizeof(Smem), stream>>>(P);
        *launched = 1;
    }
    return cudaGetLastError();
}

cudaError_t launch_lpt_build(const unsigned int* cost, int n, int ncx, int* list, cudaStream_t stream, const int* members, int parts) {
    lpt_build_kernel<<<1, 1024, 0, stream>>>(cost, n, ncx, list, members, parts);
    return cudaGetLastError();
}

}  // namespace fse
