// fse_tick_rows.cuh — "simultaneous rows" schedule of the chunk tick; included by fse_tick.cu.
//
// The reference's chunk colours, passes and bottom-up rows; a row step of a pass is executed by 128 threads (one per column) that
// first DECIDE from the state before the step and then COMMIT:
//   pass 1:  D | bar | C1 own column + publish horizontal flows/pokes | bar | C2 targets gather (left then right), hand
//            back what no longer fits | bar | refunds + C3 area effects (FIRE, water-on-lava, pair interactions; claims by
//            the lowest source column — rare)
//   pass 2:  D + destination claims (atomicMin of the source column) | bar | C winners move, liquid diff applied | bar | pokes
//   pass 3:  D + claims | C   (rows do not depend on each other: one warp per row)
// The CPU oracle restates exactly this (oracle/rows_oracle.cpp, Schedule::ROWS); tests require bit-equality.
// Kernels: tick_rows_kernel (all three passes pipelined in one CTA: small worlds), tick_pass_kernel<PASS, SKIP> + tick_pass3_kernel
// (one kernel per pass: the default), classify_rows_kernel (settled-row skipping), tick_pass2_apply_kernel (pass 2 split).
#pragma once
#include <cstddef>

namespace fse {

constexpr int ROWS_THREADS = 320;  // warps 0-3 pass 1, 4-7 pass 2 (one column per thread), warp 8 pass 3, warp 9 IO
enum Act { A_NONE = 0, A_MARK, A_REACT, A_SAND_PART, A_SAND_SWAP, A_SOUP_ZERO, A_SOUP_PART, A_SOUP_FLOW, A_SOUP_SWAPDOWN, A_GAS_UP, A_FIRE, A_INTERACT };

// decision bits
constexpr uint32_t DB_COIN = 1u << 4, DB_POKEL = 1u << 5, DB_POKER = 1u << 6, DB_MOVED = 1u << 7, DB_CHANGED = 1u << 8, DB_SWAPUP = 1u << 9,
                   DB_WL = 1u << 10, DB_BOTSOUP = 1u << 11, DB_TOPSOUP = 1u << 12, DB_LEFTSOUP = 1u << 13, DB_RIGHTSOUP = 1u << 14,
                   DB_EMBER = 1u << 15, DB_DIE = 1u << 16;
struct Dec1 {
    uint32_t bits;   // act in bits 0..3, flags above, stl_new / product / below-material in bits 24..31
    float fd_new, fD, fL, fR, fU;  // FIRE: fL carries the ignite mask
};

// per-row scratch of each pass (shared memory); the fused kernel carries all three, a per-pass kernel only its own
struct Scratch1 {
    float outL[CHUNK + 2], outR[CHUNK + 2], refL[CHUNK + 2], refR[CHUNK + 2];
    uint32_t area_arg[CHUNK];
    uint32_t hf[CHUNK + 2];  // what each column published in C1 (see commit1); entries 0 and CHUNK + 1 stay 0
    int p1_any[2], p1_area[2], p1_horiz[2];
    uint32_t area_mask[2][4];   // columns of this row with an area effect to apply (bit i = column i)
    uint8_t areaClaim[11][P8];  // serial C3: claimant column + 1 (0 = free); 4-byte aligned (cleared as words)
    uint8_t area_kind[CHUNK];
    long long dbg_phase[10];    // profiling aid (FSE_ROLE_CYCLES): cycles of D, C1, C2, area, -, D own; rows seen, active, with gather, with area
};
struct Scratch2 {
    int claimDn[2][CHUNK + 2], claimUp[2][CHUNK + 2];
    int p2_any[2], p2_poke[2];
    uint8_t poke2[CHUNK];
};
struct Scratch3 {
    int claim3[2][CHUNK + 2];
};
struct RowScratch : Scratch1, Scratch2, Scratch3 {
    long long dbg_arrival[2][4];  // profiling aid: clock at which each role finished its step (double-buffered)
};
static_assert(offsetof(Scratch1, areaClaim) % 4 == 0, "areaClaim is cleared with 32-bit stores");
static_assert(sizeof(Scratch1) % 4 == 0 && sizeof(Scratch2) % 4 == 0 && sizeof(Scratch3) % 4 == 0, "scratch is cleared with 32-bit stores");

struct __align__(128) SmemRows {
    SmemHead h;
    unsigned char ring[RING * ROW_BYTES];
    Ctx ctx;  // CTA-uniform context, kept in shared memory so the rule code needs no registers for it
    unsigned long long bar[RING];
    RowScratch rs;
};

// ring geometry from the context (the fused kernel uses 28 rows, the per-pass kernels 16)
template <int RN>
__device__ __forceinline__ int rsn(int s, int dy) {
    if ((RN & (RN - 1)) == 0) return (s - dy) & (RN - 1);
    int v = s - dy;
    if (v < 0) v += RN;
    if (v >= RN) v -= RN;
    return v;
}
template <int RN>
__device__ __forceinline__ int slotk(const Ctx& c, int k) { return (RN & (RN - 1)) == 0 ? ((k + c.koff) & (RN - 1)) : ((k + c.koff) % RN); }

// optional pass-1 phase clocks (scripts/role_cycles.py); compiled out unless FSE_ROLE_CYCLES is defined
#ifdef FSE_ROLE_CYCLES
#define FSE_P1_CLOCK(name) const long long name = clock64()
#define FSE_P1_PHASE(slot, since) do { if (t == 0) R.dbg_phase[slot] += clock64() - (since); } while (0)
#define FSE_P1_COUNT(slot) do { if (t == 0) R.dbg_phase[slot] += 1; } while (0)
#else
#define FSE_P1_COUNT(slot) do { } while (0)
#define FSE_P1_CLOCK(name) do { } while (0)
#define FSE_P1_PHASE(slot, since) do { } while (0)
#endif

// fire-and-forget float add to global memory (RED.E.ADD.F32; nothing is returned, so no latency is exposed).  Like every global f32
// atomic it flushes subnormals to zero; liquid flows are differences of amounts >= FLUID_MinValue (>= 5e-11), far above that range.
__device__ __forceinline__ void red_add_f32(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}
__device__ __forceinline__ void pass_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

__device__ __forceinline__ float clampflow(float flow, float cap, bool speed) {
    if (speed && flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
    flow = fmaxf(flow, 0.0f);
    if (flow > fminf(FLUID_MaxFlow, cap)) flow = fminf(FLUID_MaxFlow, cap);
    return flow;
}

// FIRE decision (world.cpp:1101-1146) for the fire cell at (s, jf), evaluated by the whole warp: lane i < 25 owns neighbour
// (xx, yy) = (i / 5 - 2, i % 5 - 2) — the reference's loop order, so the RNG slots are the same.  Returns the decision bits
// (A_FIRE | DB_EMBER | DB_DIE) and the ignite mask; every lane gets the same values.
template <int RN>
__device__ __forceinline__ uint32_t decide_fire_coop(const Ctx& c, int s, int jf, int xf, int y, int lane, uint32_t& ignite) {
    const uint32_t cb = rng_cell(c.rkey, xf, y);
    uint32_t bits = A_FIRE;
    ignite = 0;
    if (rng_draw(cb, S_FIRE_EMBER) % 10 == 0) bits |= DB_EMBER;
    if (rng_draw(cb, S_FIRE_DIE) % 150 == 0) return bits | DB_DIE;
    bool solid = false;
    if (lane < 25) solid = PHYS(rsn<RN>(s, lane % 5 - 2), jf + lane / 5 - 2) == P_SOLID;
    const unsigned solids = __ballot_sync(0xffffffffu, solid);
    ignite = __ballot_sync(0xffffffffu, solid && rng_draw(cb, S_FIRE_IGNITE0 + lane) % 500 == 0);
    if (!solids && rng_draw(cb, S_FIRE_DIE_ALONE) % 120 == 0) bits |= DB_DIE;
    return bits;
}

// ---- pass 1: decide (world.cpp:1089-1586, read-only) ------------------------------------------------------------------
template <int RN>
__device__ Dec1 decide1(const Ctx& c, int s, int j, int x, int y) {
    Dec1 d;
    d.bits = A_NONE;
    d.fd_new = d.fD = d.fL = d.fR = d.fU = 0.0f;
    const uint8_t f0 = FLG(s, j);
    if (f0 & F_VISITED) return d;
    const uint8_t m = MAT(s, j);
    if (c.iter >= (int)LUTP->iters[m]) {
        d.bits = A_MARK | ((uint32_t)f0 << 16);  // the caller marks the cell without reading its flag byte again
        return d;
    }
    const int type = LUTP->phys[m];
    if (type == P_AIR || type == P_SOLID) return d;
    const uint32_t cb = rng_cell(c.rkey, x, y);
    const int sb = rsn<RN>(s, 1);
    if ((int)m == c.fire) {  // 1101-1146: filled in warp-cooperatively by the caller (decide_fire_coop)
        d.bits = A_FIRE;
        return d;
    }
    if (type == P_SAND) {  // 1148-1267
        // neighbourhood up front (independent shared-memory reads, see the liquid branch)
        const uint8_t mb = MAT(sb, j), mbl = MAT(sb, j - 1), mbr = MAT(sb, j + 1);
        const uint8_t m2 = MAT(rsn<RN>(s, 2), j), m3 = MAT(rsn<RN>(s, 3), j), m4 = MAT(rsn<RN>(s, 4), j);
        const int bt = LUTP->phys[mb], btl = LUTP->phys[mbl], btr = LUTP->phys[mbr];
        const float myDens = LUTP->dens[m], densB = LUTP->dens[mb], densL = LUTP->dens[mbl], densR = LUTP->dens[mbr];
        const bool deepAir = LUTP->phys[m2] == P_AIR && LUTP->phys[m3] == P_AIR && LUTP->phys[m4] == P_AIR;
        const uint8_t mf = LUTP->mflags[m];
        if ((mf & MF_INTERACT) && has_interaction(c, m, mb)) {
            d.bits = A_INTERACT | ((uint32_t)mb << 24);
            return d;
        }
        if (mf & MF_REACT) {
            const int16_t temp = TMP(s, j);
            int prod = -1;
            if (!(mf & MF_REACT_MULTI)) {
                const Lut::Rx rx = LUTP->rx[m];
                if ((rx.type == FSE_REACT_TEMPERATURE_BELOW && temp < rx.thr) || (rx.type == FSE_REACT_TEMPERATURE_ABOVE && temp > rx.thr)) prod = rx.prod;
            } else {
                for (int i = c.T->react_off[m]; i < c.T->react_off[m + 1]; i++) {
                    const fse_interaction in = c.T->react[i];
                    if ((in.type == FSE_REACT_TEMPERATURE_BELOW && temp < in.data1) || (in.type == FSE_REACT_TEMPERATURE_ABOVE && temp > in.data1))
                        prod = (int)in.data2;
                }
            }
            if (prod >= 0) {
                d.bits = A_REACT | ((uint32_t)prod << 24);
                return d;
            }
        }
        if (!(bt == P_AIR || (bt != P_SOLID && densB < myDens))) return d;
        // the grain can fall: all five draws at once (pure functions of the cell; independent hash chains overlap)
        const uint32_t rHes = rng_draw(cb, S_SAND_HESITATE), rMov = rng_draw(cb, S_SAND_MOVED), rSelf = rng_draw(cb, S_SAND_TX_SELF),
                       rL = rng_draw(cb, S_SAND_TX_L), rR = rng_draw(cb, S_SAND_TX_R);
        const bool canL = btl == P_AIR || (btl != P_SOLID && densL < myDens), canR = btr == P_AIR || (btr != P_SOLID && densR < myDens);
        if ((canL || canR) && rHes % 20 == 0) return d;
        uint32_t bits;
        if (bt == P_AIR && deepAir) {
            bits = A_SAND_PART;
        } else {
            bits = A_SAND_SWAP;
            if (rMov % 2 == 0) bits |= DB_COIN;
        }
        if (rSelf % 2 == 0) {
            if (rL % 2 == 0) bits |= DB_POKEL;
            if (rR % 2 == 0) bits |= DB_POKER;
        }
        d.bits = bits;
        return d;
    }
    if (type == P_SOUP) {  // 1269-1537
        // the whole neighbourhood is read up front: the decision below is a chain of short branches, and with the loads hoisted
        // their shared-memory latencies overlap instead of adding up
        const int st = rsn<RN>(s, -1);
        const float fl = FL(s, j);
        const uint8_t mb0 = MAT(sb, j), ml = MAT(s, j - 1), mr = MAT(s, j + 1), mt = MAT(st, j);
        const uint8_t m2 = MAT(rsn<RN>(s, 2), j), m3 = MAT(rsn<RN>(s, 3), j), m4 = MAT(rsn<RN>(s, 4), j);
        const float bottomFl = FL(sb, j), leftFl = FL(s, j - 1), rightFl = FL(s, j + 1), topFl = FL(st, j);
        float fd = FD(s, j);
        uint8_t stl = STL(s, j);
        const int bph = LUTP->phys[mb0], lph = LUTP->phys[ml], rph = LUTP->phys[mr], tph = LUTP->phys[mt];
        const bool deepAir = LUTP->phys[m2] == P_AIR && LUTP->phys[m3] == P_AIR && LUTP->phys[m4] == P_AIR;
        if (fl == 0.0f) return d;
        if (fl < FLUID_MinValue) {
            d.bits = A_SOUP_ZERO;
            return d;
        }
        if ((double)fl > 0.005 && bph == P_AIR && deepAir) {
            d.bits = A_SOUP_PART;
            return d;
        }
        if (f0 & F_MOVED) return d;
        const float start = fl;
        float rem = fl;
        const bool airBelow = bph == P_AIR;
        uint32_t bits = A_SOUP_FLOW;
        if (bph == P_SOUP) bits |= DB_BOTSOUP;
        bool early = false;
        if ((airBelow && c.iter <= 2) || mb0 == m) {  // 1315-1334
            const float dst = bph == P_SOUP ? bottomFl : 0.0f;
            float flow = vertical_flow(start, dst) - dst;
            flow = clampflow(flow, start, bottomFl > 0);
            if (flow != 0) {
                rem -= flow;
                fd -= flow;
                d.fD = flow;
            }
        } else if (c.iter == 0 && bph == P_SOUP && mb0 != m) {  // 1335-1341
            if (rng_draw(cb, S_SOUP_SWAP_DOWN) % 10 == 0) {
                d.bits = A_SOUP_SWAPDOWN;
                return d;
            }
        }
        if (rem < FLUID_MinValue) {
            fd -= rem;
            early = true;
        }
        if (lph == P_SOUP) bits |= DB_LEFTSOUP;
        if (rph == P_SOUP) bits |= DB_RIGHTSOUP;
        const bool canL = (lph == P_AIR || ml == m) && !airBelow;
        const bool canR = (rph == P_AIR || mr == m) && !airBelow;
        if (!early && canL) {  // 1355-1375
            const float dst = lph == P_SOUP ? leftFl : 0.0f;
            const float flow = clampflow((rem - dst) / (canR ? 3.0f : 2.0f), rem, true);
            if (flow != 0) {
                rem -= flow;
                fd -= flow;
                d.fL = flow;
            }
        }
        if (!early && rem < FLUID_MinValue) {
            fd -= rem;
            early = true;
        }
        if (!early && canR) {  // 1383-1403
            const float dst = rph == P_SOUP ? rightFl : 0.0f;
            const float flow = clampflow((rem - dst) / 2.0f, rem, true);
            if (flow != 0) {
                rem -= flow;
                fd -= flow;
                d.fR = flow;
            }
        }
        if (!early && rem < FLUID_MinValue) {
            fd -= rem;
            early = true;
        }
        if (tph == P_SOUP) bits |= DB_TOPSOUP;
        bool swapUp = false;
        if (!early) {
            if (tph == P_AIR || mt == m) {  // 1413-1432
                const float dst = tph == P_SOUP ? topFl : 0.0f;
                const float flow = clampflow(rem - vertical_flow(rem, dst), rem, true);
                if (flow != 0) {
                    rem -= flow;
                    fd -= flow;
                    d.fU = flow;
                }
            } else if (c.iter == 0 && tph == P_SOUP && mt != m) {  // 1433-1439
                if (rng_draw(cb, S_SOUP_SWAP_UP) % 10 == 0) swapUp = true;
            }
        }
        if (!early && !swapUp && rem < FLUID_MinValue) {
            fd -= rem;
            early = true;
        }
        d.fd_new = fd;
        bool moved = (f0 & F_MOVED) != 0;
        if (swapUp) bits |= DB_SWAPUP;
        if (!early && !swapUp) {
            if (start == rem) {  // 1447-1451
                stl = (uint8_t)(stl + 1);
                if (stl >= 10) moved = true;
            } else {
                bits |= DB_CHANGED;
            }
            if ((int)m == c.water && (int)mb0 == c.lava) bits |= DB_WL;  // 1519
        }
        if (moved) bits |= DB_MOVED;
        d.bits = bits | ((uint32_t)stl << 24);
        return d;
    }
    if (type == P_GAS) {  // 1569-1585
        const int st = rsn<RN>(s, -1);
        if (PHYS(st, j) == P_AIR && !((PHYS(st, j - 1) == P_AIR || PHYS(st, j + 1) == P_AIR) && rng_draw(cb, S_GAS1) % 2 == 0)) d.bits = A_GAS_UP;
    }
    return d;
}

// own-column commit; k = column index in the scratch arrays (j - HX8 + 1)
template <int RN>
__device__ void commit1(const Ctx& c, Scratch1& R, const Dec1& d, int s, int j, int x, int y, int par) {
    const int k = j - HX8 + 1;
    const int act = d.bits & 15;
    float oL = 0.0f, oR = 0.0f;
    uint8_t chg = 0, pkL = 0, pkR = 0, akind = 0;
    uint32_t aarg = 0;
    const int sb = rsn<RN>(s, 1), st = rsn<RN>(s, -1);
    switch (act) {
        case A_MARK:
            FLG(s, j) = FLG(s, j) | F_VISITED;
            ROWVIS[s] = 1;
            break;
        case A_REACT: {
            const int16_t t = TMP(s, j);
            CellR n = create(c, (int)(d.bits >> 24), x, y);
            n.tmp = t;
            stc(c, s, j, n, F_DIRTY | F_VISITED);
            break;
        }
        case A_SAND_PART:
        case A_SAND_SWAP: {
            CellR tile = ldc(c, s, j);
            const CellR below = ldc(c, sb, j);
            stc(c, s, j, below, F_DIRTY);
            if (act == A_SAND_PART) {
                const uint32_t cb = rng_cell(c.rkey, x, y);
                const float vx = ((int)(rng_draw(cb, S_SAND_PART_VX) % 10) - 5) / 20.0f;
                const float vy = -((int)(rng_draw(cb, S_SAND_PART_VY) % 2) + 3) / 10.0f + 1.5f;
                emit_particle(c, tile, (float)x, (float)(y + 1), vx, vy, 0.1f, false, 0, 60, particle_id(c, x, y, 14));
            } else {
                if (d.bits & DB_COIN) tile.moved = 1;
                stc(c, sb, j, tile, F_DIRTY | F_VISITED);
            }
            pkL = (d.bits & DB_POKEL) ? 1 : 0;
            pkR = (d.bits & DB_POKER) ? 1 : 0;
            break;
        }
        case A_SOUP_ZERO:
            FL(s, j) = 0.0f;
            ROWMOD[s] = 1;
            ROWCHG[s] = 1;
            break;
        case A_SOUP_PART: {
            const CellR tile = ldc(c, s, j);
            stc(c, s, j, nothing(c), F_DIRTY);
            int n = (int)(tile.fl / 4);
            if (n < 1) n = 1;
            const uint32_t cb = rng_cell(c.rkey, x, y);
            for (int i = 0; i < n; i++) {
                CellR nt = fresh_fluid(tile);
                nt.fl = tile.fl / n;
                const float vx = ((int)(rng_draw(cb, S_SOUP_PART0 + 2 * (i & 7)) % 10) - 5) / 30.0f;
                const float vy = -((int)(rng_draw(cb, S_SOUP_PART0 + 2 * (i & 7) + 1) % 2) + 3) / 10.0f + 1.0f;
                emit_particle(c, nt, (float)x, (float)(y + 1), vx, vy, 0.1f, false, 0, 60, particle_id(c, x, y, i & 7));
            }
            break;
        }
        case A_SOUP_SWAPDOWN: {
            const CellR tile = ldc(c, s, j), bottom = ldc(c, sb, j);
            stc(c, s, j, bottom, 0);
            stc(c, sb, j, tile, 0);
            break;
        }
        case A_SOUP_FLOW: {
            CellR tile = ldc(c, s, j);
            tile.fd = d.fd_new;
            tile.stl = (uint8_t)(d.bits >> 24);
            tile.moved = (d.bits & DB_MOVED) ? 1 : 0;
            if (d.fD != 0) pour(c, sb, j, PHYS(sb, j), tile, d.fD);
            if (d.fU != 0) pour(c, st, j, PHYS(st, j), tile, d.fU);
            if (d.bits & DB_SWAPUP) {  // 1433-1439
                const CellR top = ldc(c, st, j);
                stc(c, s, j, top, 0);
                stc(c, st, j, tile, 0);
            } else {
                stc(c, s, j, tile, (d.bits & DB_CHANGED) ? F_DIRTY : 0);
                if (d.bits & DB_CHANGED) {  // 1452-1458, vertical neighbours
                    if (d.bits & DB_TOPSOUP) set_moved(c, st, j, false);
                    if (d.bits & DB_BOTSOUP) set_moved(c, sb, j, false);
                    chg = (uint8_t)(((d.bits & DB_LEFTSOUP) ? 1 : 0) | ((d.bits & DB_RIGHTSOUP) ? 2 : 0));
                }
                if (d.bits & DB_WL) akind = 2;
            }
            oL = d.fL;
            oR = d.fR;
            if (c.flowx) {  // world.cpp:1334, 1374, 1402, 1432 — the flows as decided; fire-and-forget reductions (RED.ADD.F32), same
                            // thread and address in program order, so flowY gets +down before -up and flowX -left before +right
                const size_t g = (size_t)(y - c.yoff) * c.W + x;
                if (d.fD != 0) red_add_f32(c.flowy + g, d.fD);
                if (d.fL != 0) red_add_f32(c.flowx + g, -d.fL);
                if (d.fR != 0) red_add_f32(c.flowx + g, d.fR);
                if (d.fU != 0) red_add_f32(c.flowy + g, -d.fU);
            }
            break;
        }
        case A_GAS_UP: {
            const CellR tile = ldc(c, s, j), up = ldc(c, st, j);
            stc(c, s, j, up, F_DIRTY);
            stc(c, st, j, tile, F_DIRTY | F_VISITED);
            break;
        }
        case A_FIRE: {
            if (d.bits & DB_EMBER) {  // 1109-1119
                const CellR tile = ldc(c, s, j);
                const uint32_t cb = rng_cell(c.rkey, x, y);
                const float vx = ((int)(rng_draw(cb, S_FIRE_EMBER_VX) % 10) - 5) / 20.0f;
                const float vy = -((int)(rng_draw(cb, S_FIRE_EMBER_VY) % 10) / 10.0f) / 3.0f + -0.5f;
                emit_particle(c, tile, (float)x, (float)(y - 1), vx, vy, 0.01f, true, 30, 10, particle_id(c, x, y, 15));
            }
            aarg = __float_as_uint(d.fL) | ((d.bits & DB_DIE) ? (1u << 25) : 0);
            if (aarg) akind = 1;
            break;
        }
        case A_INTERACT:
            akind = 3;
            aarg = d.bits >> 24;
            break;
        default:
            break;
    }
    // publish: bit 0/1 flow to the left/right, bit 2/3 un-settle the left/right soup neighbour, bit 4/5 poke the sand below-left/right
    const uint32_t h = (oL != 0 ? 1u : 0u) | (oR != 0 ? 2u : 0u) | ((uint32_t)chg << 2) | ((uint32_t)pkL << 4) | ((uint32_t)pkR << 5);
    R.hf[k] = h;
    if (h) {
        R.p1_horiz[par] = 1;
        if (h & 1) R.outL[k] = oL;
        if (h & 2) R.outR[k] = oR;
    }
    if (akind) {
        R.area_kind[k - 1] = akind;
        R.area_arg[k - 1] = aarg;
    }
}

// C2: column k (scratch index) of row slot s receives its horizontal inflows, un-settle flags and, for the row below, pokes
template <int RN>
__device__ __forceinline__ void gather1(const Ctx& c, Scratch1& R, int s, int k) {
    const uint32_t hl = k > 0 ? R.hf[k - 1] : 0u, hr = k < CHUNK + 1 ? R.hf[k + 1] : 0u;
    if (!((hl & (2u | 8u | 32u)) | (hr & (1u | 4u | 16u)))) return;
    const int j = k - 1 + HX8;
    const float inL = (hl & 2) ? R.outR[k - 1] : 0.0f, inR = (hr & 1) ? R.outL[k + 1] : 0.0f;
    if (inL != 0 || inR != 0) {
        const uint8_t m = MAT(s, j);
        const int ph = LUTP->phys[m];
        if (ph == P_AIR) {
            if (inL != 0) {
                CellR n = fresh_fluid(ldc(c, s, j - 1));
                n.fd = inL;
                if (inR != 0) {
                    if (MAT(s, j + 1) == n.mat) n.fd = n.fd + inR;
                    else R.refL[k + 1] = inR;
                }
                stc(c, s, j, n, 0);
            } else {
                CellR n = fresh_fluid(ldc(c, s, j + 1));
                n.fd = inR;
                stc(c, s, j, n, 0);
            }
        } else {
            if (inL != 0) {
                if (ph == P_SOUP && MAT(s, j - 1) == m) {
                    FD(s, j) = FD(s, j) + inL;
                    ROWMOD[s] = 1;
                    ROWCHG[s] = 1;
                } else {
                    R.refR[k - 1] = inL;
                }
            }
            if (inR != 0) {
                if (ph == P_SOUP && MAT(s, j + 1) == m) {
                    FD(s, j) = FD(s, j) + inR;
                    ROWMOD[s] = 1;
                    ROWCHG[s] = 1;
                } else {
                    R.refL[k + 1] = inR;
                }
            }
        }
    }
    if (((hl & 8) || (hr & 4)) && PHYS(s, j) == P_SOUP) set_moved(c, s, j, false);
    if ((hl & 32) || (hr & 16)) {
        const int sb = rsn<RN>(s, 1);
        if (PHYS(sb, j) == P_SAND) set_moved(c, sb, j, true);
    }
}

// C3: area effects (FIRE burn-out / ignition, water on lava, pair interactions).  Every source column first claims the cells
// it wants to rewrite — a contested cell goes to the lowest source column, as if the sources ran one after the other in
// ascending x — then (after a barrier) rewrites the cells it owns.  pass = 0: claim, pass = 1: apply.
__device__ __forceinline__ void area_claim_min(Scratch1& R, int tj, int dy, int i) {  // dy in -5..5 rows below(+)/above(-)
    uint8_t* cell = &R.areaClaim[dy + 5][tj];
    uint32_t* w = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(cell) & ~(uintptr_t)3);
    const int sh = (int)(reinterpret_cast<uintptr_t>(cell) & 3) * 8;
    const uint32_t v = (uint32_t)(i + 1);
    uint32_t old = *w;
    while (true) {
        const uint32_t b = (old >> sh) & 0xffu;
        if (b != 0 && b <= v) return;
        const uint32_t prev = atomicCAS(w, old, (old & ~(0xffu << sh)) | (v << sh));
        if (prev == old) return;
        old = prev;
    }
}
template <int RN>
__device__ __noinline__ void area_effects(const Ctx& c, Scratch1& R, int s, int cx, int y, int i, int pass) {
    auto claim = [&](int i_, int tj, int dy) { area_claim_min(R, tj, dy, i_); };
    auto mine = [&](int i_, int tj, int dy) { return R.areaClaim[dy + 5][tj] == (uint8_t)(i_ + 1); };
    {
        {
            const int kind = R.area_kind[i];
            if (!kind) return;
            const int j = HX8 + i, x = cx + i;
            const uint32_t arg = R.area_arg[i];
            if (kind == 1) {  // FIRE: burn out / ignite (1121-1144)
                const bool die = (arg >> 25) & 1;
                if (pass == 0) {
                    if (die) claim(i, j, 0);
                    for (int kk = 0; kk < 25; kk++)
                        if ((arg >> kk) & 1) claim(i, j + kk / 5 - 2, kk % 5 - 2);
                } else {
                    for (int kk = 0; kk < 25; kk++) {
                        const int xx = kk / 5 - 2, yy = kk % 5 - 2;
                        if (((arg >> kk) & 1) && mine(i, j + xx, yy)) stc(c, rsn<RN>(s, yy), j + xx, create(c, c.fire, x + xx, y + yy), F_DIRTY | F_VISITED);
                    }
                    if (die && mine(i, j, 0)) stc(c, s, j, nothing(c), F_DIRTY | F_VISITED);
                }
            } else if (kind == 2) {  // water on lava (1519-1537)
                if (pass == 0) {
                    for (int xx = -1; xx <= 1; xx++)
                        for (int yy = 0; yy <= 2; yy++) claim(i, j + xx, yy);
                } else {
                    if (mine(i, j, 0)) stc(c, s, j, create(c, c.steam, x, y), F_DIRTY);
                    if (mine(i, j, 1)) stc(c, rsn<RN>(s, 1), j, create(c, c.obsidian, x, y + 1), F_DIRTY | F_VISITED);
                    for (int xx = -1; xx <= 1; xx++)
                        for (int yy = 0; yy <= 2; yy++)
                            if (mine(i, j + xx, yy) && (int)MAT(rsn<RN>(s, yy), j + xx) == c.lava)
                                stc(c, rsn<RN>(s, yy), j + xx, create(c, c.obsidian, x + xx, y + yy), F_DIRTY | F_VISITED);
                }
            } else {  // pair interactions (1153-1179): the list of the material the source had when it decided
                int mb, msrc;
                if (pass == 0) {
                    msrc = MAT(s, j);
                    mb = (int)(arg & 0xff);
                    R.area_arg[i] = (uint32_t)mb | ((uint32_t)msrc << 8);
                    claim(i, j, 0);
                } else {
                    if (!mine(i, j, 0)) return;  // a lower-column effect rewrote the source: its list is void
                    mb = (int)(arg & 0xff);
                    msrc = (int)((arg >> 8) & 0xff);
                }
                const int lo = c.T->inter_off[msrc * c.nmat + mb], hi = c.T->inter_off[msrc * c.nmat + mb + 1];
                for (int q = lo; q < hi; q++) {
                    const fse_interaction in = c.T->inter[q];
                    const int rad = (int)in.data2;
                    if (in.type != FSE_INTERACT_TRANSFORM_MATERIAL && in.type != FSE_INTERACT_SPAWN_MATERIAL) continue;
                    for (int xx = in.ofs_x - rad; xx <= in.ofs_x + rad; xx++)
                        for (int yy = in.ofs_y - rad; yy <= in.ofs_y + rad; yy++) {
                            if (pass == 0) {
                                claim(i, j + xx, yy);
                            } else if (mine(i, j + xx, yy)) {
                                const int tm = MAT(rsn<RN>(s, yy), j + xx);
                                const bool hit = in.type == FSE_INTERACT_TRANSFORM_MATERIAL ? tm == mb : ((xx == 0 && yy == 0) || tm == c.air);
                                if (hit) stc(c, rsn<RN>(s, yy), j + xx, create(c, in.data1, x + xx, y + yy), F_DIRTY | F_VISITED);
                            }
                        }
                }
            }
        }
    }
}

template <int RN>
__device__ void pass1_rows(const Ctx& c, Scratch1& R, int k, int s, int cx, int cy, int t) {  // s = slot of row k
    const int y = cy + c.yoff + CHUNK - 1 - k;
    const int par = k & 1;
    const int lane = t & 31;
    const int j = HX8 + t;
    if (t == 0) {
        R.p1_any[par ^ 1] = 0;
        R.p1_area[par ^ 1] = 0;
        R.p1_horiz[par ^ 1] = 0;
        R.area_mask[par ^ 1][0] = R.area_mask[par ^ 1][1] = R.area_mask[par ^ 1][2] = R.area_mask[par ^ 1][3] = 0;
    }
    FSE_P1_CLOCK(T0);
    Dec1 d = decide1<RN>(c, s, j, cx + t, y);
    {  // FIRE cells of this warp's 32 columns, one at a time, all lanes helping
        unsigned fm = __ballot_sync(0xffffffffu, (d.bits & 15) == A_FIRE);
#pragma unroll 1
        while (fm) {
            const int src = __ffs(fm) - 1;
            fm &= fm - 1;
            const int tt = (t & ~31) + src;
            uint32_t ignite;
            const uint32_t bits = decide_fire_coop<RN>(c, s, HX8 + tt, cx + tt, y, lane, ignite);
            if (lane == src) {
                d.bits = bits;
                d.fL = __uint_as_float(ignite);
            }
        }
    }
    int a = d.bits & 15;
    if (a == A_MARK) {  // iterations exhausted (1095): only the cell's own visited bit, which no decision of this step reads
        FLG(s, j) = (uint8_t)((d.bits >> 16) | F_VISITED);
        ROWVIS[s] = 1;
        d.bits = a = A_NONE;
    }
    if (a) R.p1_any[par] = 1;
    if ((a == A_FIRE && ((d.bits & DB_DIE) || __float_as_uint(d.fL))) || a == A_INTERACT || (d.bits & DB_WL)) {
        atomicOr(&R.area_mask[par][t >> 5], 1u << (t & 31));
        R.p1_area[par] = 1;
    }
    FSE_P1_PHASE(5, T0);  // own decide time of warp 0
    pass_bar(1);
    FSE_P1_CLOCK(T1);
    FSE_P1_PHASE(0, T0);  // decide incl. waiting for the slowest warp
    FSE_P1_COUNT(6);
    if (!R.p1_any[par]) return;
    FSE_P1_COUNT(7);
    commit1<RN>(c, R, d, s, j, cx + t, y, par);
    pass_bar(1);
    FSE_P1_CLOCK(T2);
    FSE_P1_PHASE(1, T1);
    if (R.p1_horiz[par]) FSE_P1_COUNT(8);
    if (R.p1_horiz[par]) {  // somebody published a horizontal flow, an un-settle flag or a poke
        gather1<RN>(c, R, s, 1 + t);
        if (t == 0) gather1<RN>(c, R, s, 0);
        if (t == CHUNK - 1) gather1<RN>(c, R, s, CHUNK + 1);
        pass_bar(1);
        // refunds to the sources, left flow first
        const int kk = 1 + t;
        const float rl = R.refL[kk], rr = R.refR[kk];
        if (rl != 0) { FD(s, j) = FD(s, j) + rl; R.refL[kk] = 0.0f; }
        if (rr != 0) { FD(s, j) = FD(s, j) + rr; R.refR[kk] = 0.0f; }
        if (rl != 0 || rr != 0) { ROWMOD[s] = 1; ROWCHG[s] = 1; }
    }
    FSE_P1_PHASE(2, T2);
    if (R.p1_area[par]) {
        uint32_t* cl = reinterpret_cast<uint32_t*>(&R.areaClaim[0][0]);
        for (int q = t; q < 11 * P8 / 4; q += CHUNK) cl[q] = 0;
        pass_bar(1);
        FSE_P1_CLOCK(T3);
        FSE_P1_COUNT(9);
        const bool src = (R.area_mask[par][t >> 5] >> (t & 31)) & 1;
        if (src) area_effects<RN>(c, R, s, cx, y, t, 0);
        pass_bar(1);
        if (src) area_effects<RN>(c, R, s, cx, y, t, 1);
        FSE_P1_PHASE(3, T3);
    }
}

// ---- pass 2 (world.cpp:1594-1820) ----------------------------------------------------------------------------------------
// decision: bits 0..2 act (0 none, 1 moved=false, 2 slide, 3 liquid apply, 4 gas diagonal), bit 3 dir right, 4 riser, 5 restick, 6 poke
template <int RN>
__device__ uint32_t decide2(const Ctx& c, int s, int j, int x, int y) {
    const uint8_t f0 = FLG(s, j);
    if (f0 & F_VISITED) return 0;
    const uint8_t m = MAT(s, j);
    const int type = LUTP->phys[m];
    if (type == P_SAND) {
        const int sb = rsn<RN>(s, 1);
        const float myDens = LUTP->dens[m];
        const bool canL = can_sink(c, sb, j - 1, myDens), canR = can_sink(c, sb, j + 1, myDens);
        if (!(canL || canR)) return 1;
        const uint32_t cb = rng_cell(c.rkey, x, y);
        bool stopped = !(f0 & F_MOVED);
        const int slip = LUTP->slip[m];
        if (stopped) {
            int drop = 0;
#pragma unroll
            for (int pil = 0; pil < 10; pil++) {
                const int sp = rsn<RN>(s, 1 + pil);
                if (PHYS(sp, j - 1) == P_AIR || PHYS(sp, j + 1) == P_AIR) drop++;
            }
            const int dd = drop + 1 - (int)LUTP->maxstab[m];
            if (dd > 0) {
                const int chance = 1000 / dd;
                if (chance < 1000 && rng_draw(cb, S_SAND2_UNSTICK) % chance == 0) stopped = false;
            }
        }
        if (stopped) return 1;
        const bool should = rng_draw(cb, S_SAND2_SHOULD) % (2 * slip) != 0;
        uint32_t bits = 0;
        if (should && rng_draw(cb, S_SAND2_TX_SELF) % 2 == 0 && rng_draw(cb, S_SAND2_TX_OTHER) % 2 == 0) bits |= 64;
        int dir = 0;
        if (should && canL && (!canR || rng_draw(cb, S_SAND2_LR) % 2 == 0)) dir = -1;
        else if (should && canR) dir = 1;
        if (!dir) return 1 | bits;
        bits |= 2;
        if (dir > 0) bits |= 8;
        if (PHYS(s, j + dir) == P_AIR) bits |= 16;
        if (rng_draw(cb, S_SAND2_RESTICK) % (20 * slip) == 0) bits |= 32;
        return bits;
    }
    if (type == P_SOUP) return 3;
    if (type == P_GAS) {
        const int st = rsn<RN>(s, -1);
        const int aL = PHYS(st, j - 1), aR = PHYS(st, j + 1);
        if (aL == P_AIR && !(aR == P_AIR && rng_draw(rng_cell(c.rkey, x, y), S_GAS2) % 2 == 0)) return 4;
        if (aR == P_AIR) return 4 | 8;
    }
    return 0;
}

template <int RN>
__device__ void commit2(const Ctx& c, Scratch2& R, uint32_t d, int s, int j, int par) {
    const int act = d & 7, i = j - HX8, dir = (d & 8) ? 1 : -1;
    uint8_t poke = 0;
    if (act == 1) {
        set_moved(c, s, j, false);
        poke = (d & 64) ? 1 : 0;
    } else if (act == 2) {
        poke = (d & 64) ? 1 : 0;
        if (R.claimDn[par][i + 1 + dir] == i) {
            const int sb = rsn<RN>(s, 1), jd = j + dir;
            CellR tile = ldc(c, s, j);
            const CellR diag = ldc(c, sb, jd);
            if (d & 16) {
                stc(c, s, jd, diag, dir < 0 ? (uint8_t)(F_DIRTY | F_VISITED) : F_DIRTY);
                stc(c, s, j, nothing(c), F_DIRTY);
            } else {
                stc(c, s, j, diag, F_DIRTY | F_VISITED);
            }
            if (d & 32) tile.moved = 0;
            stc(c, sb, jd, tile, F_DIRTY | F_VISITED);
        }
    } else if (act == 3) {  // 1728-1745
        const float fd = FD(s, j);
        const float a = FL(s, j) + fd;
        if (a < FLUID_MinValue) {
            stc(c, s, j, nothing(c), F_DIRTY | F_VISITED);
        } else {
            FL(s, j) = a;
            FD(s, j) = 0.0f;
            FLG(s, j) = FLG(s, j) | F_DIRTY | F_VISITED;
            ROWMOD[s] = 1;
            if (fd != 0.0f) ROWCHG[s] = 1;
        }
    } else if (act == 4) {
        if (R.claimUp[par][i + 1 + dir] == i) {
            const int st = rsn<RN>(s, -1), jd = j + dir;
            const CellR tile = ldc(c, s, j), other = ldc(c, st, jd);
            stc(c, s, j, other, F_DIRTY);
            stc(c, st, jd, tile, F_DIRTY | F_VISITED);
        }
    }
    R.poke2[i] = poke;
}

template <int RN>
__device__ void pass2_rows(const Ctx& c, Scratch2& R, int k, int s, int cx, int cy, int t) {  // s = slot of row k
    const int y = cy + c.yoff + CHUNK - 1 - k;
    const int par = k & 1;
    const int j = HX8 + t;
    // reset next row's claim slots and flags (their last readers finished before the previous step barrier)
    R.claimDn[par ^ 1][1 + t] = 1 << 30;
    R.claimUp[par ^ 1][1 + t] = 1 << 30;
    if (t == 0) {
        R.claimDn[par ^ 1][0] = R.claimUp[par ^ 1][0] = 1 << 30;
        R.claimDn[par ^ 1][CHUNK + 1] = R.claimUp[par ^ 1][CHUNK + 1] = 1 << 30;
        R.p2_any[par ^ 1] = 0;
        R.p2_poke[par ^ 1] = 0;
    }
    const uint32_t d = decide2<RN>(c, s, j, cx + t, y);
    if (d) R.p2_any[par] = 1;
    if (d & 64) R.p2_poke[par] = 1;
    if ((d & 7) == 2) atomicMin(&R.claimDn[par][t + 1 + ((d & 8) ? 1 : -1)], t);
    if ((d & 7) == 4) atomicMin(&R.claimUp[par][t + 1 + ((d & 8) ? 1 : -1)], t);
    pass_bar(2);
    if (!R.p2_any[par]) return;
    commit2<RN>(c, R, d, s, j, par);
    if (R.p2_poke[par]) {  // 1658-1673: "moved" handed to the sand below, after the slides
        pass_bar(2);
        const int sb = rsn<RN>(s, 1);
        if (R.poke2[t] && PHYS(sb, j) == P_SAND) set_moved(c, sb, j, true);
    }
}

// ---- pass 3 (world.cpp:1828-1891) ----------------------------------------------------------------------------------------
__device__ int decide3(const Ctx& c, int s, int j, int x, int y) {  // 0 none, -1 / +1 move, 2 steam condenses
    if (FLG(s, j) & F_VISITED) return 0;
    const uint8_t m = MAT(s, j);
    if (LUTP->phys[m] != P_GAS) return 0;
    const int l = PHYS(s, j - 1), r = PHYS(s, j + 1);
    const uint32_t cb = rng_cell(c.rkey, x, y);
    if (l == P_AIR && !(r == P_AIR && rng_draw(cb, S_GAS3) % 2 == 0)) return -1;
    if (r == P_AIR) return 1;
    if ((int)m == c.steam && rng_draw(cb, S_STEAM) % 10 == 0) return 2;
    return 0;
}

__device__ void commit3(const Ctx& c, Scratch3& R, int d, int s, int j, int x, int y, int par) {
    const int i = j - HX8;
    if (d == 2) {
        stc(c, s, j, create(c, c.water, x, y), F_DIRTY);
    } else if (d != 0 && R.claim3[par][i + 1 + d] == i) {
        const CellR tile = ldc(c, s, j), other = ldc(c, s, j + d);
        stc(c, s, j, other, F_DIRTY);
        stc(c, s, j + d, tile, F_DIRTY | F_VISITED);
    }
}

template <int RN>
__device__ void pass3_rows(const Ctx& c, Scratch3& R, int k, int cx, int cy, int lane) {
    const int s = slotk<RN>(c, k);
    const int y = cy + c.yoff + CHUNK - 1 - k;
    const int par = k & 1;
    // quick vote: any unvisited GAS in this row?
    bool gas = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int j = HX8 + lane + 32 * q;
        gas |= !(FLG(s, j) & F_VISITED) && PHYS(s, j) == P_GAS;
    }
    if (!__any_sync(0xffffffffu, gas)) return;
    for (int q = lane; q < CHUNK + 2; q += 32) R.claim3[par][q] = 1 << 30;
    __syncwarp();
    int d[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int i = lane + 32 * q;
        d[q] = decide3(c, s, HX8 + i, cx + i, y);
        if (d[q] == 1 || d[q] == -1) atomicMin(&R.claim3[par][i + 1 + d[q]], i);
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int i = lane + 32 * q;
        commit3(c, R, d[q], s, HX8 + i, cx + i, y, par);
    }
}

__global__ void __launch_bounds__(ROWS_THREADS, 2) tick_rows_kernel(const __grid_constant__ TickParams P) {
    unsigned char* const smem_raw = fse_smem;
    SmemRows& S = *reinterpret_cast<SmemRows*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int role = warp < 4 ? 0 : (warp < 8 ? 1 : (warp == 8 ? 2 : 3));  // pass 1, pass 2, pass 3, IO
    const int t = tid & 127;

    int cxi, cyi;
    const int bid = (int)blockIdx.x + P.chunk_base;
    if (P.list_count && bid >= *P.list_count) return;
    if (P.chunk_list) {
        int v = P.chunk_list[bid];
        cxi = v & 0xffff;
        cyi = v >> 16;
    } else {
        cxi = bid % P.ncx;
        cyi = bid / P.ncx;
    }
    const int cx = P.x0 + cxi * 2 * CHUNK;
    const int cy = P.y0 + cyi * 2 * CHUNK;

    const DevTables* T = P.tabs;
    {
        const uint4* src = reinterpret_cast<const uint4*>(&T->lut);
        uint4* dst = reinterpret_cast<uint4*>(&S.h.lut);
        for (int i = tid; i < (int)(sizeof(Lut) / 16); i += blockDim.x) dst[i] = __ldg(src + i);
        uint32_t* z = reinterpret_cast<uint32_t*>(&S.rs);
        for (int i = tid; i < (int)(sizeof(RowScratch) / 4); i += blockDim.x) z[i] = 0;
    }
    __syncthreads();
    if (tid < CHUNK + 2) {
        for (int b = 0; b < 2; b++) {
            S.rs.claimDn[b][tid] = 1 << 30;
            S.rs.claimUp[b][tid] = 1 << 30;
            S.rs.claim3[b][tid] = 1 << 30;
        }
    }
    if (tid == 0) {
        for (int q = 0; q < RING; q++) mbar_init(&S.bar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    Ctx& c = S.ctx;
    if (tid == 0) {
        c.T = T;
        c.pbuf = P.pbuf;
        c.pcount = P.pcount;
        c.pcap = P.pcap;
        c.rkey = P.rkey;
        c.tick = P.tick;
        c.iter = P.iter;
        c.nmat = T->n;
        c.yoff = P.y_off;
        c.ringn = RING;
        c.ringmask = 0;
        c.koff = HALO_DN;
        c.air = T->air; c.fire = T->fire; c.water = T->water; c.lava = T->lava; c.steam = T->steam; c.obsidian = T->obsidian;
        c.flowx = P.flowx; c.flowy = P.flowy; c.W = P.W;
    }
    __syncthreads();

    const bool io = warp == 9;
    if (io && lane == 0) {
        for (int k = -HALO_DN; k < HALO_UP + PF; k++) issue_row_load_any(P, S.ring, S.bar, S.h.rowmod, S.h.rowchg, k, cx, cy);
    }
    for (int k = -HALO_DN; k < HALO_UP; k++) mbar_wait(&S.bar[slotk<RING>(c, k)], 0);

    bool io_modified = false, io_inert = true;
    long long dbg_acc = 0;
    for (int st = 0; st < N_STEPS; st++) {
        const int kw = st + HALO_UP;
        if (kw <= LAST_ROW) mbar_wait(&S.bar[slotk<RING>(c, kw)], (uint32_t)(((kw + HALO_DN) / RING) & 1));
        fence_proxy_async();
        __syncthreads();
        long long clk0 = 0;
        if (P.dbg) {  // the step barrier released when the slowest role of the previous step arrived
            const long long* a = S.rs.dbg_arrival[(st & 1) ^ 1];
            clk0 = max(max(a[0], a[1]), max(a[2], a[3]));
        }
        if (role == 0) {
            if (st < CHUNK) pass1_rows<RING>(c, S.rs, st, slotk<RING>(c, st), cx, cy, t);
        } else if (role == 1) {
            const int k = st - L12;
            if (k >= 0 && k < CHUNK) pass2_rows<RING>(c, S.rs, k, slotk<RING>(c, k), cx, cy, t);
        } else if (role == 2) {
            const int k = st - L12 - L23;
            if (k >= 0 && k < CHUNK) pass3_rows<RING>(c, S.rs, k, cx, cy, lane);
        } else if (io) {
            const int ks = st - STORE_LAG;
            if (ks >= -HALO_WR && ks <= LAST_ROW) {
                const int q = slotk<RING>(c, ks);
                if (P.awake && ks >= 0 && ks < CHUNK) io_inert &= row_is_inert(c, q, rsn<RING>(q, 1), lane);
                io_modified |= S.h.rowchg[q] != 0;
                if (S.h.rowmod[q]) {
                    uint32_t* fw = reinterpret_cast<uint32_t*>(S.ring + q * ROW_BYTES + OFF_FLG);
                    for (int w = lane; w < P8 / 4; w += 32) fw[w] &= 0x7f7f7f7fU;
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) issue_row_store_any(P, S.ring, ks, cx, cy);
                }
                if (lane == 0) bulk_commit();
            }
            const int kl = st + HALO_UP + PF;
            if (kl <= LAST_ROW && lane == 0) {
                bulk_wait_read<1>();
                issue_row_load_any(P, S.ring, S.bar, S.h.rowmod, S.h.rowchg, kl, cx, cy);
            }
        }
        if (P.dbg) {
            const long long now = clock64();
            if (st > 0) dbg_acc += now - clk0;
            if ((tid & 127) == 0 || tid == 256 || tid == 288) S.rs.dbg_arrival[st & 1][role] = now;
        }
    }
    if (P.dbg && (tid == 0 || tid == 128 || tid == 256 || tid == 288)) atomicAdd(&P.dbg[role], (unsigned long long)dbg_acc);
    if (P.dbg && tid == 0) {
        atomicAdd(&P.dbg[4], 1ULL);
        for (int q = 0; q < 6; q++)
            if (q != 3) atomicAdd(&P.dbg[5 + q], (unsigned long long)S.rs.dbg_phase[q]);
    }
    if (io) {
        if (P.awake) {
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

// ======================================================================================================================
// Per-pass kernels: the same rows schedule, one kernel per pass.  The fused kernel above keeps 28 rows of a chunk in
// shared memory (86 KB -> 2 CTAs/SM) and is bound by the latency of the pass-1 row chain; with one pass per kernel the window
// is 13-14 rows (42 KB -> 5 CTAs/SM), the instruction footprint is a third, and more chunks are in flight per SM.  The
// price is that a chunk streams through HBM more than once — affordable, the tick is latency bound, not bandwidth bound
// (profiles/r1_tick_ncu.md).  tickVisited marks of the chunk's own cells travel between the kernels in bit 7 of the flag
// plane; pass 3 clears them, so they never outlive a colour phase.
//   pass 1, pass 2: one CTA per chunk (4 compute warps + 1 IO warp), rows bottom-up through a shared-memory window
//   pass 3: its rows do not depend on each other (a gas cell only looks at and moves within its own row), so it runs one
//           warp per row straight on global memory
#ifndef FSE_P1_RN
#define FSE_P1_RN 14
#endif
#ifndef FSE_P2_RN
#define FSE_P2_RN 14
#endif
#ifndef FSE_PASS_PF
#define FSE_PASS_PF 2
#endif
#ifndef FSE_P2_PF
#define FSE_P2_PF FSE_PASS_PF
#endif
#ifndef FSE_PASS_MINB
#define FSE_PASS_MINB 5
#endif
template <int PASS>
struct PassGeom {
    static constexpr int KMIN = PASS == 1 ? -5 : -10;     // lowest row (below the chunk) that is read
    static constexpr int FULL_LO = PASS == 1 ? -5 : -1;   // rows >= FULL_LO carry all planes and may be written
    static constexpr int UP = PASS == 1 ? 5 : 1;          // rows above the current one that are touched
    static constexpr int LAST = CHUNK - 1 + UP;
    static constexpr int SL = UP + 1;                     // a row is final SL steps after its own step
    static constexpr int RN = PASS == 1 ? FSE_P1_RN : FSE_P2_RN;  // rows in the window: live rows + rows in flight
    static constexpr int PF = PASS == 1 ? FSE_PASS_PF : FSE_P2_PF;  // rows loaded ahead of the step that needs them
    // pass 2 is bound by its IO warp (7 bulk copies per row are issued one lane after the other, ~90 cycles each): there the
    // store side and the load side get a warp each.  Pass 1 (72 registers) stays at 5 warps to keep 5 CTAs per SM.
    static constexpr int THREADS = PASS == 2 ? 192 : 160;
};
// A slot that is loaded at step st was stored at step st + UP + PF - RN + SL.  With two IO warps the loader does not see the storer's
// bulk groups; it relies on the storer's wait_group.read 1 of the step before (ordered by the step barrier), which covers stores
// issued two or more steps ago.
static_assert(PassGeom<2>::UP + PassGeom<2>::PF - PassGeom<2>::RN + PassGeom<2>::SL <= -2, "pass 2: split IO warps need the slot's store two steps old");
// pass 1: live rows st-5..st+5; the row loaded at step st (st+7) takes the slot of row st-7, whose store was issued a step earlier
// pass 2: live rows st-10..st+1; the row loaded at step st (st+3) takes the slot of row st-11
static_assert(PassGeom<1>::UP + PassGeom<1>::PF - PassGeom<1>::RN <= -PassGeom<1>::SL, "pass 1 window");
static_assert(PassGeom<2>::UP + PassGeom<2>::PF - PassGeom<2>::RN < PassGeom<2>::KMIN, "pass 2 window");

template <int PASS> struct PassScratch { typedef Scratch1 type; };
template <> struct PassScratch<2> { typedef Scratch2 type; };

template <int PASS>
struct __align__(128) SmemPass {
    SmemHead h;
    unsigned char ring[PassGeom<PASS>::RN * ROW_BYTES];
    Ctx ctx;
    unsigned long long bar[PassGeom<PASS>::RN];
    uint32_t m_act[4];   // rows of the chunk this pass must run (classify_rows_kernel; pass 2: + what pass 1 changed)
    uint32_t m_lazy[4];  // pass 2: rows whose pass-1 tickVisited marks are still implicit
    uint32_t m_out[3][4];  // pass 1 in split mode (store warp only): rows with unvisited powder / gas, with unvisited gas, with unvisited liquid
    typename PassScratch<PASS>::type rs;
};

// The IO warp moves a row with one bulk copy per plane; lane p < 7 owns plane p (mat, flg, stl, tmp, col, fl, fd), so the
// seven copies of a row are issued side by side instead of one after the other by a single thread.
struct PlaneIO {
    unsigned char* g;     // plane base + first column of the chunk's window, in bytes
    size_t row_stride;    // bytes per world row
    uint32_t soff;        // offset of the plane inside a shared-memory row
    uint32_t bytes;       // bytes per window row
};
__device__ __forceinline__ PlaneIO plane_io(const TickParams& P, int lane, int cx) {
    PlaneIO io;
    const int es = lane < 3 ? 1 : (lane == 3 ? 2 : 4);
    unsigned char* base = lane == 0 ? (unsigned char*)P.p.mat : lane == 1 ? (unsigned char*)P.p.flg : lane == 2 ? (unsigned char*)P.p.stl
                        : lane == 3 ? (unsigned char*)P.p.tmp : lane == 4 ? (unsigned char*)P.p.col : lane == 5 ? (unsigned char*)P.p.fl
                                                                                                               : (unsigned char*)P.p.fd;
    io.g = base + (size_t)(cx - (lane < 3 ? HX8 : HXW)) * es;
    io.row_stride = (size_t)P.W * es;
    io.soff = lane == 0 ? OFF_MAT : lane == 1 ? OFF_FLG : lane == 2 ? OFF_STL : lane == 3 ? OFF_TMP : lane == 4 ? OFF_COL : lane == 5 ? OFF_FL : OFF_FD;
    io.bytes = (lane < 3 ? P8 : PW) * es;
    return io;
}

// whole IO warp: load row k of the chunk into its slot (rows below FULL_LO: material plane only)
template <int PASS>
__device__ __forceinline__ void pass_row_load(SmemPass<PASS>& S, const PlaneIO& io, int lane, int k, int cy, int q, int kb) {
    using G = PassGeom<PASS>;
    unsigned long long* bar = &S.bar[q];
    const bool mat_only = k < kb + G::FULL_LO;  // rows under the segment's writable range are only read for their material
    if (lane == 0) {
        S.h.rowmod[q] = 0;
        S.h.rowchg[q] = 0;
        S.h.rowvis[q] = 0;
        S.h.rowlazy[q] = (PASS == 1 && k >= 0 && k < kb) ? 1 : 0;  // rows under the segment's first row were skipped by pass 1
        mbar_expect_tx(bar, mat_only ? P8 : ROW_BYTES);
    }
    __syncwarp();
    if (lane < (mat_only ? 1 : 7))
        bulk_g2s(S.ring + q * ROW_BYTES + io.soff, io.g + (size_t)(cy + CHUNK - 1 - k) * io.row_stride, io.bytes, bar);
}

// ---- settled-row skipping (north_star: "warp-vote/ballot skips settled tiles", here at chunk-row granularity) -----------------------
// Before pass 1 of a colour phase classify_rows_kernel reads the material and flag planes of the phase's chunks once, at streaming
// speed, and votes per chunk row whether pass 1 / pass 2 can change anything in it for ANY random draw:
//   a cell whose material has used up its iterations (iter >= Material::iterations, world.cpp:1093-1096 — AIR and SOLID always) only
//     gets its tickVisited mark in pass 1 and is skipped by passes 2 and 3;
//   SAND acts in pass 1 if it can sink into the cell below, has a pair interaction armed with it or a temperature reaction firing, and
//     in pass 2 if it is `moved` or can sink into a lower diagonal (1609-1654);
//   SOUP acts in pass 1 unless it is settled (`moved`, 1307) over a non-AIR cell with a sane amount, and always in pass 2 (1728-1745);
//   GAS and FIRE always act.
// A row neither pass must run is not stepped: pass 1 skips row k when its bit is clear and neither row k nor row k - 1 (what the
// decisions of row k read besides the cell itself) has been changed by an earlier step of this pass.  The tickVisited marks such a
// row would have received (iterations used up) stay IMPLICIT — nothing is written — until a running row is about to write into
// it: the running row first materialises the marks of the skipped rows within its write reach (their materials are still the ones
// they had at their own step), so every positional mark is in place before any cell moves onto it.  Rows that stay implicit to the
// end of pass 1 are handed to pass 2 as a bit mask; pass 2 materialises them the same way, just before a running row can touch
// them.  A chunk with no row to run exits before loading anything.  Results are bit-identical to stepping every row.
__device__ __forceinline__ void materialize_marks(const Ctx& c, int slot, int j) {
    if (c.iter >= (int)LUTP->iters[MAT(slot, j)]) {
        FLG(slot, j) = FLG(slot, j) | F_VISITED;
        ROWVIS[slot] = 1;
    }
}

// grid = 4 CTAs per chunk of the launch, 1024 threads: warp w of CTA c votes row k = 32 * c + w (counted from the chunk's bottom row)
__global__ void __launch_bounds__(1024) classify_rows_kernel(const __grid_constant__ TickParams P) {
    __shared__ uint32_t s_a1[32], s_a2[32], s_a3[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {
        const uint4* src = reinterpret_cast<const uint4*>(&P.tabs->lut);
        uint4* dst = reinterpret_cast<uint4*>(fse_smem);
        for (int i = tid; i < (int)(sizeof(Lut) / 16); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int chunk = (int)(blockIdx.x >> 2) + P.chunk_base;
    if (P.list_count && chunk >= *P.list_count) return;
    int cxi, cyi;
    if (P.chunk_list) {
        const int v = P.chunk_list[chunk];
        cxi = v & 0xffff;
        cyi = v >> 16;
    } else {
        cxi = chunk % P.ncx;
        cyi = chunk / P.ncx;
    }
    const int cx = P.x0 + cxi * 2 * CHUNK, cy = P.y0 + cyi * 2 * CHUNK;
    const int k = (int)(blockIdx.x & 3) * 32 + warp;
    const int ym = cy + CHUNK - 1 - k;  // memory row of chunk row k
    const DevTables* T = P.tabs;
    const size_t base = (size_t)ym * P.W + cx;
    const uint32_t mw = __ldg(reinterpret_cast<const uint32_t*>(P.p.mat + base) + lane);
    const uint32_t fw = __ldg(reinterpret_cast<const uint32_t*>(P.p.flg + base) + lane);
    const uint32_t bw = __ldg(reinterpret_cast<const uint32_t*>(P.p.mat + base + P.W) + lane);  // the row below
    uint32_t bL = __shfl_up_sync(0xffffffffu, bw >> 24, 1), bR = __shfl_down_sync(0xffffffffu, bw & 0xff, 1);
    if (lane == 0) bL = __ldg(P.p.mat + base + P.W - 1);
    if (lane == 31) bR = __ldg(P.p.mat + base + P.W + CHUNK);
    bool a1 = false, a2 = false, a3 = false;  // a3: powder or gas that still acts (pass 2 needs its sequential row steps for those)
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int m = (mw >> (8 * q)) & 0xff;
        if (P.iter >= (int)LUTP->iters[m]) continue;  // marked in pass 1 (implicitly), skipped afterwards
        const int ph = LUTP->phys[m];
        if (ph == P_AIR || ph == P_SOLID) continue;
        a3 |= ph == P_SAND || ph == P_GAS;
        const bool moved = (fw >> (8 * q)) & F_MOVED;
        const int mb = (bw >> (8 * q)) & 0xff;
        const int mbl = q == 0 ? (int)bL : (int)((bw >> (8 * (q - 1))) & 0xff), mbr = q == 3 ? (int)bR : (int)((bw >> (8 * (q + 1))) & 0xff);
        if (ph == P_PASSABLE) {
            a1 |= m == T->fire;
        } else if (ph == P_SAND) {
            const float d = LUTP->dens[m];
            auto sink = [&](int mm) { const int t = LUTP->phys[mm]; return t == P_AIR || (t != P_SOLID && LUTP->dens[mm] < d); };
            const uint8_t mf = LUTP->mflags[m];
            bool act = sink(mb);
            if (mf & MF_INTERACT) {
                if (mf & MF_INTERACT_SLOW) act = act || T->inter_off[m * T->n + mb + 1] > T->inter_off[m * T->n + mb];
                else {
                    const int r = LUTP->irow[m];
                    act = act || (r != 0 && ((LUTP->ibits[r - 1][mb >> 5] >> (mb & 31)) & 1u));
                }
            }
            if (mf & MF_REACT) {
                if (mf & MF_REACT_MULTI) act = true;
                else {
                    const Lut::Rx rx = LUTP->rx[m];
                    const int16_t t = __ldg(P.p.tmp + base + 4 * lane + q);
                    act = act || (rx.type == FSE_REACT_TEMPERATURE_BELOW && t < rx.thr) || (rx.type == FSE_REACT_TEMPERATURE_ABOVE && t > rx.thr);
                }
            }
            a1 |= act;
            a2 |= moved || sink(mbl) || sink(mbr);
        } else if (ph == P_SOUP) {
            a2 = true;
            if (!moved || LUTP->phys[mb] == P_AIR) a1 = true;
            else {
                const float fl = __ldg(P.p.fl + base + 4 * lane + q);
                a1 |= fl != 0.0f && fl < FLUID_MinValue;
            }
        } else {
            a1 = a2 = true;
        }
    }
    const bool r1 = __any_sync(0xffffffffu, a1), r2 = __any_sync(0xffffffffu, a2), r3 = __any_sync(0xffffffffu, a3);
    if (lane == 0) {
        s_a1[warp] = r1 ? 1u : 0u;
        s_a2[warp] = r2 ? 1u : 0u;
        s_a3[warp] = r3 ? 1u : 0u;
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t w1 = __ballot_sync(0xffffffffu, s_a1[lane] != 0), w2 = __ballot_sync(0xffffffffu, s_a2[lane] != 0);
        const uint32_t w3 = __ballot_sync(0xffffffffu, s_a3[lane] != 0);
        if (lane == 0) {
            uint32_t* o = P.rowmask + (size_t)(cyi * P.ncx + cxi) * ROWMASK_WORDS + (blockIdx.x & 3);
            o[0] = w1;
            o[4] = w2;
            o[8] = 0;
            if (P.phase_rows) {
                atomicAdd(P.phase_rows, (unsigned int)__popc(w1 | w2));  // rows some pass must run (fse_tick's skip gate)
                if (w3) atomicAdd(P.phase_rows + 16, (unsigned int)__popc(w3));  // rows with live powder / gas (fse_tick's pass-2 split gate)
            }
        }
    }
}

// One pass over one chunk by the whole CTA (4 compute warps + IO warp).  The material LUT is already in shared memory; scratch,
// mbarriers and the context are (re)initialised here.  On return every bulk store of the pass has completed.
// SKIP = false: the instantiation without row masks (every row is stepped, one segment); the skip gate of fse_tick picks it for phases
// where too few rows are settled to pay for the classification — none of the mask bookkeeping is compiled into it.
template <int PASS, bool SKIP>
__device__ __forceinline__ void run_pass(const TickParams& P, int cx, int cy, int iter, uint32_t rkey, unsigned int* cost_slot, int mask_idx) {
    using G = PassGeom<PASS>;
    unsigned char* const smem_raw = fse_smem;
    SmemPass<PASS>& S = *reinterpret_cast<SmemPass<PASS>*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // warps 0-3 compute; warp 4 stores rows back; the last warp loads rows (the same warp in a 160-thread CTA)
    const bool io = warp >= 4, io_store = warp == 4, io_load = warp == (int)(blockDim.x >> 5) - 1;
    const long long t_begin = (PASS == 1 && cost_slot) ? clock64() : 0;
#ifdef FSE_ROLE_CYCLES
    const long long t_begin_dbg = clock64();
#endif
    const DevTables* T = P.tabs;
    // settled-row skipping (see classify_rows_kernel): the rows this pass must run; a chunk without any is done
    uint32_t* const gmask = (SKIP && P.rowmask) ? P.rowmask + (size_t)mask_idx * ROWMASK_WORDS : nullptr;
    // pass 1 without row skipping, split mode: the store warp looks at every row as it becomes final and tells pass 2 which rows still
    // hold powder or gas that has not moved (they need pass 2's sequential row steps) and which hold nothing but liquid waiting for its
    // fluidAmountDiff (tick_pass2_apply_kernel takes those, all at once)
    uint32_t* const omask = (PASS == 1 && !SKIP && P.split && P.rowmask) ? P.rowmask + (size_t)mask_idx * ROWMASK_WORDS : nullptr;
    if (omask && tid < 12) (&S.m_out[0][0])[tid] = 0u;  // (a __syncthreads follows before the first row is stored)
    if (gmask) {
        const uint32_t* ga = gmask + (PASS == 1 ? 0 : 4);
        const uint32_t a0 = ga[0], a1 = ga[1], a2 = ga[2], a3 = ga[3];
        // (with active-chunk tracking pass 2 still walks the chunk: its store warp decides whether the chunk may sleep)
        if (!(a0 | a1 | a2 | a3) && !(PASS == 2 && P.chunk_state)) {
            if (PASS == 1) {
                if (tid < 4) gmask[8 + tid] = 0xffffffffu;  // every row's marks stay implicit
                if (cost_slot && tid == 0) *cost_slot = 1u;
                if (P.chunk_state && tid == 0) P.chunk_state[((cy + P.y_off) / CHUNK) * P.acols + cx / CHUNK] = 0u;
            }
            return;
        }
        if (tid == 0) {
            S.m_act[0] = a0; S.m_act[1] = a1; S.m_act[2] = a2; S.m_act[3] = a3;
            for (int q = 0; q < 4; q++) S.m_lazy[q] = PASS == 2 ? gmask[8 + q] : 0u;
        }
    } else if (tid == 0) {
        for (int q = 0; q < 4; q++) {
            S.m_act[q] = 0xffffffffu;
            S.m_lazy[q] = 0u;
        }
    }
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(&S.rs);
        for (int i = tid; i < (int)(sizeof(S.rs) / 4); i += blockDim.x) z[i] = 0;
    }
    __syncthreads();
    if (PASS == 2) {
        Scratch2& R2 = reinterpret_cast<Scratch2&>(S.rs);
        for (int i = tid; i < CHUNK + 2; i += blockDim.x)
            for (int b = 0; b < 2; b++) R2.claimDn[b][i] = R2.claimUp[b][i] = 1 << 30;
    }
    Ctx& c = S.ctx;
    if (tid == 0) {
        for (int q = 0; q < G::RN; q++) mbar_init(&S.bar[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        c.T = T;
        c.pbuf = P.pbuf;
        c.pcount = P.pcount;
        c.pcap = P.pcap;
        c.rkey = rkey;
        c.tick = P.tick;
        c.iter = iter;
        c.nmat = T->n;
        c.yoff = P.y_off;
        c.ringn = G::RN;
        c.ringmask = 0;
        c.koff = -G::KMIN;
        c.air = T->air; c.fire = T->fire; c.water = T->water; c.lava = T->lava; c.steam = T->steam; c.obsidian = T->obsidian;
        c.flowx = P.flowx; c.flowy = P.flowy; c.W = P.W;
    }
    __syncthreads();
    // ---- segments ----------------------------------------------------------------------------------------------------------------
    // The pass walks the chunk bottom-up in SEGMENTS of consecutive rows.  A segment starts at a row classify_rows_kernel found
    // active (kb), loads its window like the start of a chunk (rows kb + KMIN ...), and goes on for as long as rows may still run:
    // up to the highest row a running row could have touched (stop_row = row + UP) and across gaps of at most SEG_GAP settled rows
    // to the next active one.  Past that the segment drains its stores (SL steps) and the pass jumps to the next active row: the
    // settled rows in between are neither loaded, stepped nor stored.  Inside a segment row k lives in slot (k - kb - KMIN) % RN and
    // is the ((k - kb - KMIN) / RN)-th user of that slot's mbarrier.  Every warp derives run / stop_row / the segment end from the
    // same shared-memory flags after the same step barrier (a skipped step writes none of them, a running step only sets them), so
    // the control flow is uniform over the CTA.  With active-chunk tracking (the store warp inspects every row) and without row
    // masks there is one segment, rows 0 .. 127.
    constexpr int SEG_GAP = 16;
    static_assert(SEG_GAP >= -G::KMIN + G::UP + G::PF + 1, "a new segment's window must not overlap rows the previous one had in flight");
    const bool jumps = SKIP && gmask != nullptr && !P.chunk_state;
    auto act_bit = [&](int k) -> bool { return (S.m_act[k >> 5] >> (k & 31)) & 1u; };
    auto next_act = [&](int k) -> int {  // first active row >= k, CHUNK if none
        while (k < CHUNK) {
            const uint32_t wbits = S.m_act[k >> 5] >> (k & 31);
            if (wbits) return k + __ffs(wbits) - 1;
            k = (k | 31) + 1;
        }
        return CHUNK;
    };
    const PlaneIO pio = plane_io(P, lane < 7 ? lane : 0, cx);
    bool io_modified = false, io_inert = true;  // active-chunk tracking (IO warp): see tick_chunk_kernel
    int last_run = -8;                          // compute warps: last row that was stepped (uniform over the CTA)
    // store warp, lane w < 4 holds word w.  pass 1: io_lazy = rows whose marks stay implicit (all of them until a row is seen at its
    // store), io_chg = rows pass 2 must run on top of its own classification; pass 2: io_lazy = rows that stopped being implicit
    uint32_t io_lazy = PASS == 1 ? 0xffffffffu : 0u, io_chg = 0;
    int kb = jumps ? next_act(0) : 0;
    bool first_segment = true;
    // Phase of every slot's mbarrier as the compute warps have seen it (bit q = parity the next wait on slot q expects).  Each load arms
    // its slot's barrier once and the compute warps wait for every loaded row exactly once, segment after segment, so the barriers are
    // initialised once per pass and never invalidated.
    static_assert(G::PF <= G::SL, "the drain of a segment must cover the rows that were loaded ahead");
    uint32_t ph = 0;
    auto wait_slot = [&](int q) {
        mbar_wait(&S.bar[q], (ph >> q) & 1u);
        ph ^= 1u << q;
    };
#pragma unroll 1
    for (;;) {
        // ---- segment start: slots free, scratch flags reset, window of row kb loaded ----
        if (!first_segment) {
            if (io_store) bulk_wait_read<0>();  // the previous segment's stores have left shared memory
            __syncthreads();
            // (the mbarriers carry on: every row a segment loads is waited for by the compute warps before the segment ends — the drain
            // lasts SL >= PF steps — so each slot's barrier has completed as many phases as the warps have counted in `ph`)
            if (!io) {  // both parities of the per-row scratch flags (a step only resets the next row's)
                if (PASS == 1) {
                    Scratch1& R = reinterpret_cast<Scratch1&>(S.rs);
                    if (tid < 2) {
                        R.p1_any[tid] = R.p1_area[tid] = R.p1_horiz[tid] = 0;
                        R.area_mask[tid][0] = R.area_mask[tid][1] = R.area_mask[tid][2] = R.area_mask[tid][3] = 0;
                    }
                } else {
                    Scratch2& R = reinterpret_cast<Scratch2&>(S.rs);
                    for (int i = tid; i < CHUNK + 2; i += CHUNK)
                        for (int bb = 0; bb < 2; bb++) R.claimDn[bb][i] = R.claimUp[bb][i] = 1 << 30;
                    if (tid < 2) R.p2_any[tid] = R.p2_poke[tid] = 0;
                }
            }
        }
        if (tid == 0) c.koff = -G::KMIN - kb;
        __syncthreads();
        first_segment = false;
        const int k0 = kb + G::KMIN;  // lowest row of the segment's window
        auto slot_of = [&](int k) -> int { return (k - k0) % G::RN; };
        if (io_load) {
#pragma unroll 1
            for (int k = k0; k < kb + G::UP + G::PF && k <= G::LAST; k++) pass_row_load<PASS>(S, pio, lane, k, cy, slot_of(k), kb);
        }
        if (!io)
            for (int k = k0; k < kb + G::UP && k <= G::LAST; k++) wait_slot(slot_of(k));

        int stop_row = kb;     // highest row that may still have to run
        bool ending = false;   // past the segment's last row: drain the stores, load nothing
        int seg_end = 0;
        // slots advance by one per step: kept as counters (the kernel is bound by dependent instructions, a modulo per use shows)
        auto inc = [](int q) -> int { return q + 1 == G::RN ? 0 : q + 1; };
        int qs = (-G::KMIN) % G::RN;                                    // slot of row st
        int qw = (-G::KMIN + G::UP) % G::RN;                            // slot of row st + UP (the row that becomes live)
        int qst = ((-G::KMIN - G::SL) % G::RN + G::RN) % G::RN;         // slot of row st - SL (stored this step)
        int ql = (-G::KMIN + G::UP + G::PF) % G::RN;                    // slot of row st + UP + PF (loaded this step)
#pragma unroll 1
        for (int st = kb;; st++) {
            const int kw = st + G::UP;
            // the IO warp never reads the row that is about to become live: only the compute warps wait for it
            if (!io && kw <= G::LAST && (!ending || kw < seg_end + G::UP + G::PF)) wait_slot(qw);
            fence_proxy_async();
            __syncthreads();
            bool run = false;
            const int q = qs, qb = qs == 0 ? G::RN - 1 : qs - 1;
            if (!ending) {
                // run the row if it was classified active, or if an earlier step of this pass changed it or the row below it
                run = st < CHUNK && (!SKIP || act_bit(st) || S.h.rowchg[q] || S.h.rowchg[qb]);
                if (run) {
                    if (st + G::UP > stop_row) stop_row = st + G::UP;
                } else if (st > stop_row) {  // nothing can run here any more unless a classified row follows closely
                    const int nA = jumps ? next_act(st + 1) : CHUNK;
                    if (!jumps && st <= G::LAST) stop_row = G::LAST;
                    else if (nA < CHUNK && nA - st <= SEG_GAP) stop_row = nA;
                    else {
                        ending = true;
                        seg_end = st;
                    }
                }
            }
            if (!io) {
                if (run) {
                    const int j = HX8 + tid;
                    if (PASS == 1) {
                        // rows skipped since the last running row, as far down as this row can write (5 rows): marks in place first
                        if (SKIP) {
                            int lo = last_run + 1;
                            if (lo < st - 5) lo = st - 5;
                            if (lo < 0) lo = 0;  // rows below the chunk belong to other chunks: never marked from here
                            for (int r = lo; r < st; r++) materialize_marks(c, slot_of(r), j);
                        }
                        pass1_rows<G::RN>(c, reinterpret_cast<Scratch1&>(S.rs), st, q, cx, cy, tid);
                    } else {
                        // rows st - 1 .. st + 1 are the ones this row can write: implicit pass-1 marks become real ones first
                        if (SKIP) {
                            int lo = last_run + 2;  // rows up to last_run + 1 were handled by that step
                            if (lo < st - 1) lo = st - 1;
                            if (lo < 0) lo = 0;
                            for (int r = lo; r <= st + 1 && r < CHUNK; r++)
                                if ((S.m_lazy[r >> 5] >> (r & 31)) & 1u) materialize_marks(c, slot_of(r), j);
                        }
                        pass2_rows<G::RN>(c, reinterpret_cast<Scratch2&>(S.rs), st, q, cx, cy, tid);
                    }
                    last_run = st;
                } else if (!ending && st < CHUNK) {
                    const int par = st & 1;  // what a running step resets for the next row (pass1_rows / pass2_rows)
                    if (PASS == 1) {
                        Scratch1& R = reinterpret_cast<Scratch1&>(S.rs);
                        if (tid == 0) {
                            R.p1_any[par ^ 1] = 0;
                            R.p1_area[par ^ 1] = 0;
                            R.p1_horiz[par ^ 1] = 0;
                            R.area_mask[par ^ 1][0] = R.area_mask[par ^ 1][1] = R.area_mask[par ^ 1][2] = R.area_mask[par ^ 1][3] = 0;
                            S.h.rowlazy[q] = 1;
                        }
                    } else {
                        Scratch2& R = reinterpret_cast<Scratch2&>(S.rs);
                        R.claimDn[par ^ 1][1 + tid] = 1 << 30;
                        R.claimUp[par ^ 1][1 + tid] = 1 << 30;
                        if (tid == 0) {
                            R.claimDn[par ^ 1][0] = R.claimUp[par ^ 1][0] = 1 << 30;
                            R.claimDn[par ^ 1][CHUNK + 1] = R.claimUp[par ^ 1][CHUNK + 1] = 1 << 30;
                            R.p2_any[par ^ 1] = 0;
                            R.p2_poke[par ^ 1] = 0;
                        }
                    }
                }
            } else {
                const int ks = st - G::SL;
                if (io_store && ks >= kb + G::FULL_LO && ks <= G::LAST) {
                    const int qs_ = qst;
                    uint32_t* fw = reinterpret_cast<uint32_t*>(S.ring + qs_ * ROW_BYTES + OFF_FLG);
                    const bool core_row = ks >= 0 && ks < CHUNK;
                    const bool all_store = S.h.rowmod[qs_] != 0;
                    const bool vis_store = S.h.rowvis[qs_] != 0;
                    if (PASS == 1 && !SKIP && omask && core_row) {  // lane l: columns 4l .. 4l + 3 of the chunk's own cells
                        const uint32_t mw4 = *reinterpret_cast<const uint32_t*>(S.ring + qs_ * ROW_BYTES + OFF_MAT + HX8 + 4 * lane);
                        const uint32_t fw4 = *reinterpret_cast<const uint32_t*>(S.ring + qs_ * ROW_BYTES + OFF_FLG + HX8 + 4 * lane);
                        bool sg = false, gs = false, sp = false;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            if ((fw4 >> (8 * b)) & F_VISITED) continue;  // pass 2 skips visited cells (world.cpp:1600)
                            const int ph = LUTP->phys[(mw4 >> (8 * b)) & 0xffu];
                            sg |= ph == P_SAND || ph == P_GAS;
                            gs |= ph == P_GAS;
                            sp |= ph == P_SOUP;
                        }
                        const bool r_sg = __any_sync(0xffffffffu, sg), r_gs = __any_sync(0xffffffffu, gs), r_sp = __any_sync(0xffffffffu, sp);
                        if (lane == 0) {  // the store warp is the only writer of these words
                            if (r_sg) S.m_out[0][ks >> 5] |= 1u << (ks & 31);
                            if (r_gs) S.m_out[1][ks >> 5] |= 1u << (ks & 31);
                            if (r_sp) S.m_out[2][ks >> 5] |= 1u << (ks & 31);
                        }
                    }
                    if (PASS == 2 && gmask && core_row && (all_store || vis_store) && lane == (ks >> 5)) io_lazy |= 1u << (ks & 31);
                    if (PASS == 1 && gmask) {
                        if (core_row && !(S.h.rowlazy[qs_] && !all_store && !vis_store) && lane == (ks >> 5)) io_lazy &= ~(1u << (ks & 31));
                        if (S.h.rowchg[qs_]) {  // pass 2 reads a row's own cells and the row below them
                            if (core_row && lane == (ks >> 5)) io_chg |= 1u << (ks & 31);
                            if (ks + 1 >= 0 && ks + 1 < CHUNK && lane == ((ks + 1) >> 5)) io_chg |= 1u << ((ks + 1) & 31);
                        }
                    }
                    // tickVisited of the chunk's own cells goes to HBM (the later passes need it); halo cells are cleared.  Only rows
                    // that are stored need it; a core row has 4 + 4 halo words (one word per lane 0..7), other rows are cleared whole
                    if (all_store || vis_store) {
                        if (core_row) {
                            if (lane < 2 * (HX8 / 4)) fw[lane < HX8 / 4 ? lane : (HX8 + CHUNK) / 4 + lane - HX8 / 4] &= 0x7f7f7f7fU;
                        } else {
                            for (int w = lane; w < P8 / 4; w += 32) fw[w] &= 0x7f7f7f7fU;
                        }
                    }
                    if (P.chunk_state) {
                        io_modified |= S.h.rowchg[qs_] != 0;
                        // rows are final for passes 1 and 2 here; pass 3 only moves GAS, which is never inert anyway
                        if (PASS == 2 && core_row) io_inert &= row_is_inert(c, qs_, rsn<G::RN>(qs_, 1), lane);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (all_store ? lane < 7 : (vis_store && lane == 1))
                        bulk_s2g(pio.g + (size_t)(cy + CHUNK - 1 - ks) * pio.row_stride, S.ring + qs_ * ROW_BYTES + pio.soff, pio.bytes);
                }
                const int kl = st + G::UP + G::PF;
                const bool load = !ending && kl <= G::LAST;
                if (io_store) {
                    bulk_commit();  // one (possibly empty) bulk group per lane and step, so wait_group counts steps
                    if (load) {
                        // each lane waits until its own stores out of the slot to fill have left shared memory (bulk groups are per
                        // thread): with one spare row in the window that store was issued a step ago and this does not stall
                        if (G::UP + G::PF - G::RN < -G::SL) bulk_wait_read<1>();
                        else bulk_wait_read<0>();
                    }
                }
                if (io_load && load) pass_row_load<PASS>(S, pio, lane, kl, cy, ql, kb);
            }
            if (ending && st >= seg_end + G::SL - 1) break;
            qs = inc(qs);
            qst = inc(qst);
            ql = inc(ql);
            qw = inc(qw);
        }
        // rows at and above the segment's end: the next classified-active one starts the next segment
        const int nxt = jumps ? next_act(seg_end) : CHUNK;
        if (nxt >= CHUNK) break;
        kb = nxt;
    }
    if (PASS == 1 && cost_slot && tid == 0) *cost_slot = (unsigned int)(clock64() - t_begin);
#ifdef FSE_ROLE_CYCLES
    if (PASS == 1 && P.dbg && tid == 0) {  // scripts/role_cycles.py: words 32..41 = pass-1 phase cycles and row counts, 30 = chunks, 31 = pass cycles
        const Scratch1& R1 = reinterpret_cast<const Scratch1&>(S.rs);
        for (int q = 0; q < 10; q++) atomicAdd(&P.dbg[32 + q], (unsigned long long)R1.dbg_phase[q]);
        atomicAdd(&P.dbg[30], 1ULL);
        atomicAdd(&P.dbg[31], (unsigned long long)(clock64() - t_begin_dbg));
    }
#endif
    if (io_store && P.chunk_state) {
        const bool inert = __all_sync(0xffffffffu, io_inert);
        const unsigned int st = (io_modified ? 1u : 0u) | ((PASS == 2 && !inert) ? 2u : 0u);
        unsigned int* slot = P.chunk_state + ((cy + P.y_off) / CHUNK) * P.acols + cx / CHUNK;
        if (lane == 0) *slot = PASS == 1 ? st : (*slot | st);  // pass 1 starts the record, pass 2 adds to it (same stream)
    }
    if (PASS == 1 && gmask && io_store && lane < 4) {
        gmask[8 + lane] = io_lazy;
        gmask[4 + lane] |= io_chg;
    }
    if (PASS == 1 && !SKIP && omask && io_store) {
        // a row of liquid may be applied ahead of the sequential rows unless gas in the row below could still move up into it (pass 2,
        // world.cpp:1799-1819, reads the row above before that row's own step)
        __syncwarp();
        if (lane < 4) {
            const uint32_t o_seq = S.m_out[0][lane], o_gas = S.m_out[1][lane], o_soup = S.m_out[2][lane];
            const uint32_t gas_below = (o_gas << 1) | (lane ? S.m_out[1][lane - 1] >> 31 : 0u);
            omask[lane] = o_soup & ~o_seq & ~gas_below;  // rows tick_pass2_apply_kernel takes
            omask[4 + lane] = o_seq | (o_soup & gas_below);  // rows pass 2 must step (liquid rows gas may enter stay in the chain)
            omask[8 + lane] = 0;                           // pass 1 stepped every row: no tickVisited mark is implicit
        }
    }
    if (PASS == 2 && gmask && io_store && lane < 4 && io_lazy) gmask[8 + lane] &= ~io_lazy;  // pass 3 applies the rule to what is left
    if (io_store) bulk_wait_all();
}

template <int PASS, bool SKIP>
__global__ void __launch_bounds__(PassGeom<PASS>::THREADS, FSE_PASS_MINB) tick_pass_kernel(const __grid_constant__ TickParams P) {
    const int tid = threadIdx.x;
    int cxi, cyi;
    const int bid = (int)blockIdx.x + P.chunk_base;
    if (P.list_count && bid >= *P.list_count) return;
    if (P.chunk_list) {
        int v = P.chunk_list[bid];
        cxi = v & 0xffff;
        cyi = v >> 16;
    } else {
        cxi = bid % P.ncx;
        cyi = bid / P.ncx;
    }
    const int cx = P.x0 + cxi * 2 * CHUNK;
    const int cy = P.y0 + cyi * 2 * CHUNK;
    {
        const uint4* src = reinterpret_cast<const uint4*>(&P.tabs->lut);
        uint4* dst = reinterpret_cast<uint4*>(fse_smem);
        for (int i = tid; i < (int)(sizeof(Lut) / 16); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    run_pass<PASS, SKIP>(P, cx, cy, P.iter, P.rkey, (PASS == 1 && P.chunk_cost) ? P.chunk_cost + cyi * P.ncx + cxi : nullptr, cyi * P.ncx + cxi);
}

// ---- pass 3 on global memory: one warp per chunk row, lane l owns columns 4l..4l+3 (world.cpp:1828-1891) ---------------------
// A gas cell moves sideways into an AIR neighbour of its own row (left first, coin flip when both are free) or, if it is
// STEAM and boxed in, condenses with probability 1/10.  All cells decide from the row as pass 2 left it; a contested AIR
// cell goes to the lower source column, i.e. a left-mover at i loses exactly when cell i-2 moves right.  Visited bits are
// dropped from the whole row afterwards (the colour phase is over).
// lazy: the row's pass-1 tickVisited marks are implicit (settled-row skipping): a cell whose material has used up its iterations
// counts as visited although its flag bit is clear.
__device__ __forceinline__ void pass3_row_global(const TickParams& P, uint32_t rkey, int cx, int ym, int lane, bool lazy) {
    const int y = ym + P.y_off;
    const DevTables* T = P.tabs;
    const uint8_t* phys = T->lut.phys;
    const size_t base = (size_t)ym * P.W + cx;
    uint32_t* flgw = reinterpret_cast<uint32_t*>(P.p.flg + base) + lane;
    uint32_t fw = __ldcg(flgw);
    const uint32_t mw = __ldcg(reinterpret_cast<const uint32_t*>(P.p.mat + base) + lane);
    const bool hadvis = (fw & 0x80808080U) != 0;
    const uint32_t fw_mem = fw;
    if (lazy) {
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (P.iter >= (int)__ldg(T->lut.iters + ((mw >> (8 * q)) & 0xff))) fw |= (uint32_t)F_VISITED << (8 * q);
    }
    int ph[4];
    bool gas = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        ph[q] = __ldg(phys + ((mw >> (8 * q)) & 0xff));
        gas |= ph[q] == P_GAS && !((fw >> (8 * q)) & F_VISITED);
    }
    if (!__any_sync(0xffffffffu, gas)) {
        if (hadvis) *flgw = fw_mem & 0x7f7f7f7fU;
        return;
    }
    // phys of the columns left of q = 0 and right of q = 3
    int phL = __shfl_up_sync(0xffffffffu, ph[3], 1), phR = __shfl_down_sync(0xffffffffu, ph[0], 1);
    if (lane == 0) phL = __ldg(phys + __ldcg(P.p.mat + base - 1));
    if (lane == 31) phR = __ldg(phys + __ldcg(P.p.mat + base + CHUNK));
    int d[4];
    uint32_t rm = 0;  // bit q: cell q moves right
#pragma unroll
    for (int q = 0; q < 4; q++) {
        d[q] = 0;
        if (ph[q] != P_GAS || ((fw >> (8 * q)) & F_VISITED)) continue;
        const int l = q == 0 ? phL : ph[q - 1], rr = q == 3 ? phR : ph[q + 1];
        const int m = (mw >> (8 * q)) & 0xff;
        const uint32_t cb = rng_cell(rkey, cx + 4 * lane + q, y);
        if (l == P_AIR && !(rr == P_AIR && rng_draw(cb, S_GAS3) % 2 == 0)) d[q] = -1;
        else if (rr == P_AIR) d[q] = 1;
        else if (m == T->steam && rng_draw(cb, S_STEAM) % 10 == 0) d[q] = 2;
        if (d[q] == 1) rm |= 1u << q;
    }
    uint32_t rmPrev = __shfl_up_sync(0xffffffffu, rm, 1);
    if (lane == 0) rmPrev = 0;
    if (hadvis) *flgw = fw_mem & 0x7f7f7f7fU;
    __syncwarp();
    if (P.chunk_state) {  // any decision changes a cell (a lost contest is rare; counting it as a change only keeps the chunk awake)
        const bool acts = d[0] | d[1] | d[2] | d[3];
        if (__any_sync(0xffffffffu, acts) && lane == 0) atomicOr(P.chunk_state + (y / CHUNK) * P.acols + cx / CHUNK, 1u);
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (d[q] == 0) continue;
        const size_t a = base + 4 * lane + q;
        if (d[q] == 2) {  // 1884-1888
            const uint64_t ct = create_color_temp(T, rkey, T->water, cx + 4 * lane + q, y);
            P.p.mat[a] = (uint8_t)T->water;
            P.p.flg[a] = F_DIRTY;
            P.p.stl[a] = 0;
            P.p.tmp[a] = (int16_t)(uint16_t)(ct >> 32);
            P.p.col[a] = (uint32_t)ct;
            P.p.fl[a] = 2.0f;
            P.p.fd[a] = 0.0f;
            continue;
        }
        if (d[q] == -1) {  // the AIR cell on the left is contested by the cell two columns to the left moving right
            const bool lost = q >= 2 ? ((rm >> (q - 2)) & 1) : ((rmPrev >> (q + 2)) & 1);
            if (lost) continue;
        }
        const size_t b = a + d[q];
        const uint8_t fa = __ldcg(P.p.flg + a), fb = __ldcg(P.p.flg + b);
        const uint8_t ma = __ldcg(P.p.mat + a), mb = __ldcg(P.p.mat + b);
        const uint8_t sa = __ldcg(P.p.stl + a), sb = __ldcg(P.p.stl + b);
        const int16_t ta = __ldcg(P.p.tmp + a), tb = __ldcg(P.p.tmp + b);
        const uint32_t ca = __ldcg(P.p.col + a), cb2 = __ldcg(P.p.col + b);
        const float la = __ldcg(P.p.fl + a), lb = __ldcg(P.p.fl + b);
        const float da = __ldcg(P.p.fd + a), db = __ldcg(P.p.fd + b);
        P.p.mat[a] = mb; P.p.mat[b] = ma;
        P.p.flg[a] = (uint8_t)(F_DIRTY | (fb & F_MOVED));
        P.p.flg[b] = (uint8_t)(F_DIRTY | (fa & F_MOVED));
        P.p.stl[a] = sb; P.p.stl[b] = sa;
        P.p.tmp[a] = tb; P.p.tmp[b] = ta;
        P.p.col[a] = cb2; P.p.col[b] = ca;
        P.p.fl[a] = lb; P.p.fl[b] = la;
        P.p.fd[a] = db; P.p.fd[b] = da;
    }
}

__global__ void __launch_bounds__(128) tick_pass3_kernel(const __grid_constant__ TickParams P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = (int)(blockIdx.x >> 5) + P.chunk_base;
    const int r = ((blockIdx.x & 31) << 2) + warp;  // memory row inside the chunk
    int cxi, cyi;
    if (P.list_count && chunk >= *P.list_count) return;
    if (P.chunk_list) {
        int v = P.chunk_list[chunk];
        cxi = v & 0xffff;
        cyi = v >> 16;
    } else {
        cxi = chunk % P.ncx;
        cyi = chunk / P.ncx;
    }
    bool lazy = false;
    if (P.rowmask) {
        const int k = CHUNK - 1 - r;
        lazy = (P.rowmask[(size_t)(cyi * P.ncx + cxi) * ROWMASK_WORDS + 8 + (k >> 5)] >> (k & 31)) & 1u;
    }
    pass3_row_global(P, P.rkey, P.x0 + cxi * 2 * CHUNK, P.y0 + cyi * 2 * CHUNK + r, lane, lazy);
}

// Pass 2 of a liquid cell (world.cpp:1728-1745) touches nothing but the cell itself: fluidAmount += fluidAmountDiff, or the cell dies.
// Rows whose unvisited cells are all liquid (and that no gas from the row below can still enter) therefore do not need pass 2's
// bottom-up row chain: one warp per row applies them straight on global memory, before the sequential rows run — a row above sees them
// applied either way, and a row below neither reads nor writes them.  Pass 1's store warp made the row lists (run_pass, omask).
__global__ void __launch_bounds__(128) tick_pass2_apply_kernel(const __grid_constant__ TickParams P) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = (int)(blockIdx.x >> 5) + P.chunk_base;
    const int r = ((blockIdx.x & 31) << 2) + warp;  // memory row inside the chunk
    int cxi, cyi;
    if (P.list_count && chunk >= *P.list_count) return;
    if (P.chunk_list) {
        int v = P.chunk_list[chunk];
        cxi = v & 0xffff;
        cyi = v >> 16;
    } else {
        cxi = chunk % P.ncx;
        cyi = chunk / P.ncx;
    }
    const int k = CHUNK - 1 - r;
    if (!((P.rowmask[(size_t)(cyi * P.ncx + cxi) * ROWMASK_WORDS + (k >> 5)] >> (k & 31)) & 1u)) return;
    const DevTables* T = P.tabs;
    const size_t base = (size_t)(P.y0 + cyi * 2 * CHUNK + r) * P.W + (P.x0 + cxi * 2 * CHUNK);
    uint32_t* flgw = reinterpret_cast<uint32_t*>(P.p.flg + base) + lane;
    const uint32_t mw = __ldcg(reinterpret_cast<const uint32_t*>(P.p.mat + base) + lane);
    uint32_t fw = __ldcg(flgw);
    const uint32_t fw0 = fw;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        if ((fw >> (8 * b)) & F_VISITED) continue;
        if (__ldg(T->lut.phys + ((mw >> (8 * b)) & 0xffu)) != P_SOUP) continue;
        const size_t g = base + 4 * lane + b;
        const float fd = __ldcg(P.p.fd + g);
        const float a = __fadd_rn(__ldcg(P.p.fl + g), fd);
        if (a < FLUID_MinValue) {  // stc(nothing, F_DIRTY | F_VISITED)
            P.p.mat[g] = (uint8_t)T->air;
            P.p.stl[g] = 0;
            P.p.tmp[g] = 0;
            P.p.col[g] = 0;
            P.p.fl[g] = 2.0f;
            P.p.fd[g] = 0.0f;
            fw = (fw & ~(0xffu << (8 * b))) | ((uint32_t)(F_DIRTY | F_VISITED) << (8 * b));
        } else {
            P.p.fl[g] = a;
            P.p.fd[g] = 0.0f;
            fw |= (uint32_t)(F_DIRTY | F_VISITED) << (8 * b);
        }
    }
    if (fw != fw0) *flgw = fw;
}

// Active-chunk bookkeeping after the three passes of a phase (per-pass kernels): wake the 3x3 chunks around a chunk whose state
// changed, put a chunk to sleep when nothing changed and every cell is provably inert (same rule as tick_chunk_kernel).
__global__ void apply_chunk_state_kernel(const TickParams P, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (P.list_count && i >= *P.list_count)) return;
    const int v = P.chunk_list[i];
    const int ci = (P.x0 + (v & 0xffff) * 2 * CHUNK) / CHUNK, cj = (P.y0 + (v >> 16) * 2 * CHUNK + P.y_off) / CHUNK;
    const unsigned int st = P.chunk_state[cj * P.acols + ci];
    if (st & 1u) {
        for (int q = 0; q < 9; q++) {
            const int ni = ci + q % 3 - 1, nj = cj + q / 3 - 1;
            if (ni >= 0 && nj >= 0 && ni < P.acols && nj < P.arows) P.awake[nj * P.acols + ni] = 1;
        }
    } else if (!(st & 2u) && !P.never_sleep) {
        P.awake[cj * P.acols + ci] = 0;
    }
}

// Longest-first launch order for the next tick: chunks of a colour are binned by the cycles pass 1 took on them this tick
// (64 bins, heaviest bin first).  A phase is ~1.3 waves of CTAs, so starting the expensive chunks first shortens its tail;
// chunks of a phase are independent, the order cannot change results.
// `members` (optional): the n chunks to order, as (cxi | cyi << 16); otherwise all chunks 0..n-1 of the colour's grid.
// parts > 1: the sorted order is dealt out to the parts of the phase like cards (rank r goes to part r % parts), so every part holds
// the same mix of heavy and light chunks, heaviest first, instead of part 0 holding all the heavy ones (part q = list entries
// [part_lo(n, parts, q), part_lo(n, parts, q + 1)), the slices launch_tick_phase launches).
__host__ __device__ __forceinline__ int part_lo(int n, int parts, int q) { return q * (n / parts) + (q < n % parts ? q : n % parts); }
__global__ void __launch_bounds__(1024) lpt_build_kernel(const unsigned int* cost, int n, int ncx, int* list, const int* members, int parts) {
    __shared__ unsigned int hist[64], base[64], maxc;
    const int tid = threadIdx.x;
    if (tid < 64) hist[tid] = 0;
    if (tid == 0) maxc = 0;
    __syncthreads();
    auto entry = [&](int i) { return members ? members[i] : ((i % ncx) | ((i / ncx) << 16)); };
    auto cost_of = [&](int v) { return cost[(v >> 16) * ncx + (v & 0xffff)]; };
    unsigned int m = 0;
    for (int i = tid; i < n; i += blockDim.x) m = max(m, cost_of(entry(i)));
    atomicMax(&maxc, m);
    __syncthreads();
    const unsigned long long scale = (unsigned long long)maxc + 1;
    for (int i = tid; i < n; i += blockDim.x) atomicAdd(&hist[63 - (int)((unsigned long long)cost_of(entry(i)) * 64 / scale)], 1u);
    __syncthreads();
    if (tid == 0) {
        unsigned int acc = 0;
        for (int b = 0; b < 64; b++) {
            base[b] = acc;
            acc += hist[b];
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const int v = entry(i);
        const unsigned int pos = atomicAdd(&base[63 - (int)((unsigned long long)cost_of(v) * 64 / scale)], 1u);
        list[parts > 1 ? part_lo(n, parts, (int)(pos % parts)) + (int)(pos / parts) : (int)pos] = v;
    }
}

}  // namespace fse
