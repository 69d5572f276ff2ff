// fse_device.cuh — device-side data layout shared by the fse kernels (sm_100a).
//
// World state lives in HBM as structure-of-arrays planes, row-major x + y*W (y grows downward,
// as in the reference: world.cpp:1001, gravity +y world.cpp:147):
//   mat  u8   material id                         (MaterialInstance::mat->id, gds.hpp:209-210)
//   flg  u8   bit0 moved (gds.hpp:214)  bit1 dirty (world::dirty[], world.hpp:131)
//             bit7 tickVisited — only ever set in shared memory, never stored to HBM
//   stl  u8   settleCount                         (gds.hpp:217)
//   tmp  i16  temperature                         (gds.hpp:213)
//   col  u32  color                               (gds.hpp:212)
//   fl   f32  fluidAmount                         (gds.hpp:215)
//   fd   f32  fluidAmountDiff                     (gds.hpp:216)
// = 17 bytes per cell (the reference's AoS MaterialInstance is 40 bytes).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fse.h"

namespace fse {

constexpr int CHUNK = 128;
constexpr uint8_t F_MOVED = 0x01;
constexpr uint8_t F_DIRTY = 0x02;
constexpr uint8_t F_VISITED = 0x80;

enum { P_AIR = 0, P_SOLID = 1, P_SAND = 2, P_SOUP = 3, P_GAS = 4, P_PASSABLE = 5 };

// material flag bits in DevTables::mflags / Lut::mflags
constexpr uint8_t MF_INTERACT = 0x01;       // Material::interact
constexpr uint8_t MF_REACT = 0x02;          // Material::react && nReactions > 0
constexpr uint8_t MF_REACT_MULTI = 0x04;    // more than one reaction: walk the list in global memory
constexpr uint8_t MF_INTERACT_SLOW = 0x08;  // partner bitmap row overflowed: look nInteractions up in global memory
constexpr int LUT_IROWS = 8;

// Hot per-material constants, copied to shared memory by the tick kernel (built on the host in fse_materials_set).
struct Lut {
    uint8_t phys[FSE_MAX_MATERIALS];
    uint8_t iters[FSE_MAX_MATERIALS];    // Material::iterations clamped to 255
    uint8_t mflags[FSE_MAX_MATERIALS];
    uint8_t slip[FSE_MAX_MATERIALS];     // Material::slipperyness ("1 to ~127", world.cpp:1614)
    uint8_t maxstab[FSE_MAX_MATERIALS];  // int(8 / sqrt(slipperyness) + 1), world.cpp:1630
    uint8_t irow[FSE_MAX_MATERIALS];     // 1-based row of ibits for an interacting material, 0 = none
    float dens[FSE_MAX_MATERIALS];
    struct Rx { int16_t thr; uint8_t type; uint8_t prod; } rx[FSE_MAX_MATERIALS];  // first reaction of the material
    uint32_t ibits[LUT_IROWS][FSE_MAX_MATERIALS / 32];  // bit b of row r: nInteractions[b] > 0
};
static_assert(sizeof(Lut) % 16 == 0, "Lut is copied with 16-byte loads");

// RNG draw sites — same numbering as the reference-derived oracle (oracle/fse_oracle.hpp)
enum Slot : uint32_t {
    S_FIRE_EMBER = 1, S_FIRE_EMBER_VX = 2, S_FIRE_EMBER_VY = 3, S_FIRE_DIE = 4, S_FIRE_IGNITE0 = 5, S_FIRE_DIE_ALONE = 31,
    S_SAND_HESITATE = 32, S_SAND_PART_VX = 33, S_SAND_PART_VY = 34, S_SAND_MOVED = 35, S_SAND_TX_SELF = 36, S_SAND_TX_L = 37,
    S_SAND_TX_R = 38, S_SOUP_PART0 = 40, S_SOUP_SWAP_DOWN = 56, S_SOUP_SWAP_UP = 57, S_GAS1 = 58, S_SAND2_UNSTICK = 59,
    S_SAND2_SHOULD = 60, S_SAND2_TX_SELF = 61, S_SAND2_TX_OTHER = 62, S_SAND2_LR = 63, S_SAND2_RESTICK = 64, S_GAS2 = 65,
    S_GAS3 = 66, S_STEAM = 67, S_CREATE_COLOR = 68, S_PROBE_X = 69, S_PROBE_Y = 70, S_BRIDGE_VX = 71, S_BRIDGE_VY = 72,
    S_EXPL_KEEP = 73, S_EXPL_VX = 74, S_EXPL_VY = 75,
};

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t v) {
    v ^= v >> 16; v *= 0x7feb352dU; v ^= v >> 15; v *= 0x846ca68bU; v ^= v >> 16;
    return v;
}
__host__ __device__ __forceinline__ uint32_t rng_key(uint32_t seed, uint32_t tick, uint32_t iter) {
    return mix32(seed ^ mix32(tick * 0x9E3779B9U + iter * 0x85EBCA77U + 0x1234567U));
}
__host__ __device__ __forceinline__ uint32_t rng_cell(uint32_t key, int x, int y) {
    return mix32(key ^ ((uint32_t)y * 0x9E3779B1U + (uint32_t)x));
}
__host__ __device__ __forceinline__ uint32_t rng_draw(uint32_t cellbase, uint32_t slot) {
    return mix32(cellbase + slot * 0x9E3779B9U) >> 1;
}
__host__ __device__ __forceinline__ uint32_t pos_hash(int x, int y) {
    return mix32((uint32_t)x * 0x9E3779B1U ^ mix32((uint32_t)y + 0x7F4A7C15U));
}

// Flattened material table in device memory (one per context).
struct DevTables {
    Lut lut;  // first member: 16-byte aligned
    int n;
    int air, fire, water, lava, steam, obsidian;
    int _pad0;
    uint8_t phys[FSE_MAX_MATERIALS];
    uint8_t alpha[FSE_MAX_MATERIALS];
    uint8_t ckind[FSE_MAX_MATERIALS], jshift[FSE_MAX_MATERIALS], jrange[FSE_MAX_MATERIALS];
    int16_t ctemp[FSE_MAX_MATERIALS];
    float density[FSE_MAX_MATERIALS];
    uint32_t color[FSE_MAX_MATERIALS];
    uint32_t add_temp[FSE_MAX_MATERIALS];
    uint32_t emit_color[FSE_MAX_MATERIALS];  // Material::emitColor (render planes, fse_render.cu)
    float cond_self[FSE_MAX_MATERIALS];
    float cond_other[FSE_MAX_MATERIALS];
    int32_t react_off[FSE_MAX_MATERIALS + 1];
    const fse_interaction* react;   // device
    const int32_t* inter_off;       // device, n*n+1
    const fse_interaction* inter;   // device
};

struct Planes {
    uint8_t* mat;
    uint8_t* flg;
    uint8_t* stl;
    int16_t* tmp;
    uint32_t* col;
    float* fl;
    float* fd;
};

// Arguments of one colour phase of one iteration (world.cpp:1057-1077).
struct TickParams {
    Planes p;
    int W, H;
    int x0, y0;        // first chunk origin of this colour (local rows)
    int y_off;         // global y of local row 0: strip worlds key the RNG on global coordinates
    int ncx, ncy;      // chunks of this colour along x / y (stride 2*CHUNK)
    int iter;
    uint32_t rkey;     // rng_key(seed, tick, iter)
    uint32_t tick;
    fse_particle* pbuf;
    unsigned int* pcount;
    unsigned int pcap;
    const DevTables* tabs;
    const int* chunk_list;  // optional list of (cxi | cyi << 16) chunk coordinates of this colour; null = all
    const int* list_count;  // optional device count of valid chunk_list entries (active-chunk pass; grid is over-provisioned)
    uint8_t* awake;         // optional per-chunk awake flags over the whole world (acols x arows); null = tracking off
    int acols, arows;
    int never_sleep;        // strip worlds keep their cut-adjacent chunk rows awake
    unsigned long long* dbg; // optional role-cycle counters (profiling aid), null otherwise
    unsigned int* chunk_state; // per-pass kernels with active tracking: bit 0 = a pass changed cell state, bit 1 = not inert (acols x arows)
    unsigned int* chunk_cost; // optional: pass 1 records the cycles each chunk took (cost[cyi * ncx + cxi]) for the next tick's ordering
    int chunk_base;         // first chunk (index into the phase's chunk grid or list) of this launch
    int fused;              // rows schedule: 1 = single fused kernel (all passes pipelined), 0 = one kernel per pass
    int fused_max_chunks;   // rows schedule: phases of at most this many chunks run in the fused kernel anyway (one wave of it)
    int schedule;           // FSE_SCHEDULE_ROWS (simultaneous rows); kept in the parameters for the launch logic
    int lpt_parts;          // parts the chunk_list was dealt into by lpt_build_kernel (0 / 1: plain longest-first order)
    unsigned int* phase_rows; // optional counter: chunk rows of this phase that pass 1 or pass 2 must run (classify_rows_kernel)
    uint32_t* rowmask;      // per-pass kernels: ROWMASK_WORDS words per chunk of the colour's grid (index cyi * ncx + cxi), see classify_rows_kernel
    int split;              // per-pass kernels without row skipping: pass 1 classifies the rows for pass 2 (see tick_pass2_apply_kernel)
    float* flowx;           // optional render-only accumulators world::flowX / flowY (world.cpp:1334, 1374, 1402, 1432), W x H floats each,
    float* flowy;           // null = not kept (fse_flow_enable)
};
// rowmask layout per chunk: [0..3] rows pass 1 must run, [4..7] rows pass 2 must run, [8..11] rows whose pass-1 tickVisited marks
// stayed implicit (bit k = row k counted from the chunk's bottom row)
constexpr int ROWMASK_WORDS = 12;

// extra streams + events for running the parts of a colour phase side by side (nullptr: single stream)
struct TickFork {
    int parts;  // 1..4 parts of a phase run side by side; part 0 on the caller's stream
    int min_chunks;  // phases with fewer chunks are launched whole (default 256; FSE_TICK_MIN_CHUNKS)
    cudaStream_t aux[3];
    cudaEvent_t ev_fork, ev_join[3];
};

// 4-connected component labels (label = lowest cell index of the component, -1 = unset cell) by min-label propagation with pointer
// jumping, Jacobi style: every sweep reads one buffer and writes the other, so no thread reads a label another thread is writing
// (compute-sanitizer racecheck clean; an in-place sweep reaches the same fixed point through a benign race).  A holds the initial
// labels (own index for set cells, -1 otherwise), B is scratch of the same size; returns the buffer with the result.  All threads of
// the CTA must call.
__device__ inline int* ccl_relax(int* A, int* B, int n, int w, int h) {
    for (;;) {
        int changed = 0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int l = A[i];
            int best = l;
            if (l >= 0) {
                const int x = i % w, y = i / w;
                if (x + 1 < w && A[i + 1] >= 0) best = min(best, A[i + 1]);
                if (x > 0 && A[i - 1] >= 0) best = min(best, A[i - 1]);
                if (y + 1 < h && A[i + w] >= 0) best = min(best, A[i + w]);
                if (y > 0 && A[i - w] >= 0) best = min(best, A[i - w]);
                best = min(best, A[best]);  // pointer jumping
                changed |= best < l;
            }
            B[i] = best;
        }
        int* t = A;
        A = B;
        B = t;
        if (!__syncthreads_or(changed)) break;  // the barrier also orders this sweep's writes before the next sweep's reads
    }
    return A;
}

}  // namespace fse
