// fse_oracle.hpp — CPU oracle for the falling-sand world tick.
//
// TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference's per-tick world
// update (cstom4994/falling_sand_engine, source/engine/world.cpp) used by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the
// checker and the timed CPU baseline.  The product path (falling_sand_engine_b200/csrc)
// never includes, links or calls anything in this directory.
//
// PARITY UNPINNED: the reference ships no test, golden vector or fixture for this path
//  and cannot be compiled in this image (needs SDL2, FMOD, OpenGL).
// The oracle is therefore pinned only by invariants derived from the reference source
// (tests/test_oracle_*.py).  Every function cites the reference lines it follows
// (paths relative to /root/reference/source/engine).
//
// Two things are substituted, as BASELINE.json's north_star prescribes:
//   * libc rand() -> a counter RNG keyed on (seed, tick, iter, x, y, slot) (RngMode::SLOT);
//     RngMode::LIBC keeps rand() for the timed CPU baseline.
//   * texture-pack colour lookups in TilesCreate -> a position hash (presentation only).
//
// Two visiting orders over the same per-cell rule code:
//   * Schedule::REFERENCE   — the reference's: 4 chunk colours x 128x128 chunks, three
//     passes per chunk, rows bottom-up, columns left-to-right (world.cpp:1057-1086).
//   * Schedule::PARTITIONED — round 1's GPU schedule (its kernel was removed from the product in round 2; the order stays here as a
//     third visiting order for the schedule-independence pins of tests/test_oracle_pins.py): identical at
//     chunk/pass/row level; inside a row the 128 columns are visited as 4 interleaved
//     classes (x mod 4 = 0,1,2,3), with FIRE cells and interacting SAND cells of a class
//     deferred to sub-phases whose members are >= 8 / >= 12 columns apart.
//     Every order-independent rule gives bit-identical results under both schedules.
//   * Schedule::ROWS        — the GPU's "simultaneous rows" schedule (rows_oracle.cpp): identical at
//     chunk/pass/row level; all 128 cells of a row decide from the state before the row step, then commit: own-column
//     writes first, horizontal liquid flows gathered by their targets in a fixed order, contested cells go to the lowest x.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/fse.h"

namespace fseo {

enum class RngMode { SLOT = 0, LIBC = 1 };
enum class Schedule { REFERENCE = 0, PARTITIONED = 1, ROWS = 2 };

// RNG draw sites.  Values are part of the cross-implementation contract
// (the CUDA kernels use the same numbers).
enum Slot : uint32_t {
    S_FIRE_EMBER = 1,       // world.cpp:1109
    S_FIRE_EMBER_VX = 2,    // world.cpp:1110
    S_FIRE_EMBER_VY = 3,    // world.cpp:1110
    S_FIRE_DIE = 4,         // world.cpp:1121
    S_FIRE_IGNITE0 = 5,     // world.cpp:1132, +(xx+2)*5+(yy+2) -> 5..29
    S_FIRE_DIE_ALONE = 31,  // world.cpp:1140
    S_SAND_HESITATE = 32,   // world.cpp:1217
    S_SAND_PART_VX = 33,    // world.cpp:1222
    S_SAND_PART_VY = 34,
    S_SAND_MOVED = 35,      // world.cpp:1231
    S_SAND_TX_SELF = 36,    // world.cpp:1244
    S_SAND_TX_L = 37,       // world.cpp:1247
    S_SAND_TX_R = 38,       // world.cpp:1258
    S_SOUP_PART0 = 40,      // world.cpp:1298, +2*(i&7) vx, +1 vy -> 40..55
    S_SOUP_SWAP_DOWN = 56,  // world.cpp:1336
    S_SOUP_SWAP_UP = 57,    // world.cpp:1434
    S_GAS1 = 58,            // world.cpp:1576
    S_SAND2_UNSTICK = 59,   // world.cpp:1635
    S_SAND2_SHOULD = 60,    // world.cpp:1656
    S_SAND2_TX_SELF = 61,   // world.cpp:1661
    S_SAND2_TX_OTHER = 62,  // world.cpp:1664
    S_SAND2_LR = 63,        // world.cpp:1675
    S_SAND2_RESTICK = 64,   // world.cpp:1688 / 1711
    S_GAS2 = 65,            // world.cpp:1804
    S_GAS3 = 66,            // world.cpp:1868
    S_STEAM = 67,           // world.cpp:1884
    S_CREATE_COLOR = 68,    // TilesCreate* colour jitter (gds.cpp:320-369,485-492), keyed on the CREATED cell
    S_PROBE_X = 69,         // world.cpp:1930
    S_PROBE_Y = 70,         // world.cpp:1931
    S_BRIDGE_VX = 71,       // game.cpp:1791
    S_BRIDGE_VY = 72,       // game.cpp:1792
    S_EXPL_KEEP = 73,       // world.cpp:2306
    S_EXPL_VX = 74,         // world.cpp:2320 / 2324
    S_EXPL_VY = 75,
};

static inline uint32_t mix32(uint32_t v) {
    v ^= v >> 16; v *= 0x7feb352dU; v ^= v >> 15; v *= 0x846ca68bU; v ^= v >> 16;
    return v;
}
static inline uint32_t rng_key(uint32_t seed, uint32_t tick, uint32_t iter) {
    return mix32(seed ^ mix32(tick * 0x9E3779B9U + iter * 0x85EBCA77U + 0x1234567U));
}
static inline uint32_t rng_cell(uint32_t key, int x, int y) {
    return mix32(key ^ ((uint32_t)y * 0x9E3779B1U + (uint32_t)x));
}
// 31-bit value standing in for one rand() call (RAND_MAX = 2^31-1 on glibc).
static inline uint32_t rng_draw(uint32_t cellbase, uint32_t slot) {
    return mix32(cellbase + slot * 0x9E3779B9U) >> 1;
}

struct Material;

// `MaterialInstance` (game_datastruct.hpp:207-225); kept as the reference's 40-byte AoS
// record with a Material pointer because the CPU baseline's speed depends on it.
struct Cell {
    uint32_t id = 0;
    Material* mat = nullptr;
    uint32_t color = 0;
    int16_t temperature = 0;
    bool moved = false;
    float fluidAmount = 2.0f;
    float fluidAmountDiff = 0.0f;
    uint8_t settleCount = 0;
};
static_assert(sizeof(Cell) == 40, "reference asserts sizeof(MaterialInstance)==40 (game_datastruct.hpp:225)");

// `Material` (game_datastruct.hpp:130-168), value fields only.
struct Material {
    uint32_t id = 0;
    int physicsType = 0;
    uint8_t alpha = 0;
    float density = 0;
    int iterations = 0;
    int emit = 0;
    uint32_t emitColor = 0;
    uint32_t color = 0;
    uint32_t addTemp = 0;
    float conductionSelf = 1.0f;
    float conductionOther = 1.0f;
    bool interact = false;
    std::vector<int> nInteractions;                        // [nMat]
    std::vector<std::vector<fse_interaction>> interactions;  // [nMat][..]
    bool react = false;
    int nReactions = 0;
    std::vector<fse_interaction> reactions;
    int slipperyness = 1;
    // TilesCreate policy
    int16_t createTemp = 0;
    uint8_t colorKind = 0, jitterShift = 0, jitterRange = 0;
};

// `CellData` (game_utils/cells.h:15-36)
struct Particle {
    Cell tile;
    float x = 0, y = 0, vx = 0, vy = 0, ax = 0, ay = 0;
    float targetX = 0, targetY = 0, targetForce = 0;
    bool phase = false, temporary = false;
    int lifetime = 0;
    int fadeTime = 60;
    uint8_t inObjectState = 0;
    bool vacuum = false;  // member of the vacuum tool's Item::vacuumCells (game.cpp:2507)
    uint64_t id = 0;
};

struct Rect { int x = 0, y = 0, w = 0, h = 0; };

// state of a particle inside one tick_particles_rounds call (kept across the stages below)
struct PrtState {
    int status = 0;  // 0 alive, 1 dead, 2 wants its start cell, 3 spiral search
    Particle adv;
    int lx = 0, ly = 0;
    int sx = 0, sy = 0, sdx = 0, sdy = -1, sj = 0;
    long cand = -1;
    bool merge = false;
};
// a deposit proposal as it travels between strip ranks
struct fseo_proposal {
    int64_t cell;   // x + y * width
    uint64_t id;    // particle id: the lowest one wins the cell
    fse_cell tile;
    int32_t merge;  // 1: add the tile's fluid to the liquid already there
    int32_t _pad;
};
static_assert(sizeof(fseo_proposal) == 48, "exchanged as raw bytes");

class ThreadPool;

class World {
public:
    World(int width, int height);
    ~World();

    // materials ------------------------------------------------------------------
    void set_materials(const fse_material* tbl, int n, const fse_special_ids& ids, const fse_interaction* inter,
                       const int32_t* inter_offsets, const fse_interaction* react, const int32_t* react_offsets);
    int n_materials() const { return (int)mats.size(); }

    // cell factories (game_datastruct.cpp:303-574) -------------------------------------
    Cell nothing() const;                                      // Tiles_NOTHING
    Cell create(uint32_t mat_id, int x, int y);  // TilesCreate(id,x,y)

    // boundary ------------------------------------------------------------------
    void write_rect(int x, int y, int w, int h, const fse_cell* cells);
    void read_rect(int x, int y, int w, int h, fse_cell* cells) const;
    void clear_dirty();
    void stats_rect(int x, int y, int w, int h, fse_stats* out) const;

    // the tick (world.cpp:1036-1948) ---------------------------------------------------
    void tick(const fse_tick_args& a, Schedule sched, RngMode rng, int threads);
    // world.cpp:1950-2004
    void tick_temperature(const Rect& zone);
    // world.cpp:2030-2195
    void tick_particles(const Rect& zone);
    // Same per-particle rules under the GPU's deterministic schedule: every particle is integrated
    // against the grid as it was at the start of the call, then deposits are resolved in rounds where the lowest
    // particle id wins a contested cell.  Conflict-free particle sets give results identical to tick_particles().
    void tick_particles_rounds(const Rect& zone, int max_rounds);
    // the same schedule in stages (strip ranks exchange proposals between prt_propose and prt_commit)
    void prt_begin(const Rect& zone);
    int prt_propose();
    void prt_commit(const fseo_proposal* ext, int n_ext, int hold_lo, int hold_hi);
    void prt_end();
    std::vector<PrtState> prt;
    std::vector<fseo_proposal> prt_props;
    void add_particle(const Particle& p) { cells.push_back(p); }

    // one chunk task under either schedule (exposed for the multi-process strip test, which
    // sequences iterations and colour phases itself: set_iteration + clear_visited per phase)
    void run_chunk(int cx, int cy, int iter, Schedule sched, std::vector<Particle>& out);
    void set_iteration(uint32_t seed, uint32_t tick, int iter, RngMode m);
    void clear_visited();

    int width, height;
    std::vector<Cell> tiles;
    std::vector<uint8_t> dirty;
    std::vector<uint8_t> visitedA, visitedB;
    uint8_t* visited = nullptr;  // current tickVisited plane
    std::vector<int32_t> newTemps;
    // render-only liquid flow accumulators (world.hpp:116-119; written at world.cpp:1334, 1374, 1402, 1432, consumed and reset by the
    // dirty -> texture loop, game.cpp:2017-2062).  Off (null) until enable_flows(): the timed CPU baseline does not carry them.
    std::vector<float> flowX, flowY, prevFlowX, prevFlowY;
    float *flowX_ = nullptr, *flowY_ = nullptr;
    void enable_flows() {
        const size_t n = (size_t)width * height;
        flowX.assign(n, 0.0f); flowY.assign(n, 0.0f); prevFlowX.assign(n, 0.0f); prevFlowY.assign(n, 0.0f);
        flowX_ = flowX.data(); flowY_ = flowY.data();
    }
    // second cell layer and background colours (world.hpp:112-113 real_layer2 / background, world.cpp:1012-1019 setTileLayer2): static
    // planes next to the grid that only the chunk merge and the renderer touch; allocated on first use
    // (kept as id / colour / temperature — what a chunk file holds of it, chunk.hpp:26-30 — so a material-table rebuild leaves it alone)
    std::vector<uint32_t> layer2Id, layer2Color, background;
    std::vector<int16_t> layer2Temp;
    std::vector<uint8_t> layer2Dirty, backgroundDirty;
    void enable_layers() {
        if (!layer2Id.empty()) return;
        const size_t n = (size_t)width * height;
        layer2Id.assign(n, (uint32_t)ids.air);  // Tiles_NOTHING (world.cpp:117-119)
        layer2Color.assign(n, 0u);
        layer2Temp.assign(n, 0);
        background.assign(n, 0u);
        layer2Dirty.assign(n, 0);
        backgroundDirty.assign(n, 0);
    }
    std::vector<Particle> cells;  // world::cells
    std::vector<Material> mats;
    fse_special_ids ids{};
    uint64_t tickCt = 0;

private:
    RngMode rngMode = RngMode::SLOT;
    uint32_t rkey = 0;       // rng_key(seed,tick,iter) of the running iteration
    uint32_t curTick = 0;
    int curIter = 0;
    ThreadPool* pool = nullptr;
    int poolThreads = 0;

    inline uint32_t draw(uint32_t slot, int x, int y) const;
    void visit1(int x, int y, int iter, std::vector<Particle>& out);
    void visit2(int x, int y);
    void visit3(int x, int y);
    void chunk_reference(int cx, int cy, int iter, std::vector<Particle>& out);
    void chunk_partitioned(int cx, int cy, int iter, std::vector<Particle>& out);
    void chunk_rows(int cx, int cy, int iter, std::vector<Particle>& out);  // rows_oracle.cpp
    friend struct RowsImpl;
    uint64_t particle_id(int x, int y, int iter, int k) const;
};

// Default material table = InitMaterials() (game_datastruct.cpp:69-280) with the ten
// rand()-generated Mat_0..9 drawn from `seed` instead of srand(time(NULL)).
struct MaterialTable {
    std::vector<fse_material> mats;
    fse_special_ids ids{};
    std::vector<fse_interaction> inter;
    std::vector<int32_t> inter_offsets;  // n*n+1
    std::vector<fse_interaction> react;
    std::vector<int32_t> react_offsets;  // n+1
};
MaterialTable default_materials(uint32_t seed);

uint64_t cell_hash(int x, int y, const fse_cell& c);

}  // namespace fseo
