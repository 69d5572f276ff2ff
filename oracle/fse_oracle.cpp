// fse_oracle.cpp — CPU oracle (TEST INFRASTRUCTURE ONLY; see fse_oracle.hpp header).
// Citations: paths relative to /root/reference/source/engine.
#include "fse_oracle.hpp"

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <future>
#include <mutex>
#include <queue>
#include <thread>

namespace fseo {

// core/const.h:19-34
static const int CHUNK = 128;
static const float FLUID_MaxValue = 0.5f;
static const float FLUID_MinValue = 0.0005f;
static const float FLUID_MaxCompression = 0.1f;
static const float FLUID_MinFlow = 0.05f;
static const float FLUID_MaxFlow = 8.0f;
static const float FLUID_FlowSpeed = 1.0f;

enum { AIR = 0, SOLID = 1, SAND = 2, SOUP = 3, GAS = 4, PASSABLE = 5, OBJECT = 5 };

// ---- thread pool (utils/utility.hpp:308-450: fixed worker count, futures per task) -------
class ThreadPool {
public:
    explicit ThreadPool(int n) {
        for (int i = 0; i < n; i++) workers.emplace_back([this] { loop(); });
    }
    ~ThreadPool() {
        {
            std::unique_lock<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : workers) t.join();
    }
    std::future<void> push(std::function<void()> f) {
        auto task = std::make_shared<std::packaged_task<void()>>(std::move(f));
        std::future<void> fut = task->get_future();
        {
            std::unique_lock<std::mutex> lk(mu);
            q.push([task] { (*task)(); });
        }
        cv.notify_one();
        return fut;
    }

private:
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [this] { return stop || !q.empty(); });
                if (stop && q.empty()) return;
                f = std::move(q.front());
                q.pop();
            }
            f();
        }
    }
    std::vector<std::thread> workers;
    std::queue<std::function<void()>> q;
    std::mutex mu;
    std::condition_variable cv;
    bool stop = false;
};

// ---- world container (world.cpp:43-172) --------------------------------------------
World::World(int w, int h) : width(w), height(h) {
    size_t n = (size_t)w * h;
    tiles.resize(n);
    dirty.assign(n, 0);
    visitedA.assign(n, 0);
    visitedB.assign(n, 0);
    visited = visitedA.data();
    newTemps.assign(n, 0);
}

World::~World() { delete pool; }

void World::set_materials(const fse_material* tbl, int n, const fse_special_ids& sid, const fse_interaction* inter,
                          const int32_t* io, const fse_interaction* react, const int32_t* ro) {
    // Re-point existing cells by id after the table is rebuilt.
    std::vector<uint32_t> oldIds(tiles.size());
    bool had = !mats.empty();
    if (had)
        for (size_t i = 0; i < tiles.size(); i++) oldIds[i] = tiles[i].mat ? tiles[i].mat->id : 0;
    mats.assign(n, Material());
    for (int i = 0; i < n; i++) {
        Material& m = mats[i];
        const fse_material& s = tbl[i];
        m.id = i;
        m.physicsType = s.physics;
        m.alpha = s.alpha;
        m.density = s.density;
        m.iterations = s.iterations;
        m.emit = s.emit;
        m.emitColor = s.emit_color;
        m.color = s.color;
        m.addTemp = s.add_temp;
        m.conductionSelf = s.conduction_self;
        m.conductionOther = s.conduction_other;
        m.interact = s.interact != 0;
        m.react = s.react != 0;
        m.slipperyness = s.slipperyness;
        m.createTemp = s.create_temp;
        m.colorKind = s.color_kind;
        m.jitterShift = s.jitter_shift;
        m.jitterRange = s.jitter_range;
        m.nInteractions.assign(n, 0);
        m.interactions.assign(n, {});
        for (int b = 0; b < n; b++) {
            int lo = io ? io[i * n + b] : 0, hi = io ? io[i * n + b + 1] : 0;
            for (int k = lo; k < hi; k++) m.interactions[b].push_back(inter[k]);
            m.nInteractions[b] = hi - lo;
        }
        int lo = ro ? ro[i] : 0, hi = ro ? ro[i + 1] : 0;
        for (int k = lo; k < hi; k++) m.reactions.push_back(react[k]);
        m.nReactions = hi - lo;
    }
    ids = sid;
    for (size_t i = 0; i < tiles.size(); i++) {
        uint32_t id = had ? oldIds[i] : (uint32_t)ids.air;
        if ((int)id >= n) id = ids.air;
        if (!had) tiles[i] = nothing();
        tiles[i].mat = &mats[id];
        tiles[i].id = id;
    }
}

// Tiles_NOTHING (game_datastruct.cpp:312): AIR, colour 0, temperature 0, fluidAmount 2.0 default.
Cell World::nothing() const {
    Cell c;
    c.mat = const_cast<Material*>(&mats[ids.air]);
    c.id = ids.air;
    c.color = 0;
    return c;
}

static inline uint32_t pos_hash(int x, int y) { return mix32((uint32_t)x * 0x9E3779B1U ^ mix32((uint32_t)y + 0x7F4A7C15U)); }

// TilesCreate(id,x,y) (game_datastruct.cpp:485-574) with the per-material policy of fse_material.
Cell World::create(uint32_t mat_id, int x, int y) {
    Material* m = &mats[mat_id];
    Cell c;
    c.mat = m;
    c.id = mat_id;
    c.temperature = m->createTemp;
    uint32_t col = m->color;
    if (m->colorKind == FSE_COLOR_JITTER) {
        uint32_t r = draw(S_CREATE_COLOR, x, y);
        col = col + ((r % (m->jitterRange ? m->jitterRange : 1)) << m->jitterShift);
    } else if (m->colorKind == FSE_COLOR_POSITIONAL) {
        col = col ^ (pos_hash(x, y) & 0x0f0f0fU);
    }
    c.color = col;
    return c;
}

inline uint32_t World::draw(uint32_t slot, int x, int y) const {
    if (rngMode == RngMode::LIBC) return (uint32_t)rand();
    return rng_draw(rng_cell(rkey, x, y), slot);
}

uint64_t World::particle_id(int x, int y, int iter, int k) const {
    return ((uint64_t)(curTick & 0xfffff) << 42) | ((uint64_t)(iter & 3) << 40) | ((uint64_t)(y & 0x3ffff) << 22) |
           ((uint64_t)(x & 0x3ffff) << 4) | (uint64_t)(k & 15);
}

// world.cpp:1021-1034
static inline float CalculateVerticalFlowValue(float remainingLiquid, float destLiquid) {
    float sum = remainingLiquid + destLiquid;
    float value = 0;
    if (sum <= FLUID_MaxValue) {
        value = FLUID_MaxValue;
    } else if (sum < 2 * FLUID_MaxValue + FLUID_MaxCompression) {
        value = (FLUID_MaxValue * FLUID_MaxValue + sum * FLUID_MaxCompression) / (FLUID_MaxValue + FLUID_MaxCompression);
    } else {
        value = (sum + FLUID_MaxCompression) / 2.0f;
    }
    return value;
}

static inline bool can_sink_into(const Cell& other, const Cell& me) {
    // world.cpp:1206 / 1214-1215 / 1609-1610
    int t = other.mat->physicsType;
    return t == AIR || (t != SOLID && other.mat->density < me.mat->density);
}

// ---- pass 1 visit of one cell (world.cpp:1089-1586) ---------------------------------
void World::visit1(int x, int y, int iter, std::vector<Particle>& out) {
    const int W = width;
    const int idx = x + y * W;
    if (visited[idx]) return;  // 1091
    if (iter >= tiles[idx].mat->iterations) {  // 1093-1096
        visited[idx] = 1;
        return;
    }
    Cell tile = tiles[idx];  // 1097: by-value copy; later stores write this copy back
    const int type = tile.mat->physicsType;

    if ((int)tile.mat->id == ids.fire) {  // 1101-1146
        if (rngMode == RngMode::LIBC) {
            // 1102-1107: colour flicker on the local copy only (never stored);
            // the draws are kept in LIBC mode so the rand() stream cost matches.
            if (rand() % 10 == 0) (void)rand();
        }
        if (draw(S_FIRE_EMBER, x, y) % 10 == 0) {  // 1109-1119
            Particle p;
            p.tile = tile;
            p.x = (float)x;
            p.y = (float)(y - 1);
            p.vx = ((int)(draw(S_FIRE_EMBER_VX, x, y) % 10) - 5) / 20.0f;
            p.vy = -((int)(draw(S_FIRE_EMBER_VY, x, y) % 10) / 10.0f) / 3.0f + -0.5f;
            p.ax = 0;
            p.ay = 0.01f;
            p.temporary = true;
            p.lifetime = 30;
            p.fadeTime = 10;
            p.id = particle_id(x, y, iter, 15);
            out.push_back(p);
        }
        if (draw(S_FIRE_DIE, x, y) % 150 == 0) {  // 1121-1125
            tiles[idx] = nothing();
            dirty[idx] = 1;
            visited[idx] = 1;
        } else {
            bool foundAny = false;  // 1127-1144
            for (int xx = -2; xx <= 2; xx++) {
                for (int yy = -2; yy <= 2; yy++) {
                    int j = (x + xx) + (y + yy) * W;
                    if (tiles[j].mat->physicsType == SOLID) {
                        foundAny = true;
                        if (draw(S_FIRE_IGNITE0 + (xx + 2) * 5 + (yy + 2), x, y) % 500 == 0) {
                            tiles[j] = create(ids.fire, x + xx, y + yy);  // TilesCreateFire()
                            dirty[j] = 1;
                            visited[j] = 1;
                        }
                    }
                }
            }
            if (!foundAny && draw(S_FIRE_DIE_ALONE, x, y) % 120 == 0) {
                tiles[idx] = nothing();
                dirty[idx] = 1;
                visited[idx] = 1;
            }
        }
    }

    if (type == SAND) {  // 1148-1267
        Cell belowTile = tiles[x + (y + 1) * W];
        int below = belowTile.mat->physicsType;

        // 1153-1179: pair interactions against the material below
        if (tile.mat->interact && belowTile.mat->id < (uint32_t)mats.size() && tile.mat->nInteractions[belowTile.mat->id] > 0) {
            for (int i = 0; i < tile.mat->nInteractions[belowTile.mat->id]; i++) {
                fse_interaction in = tile.mat->interactions[belowTile.mat->id][i];
                int rad = (int)in.data2;
                if (in.type == FSE_INTERACT_TRANSFORM_MATERIAL) {
                    for (int xx = in.ofs_x - rad; xx <= in.ofs_x + rad; xx++)
                        for (int yy = in.ofs_y - rad; yy <= in.ofs_y + rad; yy++) {
                            int j = (x + xx) + (y + yy) * W;
                            if (tiles[j].mat->id == belowTile.mat->id) {
                                tiles[j] = create((uint32_t)in.data1, x + xx, y + yy);
                                dirty[j] = 1;
                                visited[j] = 1;
                            }
                        }
                } else if (in.type == FSE_INTERACT_SPAWN_MATERIAL) {
                    for (int xx = in.ofs_x - rad; xx <= in.ofs_x + rad; xx++)
                        for (int yy = in.ofs_y - rad; yy <= in.ofs_y + rad; yy++) {
                            int j = (x + xx) + (y + yy) * W;
                            if ((xx == 0 && yy == 0) || (int)tiles[j].mat->id == ids.air) {
                                tiles[j] = create((uint32_t)in.data1, x + xx, y + yy);
                                dirty[j] = 1;
                                visited[j] = 1;
                            }
                        }
                }
            }
            return;  // 1178
        }

        // 1181-1204: temperature reactions (product keeps the temperature)
        if (tile.mat->react && tile.mat->nReactions > 0) {
            bool react = false;
            for (int i = 0; i < tile.mat->nReactions; i++) {
                fse_interaction in = tile.mat->reactions[i];
                bool fire = (in.type == FSE_REACT_TEMPERATURE_BELOW && tile.temperature < in.data1) ||
                            (in.type == FSE_REACT_TEMPERATURE_ABOVE && tile.temperature > in.data1);
                if (fire) {
                    tiles[idx] = create(in.data2, x, y);
                    tiles[idx].temperature = tile.temperature;
                    dirty[idx] = 1;
                    visited[idx] = 1;
                    react = true;
                }
            }
            if (react) return;
        }

        bool canMoveBelow = (below == AIR || (below != SOLID && belowTile.mat->density < tile.mat->density));  // 1206
        if (!canMoveBelow) return;

        const Cell& belowLTile = tiles[(x - 1) + (y + 1) * W];
        const Cell& belowRTile = tiles[(x + 1) + (y + 1) * W];
        bool canMoveBelowL = can_sink_into(belowLTile, tile);
        bool canMoveBelowR = can_sink_into(belowRTile, tile);

        bool hesitate = (canMoveBelowL || canMoveBelowR) && (draw(S_SAND_HESITATE, x, y) % 20 == 0);  // 1217
        if (!hesitate) {
            if (belowTile.mat->physicsType == AIR && tiles[x + (y + 2) * W].mat->physicsType == AIR &&
                tiles[x + (y + 3) * W].mat->physicsType == AIR && tiles[x + (y + 4) * W].mat->physicsType == AIR) {
                // 1218-1225: free fall -> loose particle
                tiles[idx] = belowTile;  // setTile(x, y, belowTile)
                dirty[idx] = 1;
                Particle p;
                p.tile = tile;
                p.x = (float)x;
                p.y = (float)(y + 1);
                p.vx = ((int)(draw(S_SAND_PART_VX, x, y) % 10) - 5) / 20.0f;
                p.vy = -((int)(draw(S_SAND_PART_VY, x, y) % 2) + 3) / 10.0f + 1.5f;
                p.ax = 0;
                p.ay = 0.1f;
                p.id = particle_id(x, y, iter, 14);
                out.push_back(p);
            } else {
                // 1227-1239: swap with the cell below
                tiles[idx] = belowTile;
                dirty[idx] = 1;
                if (draw(S_SAND_MOVED, x, y) % 2 == 0) tile.moved = true;
                tiles[x + (y + 1) * W] = tile;
                dirty[x + (y + 1) * W] = 1;
                visited[x + (y + 1) * W] = 1;
            }
            // 1242-1266: transmit movement to the diagonal-below sand
            if (draw(S_SAND_TX_SELF, x, y) % 2 == 0) {
                if (x > 0 && tiles[(x - 1) + (y + 1) * W].mat->physicsType == SAND) {
                    if (draw(S_SAND_TX_L, x, y) % 2 == 0) tiles[(x - 1) + (y + 1) * W].moved = true;
                }
                if (x < W - 1 && tiles[(x + 1) + (y + 1) * W].mat->physicsType == SAND) {
                    if (draw(S_SAND_TX_R, x, y) % 2 == 0) tiles[(x + 1) + (y + 1) * W].moved = true;
                }
            }
        }
    } else if (type == SOUP) {  // 1269-1568 (jongallant LiquidSimulator; `moved` == settled)
        if (tile.fluidAmount == 0.0f) return;  // 1275
        if (tile.fluidAmount < FLUID_MinValue) {  // 1277-1281
            tile.fluidAmount = 0.0f;
            tiles[idx] = tile;
            return;
        }
        // 1283-1305: free fall -> n particles
        if (tile.fluidAmount > 0.005 && tiles[x + (y + 1) * W].mat->physicsType == AIR && tiles[x + (y + 2) * W].mat->physicsType == AIR &&
            tiles[x + (y + 3) * W].mat->physicsType == AIR && tiles[x + (y + 4) * W].mat->physicsType == AIR) {
            tiles[idx] = nothing();
            dirty[idx] = 1;
            int n = (int)(tile.fluidAmount / 4);
            if (n < 1) n = 1;
            for (int i = 0; i < n; i++) {
                float amt = tile.fluidAmount / n;
                Cell nt;
                nt.mat = tile.mat;
                nt.id = tile.mat->id;
                nt.color = tile.color;
                nt.temperature = tile.temperature;
                nt.fluidAmount = amt;
                nt.fluidAmountDiff = 0;
                nt.moved = false;
                Particle p;
                p.tile = nt;
                p.x = (float)x;
                p.y = (float)(y + 1);
                p.vx = ((int)(draw(S_SOUP_PART0 + 2 * (i & 7), x, y) % 10) - 5) / 30.0f;
                p.vy = -((int)(draw(S_SOUP_PART0 + 2 * (i & 7) + 1, x, y) % 2) + 3) / 10.0f + 1.0f;
                p.ax = 0;
                p.ay = 0.1f;
                p.id = particle_id(x, y, iter, i & 7);
                out.push_back(p);
            }
            return;
        }
        if (tile.moved) return;  // 1307 (settled)

        float startValue = tile.fluidAmount;
        float remainingValue = tile.fluidAmount;

        Cell bottom = tiles[x + (y + 1) * W];  // 1312
        bool airBelow = bottom.mat->physicsType == AIR;
        if ((airBelow && iter <= 2) || (bottom.mat->id == tile.mat->id)) {  // 1315-1334
            float dstFl = bottom.mat->physicsType == SOUP ? bottom.fluidAmount : 0.0f;
            float flow = CalculateVerticalFlowValue(startValue, dstFl) - dstFl;
            if (bottom.fluidAmount > 0 && flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = std::max(flow, 0.0f);
            if (flow > std::min(FLUID_MaxFlow, startValue)) flow = std::min(FLUID_MaxFlow, startValue);
            if (flow != 0) {
                remainingValue -= flow;
                tile.fluidAmountDiff -= flow;
                Cell& d = tiles[x + (y + 1) * W];
                if (bottom.mat->physicsType == AIR) {
                    d = Cell();
                    d.mat = tile.mat;
                    d.id = tile.mat->id;
                    d.color = tile.color;
                    d.temperature = tile.temperature;
                    d.fluidAmount = 0.0f;
                }
                d.fluidAmountDiff += flow;
            }
            if (flowY_) flowY_[idx] += flow;  // 1334
        } else if (iter == 0 && bottom.mat->physicsType == SOUP && (bottom.mat->id != tile.mat->id)) {  // 1335-1341
            if (draw(S_SOUP_SWAP_DOWN, x, y) % 10 == 0) {
                tiles[idx] = bottom;
                tiles[x + (y + 1) * W] = tile;
                return;
            }
        }
        if (remainingValue < FLUID_MinValue) {  // 1343-1347
            tile.fluidAmountDiff -= remainingValue;
            tiles[idx] = tile;
            return;
        }

        Cell left = tiles[(x - 1) + y * W];  // 1349-1353
        bool canMoveLeft = (left.mat->physicsType == AIR || (left.mat->id == tile.mat->id)) && !airBelow;
        Cell right = tiles[(x + 1) + y * W];
        bool canMoveRight = (right.mat->physicsType == AIR || (right.mat->id == tile.mat->id)) && !airBelow;

        if (canMoveLeft) {  // 1355-1375
            float dstFl = left.mat->physicsType == SOUP ? left.fluidAmount : 0.0f;
            float flow = (remainingValue - dstFl) / (canMoveRight ? 3.0f : 2.0f);
            if (flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = std::max(flow, 0.0f);
            if (flow > std::min(FLUID_MaxFlow, remainingValue)) flow = std::min(FLUID_MaxFlow, remainingValue);
            if (flow != 0) {
                remainingValue -= flow;
                tile.fluidAmountDiff -= flow;
                Cell& d = tiles[(x - 1) + y * W];
                if (left.mat->physicsType == AIR) {
                    d = Cell();
                    d.mat = tile.mat;
                    d.id = tile.mat->id;
                    d.color = tile.color;
                    d.temperature = tile.temperature;
                    d.fluidAmount = 0.0f;
                }
                d.fluidAmountDiff += flow;
            }
            if (flowX_) flowX_[idx] -= flow;  // 1374
        }
        if (remainingValue < FLUID_MinValue) {  // 1377-1381
            tile.fluidAmountDiff -= remainingValue;
            tiles[idx] = tile;
            return;
        }
        if (canMoveRight) {  // 1383-1403 (divisor 2.0f in both arms)
            float dstFl = right.mat->physicsType == SOUP ? right.fluidAmount : 0.0f;
            float flow = (remainingValue - dstFl) / 2.0f;
            if (flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = std::max(flow, 0.0f);
            if (flow > std::min(FLUID_MaxFlow, remainingValue)) flow = std::min(FLUID_MaxFlow, remainingValue);
            if (flow != 0) {
                remainingValue -= flow;
                tile.fluidAmountDiff -= flow;
                Cell& d = tiles[(x + 1) + y * W];
                if (right.mat->physicsType == AIR) {
                    d = Cell();
                    d.mat = tile.mat;
                    d.id = tile.mat->id;
                    d.color = tile.color;
                    d.temperature = tile.temperature;
                    d.fluidAmount = 0.0f;
                }
                d.fluidAmountDiff += flow;
            }
            if (flowX_) flowX_[idx] += flow;  // 1402
        }
        if (remainingValue < FLUID_MinValue) {  // 1405-1409
            tile.fluidAmountDiff -= remainingValue;
            tiles[idx] = tile;
            return;
        }

        Cell top = tiles[x + (y - 1) * W];  // 1411
        if (top.mat->physicsType == AIR || (top.mat->id == tile.mat->id)) {  // 1413-1432
            float dstFl = top.mat->physicsType == SOUP ? top.fluidAmount : 0.0f;
            float flow = remainingValue - CalculateVerticalFlowValue(remainingValue, dstFl);
            if (flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
            flow = std::max(flow, 0.0f);
            if (flow > std::min(FLUID_MaxFlow, remainingValue)) flow = std::min(FLUID_MaxFlow, remainingValue);
            if (flow != 0) {
                remainingValue -= flow;
                tile.fluidAmountDiff -= flow;
                Cell& d = tiles[x + (y - 1) * W];
                if (top.mat->physicsType == AIR) {
                    d = Cell();
                    d.mat = tile.mat;
                    d.id = tile.mat->id;
                    d.color = tile.color;
                    d.temperature = tile.temperature;
                    d.fluidAmount = 0.0f;
                }
                d.fluidAmountDiff += flow;
            }
            if (flowY_) flowY_[idx] -= flow;  // 1432
        } else if (iter == 0 && top.mat->physicsType == SOUP && (top.mat->id != tile.mat->id)) {  // 1433-1439
            if (draw(S_SOUP_SWAP_UP, x, y) % 10 == 0) {
                tiles[idx] = top;
                tiles[x + (y - 1) * W] = tile;
                return;
            }
        }
        if (remainingValue < FLUID_MinValue) {  // 1441-1445
            tile.fluidAmountDiff -= remainingValue;
            tiles[idx] = tile;
            return;
        }

        if (startValue == remainingValue) {  // 1447-1458
            tile.settleCount++;
            if (tile.settleCount >= 10) tile.moved = true;
        } else {
            dirty[idx] = 1;
            if (top.mat->physicsType == SOUP) tiles[x + (y - 1) * W].moved = false;
            if (bottom.mat->physicsType == SOUP) tiles[x + (y + 1) * W].moved = false;
            if (left.mat->physicsType == SOUP) tiles[(x - 1) + y * W].moved = false;
            if (right.mat->physicsType == SOUP) tiles[(x + 1) + y * W].moved = false;
        }
        tiles[idx] = tile;  // 1460

        // 1519-1537: WATER directly above LAVA -> steam + obsidian crust
        const Cell& belowTile = tiles[x + (y + 1) * W];
        if ((int)tile.mat->id == ids.water && (int)belowTile.mat->id == ids.lava) {
            tiles[idx] = create(ids.steam, x, y);
            dirty[idx] = 1;
            tiles[x + (y + 1) * W] = create(ids.obsidian, x, y + 1);
            dirty[x + (y + 1) * W] = 1;
            visited[x + (y + 1) * W] = 1;
            for (int xx = -1; xx <= 1; xx++)
                for (int yy = 0; yy <= 2; yy++) {
                    int j = (x + xx) + (y + yy) * W;
                    if ((int)tiles[j].mat->id == ids.lava) {
                        tiles[j] = create(ids.obsidian, x + xx, y + yy);
                        dirty[j] = 1;
                        visited[j] = 1;
                    }
                }
            return;
        }
    } else if (type == GAS) {  // 1569-1585
        int above = tiles[x + (y - 1) * W].mat->physicsType;
        int aboveL = tiles[(x - 1) + (y - 1) * W].mat->physicsType;
        int aboveR = tiles[(x + 1) + (y - 1) * W].mat->physicsType;
        if (above == AIR && !((aboveL == AIR || aboveR == AIR) && draw(S_GAS1, x, y) % 2 == 0)) {
            tiles[idx] = tiles[x + (y - 1) * W];
            dirty[idx] = 1;
            tiles[x + (y - 1) * W] = tile;
            dirty[x + (y - 1) * W] = 1;
            visited[x + (y - 1) * W] = 1;
        }
    }
}

// ---- pass 2 visit (world.cpp:1594-1820) --------------------------------------------
void World::visit2(int x, int y) {
    const int W = width;
    const int idx = x + y * W;
    if (visited[idx]) return;  // 1596
    Cell tile = tiles[idx];    // 1598
    const int type = tile.mat->physicsType;

    if (type == SAND) {  // 1602-1727
        Cell belowLTile = tiles[(x - 1) + (y + 1) * W];
        Cell belowRTile = tiles[(x + 1) + (y + 1) * W];
        bool canMoveBelowL = can_sink_into(belowLTile, tile);
        bool canMoveBelowR = can_sink_into(belowRTile, tile);
        bool stoppedByFriction = !tile.moved;  // 1612
        int slipperyness = tile.mat->slipperyness;

        if (stoppedByFriction) {  // 1617-1645
            int drop = 0;
            for (int pil = 0; pil < 10; pil++) {
                int pilChL = tiles[(x - 1) + (y + 1 + pil) * W].mat->physicsType;
                int pilChR = tiles[(x + 1) + (y + 1 + pil) * W].mat->physicsType;
                if (pilChL == AIR || pilChR == AIR) drop++;
            }
            int maxStability = (int)(8 / sqrt((double)slipperyness) + 1);  // 1630
            if (drop + 1 - maxStability > 0) {
                int chance = 1000 / (drop + 1 - maxStability);
                if (chance < 1000) {
                    if (draw(S_SAND2_UNSTICK, x, y) % chance == 0) {
                        stoppedByFriction = false;
                        tiles[idx].moved = true;
                    }
                }
            }
        }
        if (stoppedByFriction || !(canMoveBelowL || canMoveBelowR)) {  // 1647-1654
            tiles[idx].moved = false;
            return;
        }
        bool shouldMove = draw(S_SAND2_SHOULD, x, y) % (2 * slipperyness) != 0;  // 1656
        if (shouldMove && (canMoveBelowL || canMoveBelowR)) {  // 1658-1673
            if (draw(S_SAND2_TX_SELF, x, y) % 2 == 0) {
                if (tiles[x + (y + 1) * W].mat->physicsType == SAND) {
                    if (draw(S_SAND2_TX_OTHER, x, y) % 2 == 0) tiles[x + (y + 1) * W].moved = true;
                }
            }
        }
        if (shouldMove && canMoveBelowL && (!canMoveBelowR || draw(S_SAND2_LR, x, y) % 2 == 0)) {  // 1675-1696
            if (tiles[(x - 1) + y * W].mat->physicsType == AIR) {
                tiles[(x - 1) + y * W] = belowLTile;
                dirty[(x - 1) + y * W] = 1;
                visited[(x - 1) + y * W] = 1;
                tiles[idx] = nothing();
                dirty[idx] = 1;
            } else {
                tiles[idx] = belowLTile;
                dirty[idx] = 1;
                visited[idx] = 1;
            }
            if (draw(S_SAND2_RESTICK, x, y) % (20 * slipperyness) == 0) tile.moved = false;
            tiles[(x - 1) + (y + 1) * W] = tile;
            dirty[(x - 1) + (y + 1) * W] = 1;
            visited[(x - 1) + (y + 1) * W] = 1;
        } else if (shouldMove && canMoveBelowR) {  // 1698-1719 (no visited mark on (x+1,y))
            if (tiles[(x + 1) + y * W].mat->physicsType == AIR) {
                tiles[(x + 1) + y * W] = belowRTile;
                dirty[(x + 1) + y * W] = 1;
                tiles[idx] = nothing();
                dirty[idx] = 1;
            } else {
                tiles[idx] = belowRTile;
                dirty[idx] = 1;
                visited[idx] = 1;
            }
            if (draw(S_SAND2_RESTICK, x, y) % (20 * slipperyness) == 0) tile.moved = false;
            tiles[(x + 1) + (y + 1) * W] = tile;
            dirty[(x + 1) + (y + 1) * W] = 1;
            visited[(x + 1) + (y + 1) * W] = 1;
        } else {  // 1721-1727
            tiles[idx].moved = false;
        }
    } else if (type == SOUP) {  // 1728-1745
        tile.fluidAmount += tile.fluidAmountDiff;
        tile.fluidAmountDiff = 0.0f;
        if (tile.fluidAmount < FLUID_MinValue) {
            tiles[idx] = nothing();
        } else {
            tiles[idx] = tile;
        }
        dirty[idx] = 1;
        visited[idx] = 1;
    } else if (type == GAS) {  // 1799-1819
        int aboveL = tiles[(x - 1) + (y - 1) * W].mat->physicsType;
        int aboveR = tiles[(x + 1) + (y - 1) * W].mat->physicsType;
        if (aboveL == AIR && !(aboveR == AIR && draw(S_GAS2, x, y) % 2 == 0)) {
            tiles[idx] = tiles[(x - 1) + (y - 1) * W];
            dirty[idx] = 1;
            tiles[(x - 1) + (y - 1) * W] = tile;
            dirty[(x - 1) + (y - 1) * W] = 1;
            visited[(x - 1) + (y - 1) * W] = 1;
        } else if (aboveR == AIR) {
            tiles[idx] = tiles[(x + 1) + (y - 1) * W];
            dirty[idx] = 1;
            tiles[(x + 1) + (y - 1) * W] = tile;
            dirty[(x + 1) + (y - 1) * W] = 1;
            visited[(x + 1) + (y - 1) * W] = 1;
        }
    }
}

// ---- pass 3 visit (world.cpp:1828-1891) --------------------------------------------
void World::visit3(int x, int y) {
    const int W = width;
    const int idx = x + y * W;
    if (visited[idx]) return;  // 1830
    Cell tile = tiles[idx];
    const int type = tile.mat->physicsType;
    if (type == GAS) {  // 1862-1890
        int l = tiles[(x - 1) + y * W].mat->physicsType;
        int r = tiles[(x + 1) + y * W].mat->physicsType;
        if (l == AIR && !(r == AIR && draw(S_GAS3, x, y) % 2 == 0)) {
            tiles[idx] = tiles[(x - 1) + y * W];
            dirty[idx] = 1;
            tiles[(x - 1) + y * W] = tile;
            dirty[(x - 1) + y * W] = 1;
            visited[(x - 1) + y * W] = 1;
        } else if (r == AIR) {
            tiles[idx] = tiles[(x + 1) + y * W];
            dirty[idx] = 1;
            tiles[(x + 1) + y * W] = tile;
            dirty[(x + 1) + y * W] = 1;
            visited[(x + 1) + y * W] = 1;
        } else {
            if ((int)tile.mat->id == ids.steam) {
                if (draw(S_STEAM, x, y) % 10 == 0) {
                    tiles[idx] = create(ids.water, x, y);  // TilesCreateWater()
                    dirty[idx] = 1;
                }
            }
        }
    }
}

// ---- chunk task, reference visiting order (world.cpp:1084-1892) ---------------------
void World::chunk_reference(int cx, int cy, int iter, std::vector<Particle>& out) {
    for (int dy = CHUNK - 1; dy >= 0; dy--)
        for (int dx = 0; dx < CHUNK; dx++) visit1(cx + dx, cy + dy, iter, out);
    for (int dy = CHUNK - 1; dy >= 0; dy--)
        for (int dx = 0; dx < CHUNK; dx++) visit2(cx + dx, cy + dy);
    for (int dy = CHUNK - 1; dy >= 0; dy--)
        for (int dx = 0; dx < CHUNK; dx++) visit3(cx + dx, cy + dy);
}

// ---- chunk task, partitioned visiting order --------------------------
// Same chunk, passes and bottom-up rows; within a row: for c in 0..3, the 32 cells
// x = cx + 4*l + c (l = 0..31) form one sub-step.  Cells of a sub-step are >= 4 columns
// apart, so every rule with a +-1 column footprint commutes inside it.  FIRE (+-2) and
// interacting SAND (+-FSE_MAX_REACH) are classified at the start of the sub-step and
// deferred to sub-phases in which members are >= 8 (l parity) / >= 12 (l mod 3) apart.
void World::chunk_partitioned(int cx, int cy, int iter, std::vector<Particle>& out) {
    const int W = width;
    int phase[32];
    for (int dy = CHUNK - 1; dy >= 0; dy--) {
        int y = cy + dy;
        for (int c = 0; c < 4; c++) {
            bool special = false;
            for (int l = 0; l < 32; l++) {
                int x = cx + 4 * l + c;
                const Cell& t = tiles[x + y * W];
                int ph = 0;
                if ((int)t.mat->id == ids.fire) {
                    ph = 1 + (l & 1);
                } else if (t.mat->physicsType == SAND && t.mat->interact) {
                    uint32_t b = tiles[x + (y + 1) * W].mat->id;
                    if (t.mat->nInteractions[b] > 0) ph = 3 + (l % 3);
                }
                phase[l] = ph;
                special |= ph != 0;
            }
            for (int ph = 0; ph < (special ? 6 : 1); ph++)
                for (int l = 0; l < 32; l++)
                    if (phase[l] == ph) visit1(cx + 4 * l + c, y, iter, out);
        }
    }
    for (int dy = CHUNK - 1; dy >= 0; dy--)
        for (int c = 0; c < 4; c++)
            for (int l = 0; l < 32; l++) visit2(cx + 4 * l + c, cy + dy);
    for (int dy = CHUNK - 1; dy >= 0; dy--)
        for (int c = 0; c < 4; c++)
            for (int l = 0; l < 32; l++) visit3(cx + 4 * l + c, cy + dy);
}

void World::set_iteration(uint32_t seed, uint32_t tick, int iter, RngMode m) {
    rngMode = m;
    curTick = tick;
    curIter = iter;
    rkey = rng_key(seed, tick, (uint32_t)iter);
}

void World::clear_visited() {
    visited = visitedA.data();
    std::memset(visited, 0, (size_t)width * height);
}

void World::run_chunk(int cx, int cy, int iter, Schedule sched, std::vector<Particle>& out) {
    if (sched == Schedule::REFERENCE)
        chunk_reference(cx, cy, iter, out);
    else if (sched == Schedule::PARTITIONED)
        chunk_partitioned(cx, cy, iter, out);
    else
        chunk_rows(cx, cy, iter, out);
}

// ---- world::tick() (world.cpp:1036-1948, without the physicsCheck tail) -------------
void World::tick(const fse_tick_args& a, Schedule sched, RngMode rng, int threads) {
    rngMode = rng;
    curTick = a.tick;
    const fse_rect z = a.tick_zone;
    if (threads > 1 && (!pool || poolThreads != threads)) {
        delete pool;
        pool = new ThreadPool(threads);  // world.cpp:59 uses 16
        poolThreads = threads;
    }
    const size_t N = (size_t)width * height;
    bool which = false;
    std::memset(visitedA.data(), 0, N);  // 1046
    for (int iter = 0; iter < a.cell_iter; iter++) {  // 1050
        curIter = iter;
        rkey = rng_key(a.seed, a.tick, (uint32_t)iter);
        for (int tk = 0; tk < 4; tk++) {  // 1057
            int chOfsX = tk % 2;             // 0 1 0 1
            int chOfsY = 1 - ((tk % 4) / 2);  // 1 1 0 0
            visited = which ? visitedB.data() : visitedA.data();  // 1066
            uint8_t* other = which ? visitedA.data() : visitedB.data();
            std::future<void> cleared;
            if (threads > 1)
                cleared = std::async(std::launch::async, [other, N] { std::memset(other, 0, N); });  // 1067
            else
                std::memset(other, 0, N);

            std::vector<std::pair<int, int>> chunks;
            for (int cx = z.x + chOfsX * CHUNK; cx < z.x + z.w; cx += CHUNK * 2)
                for (int cy = z.y + chOfsY * CHUNK; cy < z.y + z.h; cy += CHUNK * 2) chunks.push_back({cx, cy});  // 1073-1074

            std::vector<std::vector<Particle>> parts(chunks.size());
            if (threads > 1) {
                std::vector<std::future<void>> futs;
                for (size_t i = 0; i < chunks.size(); i++) {
                    futs.push_back(pool->push([this, i, iter, sched, &chunks, &parts] {
                        run_chunk(chunks[i].first, chunks[i].second, iter, sched, parts[i]);
                    }));
                }
                for (auto& f : futs) f.get();  // 1903-1908
            } else {
                for (size_t i = 0; i < chunks.size(); i++) run_chunk(chunks[i].first, chunks[i].second, iter, sched, parts[i]);
            }
            for (auto& pv : parts) cells.insert(cells.end(), pv.begin(), pv.end());
            if (threads > 1) cleared.get();  // 1909
            which = !which;                  // 1911
        }
    }
    tickCt++;  // 1927
}

// ---- world::tickTemperature() (world.cpp:1950-2004) ---------------------------------
void World::tick_temperature(const Rect& z) {
    const int W = width;
    for (int y = (z.y + z.h) - 1; y >= z.y; y--) {
        for (int x = z.x; x < z.x + z.w; x++) {
            float n = 0.01;
            float v = 0;
            float factor = 0;
            for (int xa = -1; xa <= 1; xa++)
                for (int ya = -1; ya <= 1; ya++) {  // FN(-1,-1) FN(-1,0) FN(-1,1) FN(0,-1) ... order, 1976-1984
                    const Cell& t = tiles[(x + xa) + (y + ya) * W];
                    if (t.temperature != 0) {
                        factor = abs(t.temperature) / 64 * t.mat->conductionOther;  // int division first
                        v += t.temperature * factor;
                        n += factor;
                    }
                }
            const Cell& s = tiles[x + y * W];
            if (v != 0) {
                newTemps[x + y * W] = s.mat->addTemp + (v / n * s.mat->conductionSelf) + (s.temperature * (1 - s.mat->conductionSelf));
            } else {
                newTemps[x + y * W] = s.mat->addTemp + s.temperature;
            }
        }
    }
    for (int y = (z.y + z.h) - 1; y >= z.y; y--)
        for (int x = z.x; x < z.x + z.w; x++) tiles[x + y * W].temperature = (int16_t)newTemps[x + y * W];
}

// ---- world::tickCells() (world.cpp:2030-2195) ----------------------------------------
void World::tick_particles(const Rect& tz) {
    const int W = width, H = height;
    auto func = [&](Particle& cur) -> bool {
        if (cur.temporary && cur.lifetime <= 0) return true;  // 2033-2037
        if (cur.targetForce != 0) {  // 2039-2051
            float tdx = cur.targetX - cur.x;
            float tdy = cur.targetY - cur.y;
            float normFac = sqrtf(tdx * tdx + tdy * tdy);
            cur.vx += tdx / normFac * cur.targetForce;
            cur.vy += tdy / normFac * cur.targetForce;
            if (normFac < 100) {
                cur.vx *= 0.95f;
                cur.vy *= 0.95f;
            }
        }
        int lx = (int)cur.x;
        int ly = (int)cur.y;
        if (cur.x < 0 || (int)(cur.x) >= W || cur.y < 0 || (int)(cur.y) >= H) return true;  // 2056
        if (!(lx >= tz.x && ly >= tz.y && lx < tz.x + tz.w && ly < tz.y + tz.h)) return false;  // 2061

        cur.vx += cur.ax;
        cur.vy += cur.ay;
        int div = (int)((fabsf(cur.vx) + fabsf(cur.vy)) + 1);  // 2066
        float dvx = cur.vx / div;
        float dvy = cur.vy / div;
        for (int i = 0; i < div; i++) {
            cur.x += dvx;
            cur.y += dvy;
            if (cur.x < 0 || (int)(cur.x) >= W || cur.y < 0 || (int)(cur.y) >= H) return true;  // 2075
            Cell& here = tiles[(int)(cur.x) + (int)(cur.y) * W];
            if (!cur.phase && here.mat->physicsType != AIR) {  // 2080
                bool isObject = here.mat->physicsType == OBJECT;
                switch (cur.inObjectState) {  // 2084-2095
                    case 0:
                        cur.inObjectState = isObject ? 1 : 2;
                        break;
                    case 1:
                        if (!isObject) cur.inObjectState = 2;
                        break;
                }
                if (!isObject || cur.inObjectState == 2) {
                    if (cur.temporary) return true;  // 2098
                    if (tiles[lx + ly * W].mat->physicsType != AIR) {  // 2104: start cell occupied -> spiral
                        bool succeeded = false;
                        int X = 32, Y = 32;
                        int x = 0, y = 0, dx = 0, dy = -1;
                        int t = std::max(X, Y);
                        int maxI = t * t;
                        for (int j = 0; j < maxI; j++) {
                            if ((-X / 2 <= x) && (x <= X / 2) && (-Y / 2 <= y) && (y <= Y / 2)) {
                                int px = (int)(cur.x + x), py = (int)(cur.y + y);
                                // the reference indexes unchecked here; the oracle skips out-of-grid probes.
                                if (px >= 0 && py >= 0 && px < W && py < H) {
                                    Cell& d = tiles[px + py * W];
                                    if (d.mat->physicsType == AIR) {
                                        d = cur.tile;
                                        dirty[px + py * W] = 1;
                                        succeeded = true;
                                        break;
                                    } else if (cur.tile.mat->physicsType == SOUP && cur.tile.mat == d.mat) {
                                        d.fluidAmount += cur.tile.fluidAmount;
                                        dirty[px + py * W] = 1;
                                        succeeded = true;
                                        break;
                                    }
                                }
                            }
                            if ((x == y) || ((x < 0) && (x == -y)) || ((x > 0) && (x == 1 - y))) {
                                t = dx;
                                dx = -dy;
                                dy = t;
                            }
                            x += dx;
                            y += dy;
                        }
                        if (succeeded) return true;
                        cur.vy = -4;  // 2155-2157
                        cur.y -= 16;
                        return false;
                    } else {  // 2159-2165: deposit at the start cell
                        tiles[lx + ly * W] = cur.tile;
                        dirty[lx + ly * W] = 1;
                        return true;
                    }
                }
            }
        }
        if (cur.lifetime > 0) cur.lifetime--;  // 2170
        return false;
    };
    cells.erase(std::remove_if(cells.begin(), cells.end(), func), cells.end());  // 2179
    cells.erase(std::remove_if(cells.begin(), cells.end(), [&](const Particle& c) { return c.y > H; }), cells.end());  // 2190
}

// ---- tickCells under the GPU's schedule (see header) ------------------------------------------
// The deposit schedule in stages, so that several strip ranks can run it together (tests/strip_particles_cpu_worker.py): every
// rank integrates and proposes for the particles it owns, the proposals that target the band around a cut are exchanged, and each
// rank commits its own and its neighbours' proposals — the lowest id wins a cell wherever it came from, and every rank writes the
// winners' cells it holds.  tick_particles_rounds() below is the single-world composition of the same stages.
void World::prt_begin(const Rect& tz) {
    const int W = width, H = height;
    typedef PrtState St;
    std::vector<St>& st = prt;
    st.assign(cells.size(), St());
    // phase 1: integrate every particle against the unmodified grid (world.cpp:2032-2174)
    for (size_t i = 0; i < cells.size(); i++) {
        St& s = st[i];
        Particle cur = cells[i];
        s.status = 0;
        do {
            if (cur.temporary && cur.lifetime <= 0) { s.status = 1; break; }
            if (cur.targetForce != 0) {
                float tdx = cur.targetX - cur.x;
                float tdy = cur.targetY - cur.y;
                float normFac = sqrtf(tdx * tdx + tdy * tdy);
                cur.vx += tdx / normFac * cur.targetForce;
                cur.vy += tdy / normFac * cur.targetForce;
                if (normFac < 100) {
                    cur.vx *= 0.95f;
                    cur.vy *= 0.95f;
                }
            }
            int lx = (int)cur.x, ly = (int)cur.y;
            s.lx = lx;
            s.ly = ly;
            if (cur.x < 0 || (int)(cur.x) >= W || cur.y < 0 || (int)(cur.y) >= H) { s.status = 1; break; }
            if (!(lx >= tz.x && ly >= tz.y && lx < tz.x + tz.w && ly < tz.y + tz.h)) break;  // alive, untouched apart from attraction
            cur.vx += cur.ax;
            cur.vy += cur.ay;
            int div = (int)((fabsf(cur.vx) + fabsf(cur.vy)) + 1);
            float dvx = cur.vx / div;
            float dvy = cur.vy / div;
            bool done = false;
            for (int k = 0; k < div && !done; k++) {
                cur.x += dvx;
                cur.y += dvy;
                if (cur.x < 0 || (int)(cur.x) >= W || cur.y < 0 || (int)(cur.y) >= H) { s.status = 1; done = true; break; }
                const Cell& here = tiles[(int)(cur.x) + (int)(cur.y) * W];
                if (!cur.phase && here.mat->physicsType != AIR) {
                    bool isObject = here.mat->physicsType == OBJECT;
                    if (cur.inObjectState == 0) cur.inObjectState = isObject ? 1 : 2;
                    else if (cur.inObjectState == 1 && !isObject) cur.inObjectState = 2;
                    if (!isObject || cur.inObjectState == 2) {
                        if (cur.temporary) { s.status = 1; done = true; break; }
                        s.status = tiles[lx + ly * W].mat->physicsType != AIR ? 3 : 2;
                        done = true;
                        break;
                    }
                }
            }
            if (done) break;
            if (cur.lifetime > 0) cur.lifetime--;
        } while (false);
        s.adv = cur;
    }
}

int World::prt_propose() {
    const int W = width, H = height;
    typedef PrtState St;
    std::vector<St>& st = prt;
    prt_props.clear();
    bool any = false;
    for (size_t i = 0; i < cells.size(); i++) {
            St& s = st[i];
            if (s.status < 2) continue;
            s.cand = -1;
            s.merge = false;
            const Particle& cur = s.adv;
            if (s.status == 2) {
                if (tiles[s.lx + s.ly * W].mat->physicsType == AIR) s.cand = s.lx + (long)s.ly * W;
                else s.status = 3;
            }
            if (s.status == 3) {
                while (s.sj < 32 * 32) {  // world.cpp:2116-2147 square spiral
                    if ((-16 <= s.sx) && (s.sx <= 16) && (-16 <= s.sy) && (s.sy <= 16)) {
                        int px = (int)(cur.x + s.sx), py = (int)(cur.y + s.sy);
                        if (px >= 0 && py >= 0 && px < W && py < H) {
                            const Cell& d = tiles[px + py * W];
                            if (d.mat->physicsType == AIR) { s.cand = px + (long)py * W; break; }
                            if (cur.tile.mat->physicsType == SOUP && cur.tile.mat == d.mat) { s.cand = px + (long)py * W; s.merge = true; break; }
                        }
                    }
                    if ((s.sx == s.sy) || ((s.sx < 0) && (s.sx == -s.sy)) || ((s.sx > 0) && (s.sx == 1 - s.sy))) {
                        int t = s.sdx;
                        s.sdx = -s.sdy;
                        s.sdy = t;
                    }
                    s.sx += s.sdx;
                    s.sy += s.sdy;
                    s.sj++;
                }
                if (s.cand < 0) {  // world.cpp:2154-2157: nothing free within the spiral -> bounce
                    s.adv.vy = -4;
                    s.adv.y -= 16;
                    s.status = 0;
                    continue;
                }
            }
            fseo_proposal pr;
            std::memset(&pr, 0, sizeof pr);
            pr.cell = s.cand;
            pr.id = cur.id;
            pr.tile.mat = (uint16_t)cur.tile.mat->id;
            pr.tile.color = cur.tile.color;
            pr.tile.temp = cur.tile.temperature;
            pr.tile.moved = cur.tile.moved;
            pr.tile.settle = cur.tile.settleCount;
            pr.tile.fluid = cur.tile.fluidAmount;
            pr.tile.fluid_diff = cur.tile.fluidAmountDiff;
            pr.merge = s.merge ? 1 : 0;
            prt_props.push_back(pr);
            any = true;
        }
    (void)any;
    (void)H;
    return (int)prt_props.size();
}

// ext: proposals of other ranks (may be null); rows [hold_lo, hold_hi) are the rows this world holds (all of them in a single world)
void World::prt_commit(const fseo_proposal* ext, int n_ext, int hold_lo, int hold_hi) {
    typedef PrtState St;
    std::vector<St>& st = prt;
    std::vector<std::pair<long, uint64_t>> claims;
    claims.reserve(prt_props.size() + (size_t)n_ext);
    for (const fseo_proposal& pr : prt_props) claims.push_back({(long)pr.cell, pr.id});
    for (int i = 0; i < n_ext; i++) claims.push_back({(long)ext[i].cell, ext[i].id});
    std::sort(claims.begin(), claims.end());
    auto winner = [&](long cell) {
        auto it = std::lower_bound(claims.begin(), claims.end(), std::make_pair(cell, (uint64_t)0));
        return it->second;
    };
    for (size_t i = 0; i < cells.size(); i++) {
        St& s = st[i];
        if (s.status < 2 || s.cand < 0) continue;
        if (winner(s.cand) != s.adv.id) continue;
        if (s.merge) tiles[s.cand].fluidAmount += s.adv.tile.fluidAmount;  // 2131-2136
        else tiles[s.cand] = s.adv.tile;                                   // 2127 / 2160
        dirty[s.cand] = 1;
        s.status = 1;
    }
    for (int i = 0; i < n_ext; i++) {  // a neighbour's particle won a cell this world holds a copy of
        const fseo_proposal& pr = ext[i];
        const int row = (int)(pr.cell / width);
        if (row < hold_lo || row >= hold_hi || winner((long)pr.cell) != pr.id) continue;
        if (pr.merge) {
            tiles[pr.cell].fluidAmount += pr.tile.fluid;
        } else {
            Cell d;
            d.mat = &mats[pr.tile.mat];
            d.id = pr.tile.mat;
            d.color = pr.tile.color;
            d.temperature = pr.tile.temp;
            d.moved = pr.tile.moved != 0;
            d.settleCount = pr.tile.settle;
            d.fluidAmount = pr.tile.fluid;
            d.fluidAmountDiff = pr.tile.fluid_diff;
            tiles[pr.cell] = d;
        }
        dirty[pr.cell] = 1;
    }
}

void World::prt_end() {
    const int H = height;
    std::vector<Particle> keep;
    for (size_t i = 0; i < cells.size(); i++) {
        if (prt[i].status == 1) continue;
        const Particle& p = prt[i].status == 0 ? prt[i].adv : cells[i];  // still pending: retried next tick from its old state
        if (p.y > H) continue;                                             // 2190
        keep.push_back(p);
    }
    cells.swap(keep);
    prt.clear();
    prt_props.clear();
}

void World::tick_particles_rounds(const Rect& tz, int max_rounds) {
    prt_begin(tz);
    for (int r = 0; r < max_rounds; r++) {
        if (prt_propose() == 0) break;
        prt_commit(nullptr, 0, 0, height);
    }
    prt_end();
}

// ---- boundary helpers ------------------------------------------------------------------
void World::write_rect(int x0, int y0, int w, int h, const fse_cell* src) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const fse_cell& s = src[x + (size_t)y * w];
            int idx = (x0 + x) + (y0 + y) * width;
            Cell& d = tiles[idx];
            d.mat = &mats[s.mat];
            d.id = s.mat;
            d.color = s.color;
            d.temperature = s.temp;
            d.moved = s.moved != 0;
            d.settleCount = s.settle;
            d.fluidAmount = s.fluid;
            d.fluidAmountDiff = s.fluid_diff;
            dirty[idx] = s.dirty;
        }
}

void World::read_rect(int x0, int y0, int w, int h, fse_cell* dst) const {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int idx = (x0 + x) + (y0 + y) * width;
            const Cell& s = tiles[idx];
            fse_cell& d = dst[x + (size_t)y * w];
            std::memset(&d, 0, sizeof d);
            d.mat = (uint16_t)s.mat->id;
            d.color = s.color;
            d.temp = s.temperature;
            d.moved = s.moved ? 1 : 0;
            d.settle = s.settleCount;
            d.fluid = s.fluidAmount;
            d.fluid_diff = s.fluidAmountDiff;
            d.dirty = dirty[idx];
        }
}

void World::clear_dirty() { std::fill(dirty.begin(), dirty.end(), 0); }

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// Order-independent state hash: sum over cells of mix64 of the packed cell and position.
uint64_t cell_hash(int x, int y, const fse_cell& c) {
    uint32_t fb, db;
    std::memcpy(&fb, &c.fluid, 4);
    std::memcpy(&db, &c.fluid_diff, 4);
    uint64_t a = ((uint64_t)(uint32_t)x << 32) | (uint32_t)y;
    uint64_t b = ((uint64_t)c.mat << 48) | ((uint64_t)(c.moved & 1) << 40) | ((uint64_t)c.settle << 32) | c.color;
    uint64_t d = ((uint64_t)(uint16_t)c.temp << 32) | fb;
    uint64_t h = mix64(a + 0x9E3779B97F4A7C15ULL);
    h = mix64(h ^ b);
    h = mix64(h ^ d);
    h = mix64(h ^ db);
    return h;
}

void World::stats_rect(int x0, int y0, int w, int h, fse_stats* out) const {
    std::memset(out, 0, sizeof *out);
    fse_cell c;
    for (int y = y0; y < y0 + h; y++)
        for (int x = x0; x < x0 + w; x++) {
            read_rect(x, y, 1, 1, &c);
            out->hash += cell_hash(x, y, c);
            out->count[c.mat]++;
            if (mats[c.mat].physicsType == SOUP) out->fluid_mass[c.mat] += (double)c.fluid + (double)c.fluid_diff;
            out->n_dirty += c.dirty;
            out->n_moved += c.moved;
        }
}

// ---- default material table (game_datastruct.cpp:69-280, Appendix C of SURVEY.md) ----
MaterialTable default_materials(uint32_t seed) {
    MaterialTable T;
    auto add = [&](int phys, int slip, int alpha, float dens, int iter, int emit, uint32_t emitColor, uint32_t color, int kind) {
        fse_material m;
        std::memset(&m, 0, sizeof m);
        m.physics = phys;
        m.slipperyness = slip;
        m.alpha = (uint8_t)alpha;
        m.density = dens;
        m.iterations = iter;
        m.emit = emit;
        m.emit_color = emitColor;
        m.color = color;
        m.conduction_self = 1.0f;
        m.conduction_other = 1.0f;
        m.color_kind = (uint8_t)kind;
        T.mats.push_back(m);
        return (int)T.mats.size() - 1;
    };
    const int FIX = FSE_COLOR_FIXED, JIT = FSE_COLOR_JITTER, POS = FSE_COLOR_POSITIONAL;
    // gds.cpp:72-78 (fallback TilesCreate colour = Material::color = 0xffffffff, ctor default gds.hpp:160)
    int air = add(AIR, 0, 255, 0, 0, 16, 0, 0x000000, FIX);
    add(SOLID, 0, 255, 1, 0, 0, 0, 0xffffffff, FIX);      // 1 GENERIC_SOLID
    add(SAND, 20, 255, 10, 2, 0, 0, 0xffffffff, FIX);     // 2 GENERIC_SAND
    add(SOUP, 0, 255, 1.5f, 3, 0, 0, 0xffffffff, FIX);    // 3 GENERIC_LIQUID
    add(GAS, 0, 255, -1, 1, 0, 0, 0xffffffff, FIX);       // 4 GENERIC_GAS
    add(PASSABLE, 0, 255, 0, 0, 0, 0, 0xffffffff, FIX);   // 5 GENERIC_PASSABLE
    add(OBJECT, 0, 255, 1000.0f, 0, 0, 0, 0xffffffff, FIX);  // 6 GENERIC_OBJECT
    // gds.cpp:81-90
    add(SOLID, 0, 255, 1, 0, 0, 0, 0x808080, POS);        // 7 STONE
    int grass = add(SAND, 20, 255, 12, 1, 0, 0, (40u << 16) + (120u << 8) + 20u, JIT);  // 8 GRASS gds.cpp:358-363
    T.mats[grass].jitter_shift = 8;
    T.mats[grass].jitter_range = 20;
    int dirt = add(SAND, 8, 255, 15, 1, 0, 0, (60u << 16) + (40u << 8) + 20u, JIT);  // 9 DIRT gds.cpp:365-370
    T.mats[dirt].jitter_shift = 16;
    T.mats[dirt].jitter_range = 10;
    add(SOLID, 0, 255, 1, 0, 0, 0, 0x888888, POS);        // 10 SMOOTH_STONE
    int cobble = add(SOLID, 0, 255, 1, 0, 0, 0, 0x6b6b6b, POS);  // 11 COBBLE_STONE
    add(SOLID, 0, 255, 1, 0, 0, 0, 0x6b4a2f, POS);        // 12 SMOOTH_DIRT
    add(SOLID, 0, 255, 1, 0, 0, 0, 0x5a3d26, POS);        // 13 COBBLE_DIRT
    add(SOLID, 0, 255, 15, 2, 0, 0, 0x7a5533, POS);       // 14 SOFT_DIRT
    // gds.cpp:92-93
    int water = add(SOUP, 0, 0x80, 1.5f, 6, 40, 0x3000AFB5, 0x00B69F, FIX);  // 15 WATER gds.cpp:427-431
    T.mats[water].create_temp = -1023;
    int lava = add(SOUP, 0, 0xC0, 2, 1, 40, 0xFFFF6900, 0xFF7C00, FIX);  // 16 LAVA gds.cpp:433-437
    T.mats[lava].create_temp = 1024;
    add(SOLID, 0, 127, 1, 0, 0, 0, 0xf0f0f0, POS);        // 17 CLOUD
    int goldOre = add(SAND, 20, 255, 20, 2, 8, 0x804000, 0xd4af37, POS);     // 18 GOLD_ORE
    int goldMolten = add(SOUP, 0, 255, 20, 2, 8, 0x6FFF9B40, 0xffc84a, POS);  // 19 GOLD_MOLTEN
    int goldSolid = add(SOLID, 0, 255, 20, 2, 8, 0, 0xffd700, POS);           // 20 GOLD_SOLID
    add(SAND, 20, 255, 20, 2, 8, 0x7F442F, 0x8a5a44, POS);  // 21 IRON_ORE
    int obsidian = add(SOLID, 0, 255, 1, 0, 0, 0, 0x2a1a3a, POS);  // 22 OBSIDIAN
    int steam = add(GAS, 0, 255, -1, 1, 0, 0, 0x666666, FIX);      // 23 STEAM gds.cpp:475
    add(SAND, 8, 255, 15, 2, 0, 0, 0xffffffff, FIX);               // 24 SOFT_DIRT_SAND (fallback colour)
    int fire = add(PASSABLE, 0, 255, 20, 1, 0, 0, (255u << 16) + (100u << 8) + 50u, JIT);  // 25 FIRE gds.cpp:477-484
    T.mats[fire].jitter_shift = 8;
    T.mats[fire].jitter_range = 50;
    add(SOLID, 0, 255, 1, 0, 0, 0, 0x707070, POS);  // 26 FLAT_COBBLE_STONE
    add(SOLID, 0, 255, 1, 0, 0, 0, 0x5e4128, POS);  // 27 FLAT_COBBLE_DIRT

    // gds.cpp:124-141
    T.mats[air].conduction_self = 0.8f;
    T.mats[air].conduction_other = 0.8f;
    T.mats[lava].conduction_self = 0.5f;
    T.mats[lava].conduction_other = 0.7f;
    T.mats[lava].add_temp = 2;
    T.mats[cobble].conduction_self = 0.01f;
    T.mats[cobble].conduction_other = 0.4f;

    // gds.cpp:172-191: ten random materials, drawn from `seed` (reference: srand(time))
    uint32_t st = mix32(seed ^ 0xA5A5A5A5U);
    auto rnd = [&]() {
        st = mix32(st + 0x9E3779B9U);
        return st >> 1;
    };
    int rand0 = (int)T.mats.size();
    for (int i = 0; i < 10; i++) {
        uint32_t rgb = rnd() % 255;
        rgb = (rgb << 8) + rnd() % 255;
        rgb = (rgb << 8) + rnd() % 255;
        int type = rnd() % 2 == 0 ? (rnd() % 2 == 0 ? SAND : GAS) : SOUP;
        float dens = 0;
        if (type == SAND)
            dens = 5 + (rnd() % 1000) / 1000.0;
        else if (type == SOUP)
            dens = 4 + (rnd() % 1000) / 1000.0;
        else
            dens = 3 + (rnd() % 1000) / 1000.0;
        int alpha = type == SAND ? 255 : (int)(rnd() % 192 + 63);
        add(type, 10, alpha, dens, (int)(rnd() % 4 + 1), 0, 0, rgb, FIX);
    }
    // gds.cpp:263-269: scriptable test materials 1001..1003
    int testSand = add(SAND, 20, 255, 10, 2, 0, 0, (220u << 16) + (155u << 8) + 100u, JIT);  // gds.cpp:320-325
    T.mats[testSand].jitter_shift = 8;
    T.mats[testSand].jitter_range = 30;
    add(SAND, 20, 255, 10, 2, 0, 0, 0xdcb464, POS);   // TEST_TEXTURED_SAND
    add(SOUP, 0, 255, 1.5f, 4, 0, 0, 0x0000ff, FIX);  // TEST_LIQUID gds.cpp:337-342

    const int n = (int)T.mats.size();
    // gds.cpp:207-239: 1..3 random interactions per random material (interact stays false, D1)
    std::vector<std::vector<fse_interaction>> pair((size_t)n * n);
    for (int i = 0; i < 10; i++) {
        int mid = rand0 + i;
        int cnt = rnd() % 3 + 1;
        for (int j = 0; j < cnt; j++) {
            for (;;) {
                int imat = rand0 + (int)(rnd() % 10);
                if (imat != mid) {
                    fse_interaction in;
                    std::memset(&in, 0, sizeof in);
                    in.type = rnd() % 2 + 1;
                    in.data1 = (int16_t)(rand0 + (int)(rnd() % 10));
                    in.data2 = rnd() % 4;
                    in.ofs_x = (int)(rnd() % 5) - 2;
                    in.ofs_y = (int)(rnd() % 5) - 2;
                    pair[(size_t)mid * n + imat].push_back(in);
                    break;
                }
            }
        }
    }
    T.inter_offsets.assign((size_t)n * n + 1, 0);
    for (size_t k = 0; k < (size_t)n * n; k++) {
        for (auto& in : pair[k]) T.inter.push_back(in);
        T.inter_offsets[k + 1] = (int32_t)T.inter.size();
    }
    // gds.cpp:246-260: temperature reactions
    std::vector<std::vector<fse_interaction>> rx(n);
    auto mkreact = [&](int m, int type, int thr, int prod) {
        fse_interaction in;
        std::memset(&in, 0, sizeof in);
        in.type = type;
        in.data1 = (int16_t)thr;
        in.data2 = (uint32_t)prod;
        rx[m].push_back(in);
        T.mats[m].react = 1;
    };
    mkreact(lava, FSE_REACT_TEMPERATURE_BELOW, 512, obsidian);
    mkreact(water, FSE_REACT_TEMPERATURE_ABOVE, 128, steam);
    mkreact(goldOre, FSE_REACT_TEMPERATURE_ABOVE, 512, goldMolten);
    mkreact(goldMolten, FSE_REACT_TEMPERATURE_BELOW, 128, goldSolid);
    T.react_offsets.assign(n + 1, 0);
    for (int m = 0; m < n; m++) {
        for (auto& in : rx[m]) T.react.push_back(in);
        T.react_offsets[m + 1] = (int32_t)T.react.size();
    }
    T.ids.air = air;
    T.ids.fire = fire;
    T.ids.water = water;
    T.ids.lava = lava;
    T.ids.steam = steam;
    T.ids.obsidian = obsidian;
    return T;
}

}  // namespace fseo
