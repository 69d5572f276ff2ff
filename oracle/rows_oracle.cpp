// rows_oracle.cpp — CPU oracle (TEST INFRASTRUCTURE ONLY): the cell rules of world::tick (reference:
// source/engine/world.cpp:1084-1892) under the GPU's "simultaneous rows" schedule (Schedule::ROWS).
//
// Chunk colours, passes and the bottom-up row order are the reference's.  Inside one row step of one pass every cell of
// the 128-wide chunk row
//   D   decides what it wants to do from the state BEFORE the step (the reference's per-cell code, read-only);
//   C1  commits what lies in its own column (itself, the cell below / above it);
//   C2  receives horizontal liquid flows: the target adds the flow from its left neighbour, then the one from its right
//       neighbour (an AIR target becomes the left source's liquid first); flows that no longer fit are handed back;
//       "moved" pokes and un-settle flags land here as well;
//   C3  applies area effects (FIRE ignition / burn-out, water-on-lava crust, pair interactions): each effect claims its
//       target cells and the source with the lowest x wins a contested cell.
// Pass 2 and pass 3 moves into another column (sand slide, gas) claim their destination the same way; a loser stays put.
// Rules that only touch their own column (vertical sand fall, reactions, liquid pass 2, gas rising, the iteration gate)
// are untouched, so they remain bit-identical to the reference order.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "fse_oracle.hpp"

namespace fseo {

namespace {
const int CHUNK = 128;
const float FLUID_MaxValue = 0.5f, FLUID_MinValue = 0.0005f, FLUID_MaxCompression = 0.1f, FLUID_MinFlow = 0.05f, FLUID_MaxFlow = 8.0f,
            FLUID_FlowSpeed = 1.0f;
enum { AIR = 0, SOLID = 1, SAND = 2, SOUP = 3, GAS = 4 };
enum Act { A_NONE = 0, A_MARK, A_REACT, A_SAND_PART, A_SAND_SWAP, A_SOUP_ZERO, A_SOUP_PART, A_SOUP_FLOW, A_SOUP_SWAPDOWN, A_GAS_UP, A_FIRE, A_INTERACT };

float vflow(float remaining, float dest) {  // world.cpp:1021-1034
    float sum = remaining + dest, value;
    if (sum <= FLUID_MaxValue) value = FLUID_MaxValue;
    else if (sum < 2 * FLUID_MaxValue + FLUID_MaxCompression)
        value = (FLUID_MaxValue * FLUID_MaxValue + sum * FLUID_MaxCompression) / (FLUID_MaxValue + FLUID_MaxCompression);
    else value = (sum + FLUID_MaxCompression) / 2.0f;
    return value;
}
inline float clampflow(float flow, float cap, bool speed) {
    if (speed && flow > FLUID_MinFlow) flow *= FLUID_FlowSpeed;
    flow = std::max(flow, 0.0f);
    if (flow > std::min(FLUID_MaxFlow, cap)) flow = std::min(FLUID_MaxFlow, cap);
    return flow;
}

struct Dec1 {
    int act = A_NONE;
    int prod = 0;
    bool coin = false, pokeL = false, pokeR = false;
    // liquid
    float fd_new = 0, flowD = 0, flowL = 0, flowR = 0, flowU = 0;
    uint8_t stl_new = 0;
    bool moved_new = false, changed = false, swapUp = false, wl = false;
    bool botSoup = false, topSoup = false, leftSoup = false, rightSoup = false;
    // fire
    bool ember = false, die = false;
    uint32_t ignite = 0;
    // interactions
    uint32_t mb = 0;
};
struct Dec2 {
    int act = 0;  // 0 none, 1 moved=false, 2 slide, 3 soup apply, 4 gas diag
    int dir = 0;  // -1 left, +1 right
    bool riser = false, restick = false, poke = false, unstick = false;
};
}  // namespace

struct RowsImpl {
    World& W;
    int iter;
    std::vector<Particle>& out;
    int w;
    RowsImpl(World& W_, int it, std::vector<Particle>& o) : W(W_), iter(it), out(o), w(W_.width) {}

    Cell& T(int x, int y) { return W.tiles[x + y * w]; }
    int phys(int x, int y) { return T(x, y).mat->physicsType; }
    uint32_t rnd(int slot, int x, int y) { return rng_draw(rng_cell(W.rkey, x, y), slot); }
    void setc(int x, int y, const Cell& c, bool dirty, bool visited) {
        T(x, y) = c;
        if (dirty) W.dirty[x + y * w] = 1;
        if (visited) W.visited[x + y * w] = 1;
    }
    bool can_sink(int x, int y, const Cell& me) {
        const Cell& o = T(x, y);
        int t = o.mat->physicsType;
        return t == AIR || (t != SOLID && o.mat->density < me.mat->density);
    }
    Cell fresh(const Cell& t) {  // MaterialInstance(tile.mat, tile.color, tile.temperature), fluidAmount 0
        Cell c;
        c.mat = t.mat;
        c.id = t.mat->id;
        c.color = t.color;
        c.temperature = t.temperature;
        c.fluidAmount = 0.0f;
        return c;
    }

    // ---------------------------------------------------------------- pass 1: decide (world.cpp:1089-1586, read-only)
    Dec1 decide1(int x, int y) {
        Dec1 d;
        const int idx = x + y * w;
        if (W.visited[idx]) return d;
        const Cell& tile = T(x, y);
        if (iter >= tile.mat->iterations) {
            d.act = A_MARK;
            return d;
        }
        const int type = tile.mat->physicsType;
        const uint32_t m = tile.mat->id;
        if ((int)m == W.ids.fire) {  // 1101-1146
            d.act = A_FIRE;
            d.ember = rnd(S_FIRE_EMBER, x, y) % 10 == 0;
            if (rnd(S_FIRE_DIE, x, y) % 150 == 0) {
                d.die = true;
            } else {
                bool found = false;
                for (int xx = -2; xx <= 2; xx++)
                    for (int yy = -2; yy <= 2; yy++)
                        if (phys(x + xx, y + yy) == SOLID) {
                            found = true;
                            int k = (xx + 2) * 5 + (yy + 2);
                            if (rnd(S_FIRE_IGNITE0 + k, x, y) % 500 == 0) d.ignite |= 1u << k;
                        }
                if (!found && rnd(S_FIRE_DIE_ALONE, x, y) % 120 == 0) d.die = true;
            }
            return d;
        }
        if (type == SAND) {  // 1148-1267
            const Cell& below = T(x, y + 1);
            const int bt = below.mat->physicsType;
            if (tile.mat->interact && tile.mat->nInteractions[below.mat->id] > 0) {
                d.act = A_INTERACT;
                d.mb = below.mat->id;
                return d;
            }
            if (tile.mat->react && tile.mat->nReactions > 0) {
                bool react = false;
                for (int i = 0; i < tile.mat->nReactions; i++) {
                    const fse_interaction& in = tile.mat->reactions[i];
                    bool hit = (in.type == FSE_REACT_TEMPERATURE_BELOW && tile.temperature < in.data1) ||
                               (in.type == FSE_REACT_TEMPERATURE_ABOVE && tile.temperature > in.data1);
                    if (hit) {
                        react = true;
                        d.prod = (int)in.data2;  // the last firing reaction wins, as the loop at 1183-1202 overwrites
                    }
                }
                if (react) {
                    d.act = A_REACT;
                    return d;
                }
            }
            bool canBelow = (bt == AIR || (bt != SOLID && below.mat->density < tile.mat->density));
            if (!canBelow) return d;
            bool canL = can_sink(x - 1, y + 1, tile), canR = can_sink(x + 1, y + 1, tile);
            if ((canL || canR) && rnd(S_SAND_HESITATE, x, y) % 20 == 0) return d;
            if (bt == AIR && phys(x, y + 2) == AIR && phys(x, y + 3) == AIR && phys(x, y + 4) == AIR) d.act = A_SAND_PART;
            else {
                d.act = A_SAND_SWAP;
                d.coin = rnd(S_SAND_MOVED, x, y) % 2 == 0;
            }
            if (rnd(S_SAND_TX_SELF, x, y) % 2 == 0) {  // pokes land on whatever is diagonally below after the swaps (C2)
                d.pokeL = rnd(S_SAND_TX_L, x, y) % 2 == 0;
                d.pokeR = rnd(S_SAND_TX_R, x, y) % 2 == 0;
            }
            return d;
        }
        if (type == SOUP) {  // 1269-1537
            if (tile.fluidAmount == 0.0f) return d;
            if (tile.fluidAmount < FLUID_MinValue) {
                d.act = A_SOUP_ZERO;
                return d;
            }
            const Cell& bottom = T(x, y + 1);
            const int bph = bottom.mat->physicsType;
            if (tile.fluidAmount > 0.005 && bph == AIR && phys(x, y + 2) == AIR && phys(x, y + 3) == AIR && phys(x, y + 4) == AIR) {
                d.act = A_SOUP_PART;
                return d;
            }
            if (tile.moved) return d;
            const float start = tile.fluidAmount;
            float rem = tile.fluidAmount;
            float fd = tile.fluidAmountDiff;
            const bool airBelow = bph == AIR;
            d.botSoup = bph == SOUP;
            d.act = A_SOUP_FLOW;
            bool early = false;
            if ((airBelow && iter <= 2) || bottom.mat->id == m) {  // 1315-1334
                float dst = bph == SOUP ? bottom.fluidAmount : 0.0f;
                float flow = vflow(start, dst) - dst;
                flow = clampflow(flow, start, bottom.fluidAmount > 0);
                if (flow != 0) {
                    rem -= flow;
                    fd -= flow;
                    d.flowD = flow;
                }
            } else if (iter == 0 && bph == SOUP && bottom.mat->id != m) {  // 1335-1341
                if (rnd(S_SOUP_SWAP_DOWN, x, y) % 10 == 0) {
                    d.act = A_SOUP_SWAPDOWN;
                    return d;
                }
            }
            if (rem < FLUID_MinValue) {
                fd -= rem;
                early = true;
            }
            const Cell& left = T(x - 1, y);
            const Cell& right = T(x + 1, y);
            d.leftSoup = left.mat->physicsType == SOUP;
            d.rightSoup = right.mat->physicsType == SOUP;
            const bool canL = (left.mat->physicsType == AIR || left.mat->id == m) && !airBelow;
            const bool canR = (right.mat->physicsType == AIR || right.mat->id == m) && !airBelow;
            if (!early && canL) {  // 1355-1375
                float dst = d.leftSoup ? left.fluidAmount : 0.0f;
                float flow = clampflow((rem - dst) / (canR ? 3.0f : 2.0f), rem, true);
                if (flow != 0) {
                    rem -= flow;
                    fd -= flow;
                    d.flowL = flow;
                }
            }
            if (!early && rem < FLUID_MinValue) {
                fd -= rem;
                early = true;
            }
            if (!early && canR) {  // 1383-1403
                float dst = d.rightSoup ? right.fluidAmount : 0.0f;
                float flow = clampflow((rem - dst) / 2.0f, rem, true);
                if (flow != 0) {
                    rem -= flow;
                    fd -= flow;
                    d.flowR = flow;
                }
            }
            if (!early && rem < FLUID_MinValue) {
                fd -= rem;
                early = true;
            }
            const Cell& top = T(x, y - 1);
            const int tph = top.mat->physicsType;
            d.topSoup = tph == SOUP;
            if (!early) {
                if (tph == AIR || top.mat->id == m) {  // 1413-1432
                    float dst = tph == SOUP ? top.fluidAmount : 0.0f;
                    float flow = clampflow(rem - vflow(rem, dst), rem, true);
                    if (flow != 0) {
                        rem -= flow;
                        fd -= flow;
                        d.flowU = flow;
                    }
                } else if (iter == 0 && tph == SOUP && top.mat->id != m) {  // 1433-1439
                    if (rnd(S_SOUP_SWAP_UP, x, y) % 10 == 0) d.swapUp = true;
                }
            }
            if (!early && !d.swapUp && rem < FLUID_MinValue) {
                fd -= rem;
                early = true;
            }
            d.fd_new = fd;
            d.stl_new = tile.settleCount;
            d.moved_new = tile.moved;
            if (!early && !d.swapUp) {
                if (start == rem) {  // 1447-1451
                    d.stl_new = (uint8_t)(tile.settleCount + 1);
                    if (d.stl_new >= 10) d.moved_new = true;
                } else {
                    d.changed = true;
                }
                d.wl = (int)m == W.ids.water && (int)bottom.mat->id == W.ids.lava;  // 1519
            }
            return d;
        }
        if (type == GAS) {  // 1569-1585
            if (phys(x, y - 1) == AIR && !((phys(x - 1, y - 1) == AIR || phys(x + 1, y - 1) == AIR) && rnd(S_GAS1, x, y) % 2 == 0)) d.act = A_GAS_UP;
        }
        return d;
    }

    // own-column commit
    void commit1(int x, int y, const Dec1& d, float* outL, float* outR, uint8_t* chg, uint8_t* pkL, uint8_t* pkR, int cx) {
        const int idx = x + y * w;
        const int k = x - cx + 1;
        switch (d.act) {
            case A_NONE:
            case A_INTERACT:
                break;
            case A_FIRE:
                if (d.ember) {  // 1109-1119
                    Particle p;
                    p.tile = T(x, y);
                    p.x = (float)x;
                    p.y = (float)(y - 1);
                    p.vx = ((int)(rnd(S_FIRE_EMBER_VX, x, y) % 10) - 5) / 20.0f;
                    p.vy = -((int)(rnd(S_FIRE_EMBER_VY, x, y) % 10) / 10.0f) / 3.0f + -0.5f;
                    p.ay = 0.01f;
                    p.temporary = true;
                    p.lifetime = 30;
                    p.fadeTime = 10;
                    p.id = W.particle_id(x, y, iter, 15);
                    out.push_back(p);
                }
                break;
            case A_MARK:
                W.visited[idx] = 1;
                break;
            case A_REACT: {
                int16_t t = T(x, y).temperature;
                Cell n = W.create(d.prod, x, y);
                n.temperature = t;
                setc(x, y, n, true, true);
                break;
            }
            case A_SAND_PART:
            case A_SAND_SWAP: {
                Cell tile = T(x, y), below = T(x, y + 1);
                setc(x, y, below, true, false);
                if (d.act == A_SAND_PART) {
                    Particle p;
                    p.tile = tile;
                    p.x = (float)x;
                    p.y = (float)(y + 1);
                    p.vx = ((int)(rnd(S_SAND_PART_VX, x, y) % 10) - 5) / 20.0f;
                    p.vy = -((int)(rnd(S_SAND_PART_VY, x, y) % 2) + 3) / 10.0f + 1.5f;
                    p.ay = 0.1f;
                    p.id = W.particle_id(x, y, iter, 14);
                    out.push_back(p);
                } else {
                    if (d.coin) tile.moved = true;
                    setc(x, y + 1, tile, true, true);
                }
                pkL[k] = d.pokeL;
                pkR[k] = d.pokeR;
                break;
            }
            case A_SOUP_ZERO:
                T(x, y).fluidAmount = 0.0f;
                break;
            case A_SOUP_PART: {
                Cell tile = T(x, y);
                setc(x, y, W.nothing(), true, false);
                int n = (int)(tile.fluidAmount / 4);
                if (n < 1) n = 1;
                for (int i = 0; i < n; i++) {
                    Cell nt = fresh(tile);
                    nt.fluidAmount = tile.fluidAmount / n;
                    Particle p;
                    p.tile = nt;
                    p.x = (float)x;
                    p.y = (float)(y + 1);
                    p.vx = ((int)(rnd(S_SOUP_PART0 + 2 * (i & 7), x, y) % 10) - 5) / 30.0f;
                    p.vy = -((int)(rnd(S_SOUP_PART0 + 2 * (i & 7) + 1, x, y) % 2) + 3) / 10.0f + 1.0f;
                    p.ay = 0.1f;
                    p.id = W.particle_id(x, y, iter, i & 7);
                    out.push_back(p);
                }
                break;
            }
            case A_SOUP_SWAPDOWN: {
                Cell tile = T(x, y), bottom = T(x, y + 1);
                T(x, y) = bottom;
                T(x, y + 1) = tile;
                break;
            }
            case A_SOUP_FLOW: {
                Cell tile = T(x, y);
                tile.fluidAmountDiff = d.fd_new;
                tile.settleCount = d.stl_new;
                tile.moved = d.moved_new;
                if (d.flowD != 0) {
                    Cell& b = T(x, y + 1);
                    if (b.mat->physicsType == AIR) {
                        b = fresh(tile);
                        b.fluidAmountDiff = d.flowD;
                    } else b.fluidAmountDiff += d.flowD;
                }
                if (d.flowU != 0) {
                    Cell& t = T(x, y - 1);
                    if (t.mat->physicsType == AIR) {
                        t = fresh(tile);
                        t.fluidAmountDiff = d.flowU;
                    } else t.fluidAmountDiff += d.flowU;
                }
                if (d.swapUp) {  // 1433-1439: the cell above comes down, this one (with its updated diff) goes up
                    Cell top = T(x, y - 1);
                    T(x, y) = top;
                    T(x, y - 1) = tile;
                } else {
                    T(x, y) = tile;
                    if (d.changed) {  // 1452-1458 (vertical neighbours here; horizontal ones in C2)
                        W.dirty[idx] = 1;
                        if (d.topSoup) T(x, y - 1).moved = false;
                        if (d.botSoup) T(x, y + 1).moved = false;
                    }
                }
                if (W.flowY_) {  // world.cpp:1334, 1374, 1402, 1432: the flows as decided (a flow handed back in C2 still counts)
                    float& fy = W.flowY_[idx];
                    float& fx = W.flowX_[idx];
                    if (d.flowD != 0) fy += d.flowD;
                    if (d.flowL != 0) fx -= d.flowL;
                    if (d.flowR != 0) fx += d.flowR;
                    if (d.flowU != 0) fy -= d.flowU;
                }
                outL[k] = d.flowL;
                outR[k] = d.flowR;
                chg[k] = d.changed ? (uint8_t)((d.leftSoup ? 1 : 0) | (d.rightSoup ? 2 : 0)) : 0;
                break;
            }
            case A_GAS_UP: {
                Cell tile = T(x, y), up = T(x, y - 1);
                setc(x, y, up, true, false);
                setc(x, y - 1, tile, true, true);
                break;
            }
        }
    }

    void pass1_row(int cx, int y) {
        Dec1 dec[CHUNK];
        bool any = false, anyArea = false;
        for (int i = 0; i < CHUNK; i++) {
            dec[i] = decide1(cx + i, y);
            any |= dec[i].act != A_NONE;
            anyArea |= dec[i].act == A_FIRE || dec[i].act == A_INTERACT || dec[i].wl;
        }
        if (!any) return;
        float outL[CHUNK + 2] = {0}, outR[CHUNK + 2] = {0}, refL[CHUNK + 2] = {0}, refR[CHUNK + 2] = {0};
        uint8_t chg[CHUNK + 2] = {0}, pkL[CHUNK + 2] = {0}, pkR[CHUNK + 2] = {0};
        for (int i = 0; i < CHUNK; i++) commit1(cx + i, y, dec[i], outL, outR, chg, pkL, pkR, cx);
        // C2: targets gather (columns cx-1 .. cx+128; index k = x - cx + 1)
        for (int k = 0; k < CHUNK + 2; k++) {
            const int x = cx - 1 + k;
            const float inL = k > 0 ? outR[k - 1] : 0.0f, inR = k < CHUNK + 1 ? outL[k + 1] : 0.0f;
            if (inL != 0 || inR != 0) {
                Cell& c = T(x, y);
                if (c.mat->physicsType == AIR) {
                    if (inL != 0) {
                        c = fresh(T(x - 1, y));
                        c.fluidAmountDiff = inL;
                        if (inR != 0) {
                            if (T(x + 1, y).mat == c.mat) c.fluidAmountDiff += inR;
                            else refL[k + 1] = inR;
                        }
                    } else {
                        c = fresh(T(x + 1, y));
                        c.fluidAmountDiff = inR;
                    }
                } else {
                    if (inL != 0) {
                        if (c.mat->physicsType == SOUP && T(x - 1, y).mat == c.mat) c.fluidAmountDiff += inL;
                        else refR[k - 1] = inL;
                    }
                    if (inR != 0) {
                        if (c.mat->physicsType == SOUP && T(x + 1, y).mat == c.mat) c.fluidAmountDiff += inR;
                        else refL[k + 1] = inR;
                    }
                }
            }
            // un-settle from a changed liquid neighbour (1456-1457): bit1 of chg[left nbr] = "my right is liquid"
            if (((k > 0 && (chg[k - 1] & 2)) || (k < CHUNK + 1 && (chg[k + 1] & 1))) && T(x, y).mat->physicsType == SOUP) T(x, y).moved = false;
            // moved-pokes on the row below (1245-1265)
            if (((k > 0 && pkR[k - 1]) || (k < CHUNK + 1 && pkL[k + 1])) && T(x, y + 1).mat->physicsType == SAND) T(x, y + 1).moved = true;
        }
        // C3: refunds, then area effects with lowest-x claims
        for (int k = 1; k <= CHUNK; k++) {
            if (refL[k] != 0) T(cx - 1 + k, y).fluidAmountDiff += refL[k];
            if (refR[k] != 0) T(cx - 1 + k, y).fluidAmountDiff += refR[k];
        }
        if (!anyArea) return;
        std::map<std::pair<int, int>, int> claim;  // (x, y) -> claiming source x
        auto claims = [&](int sx, int tx, int ty) {
            auto it = claim.find({tx, ty});
            if (it == claim.end()) claim[{tx, ty}] = sx;
        };
        auto mine = [&](int sx, int tx, int ty) { return claim[{tx, ty}] == sx; };
        for (int pass = 0; pass < 2; pass++)
            for (int i = 0; i < CHUNK; i++) {  // ascending x: the first claim is the lowest x
                const Dec1& d = dec[i];
                const int x = cx + i;
                if (d.act == A_FIRE) {
                    if (pass == 0) {
                        if (d.die) claims(x, x, y);
                        for (int kk = 0; kk < 25; kk++)
                            if (d.ignite >> kk & 1) claims(x, x + kk / 5 - 2, y + kk % 5 - 2);
                    } else {
                        for (int kk = 0; kk < 25; kk++)
                            if ((d.ignite >> kk & 1) && mine(x, x + kk / 5 - 2, y + kk % 5 - 2))
                                setc(x + kk / 5 - 2, y + kk % 5 - 2, W.create(W.ids.fire, x + kk / 5 - 2, y + kk % 5 - 2), true, true);
                        if (d.die && mine(x, x, y)) setc(x, y, W.nothing(), true, true);
                    }
                } else if (d.wl) {  // 1519-1537
                    if (pass == 0) {
                        for (int xx = -1; xx <= 1; xx++)
                            for (int yy = 0; yy <= 2; yy++) claims(x, x + xx, y + yy);
                    } else {
                        if (mine(x, x, y)) setc(x, y, W.create(W.ids.steam, x, y), true, false);
                        if (mine(x, x, y + 1)) setc(x, y + 1, W.create(W.ids.obsidian, x, y + 1), true, true);
                        for (int xx = -1; xx <= 1; xx++)
                            for (int yy = 0; yy <= 2; yy++)
                                if (mine(x, x + xx, y + yy) && (int)T(x + xx, y + yy).mat->id == W.ids.lava)
                                    setc(x + xx, y + yy, W.create(W.ids.obsidian, x + xx, y + yy), true, true);
                    }
                } else if (d.act == A_INTERACT) {  // 1153-1179
                    const Material* mm = T(x, y).mat;
                    if (pass == 1 && !mine(x, x, y)) {
                        // the source itself was transformed by a lower-x effect this step: its interaction list is void
                        // (it still owns whatever else it claimed; those cells are simply left alone)
                        continue;
                    }
                    const auto& list = mm->interactions[d.mb];
                    if (pass == 0) claims(x, x, y);
                    for (size_t q = 0; q < list.size(); q++) {
                        const fse_interaction& in = list[q];
                        const int rad = (int)in.data2;
                        if (in.type != FSE_INTERACT_TRANSFORM_MATERIAL && in.type != FSE_INTERACT_SPAWN_MATERIAL) continue;
                        for (int xx = in.ofs_x - rad; xx <= in.ofs_x + rad; xx++)
                            for (int yy = in.ofs_y - rad; yy <= in.ofs_y + rad; yy++) {
                                if (pass == 0) {
                                    claims(x, x + xx, y + yy);
                                } else if (mine(x, x + xx, y + yy)) {
                                    const uint32_t tm = T(x + xx, y + yy).mat->id;
                                    bool hit = in.type == FSE_INTERACT_TRANSFORM_MATERIAL ? tm == d.mb : ((xx == 0 && yy == 0) || (int)tm == W.ids.air);
                                    if (hit) setc(x + xx, y + yy, W.create((uint32_t)in.data1, x + xx, y + yy), true, true);
                                }
                            }
                    }
                }
            }
    }

    // ---------------------------------------------------------------- pass 2 (world.cpp:1594-1820)
    Dec2 decide2(int x, int y) {
        Dec2 d;
        const int idx = x + y * w;
        if (W.visited[idx]) return d;
        const Cell& tile = T(x, y);
        const int type = tile.mat->physicsType;
        if (type == SAND) {
            const bool canL = can_sink(x - 1, y + 1, tile), canR = can_sink(x + 1, y + 1, tile);
            if (!(canL || canR)) {
                d.act = 1;
                return d;
            }
            bool stopped = !tile.moved;
            const int slip = tile.mat->slipperyness;
            if (stopped) {
                int drop = 0;
                for (int pil = 0; pil < 10; pil++)
                    if (phys(x - 1, y + 1 + pil) == AIR || phys(x + 1, y + 1 + pil) == AIR) drop++;
                int maxStab = (int)(8 / sqrt((double)slip) + 1);
                if (drop + 1 - maxStab > 0) {
                    int chance = 1000 / (drop + 1 - maxStab);
                    if (chance < 1000 && rnd(S_SAND2_UNSTICK, x, y) % chance == 0) stopped = false;
                }
            }
            if (stopped) {
                d.act = 1;
                return d;
            }
            const bool should = rnd(S_SAND2_SHOULD, x, y) % (2 * slip) != 0;
            if (should && rnd(S_SAND2_TX_SELF, x, y) % 2 == 0 && rnd(S_SAND2_TX_OTHER, x, y) % 2 == 0) d.poke = true;
            if (should && canL && (!canR || rnd(S_SAND2_LR, x, y) % 2 == 0)) d.dir = -1;
            else if (should && canR) d.dir = 1;
            if (d.dir) {
                d.act = 2;
                d.riser = phys(x + d.dir, y) == AIR;
                d.restick = rnd(S_SAND2_RESTICK, x, y) % (20 * slip) == 0;
            } else {
                d.act = 1;
            }
        } else if (type == SOUP) {
            d.act = 3;
        } else if (type == GAS) {
            const int aL = phys(x - 1, y - 1), aR = phys(x + 1, y - 1);
            if (aL == AIR && !(aR == AIR && rnd(S_GAS2, x, y) % 2 == 0)) d.dir = -1;
            else if (aR == AIR) d.dir = 1;
            if (d.dir) d.act = 4;
        }
        return d;
    }

    void pass2_row(int cx, int y) {
        Dec2 dec[CHUNK];
        int claimDn[CHUNK + 2], claimUp[CHUNK + 2];  // destination claims in row y+1 (slides) / row y-1 (gas): lowest source x
        for (int k = 0; k < CHUNK + 2; k++) claimDn[k] = claimUp[k] = 1 << 30;
        bool any = false;
        for (int i = 0; i < CHUNK; i++) {
            dec[i] = decide2(cx + i, y);
            any |= dec[i].act != 0;
            if (dec[i].act == 2) claimDn[i + 1 + dec[i].dir] = std::min(claimDn[i + 1 + dec[i].dir], i);
            if (dec[i].act == 4) claimUp[i + 1 + dec[i].dir] = std::min(claimUp[i + 1 + dec[i].dir], i);
        }
        if (!any) return;
        uint8_t poke[CHUNK] = {0};
        for (int i = 0; i < CHUNK; i++) {
            const Dec2& d = dec[i];
            const int x = cx + i, idx = x + y * w;
            if (d.act == 1) {
                T(x, y).moved = false;  // 1647-1654 / 1721-1727
            } else if (d.act == 2) {
                poke[i] = d.poke;
                if (claimDn[i + 1 + d.dir] != i) continue;  // lost the destination: stays put
                Cell tile = T(x, y);
                Cell diag = T(x + d.dir, y + 1);
                if (d.riser) {
                    setc(x + d.dir, y, diag, true, d.dir < 0);  // the left slide marks the riser visited, the right one does not (1679 vs 1700-1704)
                    setc(x, y, W.nothing(), true, false);
                } else {
                    setc(x, y, diag, true, true);
                }
                if (d.restick) tile.moved = false;
                setc(x + d.dir, y + 1, tile, true, true);
            } else if (d.act == 3) {  // 1728-1745
                Cell tile = T(x, y);
                tile.fluidAmount += tile.fluidAmountDiff;
                tile.fluidAmountDiff = 0.0f;
                if (tile.fluidAmount < FLUID_MinValue) T(x, y) = W.nothing();
                else T(x, y) = tile;
                W.dirty[idx] = 1;
                W.visited[idx] = 1;
            } else if (d.act == 4) {
                if (claimUp[i + 1 + d.dir] != i) continue;
                Cell tile = T(x, y), other = T(x + d.dir, y - 1);
                setc(x, y, other, true, false);
                setc(x + d.dir, y - 1, tile, true, true);
            }
        }
        for (int i = 0; i < CHUNK; i++)  // 1658-1673: "moved" handed to the sand below, after the slides
            if (poke[i] && T(cx + i, y + 1).mat->physicsType == SAND) T(cx + i, y + 1).moved = true;
    }

    // ---------------------------------------------------------------- pass 3 (world.cpp:1828-1891)
    void pass3_row(int cx, int y) {
        int dir[CHUNK];
        bool steam[CHUNK];
        int claim[CHUNK + 2];
        for (int k = 0; k < CHUNK + 2; k++) claim[k] = 1 << 30;
        for (int i = 0; i < CHUNK; i++) {
            const int x = cx + i;
            dir[i] = 0;
            steam[i] = false;
            if (W.visited[x + y * w] || phys(x, y) != GAS) continue;
            const int l = phys(x - 1, y), r = phys(x + 1, y);
            if (l == AIR && !(r == AIR && rnd(S_GAS3, x, y) % 2 == 0)) dir[i] = -1;
            else if (r == AIR) dir[i] = 1;
            else if ((int)T(x, y).mat->id == W.ids.steam && rnd(S_STEAM, x, y) % 10 == 0) steam[i] = true;
            if (dir[i]) claim[i + 1 + dir[i]] = std::min(claim[i + 1 + dir[i]], i);
        }
        for (int i = 0; i < CHUNK; i++) {
            const int x = cx + i;
            if (dir[i] && claim[i + 1 + dir[i]] == i) {
                Cell tile = T(x, y), other = T(x + dir[i], y);
                setc(x, y, other, true, false);
                setc(x + dir[i], y, tile, true, true);
            } else if (steam[i]) {
                setc(x, y, W.create(W.ids.water, x, y), true, false);
            }
        }
    }
};

void World::chunk_rows(int cx, int cy, int iter, std::vector<Particle>& out) {
    RowsImpl R(*this, iter, out);
    for (int dy = CHUNK - 1; dy >= 0; dy--) R.pass1_row(cx, cy + dy);
    for (int dy = CHUNK - 1; dy >= 0; dy--) R.pass2_row(cx, cy + dy);
    for (int dy = CHUNK - 1; dy >= 0; dy--) R.pass3_row(cx, cy + dy);
}

}  // namespace fseo
