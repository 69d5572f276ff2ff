// fse_entities.cu — entities <-> grid, hand-written for sm_100a:
//   fse_entities_tick   world::tickEntities          (reference: source/engine/world.cpp:3010-3247)
//   fse_entities_stamp  WorldEntitySystem::process   (source/game/player.cpp:173-199)
//   fse_object_delete   the objectDelete loop        (source/engine/game.cpp:2128-2139)
//
// The reference handles its entities one after the other on the main thread, and inside one entity the sweeps are a strictly
// sequential walk (a kicked grain is gone for the next probe, a step-up shifts the rows the rest of the box probes).  So one CTA
// takes the entities in array order.  Per entity the whole CTA reduces the overlap push-out (world.cpp:3013-3034) and stages the
// physics types of the box's surroundings in shared memory; one thread then runs the sweeps on that tile exactly as written —
// 2 x 8 sub-steps x hw x hh probes that would each be a dependent trip to HBM otherwise — and touches global memory only where a
// grain is kicked into the particle pool.  Stamping and the object delete are independent per cell and run across the CTA / grid.
// rand() -> counter RNG keyed on the kicked / stamped cell (same slots as oracle/entity_oracle.cpp).
#include <cmath>
#include <vector>

#include "fse_internal.hpp"

namespace fse {

enum : uint32_t { S_ENT_VX = 76, S_ENT_VY = 77, S_STAMP_X = 78, S_STAMP_VX = 79, S_STAMP_VY = 80 };
constexpr int ENT_TILE_CAP = 96 * 1024;  // bytes of shared memory for the staged physics types
constexpr int ENT_THREADS = 256;

struct EntArgs {
    Planes p;
    const DevTables* T;
    int W, H;
    fse_entity* ents;
    int n;
    float lzx, lzy;
    uint32_t rkey, tick;
    fse_particle* pbuf;
    unsigned int* pcount;
    unsigned int pcap;
    int air, object_mat;
    long long* objdel;         // indices stamped since the last fse_object_delete
    unsigned int* objdel_cnt;
    unsigned int objdel_cap;
    // strip worlds: the kernels work in GLOBAL rows — W x H is the whole world, the plane pointers are moved back by the rows above this
    // rank's window, so (sy * W + sx) addresses the window where it holds the row.  [ylo, yhi): rows of the window, [own_lo, own_hi): rows
    // this rank owns (their kicked grains are its particles), exec[i] (null on one world): the rank that runs entity i
    int ylo, yhi, own_lo, own_hi;
    const int* exec;
    int rank;
};

__device__ __forceinline__ uint64_t entity_particle_id(uint32_t tick, int kind, int x, int y) {
    return (2ULL << 62) | (1ULL << 61) | ((uint64_t)(kind & 1) << 60) | ((uint64_t)(tick & 0xfffff) << 40) | ((uint64_t)(y & 0xfffff) << 20) | (uint64_t)(x & 0xfffff);
}
__device__ __forceinline__ bool blocks_type(int t) { return t == P_SOLID || t == P_SAND || t == P_PASSABLE; }  // OBJECT == PASSABLE == 5

__device__ void ent_emit(const EntArgs& a, size_t g, float px, float py, float vx, float vy, uint64_t id) {
    const unsigned int i = atomicAdd(a.pcount, 1u);
    if (i >= a.pcap) return;  // counted; fse_particles_dropped reports it (the host made room for every cell under the boxes)
    fse_particle q;
    memset(&q, 0, sizeof q);
    const uint8_t f = a.p.flg[g];
    q.tile.mat = a.p.mat[g];
    q.tile.moved = (f & F_MOVED) ? 1 : 0;
    q.tile.settle = a.p.stl[g];
    q.tile.color = a.p.col[g];
    q.tile.temp = a.p.tmp[g];
    q.tile.fluid = a.p.fl[g];
    q.tile.fluid_diff = a.p.fd[g];
    q.x = px; q.y = py; q.vx = vx; q.vy = vy; q.ay = 0.1f;
    q.fade_time = 60;
    q.id = id;
    a.pbuf[i] = q;
}
__device__ __forceinline__ void ent_write(const EntArgs& a, size_t g, int mat, uint8_t flg, uint32_t col) {
    a.p.mat[g] = (uint8_t)mat;
    a.p.flg[g] = flg;
    a.p.stl[g] = 0;
    a.p.tmp[g] = 0;
    a.p.col[g] = col;
    a.p.fl[g] = 2.0f;
    a.p.fd[g] = 0.0f;
}

struct EntShared {
    fse_entity e;
    int nInter, avX, avY;
    int tx0, ty0, tw, th;  // staged tile: world cells [tx0, tx0 + tw) x [ty0, ty0 + th); tw == 0: no tile (box too large), probe HBM
};

__global__ void __launch_bounds__(ENT_THREADS) entities_tick_kernel(EntArgs a) {
    extern __shared__ unsigned char ent_tile[];
    __shared__ EntShared S;
    const int tid = threadIdx.x;
    // physics type of world cell (sx, sy), from the tile where it covers the cell
    auto type_at = [&](int sx, int sy) -> int {
        const int lx = sx - S.tx0, ly = sy - S.ty0;
        if (S.tw > 0 && lx >= 0 && ly >= 0 && lx < S.tw && ly < S.th) return ent_tile[ly * S.tw + lx];
        return a.T->phys[a.p.mat[(size_t)sy * a.W + sx]];
    };
    // addCell(new CellData(tp, sx, sy, vx, vy, 0, 0.1f)); real_tiles[...] = Tiles_NOTHING; dirty[...] = true (world.cpp:3070-3072)
    auto kick = [&](int sx, int sy, float bx, float by) {
        const size_t g = (size_t)sy * a.W + sx;
        const uint32_t cb = rng_cell(a.rkey, sx, sy);
        const float vx = ((int)(rng_draw(cb, S_ENT_VX) % 10) - 5) / 10.0f + bx;
        const float vy = ((int)(rng_draw(cb, S_ENT_VY) % 10) - 5) / 10.0f + by;
        ent_emit(a, g, (float)sx, (float)sy, vx, vy, entity_particle_id(a.tick, 0, sx, sy));
        ent_write(a, g, a.air, F_DIRTY, 0u);
        const int lx = sx - S.tx0, ly = sy - S.ty0;
        if (S.tw > 0 && lx >= 0 && ly >= 0 && lx < S.tw && ly < S.th) ent_tile[ly * S.tw + lx] = P_AIR;
    };
    for (int ei = 0; ei < a.n; ei++) {
        if (a.exec && a.exec[ei] != a.rank) {  // strips: another rank runs this entity; its record here becomes zeros for the sum over the ranks
            uint32_t* z = reinterpret_cast<uint32_t*>(&a.ents[ei]);
            for (int q = tid; q < (int)(sizeof(fse_entity) / 4); q += ENT_THREADS) z[q] = 0u;
            continue;
        }
        if (tid == 0) {
            S.e = a.ents[ei];
            S.e.destroy = 0;
            S.nInter = S.avX = S.avY = 0;
            S.tw = 0;
        }
        __syncthreads();
        const int hw = S.e.hw, hh = S.e.hh;
        {   // overlap push-out (3013-3030): a reduction over the box
            int ni = 0, ax = 0, ay = 0;
            for (int c = tid; c < hw * hh; c += ENT_THREADS) {
                const int xx = c / hh, yy = c % hh;
                const int sx = (int)((S.e.x + xx) + a.lzx), sy = (int)((S.e.y + yy) + a.lzy);
                if (sx < 0 || sy < 0 || sx >= a.W || sy >= a.H) continue;
                if (blocks_type(a.T->phys[a.p.mat[(size_t)sy * a.W + sx]])) {
                    ni++;
                    ax += xx - hw / 2;
                    ay += yy - hh / 2;
                }
            }
            if (ni) {
                atomicAdd(&S.nInter, ni);
                atomicAdd(&S.avX, ax);
                atomicAdd(&S.avY, ay);
            }
        }
        __syncthreads();
        if (tid == 0) {
            if (S.nInter > 0) {  // 3031-3034
                S.e.x += S.avX > 0 ? -1 : (S.avX < 0 ? 1 : 0);
                S.e.y += S.avY > 0 ? -1 : (S.avY < 0 ? 1 : 0);
            }
            S.e.vy = (float)((double)S.e.vy + 0.25);  // 3036
            // tile: the box plus everything the sweeps can reach (|vx| sideways, |vy| + one step-up per sub-step vertically)
            const int m = (int)ceilf(fabsf(S.e.vx)) + (int)ceilf(fabsf(S.e.vy)) + 12;
            const long long tw = (long long)hw + 2 * m, th = (long long)hh + 2 * m;
            if (m < 4096 && tw * th <= ENT_TILE_CAP) {
                S.tx0 = (int)(S.e.x + a.lzx) - m;
                S.ty0 = (int)(S.e.y + a.lzy) - m;
                S.tw = (int)tw;
                S.th = (int)th;
            }
        }
        __syncthreads();
        for (int c = tid; c < S.tw * S.th; c += ENT_THREADS) {
            const int sx = S.tx0 + c % S.tw, sy = S.ty0 + c / S.tw;
            ent_tile[c] = (sx >= 0 && sy >= 0 && sx < a.W && sy < a.H) ? a.T->phys[a.p.mat[(size_t)sy * a.W + sx]] : (uint8_t)P_AIR;
        }
        __syncthreads();
        if (tid == 0) {
            fse_entity& cur = S.e;
            const int width = a.W, height = a.H;
            const float lzx = a.lzx, lzy = a.lzy;
            const int dir = cur.vx > 0.001 ? 0 : (cur.vx < -0.001 ? 1 : -1);
            if (dir >= 0) {  // 3038-3151
                const float stx = cur.x;
                for (float dx = 0; dir == 0 ? dx < cur.vx : dx > cur.vx; dx = (float)((double)dx + (double)cur.vx / 8.0)) {
                    const float nx = stx + dx;
                    float ny = cur.y;
                    bool collide = false;
                    for (int xx = 0; xx < hw; xx++)
                        for (int yy = 0; yy < hh; yy++) {
                            const int sx = (int)((nx + xx) + lzx), sy = (int)((ny + yy) + lzy);
                            if (!(sx >= 0 && sy >= 0 && sx < width && sy < height)) continue;
                            const int t = type_at(sx, sy);
                            if (!blocks_type(t)) continue;
                            if (yy == hh - 1) {  // 3052-3066
                                for (int xx1 = 0; xx1 < hw; xx1++)
                                    for (int yy1 = 0; yy1 < hh; yy1++) {
                                        const int sx1 = (int)((nx + xx1) + lzx), sy1 = (int)((ny + yy1) + lzy - 1);
                                        if (sx1 >= 0 && sy1 >= 0 && sx1 < width && sy1 < height && blocks_type(type_at(sx1, sy1))) collide = true;
                                    }
                                if (!collide) ny--;
                            } else if (t == P_SAND) {  // 3068-3074
                                kick(sx, sy, dir == 0 ? 0.5f : -0.5f, 0.0f);
                                cur.vx = (float)((double)cur.vx * 0.99);
                            } else {
                                collide = true;
                            }
                        }
                    if (!collide) {
                        cur.x = nx;
                        cur.y = ny;
                    } else {
                        cur.vx /= 2;
                        break;
                    }
                }
            }
            cur.ground = 0;  // 3153
            if (cur.vy > 0.001) {  // 3155-3183
                const float sty = cur.y;
                for (float dy = 0; dy < cur.vy; dy = (float)((double)dy + (double)cur.vy / 8.0)) {
                    const float ny = sty + dy, nx = cur.x;
                    bool collide = false;
                    for (int xx = 0; xx < hw && !collide; xx++)
                        for (int yy = 0; yy < hh; yy++) {
                            const int sx = (int)((nx + xx) + lzx), sy = (int)((ny + yy) + lzy);
                            if (sx >= 0 && sy >= 0 && sx < width && sy < height && blocks_type(type_at(sx, sy))) {
                                collide = true;  // nothing is kicked on the way down: the first hit settles it
                                break;
                            }
                        }
                    if (!collide) {
                        cur.y = ny;
                    } else {
                        cur.vy /= 2;
                        cur.ground = 1;
                        break;
                    }
                }
            } else if (cur.vy < -0.001) {  // 3184-3220
                const float sty = cur.y;
                for (float dy = 0; dy > cur.vy; dy = (float)((double)dy + (double)cur.vy / 8.0)) {
                    const float ny = sty + dy, nx = cur.x;
                    bool collide = false;
                    for (int xx = 0; xx < hw; xx++)
                        for (int yy = 0; yy < hh; yy++) {
                            const int sx = (int)((nx + xx) + lzx), sy = (int)((ny + yy) + lzy);
                            if (!(sx >= 0 && sy >= 0 && sx < width && sy < height)) continue;
                            const int t = type_at(sx, sy);
                            if (!blocks_type(t)) continue;
                            if (t == P_SAND) {
                                kick(sx, sy, 0.0f, -0.5f);
                                cur.vy = (float)((double)cur.vy * 0.99);
                            } else {
                                collide = true;
                            }
                        }
                    if (!collide) {
                        cur.y = ny;
                    } else {
                        cur.vy /= 2;
                        cur.ground = 1;
                        break;
                    }
                }
            }
            if (fabsf(cur.vx) >= 1024.0f || fabsf(cur.vy) >= 1024.0f) {  // 3222-3225
                cur.destroy = 1;
            } else {
                cur.vx = (float)((double)cur.vx * 0.99);  // 3227-3229
                cur.vy = (float)((double)cur.vy * 0.99);
            }
            a.ents[ei] = cur;
            __threadfence();
        }
        __syncthreads();  // the next entity sees this one's kicks
    }
}

// "entity fluid displacement & make solid" (game/player.cpp:176-196): cells are independent inside one entity, entities in order
__global__ void __launch_bounds__(ENT_THREADS) entities_stamp_kernel(EntArgs a) {
    const int tid = threadIdx.x;
    for (int ei = 0; ei < a.n; ei++) {
        const fse_entity pl = a.ents[ei];
        for (int c = tid; c < pl.hw * pl.hh; c += ENT_THREADS) {
            const int tx = c / pl.hh, ty = c % pl.hh;
            const int wx = (int)(tx + pl.x + a.lzx), wy = (int)(ty + pl.y + a.lzy);
            if (wx < 0 || wy < 0 || wx >= a.W || wy >= a.H) continue;
            if (wy < a.ylo || wy >= a.yhi) continue;  // strips: a rank stamps the cells of its window (cells are independent), ghost rows included
            const size_t g = (size_t)wy * a.W + wx;
            const int t = a.T->phys[a.p.mat[g]];
            if (t != P_AIR && t != P_SAND && t != P_SOUP) continue;
            uint8_t flg = a.p.flg[g] & F_DIRTY;
            if (t != P_AIR && !(wy >= a.own_lo && wy < a.own_hi)) flg = F_DIRTY;  // a ghost cell: its owner makes the particle
            else if (t != P_AIR) {
                const uint32_t cb = rng_cell(a.rkey, wx, wy);
                const float px = (float)(wx + (int)(rng_draw(cb, S_STAMP_X) % 3) - 1 - pl.vx);
                const float py = (float)(wy - fabsf(pl.vy));
                const float vx = (float)(-pl.vx / 4 + ((int)(rng_draw(cb, S_STAMP_VX) % 10) - 5) / 5.0f);
                const float vy = (float)(-pl.vy / 4 + -((int)(rng_draw(cb, S_STAMP_VY) % 5) + 5) / 5.0f);
                ent_emit(a, g, px, py, vx, vy, entity_particle_id(a.tick, 1, wx, wy));
                flg = F_DIRTY;
            }
            ent_write(a, g, a.object_mat, flg, 0x00ff00u);  // Tiles_OBJECT (gds.cpp:318)
            const unsigned int k = atomicAdd(a.objdel_cnt, 1u);
            if (k < a.objdel_cap) a.objdel[k] = (long long)g;
        }
        __syncthreads();
    }
}

__global__ void object_delete_kernel(Planes p, const long long* list, unsigned int n, int air) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t g = (size_t)list[i];
    p.mat[g] = (uint8_t)air;
    p.flg[g] = p.flg[g] & F_DIRTY;  // dirty[] is not touched by the loop (game.cpp:2136-2138)
    p.stl[g] = 0;
    p.tmp[g] = 0;
    p.col[g] = 0;
    p.fl[g] = 2.0f;
    p.fd[g] = 0.0f;
}

struct EntityBufs {
    fse_entity* d_ents = nullptr;
    int cap = 0;
    long long* d_objdel = nullptr;
    unsigned int* d_cnt = nullptr;
    int* d_exec = nullptr;                    // strips: runner of every entity of the current fse_entities_tick
    int exec_cap = 0;
    unsigned int objdel_cap = 0;
    unsigned int stamped = 0;                 // upper bound of the entries in d_objdel
    std::vector<fse_rect> rects;              // boxes stamped since the last delete (active-chunk wake-up)
};

void entities_free(fse_world* w) {
    EntityBufs* b = (EntityBufs*)w->entity_bufs;
    if (!b) return;
    cudaFree(b->d_ents);
    cudaFree(b->d_objdel);
    cudaFree(b->d_cnt);
    cudaFree(b->d_exec);
    delete b;
    w->entity_bufs = nullptr;
}

}  // namespace fse

using namespace fse;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return fail(FSE_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

static int entity_setup(fse_world* w, const fse_entity* ents, int32_t n, const char* who, EntityBufs** out) {
    if (!w || (!ents && n > 0) || n < 0) return fail(FSE_EINVAL, "%s: bad argument", who);
    for (int i = 0; i < n; i++)
        if (ents[i].hw < 0 || ents[i].hh < 0 || ents[i].hw > 4096 || ents[i].hh > 4096 || !std::isfinite(ents[i].x) || !std::isfinite(ents[i].y) ||
            !std::isfinite(ents[i].vx) || !std::isfinite(ents[i].vy))
            return fail(FSE_EINVAL, "%s: entity %d has a bad box or a non-finite position / velocity", who, i);
    CK(cudaSetDevice(w->ctx->device));
    EntityBufs* b = (EntityBufs*)w->entity_bufs;
    if (!b) {
        b = new EntityBufs();
        w->entity_bufs = b;
        CK(cudaMalloc((void**)&b->d_cnt, sizeof(unsigned int)));
        CK(cudaMemsetAsync(b->d_cnt, 0, sizeof(unsigned int), w->stream));
    }
    if (n > b->cap) {
        CK(cudaStreamSynchronize(w->stream));
        cudaFree(b->d_ents);
        b->d_ents = nullptr;
        b->cap = 0;
        CK(cudaMalloc((void**)&b->d_ents, sizeof(fse_entity) * (size_t)(n + 16)));
        b->cap = n + 16;
    }
    if (n) CK(cudaMemcpyAsync(b->d_ents, ents, sizeof(fse_entity) * (size_t)n, cudaMemcpyHostToDevice, w->stream));
    *out = b;
    return FSE_OK;
}

static void entity_args(fse_world* w, EntityBufs* b, int32_t n, float lx, float ly, uint32_t tick, uint32_t seed, EntArgs* a) {
    memset(a, 0, sizeof *a);
    a->p = w->p;
    a->T = w->ctx->d_tabs;
    a->W = w->W;
    a->H = w->H;
    a->ylo = 0; a->yhi = w->H; a->own_lo = 0; a->own_hi = w->H;
    if (w->strip) {  // global rows (see EntArgs)
        const size_t back = (size_t)w->y_off * w->W;
        a->p.mat -= back; a->p.flg -= back; a->p.stl -= back; a->p.tmp -= back; a->p.col -= back; a->p.fl -= back; a->p.fd -= back;
        a->H = w->Hglobal;
        a->ylo = w->y_off; a->yhi = w->y_off + w->H; a->own_lo = w->own_lo; a->own_hi = w->own_hi;
    }
    a->rank = w->ctx->rank;
    a->ents = b->d_ents;
    a->n = n;
    a->lzx = lx;
    a->lzy = ly;
    a->rkey = rng_key(seed, tick, 8u);  // the tick's own iterations use 0..3, bridge and explosion 7
    a->tick = tick;
    a->pbuf = w->pbuf;
    a->pcount = w->pcount;
    a->pcap = w->pcap;
    a->air = w->ctx->h_tabs.air;
    a->objdel = b->d_objdel;
    a->objdel_cnt = b->d_cnt;
    a->objdel_cap = b->objdel_cap;
}

extern "C" FSE_API int fse_entities_tick(fse_world* w, fse_entity* ents, int32_t n, float load_x, float load_y, uint32_t tick, uint32_t seed) {
    EntityBufs* b = nullptr;
    if (int r = entity_setup(w, ents, n, "fse_entities_tick", &b)) return r;
    if (n == 0) return FSE_OK;
    long long cells = 0;
    for (int i = 0; i < n; i++) cells += 16LL * ents[i].hw * ents[i].hh;  // 2 x 8 sub-steps, each can kick every grain under the box
    if (int r = particles_headroom(w, cells + 1024, false)) return r;
    static bool configured = false;
    if (!configured) {
        CK(cudaFuncSetAttribute(entities_tick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ENT_TILE_CAP));
        configured = true;
    }
    EntArgs a;
    entity_args(w, b, n, load_x, load_y, tick, seed, &a);
    const bool multi = w->strip && w->ctx->nranks > 1;
    std::vector<int4> box(n);
    for (int i = 0; i < n; i++) {  // whatever a sweep can reach (the kernel's tile, + the push-out step and the gravity it adds first)
        const int m = (int)std::ceil(std::fabs(ents[i].vx)) + (int)std::ceil(std::fabs(ents[i].vy)) + 12 + 2;
        const int x0 = (int)(ents[i].x + load_x) - m, y0 = (int)(ents[i].y + load_y) - m;
        box[i] = make_int4(x0, y0, x0 + ents[i].hw + 2 * m, y0 + ents[i].hh + 2 * m);
    }
    std::vector<int> exec;
    if (multi) {
        // multi-rank strips: every rank makes the call with the same entities; an entity — with every entity whose reach overlaps its own, they
        // see each other's kicks — is run by the rank that holds its reach box (ghost rows refreshed first); the records are summed over the
        // ranks afterwards (a rank zeroes the ones it did not run) and the boxes travel to the neighbours they reach into
        if (int r = strip_group_runners(w, box, "fse_entities_tick", "entity", exec)) return r;
        if (int r = strip_refresh(w, w->stream, STRIP_GHOST)) return r;
        if (n > b->exec_cap) {
            CK(cudaStreamSynchronize(w->stream));
            cudaFree(b->d_exec);
            b->d_exec = nullptr;
            b->exec_cap = 0;
            CK(cudaMalloc((void**)&b->d_exec, sizeof(int) * (size_t)(n + 16)));
            b->exec_cap = n + 16;
        }
        CK(cudaMemcpyAsync(b->d_exec, exec.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, w->stream));
        a.exec = b->d_exec;
    } else {
        for (int i = 0; i < n; i++) {  // active-chunk tracking: whatever a sweep can reach is awake afterwards
            int x0 = box[i].x, y0 = box[i].y - w->y_off, x1 = box[i].z, y1 = box[i].w - w->y_off;
            x0 = x0 < 0 ? 0 : x0;
            y0 = y0 < 0 ? 0 : y0;
            x1 = x1 > w->W ? w->W : x1;
            y1 = y1 > w->H ? w->H : y1;
            if (x1 > x0 && y1 > y0)
                if (int r = fse_wake_rect(w, x0, y0, x1 - x0, y1 - y0)) return r;
        }
    }
    entities_tick_kernel<<<1, ENT_THREADS, ENT_TILE_CAP, w->stream>>>(a);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    if (multi) {
        if (int r = strip_allreduce_u32(w, (unsigned int*)b->d_ents, (size_t)n * (sizeof(fse_entity) / 4), w->stream)) return r;
        std::vector<int4> rect[4];
        for (int i = 0; i < n; i++) strip_rects_of_box(w, exec[i], box[i].x, box[i].y, box[i].z, box[i].w, rect);
        if (int r = strip_push_rects(w, rect, w->stream)) return r;
    }
    CK(cudaMemcpyAsync(ents, b->d_ents, sizeof(fse_entity) * (size_t)n, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    return FSE_OK;
}

extern "C" FSE_API int fse_entities_stamp(fse_world* w, const fse_entity* ents, int32_t n, float load_x, float load_y, int32_t object_mat, uint32_t tick,
                                          uint32_t seed) {
    EntityBufs* b = nullptr;
    if (int r = entity_setup(w, ents, n, "fse_entities_stamp", &b)) return r;
    if (object_mat < 0 || object_mat >= w->ctx->h_tabs.n) return fail(FSE_EINVAL, "fse_entities_stamp: object material %d out of range", object_mat);
    if (n == 0) return FSE_OK;
    long long cells = 0;
    for (int i = 0; i < n; i++) cells += (long long)ents[i].hw * ents[i].hh;
    if (int r = particles_headroom(w, cells, false)) return r;
    if ((long long)b->stamped + cells > (long long)b->objdel_cap) {  // grow the list, keeping what earlier calls of this tick stamped
        CK(cudaStreamSynchronize(w->stream));
        const unsigned int cap = (unsigned int)((long long)b->stamped + cells) * 2 + 1024;
        long long* nb = nullptr;
        CK(cudaMalloc((void**)&nb, sizeof(long long) * (size_t)cap));
        if (b->stamped) CK(cudaMemcpy(nb, b->d_objdel, sizeof(long long) * (size_t)b->stamped, cudaMemcpyDeviceToDevice));
        cudaFree(b->d_objdel);
        b->d_objdel = nb;
        b->objdel_cap = cap;
    }
    b->stamped += (unsigned int)cells;
    EntArgs a;
    entity_args(w, b, n, load_x, load_y, tick, seed, &a);
    a.object_mat = object_mat;
    if (w->strip && w->ctx->nranks > 1)  // every rank stamps the cells of its window (they are independent); ghost rows must be the owner's first
        if (int r = strip_refresh(w, w->stream, STRIP_GHOST)) return r;
    for (int i = 0; i < n; i++) {
        fse_rect r{(int)(ents[i].x + load_x) - 1, (int)(ents[i].y + load_y) - 1 - w->y_off, ents[i].hw + 2, ents[i].hh + 2};
        if (r.x < 0) r.x = 0;
        if (r.y < 0) r.y = 0;
        if (r.x + r.w > w->W) r.w = w->W - r.x;
        if (r.y + r.h > w->H) r.h = w->H - r.y;
        if (r.w <= 0 || r.h <= 0) continue;
        b->rects.push_back(r);
        if (int rr = fse_wake_rect(w, r.x, r.y, r.w, r.h)) return rr;
    }
    entities_stamp_kernel<<<1, ENT_THREADS, 0, w->stream>>>(a);
    CK(cudaGetLastError());
    w->ctx->launches += 1;
    return FSE_OK;
}

extern "C" FSE_API int fse_object_delete(fse_world* w) {
    if (!w) return fail(FSE_EINVAL, "fse_object_delete: null world");
    EntityBufs* b = (EntityBufs*)w->entity_bufs;
    if (!b || b->stamped == 0) return FSE_OK;
    CK(cudaSetDevice(w->ctx->device));
    unsigned int n = 0;
    CK(cudaMemcpyAsync(&n, b->d_cnt, sizeof n, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    if (n > b->objdel_cap) n = b->objdel_cap;
    if (n) {
        Planes p = w->p;
        if (w->strip) {  // the list holds global cell indices (see EntArgs)
            const size_t back = (size_t)w->y_off * w->W;
            p.mat -= back; p.flg -= back; p.stl -= back; p.tmp -= back; p.col -= back; p.fl -= back; p.fd -= back;
        }
        object_delete_kernel<<<(n + 255) / 256, 256, 0, w->stream>>>(p, b->d_objdel, n, w->ctx->h_tabs.air);
        CK(cudaGetLastError());
        w->ctx->launches += 1;
    }
    CK(cudaMemsetAsync(b->d_cnt, 0, sizeof(unsigned int), w->stream));
    for (const fse_rect& r : b->rects)
        if (int rr = fse_wake_rect(w, r.x, r.y, r.w, r.h)) return rr;
    b->rects.clear();
    b->stamped = 0;
    return FSE_OK;
}
