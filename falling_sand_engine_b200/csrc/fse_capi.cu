// fse_capi.cu — the C ABI of include/fse.h on top of the sm_100a kernels.
// Host logic only: validation, device memory, launch sequencing (the 4 colours x cell_iter schedule of
// world::tick, world.cpp:1050-1077), error mapping.  No CPU fallback exists: every entry point needs the GPU.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <algorithm>
#include <string>
#include <vector>

#include "fse_device.cuh"
#include "fse_internal.hpp"

namespace fse {
thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
}  // namespace fse

using namespace fse;

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return fail(FSE_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

extern "C" {

FSE_API const char* fse_last_error(void) { return g_err.c_str(); }
FSE_API const char* fse_version(void) { return "fse-b200 0.1 (sm_100a)"; }

FSE_API int fse_abi_sizeof(int which) {
    switch (which) {
        case 0: return (int)sizeof(fse_material);
        case 1: return (int)sizeof(fse_interaction);
        case 2: return (int)sizeof(fse_special_ids);
        case 3: return (int)sizeof(fse_cell);
        case 4: return (int)sizeof(fse_rect);
        case 5: return (int)sizeof(fse_tick_args);
        case 6: return (int)sizeof(fse_particle);
        case 7: return (int)sizeof(fse_stats);
        case 8: return (int)sizeof(fse_render_stats);
    }
    return -1;
}

FSE_API int fse_ctx_create(int device, fse_ctx** out) {
    if (!out) return fail(FSE_EINVAL, "fse_ctx_create: out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(FSE_ECUDA, "fse_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(FSE_EINVAL, "fse_ctx_create: device %d out of range [0,%d)", device, n);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(FSE_ECUDA, "fse_ctx_create: device is sm_%d%d; kernels are built for sm_100a only", prop.major, prop.minor);
    fse_ctx* c = new fse_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    CK(cudaMalloc(&c->d_tabs, sizeof(DevTables)));
    *out = c;
    return FSE_OK;
}

FSE_API void fse_ctx_destroy(fse_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaFree(c->d_tabs);
    cudaFree(c->d_inter);
    cudaFree(c->d_inter_off);
    cudaFree(c->d_react);
    delete c;
}

FSE_API int64_t fse_launch_count(fse_ctx* c) { return c ? c->launches.load() : 0; }

FSE_API int fse_materials_set(fse_ctx* c, const fse_material* tbl, int n, const fse_special_ids* ids, const fse_interaction* inter,
                              const int32_t* inter_offsets, const fse_interaction* react, const int32_t* react_offsets) {
    if (!c || !tbl || !ids) return fail(FSE_EINVAL, "fse_materials_set: null argument");
    if (n < 1 || n > FSE_MAX_MATERIALS) return fail(FSE_EINVAL, "fse_materials_set: %d materials (device plane is u8: 1..%d)", n, FSE_MAX_MATERIALS);
    const int sid[6] = {ids->air, ids->fire, ids->water, ids->lava, ids->steam, ids->obsidian};
    for (int i = 0; i < 6; i++)
        if (sid[i] < 0 || sid[i] >= n) return fail(FSE_EINVAL, "fse_materials_set: special id #%d = %d out of range", i, sid[i]);
    if (tbl[ids->air].physics != FSE_AIR) return fail(FSE_EINVAL, "fse_materials_set: ids.air must have physics AIR");
    if (tbl[ids->fire].physics != FSE_PASSABLE) return fail(FSE_EINVAL, "fse_materials_set: ids.fire must be PASSABLE (gds.cpp:107)");
    CK(cudaSetDevice(c->device));
    DevTables& h = c->h_tabs;
    memset(&h, 0, sizeof h);
    h.n = n;
    h.air = ids->air; h.fire = ids->fire; h.water = ids->water; h.lava = ids->lava; h.steam = ids->steam; h.obsidian = ids->obsidian;
    int n_inter = inter_offsets ? inter_offsets[n * n] : 0;
    int n_react = react_offsets ? react_offsets[n] : 0;
    int max_reach = 0;
    int n_irows = 0;
    for (int i = 0; i < n; i++) {
        const fse_material& m = tbl[i];
        if (m.physics < 0 || m.physics > 5) return fail(FSE_EINVAL, "material %d: physics %d", i, m.physics);
        if (m.physics == FSE_SAND && (m.slipperyness < 1 || m.slipperyness > 255))
            return fail(FSE_EINVAL, "material %d: SAND needs 1 <= slipperyness <= 255 (reference divides by it, world.cpp:1656)", i);
        if (m.iterations < 0) return fail(FSE_EINVAL, "material %d: negative iterations", i);
        if (!std::isfinite(m.conduction_self) || !std::isfinite(m.conduction_other) || !std::isfinite(m.density))
            return fail(FSE_EINVAL, "material %d: conduction / density must be finite (the temperature kernel adds the zero contribution of a "
                                    "neighbour at temperature 0 instead of skipping it, world.cpp:1976)", i);
        Lut& L = h.lut;
        h.phys[i] = (uint8_t)m.physics;
        L.phys[i] = (uint8_t)m.physics;
        L.iters[i] = (uint8_t)(m.iterations > 255 ? 255 : m.iterations);
        int nr = react_offsets ? react_offsets[i + 1] - react_offsets[i] : 0;
        L.mflags[i] = (uint8_t)((m.interact ? MF_INTERACT : 0) | ((m.react && nr > 0) ? MF_REACT : 0) | (nr > 1 ? MF_REACT_MULTI : 0));
        L.slip[i] = (uint8_t)(m.slipperyness < 0 ? 0 : (m.slipperyness > 255 ? 255 : m.slipperyness));
        L.maxstab[i] = m.slipperyness >= 1 ? (uint8_t)(int)(8 / sqrt((double)m.slipperyness) + 1) : 0;  // world.cpp:1630
        L.dens[i] = m.density;
        if (nr > 0) {
            const fse_interaction& r0 = react[react_offsets[i]];
            L.rx[i].thr = r0.data1;
            L.rx[i].type = (uint8_t)r0.type;
            L.rx[i].prod = (uint8_t)r0.data2;
        }
        if (m.interact && inter_offsets) {  // partner bitmap: which materials below trigger an interaction list
            bool any = false;
            for (int b = 0; b < n; b++) any |= inter_offsets[i * n + b + 1] > inter_offsets[i * n + b];
            if (any) {
                if (n_irows < LUT_IROWS) {
                    for (int b = 0; b < n; b++)
                        if (inter_offsets[i * n + b + 1] > inter_offsets[i * n + b]) L.ibits[n_irows][b >> 5] |= 1u << (b & 31);
                    L.irow[i] = (uint8_t)(++n_irows);
                } else {
                    L.mflags[i] |= MF_INTERACT_SLOW;
                }
            }
        }
        h.alpha[i] = m.alpha;
        h.ckind[i] = m.color_kind;
        h.jshift[i] = m.jitter_shift;
        h.jrange[i] = m.jitter_range;
        h.ctemp[i] = m.create_temp;
        h.density[i] = m.density;
        h.color[i] = m.color;
        h.add_temp[i] = m.add_temp;
        h.emit_color[i] = m.emit_color;
        h.cond_self[i] = m.conduction_self;
        h.cond_other[i] = m.conduction_other;
        h.react_off[i] = react_offsets ? react_offsets[i] : 0;
    }
    h.react_off[n] = n_react;
    for (int i = n + 1; i <= FSE_MAX_MATERIALS; i++) h.react_off[i] = n_react;
    for (int k = 0; k < n_react; k++) {
        if (react[k].data2 >= (uint32_t)n) return fail(FSE_EINVAL, "reaction %d: product %u out of range", k, react[k].data2);
    }
    for (int a = 0; a < n && inter_offsets; a++)
        for (int b = 0; b < n; b++)
            for (int k = inter_offsets[a * n + b]; k < inter_offsets[a * n + b + 1]; k++) {
                const fse_interaction& in = inter[k];
                if (in.type != FSE_INTERACT_TRANSFORM_MATERIAL && in.type != FSE_INTERACT_SPAWN_MATERIAL) continue;
                if (in.data1 < 0 || in.data1 >= n) return fail(FSE_EINVAL, "interaction %d: product %d out of range", k, in.data1);
                if (!tbl[a].interact) continue;  // dead list: never consulted
                int reach = (int)in.data2 + (abs(in.ofs_x) > abs(in.ofs_y) ? abs(in.ofs_x) : abs(in.ofs_y));
                if (reach > max_reach) max_reach = reach;
            }
    if (max_reach > FSE_MAX_REACH)
        return fail(FSE_EINVAL, "fse_materials_set: interaction reach %d exceeds FSE_MAX_REACH=%d (|ofs|+radius)", max_reach, FSE_MAX_REACH);
    cudaFree(c->d_inter); c->d_inter = nullptr;
    cudaFree(c->d_inter_off); c->d_inter_off = nullptr;
    cudaFree(c->d_react); c->d_react = nullptr;
    CK(cudaMalloc(&c->d_inter_off, sizeof(int32_t) * ((size_t)n * n + 1)));
    if (inter_offsets) {
        CK(cudaMemcpy(c->d_inter_off, inter_offsets, sizeof(int32_t) * ((size_t)n * n + 1), cudaMemcpyHostToDevice));
    } else {
        CK(cudaMemset(c->d_inter_off, 0, sizeof(int32_t) * ((size_t)n * n + 1)));
    }
    CK(cudaMalloc(&c->d_inter, sizeof(fse_interaction) * (size_t)(n_inter > 0 ? n_inter : 1)));
    if (n_inter > 0) CK(cudaMemcpy(c->d_inter, inter, sizeof(fse_interaction) * n_inter, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&c->d_react, sizeof(fse_interaction) * (size_t)(n_react > 0 ? n_react : 1)));
    if (n_react > 0) CK(cudaMemcpy(c->d_react, react, sizeof(fse_interaction) * n_react, cudaMemcpyHostToDevice));
    h.react = c->d_react;
    h.inter_off = c->d_inter_off;
    h.inter = c->d_inter;
    CK(cudaMemcpy(c->d_tabs, &h, sizeof h, cudaMemcpyHostToDevice));
    c->has_materials = true;
    c->max_reach = max_reach;
    return FSE_OK;
}

// ---- world ---------------------------------------------------------------------------------------------
static void free_world(fse_world* w) {
    cudaFree(w->p.mat); cudaFree(w->p.flg); cudaFree(w->p.stl); cudaFree(w->p.tmp); cudaFree(w->p.col); cudaFree(w->p.fl); cudaFree(w->p.fd);
    cudaFree(w->tmp_scratch); cudaFree(w->pbuf); cudaFree(w->pcount); cudaFree(w->d_stats); cudaFree(w->d_stage);
    cudaFree(w->pbuf2); cudaFree(w->part_scratch); cudaFree(w->claim_keys); cudaFree(w->claim_vals);
    if (w->h_stats) cudaFreeHost(w->h_stats);
    for (auto& ev : w->kt_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    if (w->ev0) cudaEventDestroy(w->ev0);
    if (w->ev1) cudaEventDestroy(w->ev1);
    for (int q = 0; q < 3; q++) {
        if (w->fork.aux[q]) cudaStreamDestroy(w->fork.aux[q]);
        if (w->fork.ev_join[q]) cudaEventDestroy(w->fork.ev_join[q]);
    }
    if (w->fork.ev_fork) cudaEventDestroy(w->fork.ev_fork);
    if (w->stream) cudaStreamDestroy(w->stream);
    if (w->comm_stream) cudaStreamDestroy(w->comm_stream);
    if (w->ev_boundary) cudaEventDestroy(w->ev_boundary);
    if (w->ev_comm) cudaEventDestroy(w->ev_comm);
    cudaFree(w->d_chunk_lists);
    cudaFree(w->d_awake); cudaFree(w->d_active_list); cudaFree(w->d_active_count);
    cudaFree(w->part_list);
    render_free(w);
    if (w->outline_pinned) cudaFreeHost(w->outline_pinned);
    if (w->outline_pinned2) cudaFreeHost(w->outline_pinned2);
    for (int q = 0; q < 4; q++) cudaFree(w->halo_stage[q]);
    cudaFree(w->d_push_rects);
    cudaFree(w->d_push_off);
    for (cudaEvent_t e_ : w->timeline_ev) cudaEventDestroy(e_);
    cudaFree(w->d_phase_rows);
    if (w->h_phase_rows) cudaFreeHost(w->h_phase_rows);
    if (w->ev_phase_rows) cudaEventDestroy(w->ev_phase_rows);
    cudaFree(w->d_lpt_cost); cudaFree(w->d_lpt_list); cudaFree(w->d_chunk_state); cudaFree(w->d_rowmask);
    fse_bodies_free(w);
    particles_strip_free(w);
    entities_free(w);
    cudaFree(w->tool_scratch);
    cudaFree(w->outline_scratch);
    delete w;
}

static int make_world(fse_ctx* c, int32_t width, int32_t height, fse_world** out) {
    CK(cudaSetDevice(c->device));
    fse_world* w = new fse_world();
    w->ctx = c;
    w->W = width;
    w->H = height;
    w->Hglobal = height;
    w->own_lo = 0;
    w->own_hi = height;
    const size_t n = (size_t)width * height;
    cudaError_t e = cudaSuccess;
    auto A = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void**)&w->p.mat, n); A((void**)&w->p.flg, n); A((void**)&w->p.stl, n); A((void**)&w->p.tmp, n * 2); A((void**)&w->p.col, n * 4);
    A((void**)&w->p.fl, n * 4); A((void**)&w->p.fd, n * 4); A((void**)&w->tmp_scratch, n * 2);
    w->pcap = 1u << 20;
    A((void**)&w->pbuf, sizeof(fse_particle) * (size_t)w->pcap);
    A((void**)&w->pcount, 64);
    A((void**)&w->d_stats, dev_stats_bytes());
    if (e == cudaSuccess) e = cudaMallocHost(&w->h_stats, dev_stats_bytes());
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
    if (const char* env = getenv("FSE_TICK_LPT")) w->lpt_on = atoi(env) != 0;
    if (const char* env = getenv("FSE_LPT_DEAL")) w->lpt_deal = atoi(env) != 0;
    if (const char* env = getenv("FSE_STRIP_TIMELINE")) w->timeline_on = atoi(env) != 0;
    if (const char* env = getenv("FSE_STRIP_BOUNDARY_MIN")) w->strip_boundary_min = atoi(env);
    if (const char* env = getenv("FSE_ROW_SKIP")) w->rowskip_mode = atoi(env);
    if (const char* env = getenv("FSE_P2_SPLIT")) w->p2_split = atoi(env);
    if (const char* env = getenv("FSE_P2_SPLIT_MAX_SEQ")) w->p2_split_max_seq = (float)atof(env);
    if (const char* env = getenv("FSE_ROW_SKIP_MAX_ACTIVE")) w->rowskip_max_active = (float)atof(env);
    for (int q = 0; q < 16; q++) {
        w->phase_active[q] = -1.0f;
        w->phase_seq[q] = -1.0f;
        w->phase_rows_total[q] = 0;
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&w->d_phase_rows, 32 * sizeof(unsigned int));  // [0..15] active rows, [16..31] rows with live powder / gas
    if (e == cudaSuccess) e = cudaMallocHost((void**)&w->h_phase_rows, 32 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->ev_phase_rows, cudaEventDisableTiming);
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
        w->fused_max_chunks = 2 * sms;  // tick_rows_kernel: 2 CTAs per SM
        if (const char* env = getenv("FSE_FUSED_MAX_CHUNKS")) w->fused_max_chunks = atoi(env);
    }
    w->fork.parts = 3;
    w->fork.min_chunks = 256;
    if (const char* env = getenv("FSE_TICK_MIN_CHUNKS")) w->fork.min_chunks = std::max(1, atoi(env));
    if (const char* env = getenv("FSE_TICK_PARTS")) w->fork.parts = std::min(4, std::max(1, atoi(env)));
    for (int q = 0; q < 3; q++) {
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->fork.aux[q], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->fork.ev_join[q], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->fork.ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&w->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&w->ev1);
    if (e == cudaSuccess) e = cudaMemsetAsync(w->pcount, 0, 64, w->stream);
    if (e == cudaSuccess) e = launch_fill_air(w->p, n, (uint8_t)c->h_tabs.air, w->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(w->stream);
    if (e != cudaSuccess) {
        free_world(w);
        return fail(e == cudaErrorMemoryAllocation ? FSE_ENOMEM : FSE_ECUDA, "fse_world_create(%d,%d): %s", width, height, cudaGetErrorString(e));
    }
    c->launches += 1;
    *out = w;
    return FSE_OK;
}

FSE_API int fse_world_create(fse_ctx* c, int32_t width, int32_t height, fse_world** out) {
    if (!c || !out) return fail(FSE_EINVAL, "fse_world_create: null argument");
    if (!c->has_materials) return fail(FSE_ESTATE, "fse_world_create: call fse_materials_set first (materials_init/push, game.lua:65-66)");
    if (width < 3 * CHUNK || height < 3 * CHUNK || width % 16 || width > (1 << 18) || height > (1 << 18))
        return fail(FSE_EINVAL, "fse_world_create: %dx%d (need >= %d, width multiple of 16, <= 262144)", width, height, 3 * CHUNK);
    return make_world(c, width, height, out);
}

// One strip of a world partitioned across the ranks of fse_comm_init.  Cuts fall on chunk-row
// boundaries of the default tickZone (y = 128 + 128*j); every rank keeps GHOST rows of its neighbours.
static const int GHOST = 32;
FSE_API int fse_strip_create(fse_ctx* c, int32_t width, int32_t height_global, fse_world** out) {
    if (!c || !out) return fail(FSE_EINVAL, "fse_strip_create: null argument");
    if (!c->has_materials) return fail(FSE_ESTATE, "fse_strip_create: call fse_materials_set first");
    if (!c->nccl_comm && c->nranks > 1) return fail(FSE_ESTATE, "fse_strip_create: call fse_comm_init first");
    if (width < 3 * CHUNK || width % 16 || width > (1 << 18) || height_global > (1 << 18) || height_global % CHUNK)
        return fail(FSE_EINVAL, "fse_strip_create: %dx%d (width multiple of 16, height multiple of %d)", width, height_global, CHUNK);
    const int nz = (height_global - 2 * CHUNK) / CHUNK;  // chunk rows of the tickZone
    if (nz < c->nranks) return fail(FSE_EINVAL, "fse_strip_create: %d chunk rows cannot be split over %d ranks", nz, c->nranks);
    const int j0 = (int)((int64_t)nz * c->rank / c->nranks), j1 = (int)((int64_t)nz * (c->rank + 1) / c->nranks);
    const int own_lo = c->rank == 0 ? 0 : CHUNK + CHUNK * j0;
    const int own_hi = c->rank == c->nranks - 1 ? height_global : CHUNK + CHUNK * j1;
    const int y_off = own_lo - GHOST < 0 ? 0 : own_lo - GHOST;
    const int y_end = own_hi + GHOST > height_global ? height_global : own_hi + GHOST;
    fse_world* w = nullptr;
    if (int r = make_world(c, width, y_end - y_off, &w)) return r;
    w->strip = true;
    w->y_off = y_off;
    w->Hglobal = height_global;
    w->own_lo = own_lo;
    w->own_hi = own_hi;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaError_t e = cudaStreamCreateWithPriority(&w->comm_stream, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->ev_boundary, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->ev_comm, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        free_world(w);
        return fail(FSE_ECUDA, "fse_strip_create: %s", cudaGetErrorString(e));
    }
    *out = w;
    return FSE_OK;
}

FSE_API int fse_strip_rows(fse_world* w, int32_t* own_lo, int32_t* own_hi, int32_t* held_lo, int32_t* held_hi) {
    if (!w) return fail(FSE_EINVAL, "fse_strip_rows: null world");
    if (own_lo) *own_lo = w->own_lo;
    if (own_hi) *own_hi = w->own_hi;
    if (held_lo) *held_lo = w->y_off;
    if (held_hi) *held_hi = w->y_off + w->H;
    return FSE_OK;
}

// Owner-authoritative refresh of the ghost rows (after edits outside fse_tick: write_rect, temperature, particles).
FSE_API int fse_strip_refresh(fse_world* w) {
    if (!w) return fail(FSE_EINVAL, "fse_strip_refresh: null world");
    if (!w->strip || w->ctx->nranks == 1) return FSE_OK;
    CK(cudaSetDevice(w->ctx->device));
    return strip_refresh(w, w->stream);
}

FSE_API void fse_world_destroy(fse_world* w) {
    if (!w) return;
    cudaSetDevice(w->ctx->device);
    cudaStreamSynchronize(w->stream);
    free_world(w);
}

FSE_API void fse_probe_position(uint32_t seed, uint32_t tick, const fse_rect* z, int32_t* x, int32_t* y) {
    if (!z) return;
    const uint32_t cb = rng_cell(rng_key(seed, tick, 15u), 0, 0);
    if (x) *x = z->x + (int32_t)(rng_draw(cb, S_PROBE_X) % (uint32_t)(z->w > 0 ? z->w : 1));
    if (y) *y = z->y + (int32_t)(rng_draw(cb, S_PROBE_Y) % (uint32_t)(z->h > 0 ? z->h : 1));
}

FSE_API int fse_sync(fse_world* w) {
    if (!w) return fail(FSE_EINVAL, "fse_sync: null world");
    CK(cudaStreamSynchronize(w->stream));
    return FSE_OK;
}

static const size_t STAGE_MAX_CELLS = (size_t)16 << 20;  // 320 MB of AoS staging at most
static int wake_rect(fse_world* w, int x, int y_local, int rw, int rh);

FSE_API int fse_write_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, const fse_cell* cells) {
    if (!w || !cells) return fail(FSE_EINVAL, "fse_write_rect: null argument");
    if (int r = check_rect(w, x, y, rw, rh, "fse_write_rect")) return r;
    y -= w->y_off;
    CK(cudaSetDevice(w->ctx->device));
    const int nmat = w->ctx->h_tabs.n;
    int band = (int)(STAGE_MAX_CELLS / (size_t)rw);
    if (band < 1) band = 1;
    if (band > rh) band = rh;
    if (int r = ensure_stage(w, (size_t)band * rw)) return r;
    for (int yy = 0; yy < rh; yy += band) {
        int hb = rh - yy < band ? rh - yy : band;
        const fse_cell* src = cells + (size_t)yy * rw;
        {   // every cell: ids beyond the table would alias zero-filled LUT rows and index the interaction offsets out of range
            const size_t nb = (size_t)hb * rw;
            unsigned int worst = 0;
            for (size_t i = 0; i < nb; i++) worst = src[i].mat > worst ? src[i].mat : worst;
            if ((int)worst >= nmat) return fail(FSE_EINVAL, "fse_write_rect: cell material %u >= %d (material table size)", worst, nmat);
        }
        CK(cudaMemcpyAsync(w->d_stage, src, (size_t)hb * rw * sizeof(fse_cell), cudaMemcpyHostToDevice, w->stream));
        CK(launch_write_rect(w->p, w->W, x, y + yy, rw, hb, w->d_stage, w->stream));
        if (int r = wake_rect(w, x, y + yy, rw, hb)) return r;
        w->ctx->launches += 1;
        if (yy + band < rh) CK(cudaStreamSynchronize(w->stream));  // the staging buffer is reused
    }
    return FSE_OK;
}

FSE_API int fse_read_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, fse_cell* cells) {
    if (!w || !cells) return fail(FSE_EINVAL, "fse_read_rect: null argument");
    if (int r = check_rect(w, x, y, rw, rh, "fse_read_rect")) return r;
    y -= w->y_off;
    CK(cudaSetDevice(w->ctx->device));
    int band = (int)(STAGE_MAX_CELLS / (size_t)rw);
    if (band < 1) band = 1;
    if (band > rh) band = rh;
    if (int r = ensure_stage(w, (size_t)band * rw)) return r;
    for (int yy = 0; yy < rh; yy += band) {
        int hb = rh - yy < band ? rh - yy : band;
        CK(launch_read_rect(w->p, w->W, x, y + yy, rw, hb, w->d_stage, w->stream));
        w->ctx->launches += 1;
        CK(cudaMemcpyAsync(cells + (size_t)yy * rw, w->d_stage, (size_t)hb * rw * sizeof(fse_cell), cudaMemcpyDeviceToHost, w->stream));
        CK(cudaStreamSynchronize(w->stream));
    }
    return FSE_OK;
}

FSE_API int fse_clear_dirty(fse_world* w) {
    if (!w) return fail(FSE_EINVAL, "fse_clear_dirty: null world");
    CK(cudaSetDevice(w->ctx->device));
    CK(launch_clear_dirty(w->p, (size_t)w->W * w->H, w->stream));
    w->ctx->launches += 1;
    return FSE_OK;
}

FSE_API int fse_stats_rect(fse_world* w, int32_t x, int32_t y, int32_t rw, int32_t rh, fse_stats* out) {
    if (!w || !out) return fail(FSE_EINVAL, "fse_stats_rect: null argument");
    if (int r = check_rect(w, x, y, rw, rh, "fse_stats_rect")) return r;
    CK(cudaSetDevice(w->ctx->device));
    CK(launch_stats(w->p, w->W, x, y - w->y_off, rw, rh, w->y_off, w->ctx->d_tabs, w->d_stats, w->stream));
    w->ctx->launches += 1;
    CK(cudaMemcpyAsync(w->h_stats, w->d_stats, dev_stats_bytes(), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    static_assert(sizeof(fse_stats) == 8 + 8 * FSE_MAX_MATERIALS * 2 + 16, "fse_stats layout");
    memcpy(out, w->h_stats, sizeof(fse_stats));  // DevStats and fse_stats share one layout
    return FSE_OK;
}

// ---- active-region tracking ---------------------------------------------------------------------------------
// wake every chunk a (local-coordinate) rect touches, with a 16-cell margin (footprints of the neighbours' rules)
static int wake_rect(fse_world* w, int x, int y_local, int rw, int rh) {
    if (!w->active_on) return FSE_OK;
    const int yg = y_local + w->y_off;
    int i0 = (x - 16) / CHUNK, i1 = (x + rw + 15) / CHUNK, j0 = (yg - 16) / CHUNK, j1 = (yg + rh + 15) / CHUNK;
    if (i0 < 0) i0 = 0;
    if (j0 < 0) j0 = 0;
    if (i1 >= w->acols) i1 = w->acols - 1;
    if (j1 >= w->arows) j1 = w->arows - 1;
    if (i1 < i0 || j1 < j0) return FSE_OK;
    CK(cudaMemset2DAsync(w->d_awake + (size_t)j0 * w->acols + i0, w->acols, 1, i1 - i0 + 1, j1 - j0 + 1, w->stream));
    return FSE_OK;
}

// Turn per-chunk sleeping on/off.  While on, fse_tick compacts the awake chunks of each colour on the device and only
// those run; the cell state is identical to a full sweep (dirty flags of sleeping chunks are not refreshed).
FSE_API int fse_active_enable(fse_world* w, int enable) {
    if (!w) return fail(FSE_EINVAL, "fse_active_enable: null world");
    CK(cudaSetDevice(w->ctx->device));
    if (enable && !w->d_awake) {
        w->acols = w->W / CHUNK;
        w->arows = w->Hglobal / CHUNK;
        CK(cudaMalloc(&w->d_awake, (size_t)w->acols * w->arows));
        CK(cudaMalloc(&w->d_active_list, sizeof(int) * ((size_t)w->acols * w->arows / 4 + w->acols + w->arows + 4)));
        CK(cudaMalloc(&w->d_active_count, 64));
        CK(cudaMalloc((void**)&w->d_chunk_state, sizeof(unsigned int) * (size_t)w->acols * w->arows));
        if (const char* env = getenv("FSE_ACTIVE_FUSED")) w->active_fused = atoi(env) != 0;
    }
    if (enable) CK(cudaMemsetAsync(w->d_awake, 1, (size_t)w->acols * w->arows, w->stream));
    w->active_on = enable != 0;
    return FSE_OK;
}

FSE_API int fse_active_stats(fse_world* w, int64_t* awake, int64_t* total) {
    if (!w || !awake || !total) return fail(FSE_EINVAL, "fse_active_stats: null argument");
    *total = 0;
    *awake = 0;
    if (!w->d_awake) return FSE_OK;
    CK(cudaSetDevice(w->ctx->device));
    std::vector<uint8_t> h((size_t)w->acols * w->arows);
    CK(cudaMemcpyAsync(h.data(), w->d_awake, h.size(), cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    *total = (int64_t)h.size();
    for (uint8_t v : h) *awake += v != 0;
    return FSE_OK;
}

// ---- the tick: 4 colours x cell_iter (world.cpp:1050-1077) ----------------------------------------------------
static int check_zone(fse_world* w, const fse_rect& z, const char* who) {
    if (z.w <= 0 || z.h <= 0 || z.w % CHUNK || z.h % CHUNK)
        return fail(FSE_EINVAL, "%s: tick zone %dx%d must be a positive multiple of %d (chunk tasks are whole chunks, world.cpp:1073-1086)", who,
                    z.w, z.h, CHUNK);
    if (z.x % 16 || z.x < 16 || z.y < 16 || z.x + z.w + 16 > w->W || z.y + z.h + 16 > w->Hglobal)
        return fail(FSE_EINVAL, "%s: tick zone (%d,%d,%d,%d) needs x %% 16 == 0 and a >=16-cell margin inside the %dx%d world", who, z.x, z.y, z.w,
                    z.h, w->W, w->Hglobal);
    if (w->strip && (z.y - CHUNK) % CHUNK)
        return fail(FSE_EINVAL, "%s: strip worlds are cut at y = 128 + 128*j; tick zone y=%d is not on that grid", who, z.y);
    return FSE_OK;
}

struct KtScope {  // CUDA events around one launch of the dominant kernel (bench.py roofline)
    fse_world* w;
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    int begin(cudaStream_t s) {
        if (!w->kt_enabled) return FSE_OK;
        if (w->kt_used < w->kt_events.size()) {
            ev = w->kt_events[w->kt_used];
        } else {
            CK(cudaEventCreate(&ev.first));
            CK(cudaEventCreate(&ev.second));
            w->kt_events.push_back(ev);
        }
        w->kt_used++;
        CK(cudaEventRecord(ev.first, s));
        return FSE_OK;
    }
    int end(cudaStream_t s) {
        if (w->kt_enabled) CK(cudaEventRecord(ev.second, s));
        return FSE_OK;
    }
};

// Per-phase chunk lists of a strip: [tk][0] = chunk rows touching a cut (launched first), [tk][1] = the rest.
static int build_strip_lists(fse_world* w, const fse_rect& z, int j0, int j1) {
    if (w->d_chunk_lists && memcmp(&w->list_zone, &z, sizeof z) == 0) return FSE_OK;
    const int nx = z.w / CHUNK;
    std::vector<int> all;
    const bool up = w->ctx->rank > 0, down = w->ctx->rank + 1 < w->ctx->nranks;
    for (int tk = 0; tk < 4; tk++) {
        const int ofx = tk % 2, ofy = 1 - (tk / 2);
        for (int part = 0; part < 2; part++) {
            w->list_off[tk][part] = (int)all.size();
            for (int j = j0; j < j1; j++) {
                if ((j % 2) != ofy) continue;
                const bool boundary = (up && j == j0) || (down && j == j1 - 1);
                if (boundary != (part == 0)) continue;
                for (int i = ofx; i < nx; i += 2) all.push_back((i / 2) | ((j / 2) << 16));
            }
            w->list_cnt[tk][part] = (int)all.size() - w->list_off[tk][part];
        }
    }
    cudaFree(w->d_chunk_lists);
    w->d_chunk_lists = nullptr;
    CK(cudaMalloc(&w->d_chunk_lists, sizeof(int) * (all.size() + 1)));
    if (!all.empty()) CK(cudaMemcpy(w->d_chunk_lists, all.data(), sizeof(int) * all.size(), cudaMemcpyHostToDevice));
    w->list_zone = z;
    memset(w->lpt_sig, 0, sizeof(w->lpt_sig));
    return FSE_OK;
}

// Profiling aid (not part of the reference surface): accumulate, per warp role of tick_rows_kernel, the cycles spent working
// between step barriers.  out[0..3] = pass1, pass2, pass3, IO cycles; out[4] = chunks processed.
FSE_API int fse_debug_role_cycles(fse_world* w, int enable, unsigned long long* out) {
    if (!w) return fail(FSE_EINVAL, "fse_debug_role_cycles: null world");
    CK(cudaSetDevice(w->ctx->device));
    if (enable && !w->d_dbg) {
        CK(cudaMalloc(&w->d_dbg, 512));
        CK(cudaMemset(w->d_dbg, 0, 512));
    }
    if (out && w->d_dbg) {
        CK(cudaStreamSynchronize(w->stream));
        CK(cudaMemcpy(out, w->d_dbg, 512, cudaMemcpyDeviceToHost));
        CK(cudaMemset(w->d_dbg, 0, 512));
    }
    if (!enable && w->d_dbg) {
        cudaFree(w->d_dbg);
        w->d_dbg = nullptr;
    }
    return FSE_OK;
}

FSE_API int fse_set_schedule(fse_world* w, int schedule) {
    if (!w || schedule < FSE_SCHEDULE_ROWS || schedule > FSE_SCHEDULE_ROWS_FUSED)
        return fail(FSE_EINVAL, "fse_set_schedule: schedule %d (1 = rows, 2 = rows fused; the classes schedule 0 was removed)", schedule);
    w->schedule = schedule == FSE_SCHEDULE_ROWS_FUSED ? FSE_SCHEDULE_ROWS : schedule;
    w->fused = schedule == FSE_SCHEDULE_ROWS_FUSED;
    return FSE_OK;
}

FSE_API int fse_tick(fse_world* w, const fse_tick_args* a) {
    if (!w || !a) return fail(FSE_EINVAL, "fse_tick: null argument");
    if (a->cell_iter < 0 || a->cell_iter > 4) return fail(FSE_EINVAL, "fse_tick: cell_iter %d (0..4; particle ids carry 2 bits)", a->cell_iter);
    if (int r = check_zone(w, a->tick_zone, "fse_tick")) return r;
    CK(cudaSetDevice(w->ctx->device));
    const fse_rect z = a->tick_zone;
    const int nx = z.w / CHUNK, ny = z.h / CHUNK;
    // strip worlds: owned chunk rows [j0, j1) of the zone
    int j0 = 0, j1 = ny;
    if (w->strip) {
        const int lo = w->ctx->rank == 0 ? z.y : w->own_lo, hi = w->ctx->rank == w->ctx->nranks - 1 ? z.y + z.h : w->own_hi;
        j0 = (lo - z.y) / CHUNK;
        j1 = (hi - z.y) / CHUNK;
        if (j0 < 0) j0 = 0;
        if (j1 > ny) j1 = ny;
        if (int r = build_strip_lists(w, z, j0, j1)) return r;
    }
    const bool multi = w->strip && w->ctx->nranks > 1;
    // falling sand and liquid leave the grid as loose particles: keep an eighth of the pool free per tick (the live count is read
    // back only when the promises since the last read could exceed the pool)
    if (int r = particles_headroom(w, (int64_t)(w->pcap / 8), false)) return r;
    KtScope kt{w};
    // skip gate: collect the counts of the last tick that finished, then zero the counters for this one
    if (w->phase_rows_pending && cudaEventQuery(w->ev_phase_rows) == cudaSuccess) {
        for (int q = 0; q < 16; q++)
            if (((w->phase_rows_valid >> q) & 1u) && w->phase_rows_total[q] > 0) {
                w->phase_active[q] = (float)w->h_phase_rows[q] / (float)w->phase_rows_total[q];
                w->phase_seq[q] = (float)w->h_phase_rows[16 + q] / (float)w->phase_rows_total[q];
            }
        w->phase_rows_pending = false;
    }
    const bool gate_probe = (w->ticks % 32) == 0;
    const bool gate_collect = !w->phase_rows_pending && w->rowskip_mode == 2;  // one copy in flight at a time
    if (gate_collect) {
        CK(cudaMemsetAsync(w->d_phase_rows, 0, 32 * sizeof(unsigned int), w->stream));
        w->phase_rows_counted = 0;
    }
    for (int iter = 0; iter < a->cell_iter; iter++) {
        for (int tk = 0; tk < 4; tk++) {
            const int ofx = tk % 2;              // 0 1 0 1   (world.cpp:1059)
            const int ofy = 1 - ((tk % 4) / 2);  // 1 1 0 0   (world.cpp:1060)
            TickParams P;
            P.p = w->p;
            P.W = w->W;
            P.H = w->H;
            P.y_off = w->y_off;
            P.x0 = z.x + ofx * CHUNK;
            P.y0 = z.y - w->y_off + ofy * CHUNK;
            P.ncx = (nx - ofx + 1) / 2;
            P.ncy = (ny - ofy + 1) / 2;
            P.iter = iter;
            P.rkey = rng_key(a->seed, a->tick, (uint32_t)iter);
            P.tick = a->tick;
            P.pbuf = w->pbuf;
            P.pcount = w->pcount;
            P.pcap = w->pcap;
            P.tabs = w->ctx->d_tabs;
            P.chunk_list = nullptr;
            P.list_count = nullptr;
            P.awake = nullptr;
            P.acols = w->acols;
            P.arows = w->arows;
            P.never_sleep = 0;
            P.schedule = w->schedule;
            P.dbg = w->d_dbg;
            P.fused = w->fused;
            P.fused_max_chunks = w->fused_max_chunks;
            P.chunk_base = 0;
            P.chunk_cost = nullptr;
            P.chunk_state = nullptr;
            P.rowmask = nullptr;
            P.flowx = w->d_flow;
            P.flowy = w->d_flow ? w->d_flow + (size_t)w->W * w->H : nullptr;
            P.phase_rows = nullptr;
            P.lpt_parts = 0;
            const int ph = iter * 4 + tk;
            bool skip_here = w->rowskip_mode == 1;
            if (w->rowskip_mode == 2) skip_here = gate_probe || w->phase_active[ph] < 0.0f || w->phase_active[ph] < w->rowskip_max_active;
            // without row skipping pass 1 still sorts the rows for pass 2 (liquid-only rows are applied by a row-parallel kernel, pass 2
            // steps only rows with powder or gas: tick_pass2_apply_kernel); not with active-chunk tracking, whose pass 2 walks every row
            // Measured on the 8192^2 mixed world: phases in which powder still acts lose 3-8 % to the split (the row-skipping instantiation of
            // pass 2 and the extra launch cost more than the skipped rows save), phases in which only liquid acts gain 10 %.  So the split
            // is used where the last classification of the phase (the skip gate's probe ticks) found live powder or gas in fewer than
            // p2_split_max_seq of the rows (FSE_P2_SPLIT = 0 never, 1 = this rule, 2 = always).  Correct in any phase either way.
            const bool split_here = !skip_here && !w->active_on &&
                                    (w->p2_split == 2 || (w->p2_split == 1 && w->phase_seq[ph] >= 0.0f && w->phase_seq[ph] < w->p2_split_max_seq));
            P.split = split_here ? 1 : 0;
            if ((skip_here || split_here) && w->schedule == FSE_SCHEDULE_ROWS && !w->fused) {
                const int need = ((nx + 1) / 2) * ((ny + 1) / 2);
                if (need > w->rowmask_cap) {
                    CK(cudaStreamSynchronize(w->stream));
                    cudaFree(w->d_rowmask);
                    w->d_rowmask = nullptr;
                    w->rowmask_cap = 0;
                    CK(cudaMalloc((void**)&w->d_rowmask, sizeof(uint32_t) * ROWMASK_WORDS * (size_t)need));
                    w->rowmask_cap = need;
                }
                P.rowmask = w->d_rowmask;
                if (gate_collect && skip_here) {
                    P.phase_rows = w->d_phase_rows + ph;
                    w->phase_rows_counted |= 1u << ph;
                    // rows the classification covers: the chunks of this colour this rank launches on the per-pass kernels
                    long long chunks = (long long)((nx - ofx + 1) / 2) * ((ny - ofy + 1) / 2);
                    if (w->strip) chunks = (long long)w->list_cnt[tk][0] + w->list_cnt[tk][1];
                    w->phase_rows_total[ph] = chunks * CHUNK;
                }
            }
            if (!w->strip) {
                const int n_chunks = P.ncx * P.ncy;
                if (n_chunks <= 0) continue;
                if (w->active_on && z.x % CHUNK == 0 && z.y % CHUNK == 0) {
                    CK(launch_compact_active(w->d_awake, w->acols, P.x0 / CHUNK, P.y0 / CHUNK, P.ncx, P.ncy, w->d_active_list,
                                             w->d_active_count, w->stream));
                    w->ctx->launches += 1;
                    P.chunk_list = w->d_active_list;
                    P.list_count = w->d_active_count;
                    P.awake = w->d_awake;
                    // per-pass kernels record what they saw per chunk and a small kernel updates the flags after pass 3;
                    // FSE_ACTIVE_FUSED=1 keeps the single fused kernel (shorter chain when only a few chunks are awake)
                    if (w->schedule == FSE_SCHEDULE_ROWS && !w->fused && !w->active_fused) P.chunk_state = w->d_chunk_state;
                }
                // longest-first order from the cycles pass 1 took on each chunk of this colour the last time (per-pass kernels)
                const bool lpt = w->lpt_on && !P.chunk_list && w->schedule == FSE_SCHEDULE_ROWS && !w->fused && n_chunks >= w->fork.min_chunks;
                if (lpt) {
                    if (n_chunks > w->lpt_cap) {
                        cudaFree(w->d_lpt_cost);
                        cudaFree(w->d_lpt_list);
                        w->d_lpt_cost = nullptr;
                        w->d_lpt_list = nullptr;
                        w->lpt_cap = 0;
                        CK(cudaMalloc((void**)&w->d_lpt_cost, sizeof(unsigned int) * 4 * (size_t)n_chunks));
                        CK(cudaMalloc((void**)&w->d_lpt_list, sizeof(int) * 4 * (size_t)n_chunks));
                        w->lpt_cap = n_chunks;
                        memset(w->lpt_sig, 0, sizeof(w->lpt_sig));
                    }
                    const int sig[4] = {P.x0, P.y0, P.ncx, P.ncy};
                    if (memcmp(sig, w->lpt_sig[tk], sizeof(sig)) == 0) {
                        P.chunk_list = w->d_lpt_list + (size_t)tk * w->lpt_cap;
                        P.lpt_parts = w->lpt_parts[tk];
                    }
                    P.chunk_cost = w->d_lpt_cost + (size_t)tk * w->lpt_cap;
                }
                if (int r = kt.begin(w->stream)) return r;
                int nl = 0;
                CK(launch_tick_phase(P, n_chunks, w->stream, &nl, &w->fork));
                if (lpt) {
                    const int deal = (w->lpt_deal && n_chunks >= w->fork.min_chunks) ? w->fork.parts : 1;
                    CK(launch_lpt_build(P.chunk_cost, n_chunks, P.ncx, w->d_lpt_list + (size_t)tk * w->lpt_cap, w->stream, nullptr, deal));
                    w->lpt_parts[tk] = deal;
                    nl += 1;
                    const int sig[4] = {P.x0, P.y0, P.ncx, P.ncy};
                    memcpy(w->lpt_sig[tk], sig, sizeof(sig));
                }
                if (int r = kt.end(w->stream)) return r;
                w->ctx->launches += nl;
                continue;
            }
            // strip: boundary chunk rows first, halo rows over NCCL on the side stream, interior chunk rows meanwhile
            P.y0 = z.y - w->y_off;  // chunk lists carry absolute (cxi, cyi) pairs: cy = y0 + cyi*256 (+128 via list parity)
            P.x0 = z.x + ofx * CHUNK;
            P.y0 += ofy * CHUNK;
            if (int r = kt.begin(w->stream)) return r;
            // boundary chunks and the halo exchange run on the (high-priority) side stream, next to the interior chunks: all
            // chunks of a colour phase are independent, only the next phase needs both
            CK(cudaEventRecord(w->ev_boundary, w->stream));
            CK(cudaStreamWaitEvent(w->comm_stream, w->ev_boundary, 0));
            cudaEvent_t* tl = nullptr;  // FSE_STRIP_TIMELINE=1: phase start | boundary chunks done | halo exchange done | interior chunks done
            if (w->timeline_on) {
                if (w->timeline_used + 4 > w->timeline_ev.size())
                    for (int q = 0; q < 4; q++) {
                        cudaEvent_t e_;
                        CK(cudaEventCreate(&e_));
                        w->timeline_ev.push_back(e_);
                    }
                tl = &w->timeline_ev[w->timeline_used];
                w->timeline_used += 4;
                CK(cudaEventRecord(tl[0], w->stream));
            }
            if (w->list_cnt[tk][0] > 0) {
                P.chunk_list = w->d_chunk_lists + w->list_off[tk][0];
                int nl = 0;
                // A full chunk row of cut-adjacent chunks runs on the per-pass kernels like the interior (FSE_STRIP_BOUNDARY_MIN, default
                // 64 chunks): next to the interior's CTAs a fused launch (86 KB, 320 threads per chunk) took 1.16 ms of a 1.75 ms phase
                // at 8 GPUs and slowed the interior down; short rows (small worlds) keep the fused kernel's shorter chain.
                TickParams B = P;
                if (w->list_cnt[tk][0] >= w->strip_boundary_min) B.fused_max_chunks = 0;
                CK(launch_tick_phase(B, w->list_cnt[tk][0], w->comm_stream, &nl, nullptr));
                w->ctx->launches += nl;
            }
            if (tl) CK(cudaEventRecord(tl[1], w->comm_stream));
            if (multi) {
                if (int r = strip_exchange(w, ofy, j0, j1, z.y - w->y_off, w->comm_stream)) {
                    // the boundary kernels are already enqueued on the side stream: the next call must not race with them
                    cudaEventRecord(w->ev_comm, w->comm_stream);
                    cudaStreamWaitEvent(w->stream, w->ev_comm, 0);
                    return r;
                }
            }
            CK(cudaEventRecord(w->ev_comm, w->comm_stream));
            if (tl) CK(cudaEventRecord(tl[2], w->comm_stream));
            if (w->list_cnt[tk][1] > 0) {
                const int n_int = w->list_cnt[tk][1], n_grid = P.ncx * P.ncy;
                const int* base_list = w->d_chunk_lists + w->list_off[tk][1];
                P.chunk_list = base_list;
                // longest-first order of the interior chunks, as on plain worlds: the cost array covers the colour's whole chunk grid,
                // the sorted list only this rank's interior chunks
                const bool lpt = w->lpt_on && w->schedule == FSE_SCHEDULE_ROWS && !w->fused && n_int >= w->fork.min_chunks && n_int > w->fused_max_chunks;
                if (lpt) {
                    if (n_grid > w->lpt_cap) {
                        cudaFree(w->d_lpt_cost);
                        cudaFree(w->d_lpt_list);
                        w->d_lpt_cost = nullptr;
                        w->d_lpt_list = nullptr;
                        w->lpt_cap = 0;
                        CK(cudaMalloc((void**)&w->d_lpt_cost, sizeof(unsigned int) * 4 * (size_t)n_grid));
                        CK(cudaMalloc((void**)&w->d_lpt_list, sizeof(int) * 4 * (size_t)n_grid));
                        CK(cudaMemsetAsync(w->d_lpt_cost, 0, sizeof(unsigned int) * 4 * (size_t)n_grid, w->stream));
                        w->lpt_cap = n_grid;
                        memset(w->lpt_sig, 0, sizeof(w->lpt_sig));
                    }
                    const int sig[4] = {P.x0 ^ (w->list_off[tk][1] << 8), P.y0, P.ncx, n_int};
                    if (memcmp(sig, w->lpt_sig[tk], sizeof(sig)) == 0) {
                        P.chunk_list = w->d_lpt_list + (size_t)tk * w->lpt_cap;
                        P.lpt_parts = w->lpt_parts[tk];
                    }
                    P.chunk_cost = w->d_lpt_cost + (size_t)tk * w->lpt_cap;
                }
                int nl = 0;
                CK(launch_tick_phase(P, n_int, w->stream, &nl, &w->fork));
                if (lpt) {
                    const int deal = (w->lpt_deal && n_int >= w->fork.min_chunks) ? w->fork.parts : 1;
                    CK(launch_lpt_build(P.chunk_cost, n_int, P.ncx, w->d_lpt_list + (size_t)tk * w->lpt_cap, w->stream, base_list, deal));
                    w->lpt_parts[tk] = deal;
                    nl += 1;
                    const int sig[4] = {P.x0 ^ (w->list_off[tk][1] << 8), P.y0, P.ncx, n_int};
                    memcpy(w->lpt_sig[tk], sig, sizeof(sig));
                    P.chunk_cost = nullptr;
                }
                w->ctx->launches += nl;
            }
            if (int r = kt.end(w->stream)) return r;
            if (tl) CK(cudaEventRecord(tl[3], w->stream));
            CK(cudaStreamWaitEvent(w->stream, w->ev_comm, 0));
        }
    }
    if (gate_collect && w->phase_rows_counted) {
        CK(cudaMemcpyAsync(w->h_phase_rows, w->d_phase_rows, 32 * sizeof(unsigned int), cudaMemcpyDeviceToHost, w->stream));
        CK(cudaEventRecord(w->ev_phase_rows, w->stream));
        w->phase_rows_valid = w->phase_rows_counted;
        w->phase_rows_pending = true;
    }
    w->ticks++;
    return FSE_OK;
}

FSE_API int fse_tick_temperature(fse_world* w, const fse_rect* zg) {
    if (!w || !zg) return fail(FSE_EINVAL, "fse_tick_temperature: null argument");
    fse_rect zz = *zg;
    if (w->strip) {  // own rows only; ghost rows are refreshed afterwards (owner-authoritative)
        const int lo = zz.y > w->own_lo ? zz.y : w->own_lo, hi = zz.y + zz.h < w->own_hi ? zz.y + zz.h : w->own_hi;
        zz.y = lo;
        zz.h = hi - lo;
    }
    const fse_rect* z = &zz;
    if (z->w <= 0 || z->h <= 0 || z->x < 1 || z->y - w->y_off < 1 || z->x + z->w + 1 > w->W || z->y - w->y_off + z->h + 1 > w->H)
        return fail(FSE_EINVAL, "fse_tick_temperature: zone (%d,%d,%d,%d) needs a 1-cell margin inside the world", z->x, z->y, z->w, z->h);
    CK(cudaSetDevice(w->ctx->device));
    CK(launch_temperature(w->p, w->tmp_scratch, w->W, w->H, z->x, z->y - w->y_off, z->w, z->h, w->ctx->d_tabs, w->active_on ? w->d_awake : nullptr,
                          w->acols, w->y_off, w->stream));
    std::swap(w->p.tmp, w->tmp_scratch);  // the new plane is the temperature plane from here on (later launches read w->p on this stream)
    w->ctx->launches += 2;
    if (w->strip && w->ctx->nranks > 1) return strip_refresh(w, w->stream);
    return FSE_OK;
}

// ---- particles (container part; the integrator lives in fse_particles.cu) -------------------------------
FSE_API int fse_particles_count(fse_world* w, int64_t* out) {
    if (!w || !out) return fail(FSE_EINVAL, "fse_particles_count: null argument");
    CK(cudaSetDevice(w->ctx->device));
    unsigned int n = 0;
    CK(cudaMemcpyAsync(&n, w->pcount, sizeof n, cudaMemcpyDeviceToHost, w->stream));
    CK(cudaStreamSynchronize(w->stream));
    if (n > w->pcap) {  // the kernels counted particles they had no slot for: the pool holds pcap, the rest is reported as dropped
        w->particles_dropped += n - w->pcap;
        n = w->pcap;
        CK(cudaMemcpyAsync(w->pcount, &n, sizeof n, cudaMemcpyHostToDevice, w->stream));
        CK(cudaStreamSynchronize(w->stream));
    }
    w->particles_seen = n;
    *out = n;
    return FSE_OK;
}

// Cumulative number of particles the kernels could not store (pool full).  The pool grows on its own (particles_headroom), so
// this stays 0 unless one call spawns more than the head-room it was given.
FSE_API int fse_particles_dropped(fse_world* w, int64_t* out) {
    if (!w || !out) return fail(FSE_EINVAL, "fse_particles_dropped: null argument");
    int64_t n = 0;
    if (int r = fse_particles_count(w, &n)) return r;
    *out = (int64_t)w->particles_dropped;
    return FSE_OK;
}



FSE_API int fse_particles_read(fse_world* w, fse_particle* out, int64_t cap, int64_t* n_out) {
    if (!w || !n_out) return fail(FSE_EINVAL, "fse_particles_read: null argument");
    int64_t n = 0;
    if (int r = fse_particles_count(w, &n)) return r;
    int64_t m = n < cap ? n : cap;
    if (m > 0 && out) CK(cudaMemcpy(out, w->pbuf, sizeof(fse_particle) * (size_t)m, cudaMemcpyDeviceToHost));
    *n_out = m;
    return FSE_OK;
}

FSE_API int fse_particles_clear(fse_world* w) {
    if (!w) return fail(FSE_EINVAL, "fse_particles_clear: null world");
    CK(cudaSetDevice(w->ctx->device));
    CK(cudaMemsetAsync(w->pcount, 0, sizeof(unsigned int), w->stream));
    return FSE_OK;
}

FSE_API int fse_particles_reserve(fse_world* w, int64_t cap) {
    if (!w || cap < 1 || cap > ((int64_t)1 << 30)) return fail(FSE_EINVAL, "fse_particles_reserve: bad capacity");
    CK(cudaSetDevice(w->ctx->device));
    int64_t n = 0;
    if (int r = fse_particles_count(w, &n)) return r;
    if (cap < n) return fail(FSE_EINVAL, "fse_particles_reserve: %lld live particles", (long long)n);
    fse_particle* nb = nullptr;
    CK(cudaMalloc(&nb, sizeof(fse_particle) * (size_t)cap));
    if (n) CK(cudaMemcpy(nb, w->pbuf, sizeof(fse_particle) * (size_t)n, cudaMemcpyDeviceToDevice));
    cudaFree(w->pbuf);
    cudaFree(w->pbuf2);
    w->pbuf2 = nullptr;
    w->pbuf2_bytes = 0;
    w->pbuf = nb;
    w->pcap = (unsigned int)cap;
    return FSE_OK;
}

FSE_API int fse_particles_add(fse_world* w, const fse_particle* p, int32_t n) {
    if (!w || (!p && n > 0) || n < 0) return fail(FSE_EINVAL, "fse_particles_add: bad argument");
    if (n == 0) return FSE_OK;
    CK(cudaSetDevice(w->ctx->device));
    int64_t have = 0;
    if (int r = fse_particles_count(w, &have)) return r;
    if (have + n > (int64_t)w->pcap) return fail(FSE_ENOMEM, "fse_particles_add: pool capacity %u exceeded", w->pcap);
    std::vector<fse_particle> tmp(p, p + n);
    for (int i = 0; i < n; i++) {
        if (tmp[i].tile.mat >= w->ctx->h_tabs.n) return fail(FSE_EINVAL, "fse_particles_add: particle %d material out of range", i);
        if (tmp[i].id == 0) tmp[i].id = (1ULL << 62) | (w->next_user_particle++);  // id bits 63..62: 00 tick kernels, 01 caller, 10 bridge, 11 explosion
    }
    CK(cudaMemcpy(w->pbuf + have, tmp.data(), sizeof(fse_particle) * (size_t)n, cudaMemcpyHostToDevice));
    unsigned int total = (unsigned int)(have + n);
    CK(cudaMemcpy(w->pcount, &total, sizeof total, cudaMemcpyHostToDevice));
    return FSE_OK;
}

// ---- measurement ----------------------------------------------------------------------------------------
FSE_API int fse_timer_start(fse_world* w) {
    if (!w) return fail(FSE_EINVAL, "fse_timer_start: null world");
    CK(cudaSetDevice(w->ctx->device));
    CK(cudaEventRecord(w->ev0, w->stream));
    return FSE_OK;
}

FSE_API int fse_timer_stop(fse_world* w, float* ms) {
    if (!w || !ms) return fail(FSE_EINVAL, "fse_timer_stop: null argument");
    CK(cudaSetDevice(w->ctx->device));
    CK(cudaEventRecord(w->ev1, w->stream));
    CK(cudaEventSynchronize(w->ev1));
    CK(cudaEventElapsedTime(ms, w->ev0, w->ev1));
    return FSE_OK;
}

FSE_API int fse_kernel_timing_enable(fse_world* w, int enable) {
    if (!w) return fail(FSE_EINVAL, "fse_kernel_timing_enable: null world");
    w->kt_enabled = enable != 0;
    w->kt_used = 0;
    return FSE_OK;
}

// each timed colour phase on its own, in launch order (12 per tick at cell_iter = 3); does not reset the record
FSE_API int fse_kernel_timing_phases(fse_world* w, float* out_ms, int64_t cap, int64_t* n_out) {
    if (!w || !n_out || (!out_ms && cap > 0)) return fail(FSE_EINVAL, "fse_kernel_timing_phases: null argument");
    CK(cudaSetDevice(w->ctx->device));
    CK(cudaStreamSynchronize(w->stream));
    int64_t n = 0;
    for (size_t i = 0; i < w->kt_used && n < cap; i++, n++) CK(cudaEventElapsedTime(&out_ms[i], w->kt_events[i].first, w->kt_events[i].second));
    *n_out = n;
    return FSE_OK;
}

// Strip worlds, FSE_STRIP_TIMELINE=1 at creation: per colour phase since the last read, the times (ms after the phase started on the
// main stream) at which the cut-adjacent chunks finished, the halo exchange finished (both on the side stream) and the interior
// chunks finished (main stream).  out = n x 3 floats.  The stand-in for an nsys timeline (nsys is not in this image).
FSE_API int fse_strip_timeline_read(fse_world* w, float* out, int64_t cap_phases, int64_t* n_out) {
    if (!w || !n_out || (!out && cap_phases > 0)) return fail(FSE_EINVAL, "fse_strip_timeline_read: null argument");
    CK(cudaSetDevice(w->ctx->device));
    CK(cudaStreamSynchronize(w->stream));
    CK(cudaStreamSynchronize(w->comm_stream));
    int64_t n = 0;
    for (size_t i = 0; i + 3 < w->timeline_used && n < cap_phases; i += 4, n++)
        for (int q = 0; q < 3; q++) CK(cudaEventElapsedTime(&out[3 * n + q], w->timeline_ev[i], w->timeline_ev[i + 1 + q]));
    *n_out = n;
    w->timeline_used = 0;
    return FSE_OK;
}

FSE_API int fse_kernel_timing_read(fse_world* w, double* total_ms, int64_t* launches) {
    if (!w || !total_ms || !launches) return fail(FSE_EINVAL, "fse_kernel_timing_read: null argument");
    CK(cudaSetDevice(w->ctx->device));
    CK(cudaStreamSynchronize(w->stream));
    double tot = 0;
    for (size_t i = 0; i < w->kt_used; i++) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, w->kt_events[i].first, w->kt_events[i].second));
        tot += ms;
    }
    *total_ms = tot;
    *launches = (int64_t)w->kt_used;
    w->kt_used = 0;
    return FSE_OK;
}

}  // extern "C"

namespace fse {
// Rects are in global coordinates; a strip world holds rows [y_off, y_off + H).
int check_rect(fse_world* w, int x, int y, int rw, int rh, const char* who) {
    if (rw <= 0 || rh <= 0 || x < 0 || y < w->y_off || (int64_t)x + rw > w->W || (int64_t)y + rh > (int64_t)w->y_off + w->H)
        return fail(FSE_EINVAL, "%s: rect (%d,%d,%d,%d) outside the held rows [%d,%d) of the %dx%d world", who, x, y, rw, rh, w->y_off,
                    w->y_off + w->H, w->W, w->Hglobal);
    return FSE_OK;
}

int ensure_stage(fse_world* w, size_t cells) {
    if (w->stage_cells >= cells) return FSE_OK;
    cudaFree(w->d_stage);
    w->d_stage = nullptr;
    w->stage_cells = 0;
    CK(cudaMalloc(&w->d_stage, cells * sizeof(fse_cell)));
    w->stage_cells = cells;
    return FSE_OK;
}
}  // namespace fse


namespace fse {
// Make room for `need` more particles before a call that spawns them.  exact: read the live count (one stream sync; rare calls
// such as fse_explosion); otherwise use the count seen by the last call that read it plus what was promised since.
int particles_headroom(fse_world* w, int64_t need, bool exact) {
    int64_t have = (int64_t)w->particles_seen + w->particles_promised;
    if (exact || have + need > (int64_t)w->pcap) {
        int64_t n = 0;
        if (int r = fse_particles_count(w, &n)) return r;
        w->particles_promised = 0;
        have = n;
    }
    if (have + need > (int64_t)w->pcap) {
        int64_t cap = (int64_t)w->pcap * 2;
        while (cap < have + need) cap *= 2;
        if (cap > ((int64_t)1 << 30)) cap = (int64_t)1 << 30;
        if (cap > (int64_t)w->pcap)
            if (int r = fse_particles_reserve(w, cap)) return r;
    }
    w->particles_promised += need;
    return FSE_OK;
}
}  // namespace fse

namespace fse {
int fse_wake_rect(fse_world* w, int x, int y_local, int rw, int rh) { return ::wake_rect(w, x, y_local, rw, rh); }
}  // namespace fse
