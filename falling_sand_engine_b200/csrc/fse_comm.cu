// fse_comm.cu — multi-GPU: horizontal strips + NCCL halo-row exchange over NVLink.
//
// The reference is single-process; this has no reference counterpart beyond "same result as one grid".  The world is
// cut at chunk-row boundaries into one strip per rank.  Chunk rows on either side of a cut have opposite colour parity
// (world.cpp:1060), so they are never processed in the same colour phase; after a phase, the rank whose boundary chunk
// row just ran owns the freshest copy of the rows around the cut and sends them to its neighbour:
//     upper rank ran its last chunk row   -> sends rows [cut-5, cut+5)   down   (it can write 5 rows below the cut)
//     lower rank ran its first chunk row  -> sends rows [cut-5, cut+10)  up     (the upper rank reads 10 rows below)
// One direction per cut per phase, 7 planes each, grouped in one ncclGroup on a side stream; the boundary chunk rows are
// launched first and the interior chunk rows overlap the transfer.  The schedule is the global one, so the result is
// bit-identical for any number of strips (tests/test_strips_*.py).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process, e.g. torch's) so the library has no
// link-time NCCL dependency and never mixes two NCCL builds in one process.
#include <dlfcn.h>

#include "fse_internal.hpp"

namespace fse {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8_t_ = 1 };

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.h) return FSE_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) return fail(FSE_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
#define SYM(f, name)                                                          \
    *(void**)(&g_nccl.f) = dlsym(h, name);                                    \
    if (!g_nccl.f) return fail(FSE_ENCCL, "libnccl.so.2 lacks %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.h = h;
    return FSE_OK;
}

#define NK(call)                                                                                           \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != 0) return fail(FSE_ENCCL, "%s: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); \
    } while (0)
#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) return fail(FSE_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

// Send or receive local rows [y_lo, y_hi) of all seven planes to/from `peer` (inside an open ncclGroup).
static int xfer_rows(fse_world* w, bool send, int peer, int y_lo, int y_hi, cudaStream_t s) {
    ncclComm_t comm = (ncclComm_t)w->ctx->nccl_comm;
    const size_t off = (size_t)y_lo * w->W, cnt = (size_t)(y_hi - y_lo) * w->W;
    struct { void* base; size_t es; } pl[7] = {{w->p.mat, 1}, {w->p.flg, 1}, {w->p.stl, 1}, {w->p.tmp, 2}, {w->p.col, 4}, {w->p.fl, 4}, {w->p.fd, 4}};
    for (int i = 0; i < 7; i++) {
        char* ptr = (char*)pl[i].base + off * pl[i].es;
        if (send) NK(g_nccl.Send(ptr, cnt * pl[i].es, ncclUint8_t_, peer, comm, s));
        else NK(g_nccl.Recv(ptr, cnt * pl[i].es, ncclUint8_t_, peer, comm, s));
    }
    return FSE_OK;
}

// Exchange after the boundary chunk rows of colour-row parity `ofy` ran.  j0/j1: owned chunk rows (zone-relative),
// cut rows are local y coordinates.  Called with comm_stream already waiting on the boundary launch.
int strip_exchange(fse_world* w, int ofy, int j0, int j1, int zone_y_local, cudaStream_t s) {
    fse_ctx* c = w->ctx;
    const bool up = c->rank > 0, down = c->rank + 1 < c->nranks;
    if (!up && !down) return FSE_OK;
    NK(g_nccl.GroupStart());
    if (up) {  // cut above my first owned chunk row j0; the rank above owns j0-1
        const int cut = zone_y_local + j0 * CHUNK;
        const bool mine_ran = (j0 % 2) == ofy;
        if (mine_ran) { if (int r = xfer_rows(w, true, c->rank - 1, cut - 5, cut + 10, s)) return r; }
        else          { if (int r = xfer_rows(w, false, c->rank - 1, cut - 5, cut + 5, s)) return r; }
    }
    if (down) {  // cut below my last owned chunk row j1-1; the rank below owns j1
        const int cut = zone_y_local + j1 * CHUNK;
        const bool mine_ran = ((j1 - 1) % 2) == ofy;
        if (mine_ran) { if (int r = xfer_rows(w, true, c->rank + 1, cut - 5, cut + 5, s)) return r; }
        else          { if (int r = xfer_rows(w, false, c->rank + 1, cut - 5, cut + 10, s)) return r; }
    }
    NK(g_nccl.GroupEnd());
    return FSE_OK;
}

// Owner-authoritative refresh: every rank sends the 16 owned rows next to each cut and receives its neighbour's.
int strip_refresh(fse_world* w, cudaStream_t s) {
    fse_ctx* c = w->ctx;
    if (c->nranks == 1) return FSE_OK;
    const int R = 16;
    NK(g_nccl.GroupStart());
    if (c->rank > 0) {
        const int cut = w->own_lo - w->y_off;
        if (int r = xfer_rows(w, true, c->rank - 1, cut, cut + R, s)) return r;
        if (int r = xfer_rows(w, false, c->rank - 1, cut - R, cut, s)) return r;
    }
    if (c->rank + 1 < c->nranks) {
        const int cut = w->own_hi - w->y_off;
        if (int r = xfer_rows(w, true, c->rank + 1, cut - R, cut, s)) return r;
        if (int r = xfer_rows(w, false, c->rank + 1, cut, cut + R, s)) return r;
    }
    NK(g_nccl.GroupEnd());
    return FSE_OK;
}

}  // namespace fse

using namespace fse;

extern "C" {

// 128-byte ncclUniqueId for rank 0 to broadcast (any out-of-band channel: torch.distributed, MPI, a file ...).
FSE_API int fse_comm_unique_id(void* out128) {
    if (!out128) return fail(FSE_EINVAL, "fse_comm_unique_id: null");
    if (int r = load_nccl()) return r;
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof id);
    return FSE_OK;
}

FSE_API int fse_comm_init(fse_ctx* c, int rank, int nranks, const void* id128) {
    if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(FSE_EINVAL, "fse_comm_init: bad argument");
    if (int r = load_nccl()) return r;
    CK(cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    NK(g_nccl.CommInitRank(&comm, nranks, id, rank));
    c->nccl_comm = comm;
    c->rank = rank;
    c->nranks = nranks;
    return FSE_OK;
}

FSE_API int fse_comm_destroy(fse_ctx* c) {
    if (!c) return fail(FSE_EINVAL, "fse_comm_destroy: null");
    if (c->nccl_comm) {
        g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
        c->nccl_comm = nullptr;
    }
    c->rank = 0;
    c->nranks = 1;
    return FSE_OK;
}

}  // extern "C"
