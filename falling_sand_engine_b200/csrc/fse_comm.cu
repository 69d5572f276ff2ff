// fse_comm.cu — multi-GPU: horizontal strips + NCCL halo-row exchange over NVLink.
//
// The reference is single-process; this has no reference counterpart beyond "same result as one grid".  The world is
// cut at chunk-row boundaries into one strip per rank.  Chunk rows on either side of a cut have opposite colour parity
// (world.cpp:1060), so they are never processed in the same colour phase; after a phase, the rank whose boundary chunk
// row just ran owns the freshest copy of the rows around the cut and sends them to its neighbour:
//     upper rank ran its last chunk row   -> sends rows [cut-5, cut+5)   down   (it can write 5 rows below the cut)
//     lower rank ran its first chunk row  -> sends rows [cut-5, cut+10)  up     (the upper rank reads 10 rows below)
// One direction per cut per phase, the seven planes packed into one message, both cuts in one ncclGroup on a side stream; the
// boundary chunk rows are launched first and the interior chunk rows overlap the transfer.  The schedule is the global one, so the result is
// bit-identical for any number of strips (tests/test_strips_*.py).
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 already in the process, e.g. torch's) so the library has no
// link-time NCCL dependency and never mixes two NCCL builds in one process.
#include <dlfcn.h>

#include <algorithm>
#include <vector>

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
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
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
    SYM(AllReduce, "ncclAllReduce")
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

// ---- packed halo messages --------------------------------------------------------------------------------------------------------
// The rows [y_lo, y_hi) of the seven planes travel as ONE message per cut and direction: a pack kernel gathers them into a staging
// buffer (plane after plane, 17 bytes per cell), one ncclSend / ncclRecv moves it, an unpack kernel scatters it on the other side.
// Full-width rows are contiguous in every plane, so both kernels are straight 16-byte copies.
struct PackArgs {
    const unsigned char* src[7];
    unsigned char* dst[7];
    size_t bytes[7];  // per plane, multiples of 16 (W is a multiple of 128)
};
__global__ void halo_copy_kernel(const PackArgs a) {
    const int pl = blockIdx.y;
    const uint4* s = reinterpret_cast<const uint4*>(a.src[pl]);
    uint4* d = reinterpret_cast<uint4*>(a.dst[pl]);
    const size_t n = a.bytes[pl] / 16;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}
static size_t halo_bytes(const fse_world* w, int rows) { return (size_t)rows * w->W * 17; }
static void halo_args(fse_world* w, unsigned char* stage, int y_lo, int y_hi, bool pack, PackArgs* a) {
    const size_t off = (size_t)y_lo * w->W, cnt = (size_t)(y_hi - y_lo) * w->W;
    struct { void* base; size_t es; } pl[7] = {{w->p.mat, 1}, {w->p.flg, 1}, {w->p.stl, 1}, {w->p.tmp, 2}, {w->p.col, 4}, {w->p.fl, 4}, {w->p.fd, 4}};
    size_t so = 0;
    for (int i = 0; i < 7; i++) {
        unsigned char* g = (unsigned char*)pl[i].base + off * pl[i].es;
        a->src[i] = pack ? g : stage + so;
        a->dst[i] = pack ? stage + so : g;
        a->bytes[i] = cnt * pl[i].es;
        so += cnt * pl[i].es;
    }
}

// One side of a cut: what this rank sends and/or receives in one exchange (rows are local y coordinates; lo == hi: nothing)
struct CutXfer {
    int peer;
    int send_lo, send_hi, recv_lo, recv_hi;
    int slot;  // staging buffers 2 * slot (send), 2 * slot + 1 (recv)
};

static int ensure_stage(fse_world* w, int idx, size_t bytes) {
    if (w->halo_stage_bytes[idx] >= bytes) return FSE_OK;
    CK(cudaStreamSynchronize(w->comm_stream));
    CK(cudaStreamSynchronize(w->stream));
    cudaFree(w->halo_stage[idx]);
    w->halo_stage[idx] = nullptr;
    w->halo_stage_bytes[idx] = 0;
    CK(cudaMalloc(&w->halo_stage[idx], bytes));
    w->halo_stage_bytes[idx] = bytes;
    return FSE_OK;
}

// pack -> one grouped send/recv per cut -> unpack, all on stream s.  The NCCL group is closed on every path.
static int exchange_cuts(fse_world* w, const CutXfer* cuts, int n_cuts, cudaStream_t s) {
    ncclComm_t comm = (ncclComm_t)w->ctx->nccl_comm;
    for (int i = 0; i < n_cuts; i++) {
        const CutXfer& c = cuts[i];
        if (c.send_hi > c.send_lo) {
            if (int r = ensure_stage(w, 2 * c.slot, halo_bytes(w, c.send_hi - c.send_lo))) return r;
            PackArgs a;
            halo_args(w, (unsigned char*)w->halo_stage[2 * c.slot], c.send_lo, c.send_hi, true, &a);
            halo_copy_kernel<<<dim3(64, 7), 256, 0, s>>>(a);
            CK(cudaGetLastError());
            w->ctx->launches += 1;
        }
        if (c.recv_hi > c.recv_lo)
            if (int r = ensure_stage(w, 2 * c.slot + 1, halo_bytes(w, c.recv_hi - c.recv_lo))) return r;
    }
    NK(g_nccl.GroupStart());
    ncclResult_t bad = 0;
    for (int i = 0; i < n_cuts && !bad; i++) {
        const CutXfer& c = cuts[i];
        if (c.send_hi > c.send_lo) bad = g_nccl.Send(w->halo_stage[2 * c.slot], halo_bytes(w, c.send_hi - c.send_lo), ncclUint8_t_, c.peer, comm, s);
        if (!bad && c.recv_hi > c.recv_lo) bad = g_nccl.Recv(w->halo_stage[2 * c.slot + 1], halo_bytes(w, c.recv_hi - c.recv_lo), ncclUint8_t_, c.peer, comm, s);
    }
    const ncclResult_t end = g_nccl.GroupEnd();  // always: an open group would swallow every later NCCL call of this thread
    if (bad) return fail(FSE_ENCCL, "ncclSend/ncclRecv: %s", g_nccl.GetErrorString(bad));
    if (end) return fail(FSE_ENCCL, "ncclGroupEnd: %s", g_nccl.GetErrorString(end));
    for (int i = 0; i < n_cuts; i++) {
        const CutXfer& c = cuts[i];
        if (c.recv_hi > c.recv_lo) {
            PackArgs a;
            halo_args(w, (unsigned char*)w->halo_stage[2 * c.slot + 1], c.recv_lo, c.recv_hi, false, &a);
            halo_copy_kernel<<<dim3(64, 7), 256, 0, s>>>(a);
            CK(cudaGetLastError());
            w->ctx->launches += 1;
        }
    }
    return FSE_OK;
}

// Exchange after the boundary chunk rows of colour-row parity `ofy` ran.  j0/j1: owned chunk rows (zone-relative),
// cut rows are local y coordinates.  Called with comm_stream already waiting on the boundary launch.
int strip_exchange(fse_world* w, int ofy, int j0, int j1, int zone_y_local, cudaStream_t s) {
    fse_ctx* c = w->ctx;
    const bool up = c->rank > 0, down = c->rank + 1 < c->nranks;
    if (!up && !down) return FSE_OK;
    CutXfer cuts[2];
    int n = 0;
    if (up) {  // cut above my first owned chunk row j0; the rank above owns j0-1
        const int cut = zone_y_local + j0 * CHUNK;
        const bool mine_ran = (j0 % 2) == ofy;
        cuts[n++] = mine_ran ? CutXfer{c->rank - 1, cut - 5, cut + 10, 0, 0, 0} : CutXfer{c->rank - 1, 0, 0, cut - 5, cut + 5, 0};
    }
    if (down) {  // cut below my last owned chunk row j1-1; the rank below owns j1
        const int cut = zone_y_local + j1 * CHUNK;
        const bool mine_ran = ((j1 - 1) % 2) == ofy;
        cuts[n++] = mine_ran ? CutXfer{c->rank + 1, cut - 5, cut + 5, 0, 0, 1} : CutXfer{c->rank + 1, 0, 0, cut - 5, cut + 10, 1};
    }
    return exchange_cuts(w, cuts, n, s);
}

// Owner-authoritative refresh: every rank sends the R owned rows next to each cut and receives its neighbour's (R <= ghost rows).
int strip_refresh(fse_world* w, cudaStream_t s, int R) {
    fse_ctx* c = w->ctx;
    if (c->nranks == 1) return FSE_OK;
    CutXfer cuts[2];
    int n = 0;
    if (c->rank > 0) {
        const int cut = w->own_lo - w->y_off;
        cuts[n++] = CutXfer{c->rank - 1, cut, cut + R, cut - R, cut, 0};
    }
    if (c->rank + 1 < c->nranks) {
        const int cut = w->own_hi - w->y_off;
        cuts[n++] = CutXfer{c->rank + 1, cut - R, cut, cut, cut + R, 1};
    }
    return exchange_cuts(w, cuts, n, s);
}

// Generic neighbour exchange on stream s: up to one send and one receive with each neighbour (null / 0 bytes: none), one ncclGroup.
int strip_sendrecv(fse_world* w, const void* up_send, size_t up_send_bytes, void* up_recv, size_t up_recv_bytes, const void* down_send,
                   size_t down_send_bytes, void* down_recv, size_t down_recv_bytes, cudaStream_t s) {
    fse_ctx* c = w->ctx;
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    const bool up = c->rank > 0, down = c->rank + 1 < c->nranks;
    NK(g_nccl.GroupStart());
    ncclResult_t bad = 0;
    if (up && up_send_bytes && !bad) bad = g_nccl.Send(up_send, up_send_bytes, ncclUint8_t_, c->rank - 1, comm, s);
    if (up && up_recv_bytes && !bad) bad = g_nccl.Recv(up_recv, up_recv_bytes, ncclUint8_t_, c->rank - 1, comm, s);
    if (down && down_send_bytes && !bad) bad = g_nccl.Send(down_send, down_send_bytes, ncclUint8_t_, c->rank + 1, comm, s);
    if (down && down_recv_bytes && !bad) bad = g_nccl.Recv(down_recv, down_recv_bytes, ncclUint8_t_, c->rank + 1, comm, s);
    const ncclResult_t end = g_nccl.GroupEnd();
    if (bad) return fail(FSE_ENCCL, "ncclSend/ncclRecv: %s", g_nccl.GetErrorString(bad));
    if (end) return fail(FSE_ENCCL, "ncclGroupEnd: %s", g_nccl.GetErrorString(end));
    return FSE_OK;
}

// Vertical camera scroll: the local rows [send_lo, send_hi) of the seven planes, packed plane after plane like a halo message, travel to
// the neighbour below (send_down) or above, while the same number of rows arrives from the neighbour on the other side.  *recv_out = the
// packed rows that arrived (staging buffer, valid until the next exchange on this world), null when there is no neighbour on that side.
int strip_shift_rows(fse_world* w, int send_lo, int send_hi, bool send_down, unsigned char** recv_out, cudaStream_t s) {
    fse_ctx* c = w->ctx;
    const bool up = c->rank > 0, down = c->rank + 1 < c->nranks;
    const bool sends = send_down ? down : up, recvs = send_down ? up : down;
    const size_t bytes = halo_bytes(w, send_hi - send_lo);
    *recv_out = nullptr;
    if (sends) {
        if (int r = ensure_stage(w, 0, bytes)) return r;
        PackArgs a;
        halo_args(w, (unsigned char*)w->halo_stage[0], send_lo, send_hi, true, &a);
        halo_copy_kernel<<<dim3(64, 7), 256, 0, s>>>(a);
        CK(cudaGetLastError());
        w->ctx->launches += 1;
    }
    if (recvs)
        if (int r = ensure_stage(w, 1, bytes)) return r;
    const void* snd = sends ? w->halo_stage[0] : nullptr;
    void* rcv = recvs ? w->halo_stage[1] : nullptr;
    int r;
    if (send_down) r = strip_sendrecv(w, nullptr, 0, rcv, recvs ? bytes : 0, snd, sends ? bytes : 0, nullptr, 0, s);
    else r = strip_sendrecv(w, snd, sends ? bytes : 0, nullptr, 0, nullptr, 0, rcv, recvs ? bytes : 0, s);
    if (r) return r;
    *recv_out = (unsigned char*)rcv;
    return FSE_OK;
}

// ---- edits by ONE rank that reach into a neighbour's rows (rigid bodies, tools) -----------------------------------------------------
// Every rank makes the call with the same arguments and derives the same plan: the rank that holds the middle row of the edit's box
// runs it on its own rows + ghost rows (refreshed first); afterwards the part of the box that lies in a neighbour's rows travels
// there as a rectangle of cells, so both ranks agree on every row they share again.
void strip_rows_of(int Hglobal, int rank, int nranks, int* own_lo, int* own_hi, int* held_lo, int* held_hi) {
    const int nz = (Hglobal - 2 * CHUNK) / CHUNK;
    const int j0 = (int)((int64_t)nz * rank / nranks), j1 = (int)((int64_t)nz * (rank + 1) / nranks);
    const int lo = rank == 0 ? 0 : CHUNK + CHUNK * j0, hi = rank == nranks - 1 ? Hglobal : CHUNK + CHUNK * j1;
    if (own_lo) *own_lo = lo;
    if (own_hi) *own_hi = hi;
    if (held_lo) *held_lo = lo - STRIP_GHOST < 0 ? 0 : lo - STRIP_GHOST;
    if (held_hi) *held_hi = hi + STRIP_GHOST > Hglobal ? Hglobal : hi + STRIP_GHOST;
}
// runner of an edit whose box spans the global rows [ya, yb]: the owner of the middle row.  who != null: fails when the box does not
// lie inside the rows the runner holds.
int strip_runner_of_rows(fse_world* w, int ya, int yb, const char* who, int* exec) {
    const int nranks = w->ctx->nranks, Hg = w->Hglobal;
    int mid = (ya + yb) / 2;
    mid = mid < 0 ? 0 : (mid >= Hg ? Hg - 1 : mid);
    int e = 0, hi = 0;
    for (;; e++) {
        strip_rows_of(Hg, e, nranks, nullptr, &hi, nullptr, nullptr);
        if (e + 1 >= nranks || mid < hi) break;
    }
    *exec = e;
    if (who) {
        int hlo, hhi;
        strip_rows_of(Hg, e, nranks, nullptr, nullptr, &hlo, &hhi);
        const int a = ya < 0 ? 0 : ya, b = yb >= Hg ? Hg - 1 : yb;
        if (a <= b && (a < hlo || b >= hhi))
            return fail(FSE_ESTATE, "%s: rows %d..%d do not fit the rows rank %d holds (%d..%d): on multi-rank strips the call must stay within %d rows of one strip",
                        who, a, b, e, hlo, hhi - 1, STRIP_GHOST);
    }
    return FSE_OK;
}
// Runners of n edits with boxes (x0, y0, x1, y1; global, inclusive) that are applied in index order: boxes that overlap (transitively)
// form a group that one rank must run — the order between them matters — and a group is run by the owner of the middle row of its
// lowest-index box; edits that do not overlap commute.  Fails when a box does not fit the rows its runner holds.
int strip_group_runners(fse_world* w, const std::vector<int4>& box, const char* who, const char* what, std::vector<int>& exec) {
    const int n = (int)box.size(), nranks = w->ctx->nranks, Hg = w->Hglobal;
    std::vector<int> parent(n), order(n), active;
    for (int i = 0; i < n; i++) parent[i] = order[i] = i;
    auto find = [&](int i) {
        while (parent[i] != i) i = parent[i] = parent[parent[i]];
        return i;
    };
    std::sort(order.begin(), order.end(), [&](int a_, int b_) { return box[a_].y < box[b_].y; });
    for (int oi = 0; oi < n; oi++) {
        const int i = order[oi];
        size_t keep = 0;
        for (size_t k = 0; k < active.size(); k++) {
            const int j = active[k];
            if (box[j].w < box[i].y) continue;  // ends above: never overlaps anything that starts later
            active[keep++] = j;
            if (box[j].x <= box[i].z && box[i].x <= box[j].z) {
                const int ri = find(i), rj = find(j);
                if (ri != rj) parent[ri > rj ? ri : rj] = ri > rj ? rj : ri;  // the root of a group is its lowest index
            }
        }
        active.resize(keep);
        active.push_back(i);
    }
    exec.assign(n, 0);
    for (int b = 0; b < n; b++) {
        const int root = find(b);
        int e = 0;
        if (root == b) {
            if (int r = strip_runner_of_rows(w, box[b].y, box[b].w, nullptr, &e)) return r;
        } else {
            e = exec[root];  // root < b: already known
        }
        exec[b] = e;
        int hlo, hhi;
        strip_rows_of(Hg, e, nranks, nullptr, nullptr, &hlo, &hhi);
        const int ya = box[b].y < 0 ? 0 : box[b].y, yb = box[b].w >= Hg ? Hg - 1 : box[b].w;
        if (ya <= yb && (ya < hlo || yb >= hhi))
            return fail(FSE_ESTATE, "%s: %s %d (rows %d..%d, with the ones it overlaps) does not fit the rows rank %d holds (%d..%d): on multi-rank strips a group of "
                        "overlapping ones must lie within %d rows of one strip", who, what, b, ya, yb, e, hlo, hhi - 1, STRIP_GHOST);
    }
    return FSE_OK;
}
// the rectangles of box [xa, xb] x [ya, yb] (global, inclusive) run by rank `exec` that this rank sends or receives:
// rect[0] up send, [1] up recv, [2] down send, [3] down recv; (x0, y0 in local rows, w, h)
void strip_rects_of_box(fse_world* w, int exec, int xa, int ya, int xb, int yb, std::vector<int4> rect[4]) {
    const int nranks = w->ctx->nranks, me = w->ctx->rank, Hg = w->Hglobal;
    xa = xa < 0 ? 0 : xa;
    xb = xb >= w->W ? w->W - 1 : xb;
    if (xa > xb) return;
    for (int side = 0; side < 2; side++) {  // 0: the neighbour above the runner, 1: the one below
        const int nb = side == 0 ? exec - 1 : exec + 1;
        if (nb < 0 || nb >= nranks || (me != exec && me != nb)) continue;
        int hlo, hhi;
        strip_rows_of(Hg, nb, nranks, nullptr, nullptr, &hlo, &hhi);
        int a = ya < 0 ? 0 : ya, b = yb >= Hg ? Hg - 1 : yb;
        if (a < hlo) a = hlo;
        if (b >= hhi) b = hhi - 1;
        if (a > b) continue;
        const int4 r = make_int4(xa, a - w->y_off, xb - xa + 1, b - a + 1);
        if (me == exec) rect[side == 0 ? 0 : 2].push_back(r);  // I ran it: send towards that neighbour
        else rect[side == 0 ? 3 : 1].push_back(r);             // my neighbour ran it: the runner is below me (side 0) or above me (side 1)
    }
}
struct RectArgs {
    Planes p;
    int W;
    const int4* rects;     // x0, y0 (local rows), w, h
    const int* cell_off;   // first cell of each rectangle in the message
    unsigned char* stage;  // per rectangle: its cells plane after plane (17 bytes per cell)
    int pack;
};
__global__ void rect_copy_kernel(RectArgs a) {
    const int4 r = a.rects[blockIdx.x];
    const int n = r.z * r.w;
    unsigned char* st = a.stage + (size_t)a.cell_off[blockIdx.x] * 17;
    unsigned char* pl[7] = {(unsigned char*)a.p.mat, (unsigned char*)a.p.flg, (unsigned char*)a.p.stl, (unsigned char*)a.p.tmp,
                            (unsigned char*)a.p.col, (unsigned char*)a.p.fl, (unsigned char*)a.p.fd};
    const int es[7] = {1, 1, 1, 2, 4, 4, 4};
    size_t so = 0;
    for (int q = 0; q < 7; q++) {
        for (int i = threadIdx.x; i < n * es[q]; i += blockDim.x) {  // byte-wise: the message is not aligned for the wider planes
            const int c = i / es[q], bq = i % es[q];
            const size_t g = ((size_t)(r.y + c / r.z) * a.W + (size_t)(r.x + c % r.z)) * es[q] + bq;
            if (a.pack) st[so + i] = pl[q][g];
            else pl[q][g] = st[so + i];
        }
        so += (size_t)n * es[q];
    }
}
// pack -> one send / receive with each neighbour -> unpack, on stream s.  Both sides of a cut derive the same rectangles in the same
// order, so the messages need no header.
int strip_push_rects(fse_world* w, const std::vector<int4> rect[4], cudaStream_t s) {
    size_t cells[4] = {0, 0, 0, 0}, total = 0, first[4];
    std::vector<int4> all_r;
    std::vector<int> all_o;
    for (int q = 0; q < 4; q++) {
        first[q] = all_r.size();
        int o = 0;
        for (const int4& r : rect[q]) {
            all_r.push_back(r);
            all_o.push_back(o);
            o += r.z * r.w;
        }
        cells[q] = (size_t)o;
        total += rect[q].size();
        if (cells[q])
            if (int r = ensure_stage(w, q, cells[q] * 17)) return r;
    }
    if (total > w->push_rects_cap) {
        CK(cudaStreamSynchronize(s));
        cudaFree(w->d_push_rects);
        cudaFree(w->d_push_off);
        w->d_push_rects = nullptr;
        w->d_push_off = nullptr;
        w->push_rects_cap = 0;
        CK(cudaMalloc((void**)&w->d_push_rects, sizeof(int4) * (total + 64)));
        CK(cudaMalloc((void**)&w->d_push_off, sizeof(int) * (total + 64)));
        w->push_rects_cap = total + 64;
    }
    if (total) {
        CK(cudaMemcpyAsync(w->d_push_rects, all_r.data(), sizeof(int4) * total, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(w->d_push_off, all_o.data(), sizeof(int) * total, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));  // all_r / all_o are locals
    }
    RectArgs ra;
    ra.p = w->p;
    ra.W = w->W;
    for (int q = 0; q < 4; q += 2)  // pack what goes up (0) and down (2)
        if (!rect[q].empty()) {
            ra.rects = (const int4*)w->d_push_rects + first[q]; ra.cell_off = w->d_push_off + first[q]; ra.stage = (unsigned char*)w->halo_stage[q]; ra.pack = 1;
            rect_copy_kernel<<<(unsigned)rect[q].size(), 128, 0, s>>>(ra);
            w->ctx->launches += 1;
        }
    CK(cudaGetLastError());
    if (int r = strip_sendrecv(w, w->halo_stage[0], cells[0] * 17, w->halo_stage[1], cells[1] * 17, w->halo_stage[2], cells[2] * 17, w->halo_stage[3], cells[3] * 17, s))
        return r;
    for (int q = 1; q < 4; q += 2)  // unpack what came from above (1) and from below (3)
        if (!rect[q].empty()) {
            ra.rects = (const int4*)w->d_push_rects + first[q]; ra.cell_off = w->d_push_off + first[q]; ra.stage = (unsigned char*)w->halo_stage[q]; ra.pack = 0;
            rect_copy_kernel<<<(unsigned)rect[q].size(), 128, 0, s>>>(ra);
            w->ctx->launches += 1;
        }
    CK(cudaGetLastError());
    return FSE_OK;
}

// In-place sum of `count` 32-bit unsigned integers over all ranks.
int strip_allreduce_u32(fse_world* w, unsigned int* dev, size_t count, cudaStream_t s) {
    ncclComm_t comm = (ncclComm_t)w->ctx->nccl_comm;
    NK(g_nccl.AllReduce(dev, dev, count, /*ncclUint32*/ 3, /*ncclSum*/ 0, comm, s));
    return FSE_OK;
}

}  // namespace fse

using namespace fse;

extern "C" {

// 128-byte ncclUniqueId for rank 0 to broadcast (any out-of-band channel: torch.distributed, MPI, a file ...).
FSE_API int fse_strip_plan(int32_t width, int32_t height_global, int32_t rank, int32_t nranks, const int32_t* boxes, int32_t n, int32_t* runner,
                           int32_t* rects_out, int32_t cap_rects, int32_t* n_rects) {
    if (!boxes || !runner || !rects_out || !n_rects || n < 0 || cap_rects < 0 || nranks < 1 || rank < 0 || rank >= nranks || width < 1 ||
        height_global % CHUNK || (height_global - 2 * CHUNK) / CHUNK < nranks)
        return fail(FSE_EINVAL, "fse_strip_plan: bad argument");
    fse_ctx c;  // what the planning functions read of a strip world; nothing here touches a device
    c.rank = rank;
    c.nranks = nranks;
    fse_world w;
    w.ctx = &c;
    w.W = width;
    w.strip = true;
    w.Hglobal = height_global;
    int held_lo, held_hi;
    strip_rows_of(height_global, rank, nranks, &w.own_lo, &w.own_hi, &held_lo, &held_hi);
    w.y_off = held_lo;
    w.H = held_hi - held_lo;
    std::vector<int4> box(n);
    for (int i = 0; i < n; i++) box[i] = make_int4(boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]);
    std::vector<int> exec;
    if (int r = strip_group_runners(&w, box, "fse_strip_plan", "box", exec)) return r;
    std::vector<int4> rect[4];
    for (int i = 0; i < n; i++) {
        runner[i] = exec[i];
        strip_rects_of_box(&w, exec[i], box[i].x, box[i].y, box[i].z, box[i].w, rect);
    }
    for (int q = 0; q < 4; q++) {
        n_rects[q] = (int32_t)rect[q].size();
        for (int k = 0; k < (int)rect[q].size() && k < cap_rects; k++) {
            int32_t* o = rects_out + ((size_t)q * cap_rects + k) * 4;
            o[0] = rect[q][k].x; o[1] = rect[q][k].y; o[2] = rect[q][k].z; o[3] = rect[q][k].w;
        }
    }
    return FSE_OK;
}

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
