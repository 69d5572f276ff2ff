// fse_internal.hpp — host-side state behind the opaque handles of include/fse.h.
#pragma once
#include <atomic>
#include <string>
#include <utility>
#include <vector>

#include "fse_device.cuh"

struct fse_ctx {
    int device = 0;
    int sm_count = 0;
    fse::DevTables h_tabs{};
    fse::DevTables* d_tabs = nullptr;
    fse_interaction* d_inter = nullptr;
    int32_t* d_inter_off = nullptr;
    fse_interaction* d_react = nullptr;
    bool has_materials = false;
    int max_reach = 0;
    std::atomic<int64_t> launches{0};
    // multi-GPU (fse_comm.cu)
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    bool spiral_ready = false;  // fse_particles.cu: the deposit spiral's offset table is in constant memory
};

struct fse_world {
    fse_ctx* ctx = nullptr;
    int W = 0, H = 0;  // local planes: H rows starting at global row y_off
    // strip worlds (fse_strip_create): the world is Hglobal rows tall, this rank owns global rows [own_lo, own_hi)
    // and holds [y_off, y_off + H) (owned rows + ghost rows).  Plain worlds: y_off = 0, Hglobal = H.
    bool strip = false;
    int y_off = 0, Hglobal = 0, own_lo = 0, own_hi = 0;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_boundary = nullptr, ev_comm = nullptr;
    int* d_chunk_lists = nullptr;  // cached per-phase boundary / interior chunk lists
    fse_rect list_zone{0, 0, 0, 0};
    int list_off[4][2]{}, list_cnt[4][2]{};
    fse::Planes p{};
    int16_t* tmp_scratch = nullptr;
    cudaStream_t stream = nullptr;
    fse::TickFork fork{};
    // loose particles (world::cells)
    fse_particle* pbuf = nullptr;
    unsigned int* pcount = nullptr;  // [0] live count, [1..] scratch counters
    unsigned int pcap = 0;
    uint64_t next_user_particle = 1;
    // particle settle scratch (fse_particles.cu)
    void* part_scratch = nullptr;
    size_t part_scratch_bytes = 0;
    void* part_list = nullptr;  // indices of the particles that take part in the deposit rounds
    size_t part_list_bytes = 0;
    fse_particle* pbuf2 = nullptr;  // compaction target, swapped with pbuf every fse_particles_tick
    size_t pbuf2_bytes = 0;
    void* tool_scratch = nullptr;    // fse_tools.cu: cell lists and results of the interactive tools
    size_t tool_scratch_bytes = 0;
    void* entity_bufs = nullptr;     // fse_entities.cu EntityBufs
    void* particle_strip = nullptr;  // strips: exchange buffers of the particle protocol (fse_particles.cu StripBufs)
    void* claim_keys = nullptr;
    size_t claim_keys_bytes = 0;
    void* claim_vals = nullptr;
    size_t claim_vals_bytes = 0;
    // active-region tracking (fse_active_enable): awake flag per 128x128 chunk of the whole world
    bool active_on = false;
    uint8_t* d_awake = nullptr;
    int acols = 0, arows = 0;
    int* d_active_list = nullptr;   // compacted (cxi | cyi << 16) of the phase being launched
    int* d_active_count = nullptr;
    unsigned int* d_chunk_state = nullptr;  // per-pass kernels: what the passes of the running phase saw in each chunk
    int fused_max_chunks = 296;            // phases this small run in the fused kernel (FSE_FUSED_MAX_CHUNKS)
    bool active_fused = false;             // FSE_ACTIVE_FUSED=1: active-chunk phases use the fused kernel
    // longest-first chunk order of the per-pass tick kernels: per colour, last tick's pass-1 cycles and the list built from them
    unsigned int* d_lpt_cost = nullptr;
    int* d_lpt_list = nullptr;
    int lpt_cap = 0;            // chunks per colour the buffers hold
    int lpt_sig[4][4]{};        // (x0, y0, ncx, ncy) the list of a colour was built for; ncx = 0: no list yet
    bool lpt_on = true;
    bool lpt_deal = false;      // FSE_LPT_DEAL=1: the longest-first order is dealt out over the parts of a phase (lpt_build_kernel)
    int lpt_parts[4] = {0, 0, 0, 0};  // parts the list of a colour was dealt into (0 / 1: plain order)
    // settled-row skipping of the per-pass kernels (classify_rows_kernel): ROWMASK_WORDS words per chunk of a colour's grid
    uint32_t* d_rowmask = nullptr;
    int rowmask_cap = 0;
    int p2_split = 1;        // FSE_P2_SPLIT: 0 never, 1 (default) from the iteration on in which no powder / gas acts, 2 always: pass 2 as a row-parallel liquid apply + sequential powder / gas rows
    int rowskip_mode = 2;    // FSE_ROW_SKIP: 0 = step every row, 1 = always skip settled rows, 2 (default) = per phase, by the gate below
    // Skip gate.  A phase is bound by the row chain of its busiest chunk, so skipping settled rows only pays when most rows are
    // settled; below that the classification costs more than it saves.  Each classified phase counts its active rows on the device;
    // the counts come back asynchronously (pinned memory + event, never a sync) and a phase uses the skip while less than
    // rowskip_max_active of its rows were active the last time it was classified.  Every 32nd tick classifies all phases again.
    unsigned int* d_phase_rows = nullptr;   // 16 counters (4 colours x cell_iter <= 4)
    unsigned int* h_phase_rows = nullptr;   // pinned copy of the last finished tick
    cudaEvent_t ev_phase_rows = nullptr;
    bool phase_rows_pending = false;
    uint32_t phase_rows_valid = 0;          // bit p: h_phase_rows[p] was counted in the tick that was copied
    uint32_t phase_rows_counted = 0;        // bit p: phase p is being classified in the tick in flight
    float phase_active[16];                 // last known active fraction per phase, < 0 = unknown
    float phase_seq[16];                    // last known fraction of rows with live powder / gas per phase, < 0 = unknown (pass-2 split gate)
    float p2_split_max_seq = 0.10f;         // FSE_P2_SPLIT_MAX_SEQ
    long long phase_rows_total[16];
    float rowskip_max_active = 0.25f;       // FSE_ROW_SKIP_MAX_ACTIVE
    // particle pool bookkeeping on the host: count seen by the last call that read it, particles promised to calls since, drops
    unsigned int particles_seen = 0;
    int64_t particles_promised = 0;
    uint64_t particles_dropped = 0;
    // strips: staging buffers of the packed halo messages (fse_comm.cu): [2 * cut] send, [2 * cut + 1] receive
    void* halo_stage[4] = {nullptr, nullptr, nullptr, nullptr};
    void* d_push_rects = nullptr;  // rectangles of strip_push_rects (int4) and their cell offsets
    int* d_push_off = nullptr;
    size_t push_rects_cap = 0;
    size_t halo_stage_bytes[4] = {0, 0, 0, 0};
    // rigid-body bridge (fse_bodies.cu) and outline scratch (fse_outline.cu)
    struct fse_bodies* bodies = nullptr;
    int last_bridge_rounds = 0;
    void* outline_scratch = nullptr;
    size_t outline_scratch_bytes = 0;
    void *outline_pinned = nullptr, *outline_pinned2 = nullptr;  // pinned host staging of fse_mask_outline
    size_t outline_pinned_bytes = 0, outline_pinned2_bytes = 0;
    // render planes (fse_render.cu): main | fire | emission RGBA, W*H words each; scroll scratch (largest plane)
    uint32_t* d_pixels = nullptr;
    void* d_render_stats = nullptr;
    // scroll: a second set of planes; fse_scroll writes the shifted world into it and swaps the two sets (no copy back)
    fse::Planes p_shadow{};
    // liquid flow accumulators (fse_flow_enable): flowX | flowY | prevFlowX | prevFlowY, W*H floats each, + the flow texture
    float* d_flow = nullptr;
    uint32_t* d_pixels_flow = nullptr;
    // second cell layer and background colours (fse_layer2_* / fse_background_*), allocated on first use; layer_dirty: bit 0 layer2Dirty,
    // bit 1 backgroundDirty; textures: layer 2 | background
    uint8_t* l2_mat = nullptr;
    int16_t* l2_tmp = nullptr;
    uint32_t *l2_col = nullptr, *bg_col = nullptr;
    uint8_t* layer_dirty = nullptr;
    uint32_t* d_pixels_layers = nullptr;
    uint8_t* l2_mat_s = nullptr;   // scroll targets of the layer planes (swapped like p_shadow)
    int16_t* l2_tmp_s = nullptr;
    uint32_t *l2_col_s = nullptr, *bg_col_s = nullptr;
    // stats / staging
    void* d_stats = nullptr;
    void* h_stats = nullptr;
    fse_cell* d_stage = nullptr;
    size_t stage_cells = 0;
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool kt_enabled = false;
    size_t kt_used = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kt_events;
    int strip_boundary_min = 64;            // cut-adjacent chunk rows of at least this many chunks use the per-pass kernels
    bool timeline_on = false;               // FSE_STRIP_TIMELINE=1: events per strip phase (fse_strip_timeline_read)
    std::vector<cudaEvent_t> timeline_ev;
    size_t timeline_used = 0;
    uint64_t ticks = 0;
    int schedule = FSE_SCHEDULE_ROWS;
    unsigned long long* d_dbg = nullptr;
    int fused = 0;  // rows schedule: fused single kernel instead of one kernel per pass (FSE_ROWS_FUSED=1 or profiling)
};

struct fse_bodies;
void fse_bodies_free(fse_world* w);

#define FSE_PARTICLE_ROUNDS 16  // deposit-resolution rounds per fse_particles_tick (losers retry next tick)

namespace fse {
extern thread_local std::string g_err;
int fail(int code, const char* fmt, ...);
int fse_wake_rect(fse_world* w, int x, int y_local, int rw, int rh);  // wake the chunks under a rect of local rows (active tracking)

int check_rect(fse_world* w, int x, int y, int rw, int rh, const char* who);  // rect in global coordinates inside the held rows
int ensure_stage(fse_world* w, size_t cells);                                 // w->d_stage holds at least `cells` fse_cell
void render_free(fse_world* w);                                               // planes owned by fse_render.cu
int particles_headroom(fse_world* w, int64_t need, bool exact);  // grow the particle pool before a call that spawns up to `need`
int strip_exchange(fse_world* w, int ofy, int j0, int j1, int zone_y_local, cudaStream_t s);
void particles_strip_free(fse_world* w);
void entities_free(fse_world* w);
int strip_refresh(fse_world* w, cudaStream_t s, int rows = 16);
int strip_sendrecv(fse_world* w, const void* up_send, size_t up_send_bytes, void* up_recv, size_t up_recv_bytes, const void* down_send,
                   size_t down_send_bytes, void* down_recv, size_t down_recv_bytes, cudaStream_t s);
int strip_allreduce_u32(fse_world* w, unsigned int* dev, size_t count, cudaStream_t s);
int strip_shift_rows(fse_world* w, int send_lo, int send_hi, bool send_down, unsigned char** recv_out, cudaStream_t s);
// edits by one rank that reach into a neighbour's rows (fse_comm.cu): layout of any rank, runner of a box, rectangles to exchange, the exchange
constexpr int STRIP_GHOST = 32;  // ghost rows of a strip (fse_strip_create)
void strip_rows_of(int Hglobal, int rank, int nranks, int* own_lo, int* own_hi, int* held_lo, int* held_hi);
int strip_runner_of_rows(fse_world* w, int ya, int yb, const char* who, int* exec);
int strip_group_runners(fse_world* w, const std::vector<int4>& box, const char* who, const char* what, std::vector<int>& exec);
void strip_rects_of_box(fse_world* w, int exec, int xa, int ya, int xb, int yb, std::vector<int4> rect[4]);
int strip_push_rects(fse_world* w, const std::vector<int4> rect[4], cudaStream_t s);
size_t tick_smem_bytes();
cudaError_t launch_lpt_build(const unsigned int* cost, int n, int ncx, int* list, cudaStream_t stream, const int* members = nullptr, int parts = 1);
cudaError_t launch_tick_phase(const TickParams& P, int n_chunks, cudaStream_t stream, int* launched, const TickFork* fork);  // *launched = kernels enqueued
cudaError_t launch_compact_active(const uint8_t* awake, int acols, int ci0, int cj0, int ncx, int ncy, int* list, int* count, cudaStream_t s);

cudaError_t launch_write_rect(Planes p, int W, int x0, int y0, int rw, int rh, const fse_cell* src, cudaStream_t s);
cudaError_t launch_read_rect(Planes p, int W, int x0, int y0, int rw, int rh, fse_cell* dst, cudaStream_t s);
cudaError_t launch_fill_air(Planes p, size_t n, uint8_t air, cudaStream_t s);
cudaError_t launch_clear_dirty(Planes p, size_t n, cudaStream_t s);
cudaError_t launch_stats(Planes p, int W, int x0, int y0, int rw, int rh, int yoff, const DevTables* T, void* out, cudaStream_t s);
size_t dev_stats_bytes();
cudaError_t launch_temperature(Planes p, int16_t* scratch, int W, int H, int zx, int zy, int zw, int zh, const DevTables* T, uint8_t* awake,
                               int acols, int yoff, cudaStream_t s);
}  // namespace fse
