// batch.cu -- ensembles of independent channels stepped by BATCHED launches (BASELINE.json configs[3];
// "replicas only", no data-path communication).
//
// sphb_ensemble_run_dual (context.cu) replays one CUDA graph per channel; a dp = 0.02 channel is 68 tiles, less
// than one wave of the 148 SMs, so an ensemble stepped that way is bound by the device's kernel launch rate
// (~0.5 us per kernel, ~12 us per channel-step).  A batch instead launches every kernel of the reference step
// m:251-283 ONCE for all channels: grid = (the channel's own grid, channels), the kernel bodies are the ones of
// the single-channel path (ops_csr.cu, neighbor.cu, scan.cu: `*_batch` kernels call the same __device__ bodies),
// and the per-channel arguments -- state pointers, list pointers, mu / g / ..., the device scalars dt, t, max|v|,
// tau -- come from records in device memory.  Every channel keeps its own context (state, neighbour list, clock):
// upload / download / status / profile of a member context work as before, on the batch's stream.
//
// Each channel advances with its OWN dt (device_dt per record); the step count is common.  Requirements checked at
// creation: one device, equal n_fluid / n_total / h / DL / sort_interval, equal cell and tile geometry (same walls
// => same cell grid), every channel past its first (exact, capacity-fitting) list build.
//
// The cell sort (m:272-278, every sort_interval steps) swaps buffer pointers: it runs per channel through the
// single-channel code on the batch's stream, then the records are refreshed; the replayed graph takes only the
// record arrays as arguments and stays valid.
#include "context.cuh"

#include <limits>

struct sphb_batch {
    std::vector<sphb_ctx *> ctx;
    int device = 0;
    cudaStream_t st = nullptr;
    // per-channel records: host mirrors (pinned) and device arrays
    size_t nbr_rec = 0;
    BatchChan *h_chan = nullptr;
    unsigned char *h_nbr = nullptr;
    ScanChan *h_scan = nullptr;
    DevBuf d_chan, d_nbr, d_scan;
    // uniform launch geometry
    int n_fluid = 0, n_total = 0, cell_cap = 0, sort_interval = 0;
    size_t ell_stride = 0;
    unsigned n_tiles = 0;
    int tile_cap = 0, ell_cap = 0;   // largest over the channels
    long long outer_count = 0;
    long long launches = 0;
    // the replayed outer step
    cudaGraphExec_t graph = nullptr;
    int graph_inner = -1;
    long long graph_launches = 0;
};

namespace {

void batch_drop_graph(sphb_batch *b)
{
    if (b->graph) cudaGraphExecDestroy(b->graph);
    b->graph = nullptr;
    b->graph_inner = -1;
}

BatchLaunch phys_launch(const sphb_batch *b)
{
    BatchLaunch L;
    L.chan = b->d_chan.as<BatchChan>();
    L.n_chan = (int)b->ctx.size();
    L.n_tiles = b->n_tiles;
    L.n_fluid = b->n_fluid;
    L.tile_cap = b->tile_cap;
    L.ncx = b->ctx[0]->view().ncx;
    L.tile_w = b->ctx[0]->nw.tile_w;
    L.st = b->st;
    return L;
}

NbrBatchLaunch nbr_launch(const sphb_batch *b)
{
    NbrBatchLaunch L;
    L.chan = b->d_nbr.p;
    L.scan = b->d_scan.as<ScanChan>();
    L.n_chan = (int)b->ctx.size();
    L.n_total = b->n_total;
    L.cell_cap = b->cell_cap;
    L.ell_stride = b->ell_stride;
    L.n_tiles = b->n_tiles;
    L.tile_cap = b->tile_cap;
    L.ell_cap = b->ell_cap;
    L.grid_ready = true;   // enqueue_physics always precedes the build and fuses the grid set-up
    L.st = b->st;
    return L;
}

// Checks that the channels still share one launch geometry and rewrites + uploads the records.
int batch_refresh(sphb_batch *b)
{
    const size_t n = b->ctx.size();
    const sphb_ctx *c0 = b->ctx[0];
    b->n_fluid = c0->prm.n_fluid;
    b->n_total = c0->prm.n_total;
    b->sort_interval = c0->prm.sort_interval;
    b->cell_cap = c0->nw.cell_cap;
    b->ell_stride = c0->nw.ell_stride;
    b->n_tiles = c0->nw.n_tiles;
    b->tile_cap = 0;
    b->ell_cap = 0;
    // the host mirrors may still be the source of the previous upload
    SPHB_CUDA(cudaStreamSynchronize(b->st));
    for (size_t k = 0; k < n; ++k) {
        sphb_ctx *c = b->ctx[k];
        if (c->slab.on || c->nw.slab) return set_error("SPHB:batch", "channel %zu is a slab of a sharded channel", k);
        if (!c->planned || !c->list_sized)
            return set_error("SPHB:batch", "channel %zu has no neighbour list yet (initialise it first)", k);
        if (c->kprof.on) return set_error("SPHB:batch", "per-kernel profiling is on in channel %zu", k);
        if (c->prm.n_fluid != b->n_fluid || c->prm.n_total != b->n_total || c->prm.h != c0->prm.h ||
            c->prm.DL != c0->prm.DL || c->prm.sort_interval != b->sort_interval)
            return set_error("SPHB:batch", "channel %zu differs from channel 0 in n_fluid / n_total / h / DL / sort_interval", k);
        if (c->nw.cell_cap != b->cell_cap || c->nw.ell_stride != b->ell_stride || c->nw.n_tiles != b->n_tiles ||
            c->nw.tile_w != c0->nw.tile_w)
            return set_error("SPHB:batch", "channel %zu has another cell / tile geometry than channel 0 (%d cells, %u tiles of "
                             "width %d vs %d, %u, %d)", k, c->nw.cell_cap, c->nw.n_tiles, c->nw.tile_w, b->cell_cap, b->n_tiles,
                             c0->nw.tile_w);
        b->tile_cap = std::max(b->tile_cap, c->nw.tile_cap);
        b->ell_cap = std::max(b->ell_cap, c->nw.ell_cap);
        SPHB_TRY(c->shear_partials.reserve(2 * (size_t)c->nw.n_tiles * sizeof(double)));
        BatchChan &r = b->h_chan[k];
        r.P = c->P;
        r.S = c->view();
        r.sc = c->sc();
        r.shear_partials = c->shear_partials.as<double>();
        r.F = neighbor_grid_fuse(c->nw);
        SPHB_TRY(neighbor_batch_fill(c->nw, c->pos.as<double2>(), b->h_nbr + k * b->nbr_rec, &b->h_scan[k]));
    }
    SPHB_TRY(csr_configure(b->tile_cap));
    SPHB_TRY(csr_batch_configure(b->tile_cap));
    SPHB_CUDA(cudaMemcpyAsync(b->d_chan.p, b->h_chan, n * sizeof(BatchChan), cudaMemcpyHostToDevice, b->st));
    SPHB_CUDA(cudaMemcpyAsync(b->d_nbr.p, b->h_nbr, n * b->nbr_rec, cudaMemcpyHostToDevice, b->st));
    SPHB_CUDA(cudaMemcpyAsync(b->d_scan.p, b->h_scan, n * sizeof(ScanChan), cudaMemcpyHostToDevice, b->st));
    return 0;
}

// m:254-270 for every channel: density + KGC, fused first substep, further acoustic substeps
int enqueue_physics(sphb_batch *b, int n_inner)
{
    const double inf = std::numeric_limits<double>::infinity();
    const BatchLaunch L = phys_launch(b);
    SPHB_TRY(csr_batch_publish_dt(L, inf, 0.0));
    SPHB_TRY(csr_batch_density(L));
    SPHB_TRY(csr_batch_kgc(L));
    SPHB_TRY(csr_batch_force_fused(L, true));
    // the step's last continuity pass also reduces the y range, derives the grid and zeroes the cell counters of the
    // list build that follows (GridFuse)
    SPHB_TRY(csr_batch_continuity(L, true, true, n_inner == 1));
    b->launches += 5;
    for (int s = 1; s < n_inner; ++s) {
        SPHB_TRY(csr_batch_time_step(L, inf, 0.0));
        SPHB_TRY(csr_batch_pressure(L));
        SPHB_TRY(csr_batch_continuity(L, true, true, s == n_inner - 1));
        b->launches += 3;
    }
    return 0;
}

// m:280-283 for every channel: list rebuild + wall-shear monitor
int enqueue_list_and_shear(sphb_batch *b)
{
    SPHB_TRY(neighbor_batch_build(nbr_launch(b), &b->launches));
    SPHB_TRY(csr_batch_wall_shear(phys_launch(b)));
    b->launches += 1;
    return 0;
}

void after_outer(sphb_batch *b, int n_inner)
{
    for (sphb_ctx *c : b->ctx) {
        c->outer_count += 1;
        c->h_step += n_inner;
        c->have_list = true;
        c->nw.half_only = false;
        c->nw.half_valid = false;
    }
}

}  // namespace

extern "C" {

SPHB_API int sphb_batch_create(sphb_ctx **ctxs, int n_ctx, sphb_batch **out)
{
    clear_error();
    if (!ctxs || n_ctx < 1 || !out) return set_error("SPHB:arg", "bad batch arguments");
    if (n_ctx > 65535) return set_error("SPHB:batch", "at most 65535 channels per batch (grid.y)");
    *out = nullptr;
    for (int k = 0; k < n_ctx; ++k) {
        if (!ctxs[k]) return set_error("SPHB:arg", "null context in batch");
        if (ctxs[k]->device != ctxs[0]->device) return set_error("SPHB:batch", "the channels of a batch live on one device");
        if (ctxs[k]->outer_count != ctxs[0]->outer_count || ctxs[k]->h_step != ctxs[0]->h_step)
            return set_error("SPHB:batch", "the channels of a batch must have taken the same number of steps");
        for (int j = 0; j < k; ++j)
            if (ctxs[j] == ctxs[k]) return set_error("SPHB:batch", "context listed twice");
    }
    SPHB_TRY(ensure_device());
    sphb_batch *b = new sphb_batch();
    b->ctx.assign(ctxs, ctxs + n_ctx);
    b->device = ctxs[0]->device;
    b->outer_count = ctxs[0]->outer_count;
    const size_t n = (size_t)n_ctx;
    int rc = [&]() -> int {
        SPHB_CUDA(cudaSetDevice(b->device));
        SPHB_CUDA(cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking));
        b->nbr_rec = (neighbor_batch_record_bytes() + 15) & ~(size_t)15;
        SPHB_CUDA(cudaMallocHost((void **)&b->h_chan, n * sizeof(BatchChan)));
        SPHB_CUDA(cudaMallocHost((void **)&b->h_nbr, n * b->nbr_rec));
        SPHB_CUDA(cudaMallocHost((void **)&b->h_scan, n * sizeof(ScanChan)));
        SPHB_TRY(b->d_chan.reserve(n * sizeof(BatchChan)));
        SPHB_TRY(b->d_nbr.reserve(n * b->nbr_rec));
        SPHB_TRY(b->d_scan.reserve(n * sizeof(ScanChan)));
        // the members launch on the batch's stream from now on (like sphb_ctx_set_stream)
        for (sphb_ctx *c : b->ctx) {
            SPHB_CUDA(cudaStreamSynchronize(c->st));
            ctx_drop_graphs(c);
            if (!c->own_st) c->own_st = c->st;
            c->st = b->st;
        }
        SPHB_TRY(batch_refresh(b));
        SPHB_CUDA(cudaStreamSynchronize(b->st));
        return 0;
    }();
    if (rc) {
        sphb_batch_destroy(b);
        return rc;
    }
    *out = b;
    return 0;
}

// Destroy the batch BEFORE its member contexts; they get their own streams back.
SPHB_API void sphb_batch_destroy(sphb_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->st) cudaStreamSynchronize(b->st);
    batch_drop_graph(b);
    for (sphb_ctx *c : b->ctx) {
        if (c->st == b->st && c->own_st) {
            ctx_drop_graphs(c);
            c->st = c->own_st;
        }
    }
    b->d_chan.release();
    b->d_nbr.release();
    b->d_scan.release();
    if (b->h_chan) cudaFreeHost(b->h_chan);
    if (b->h_nbr) cudaFreeHost(b->h_nbr);
    if (b->h_scan) cudaFreeHost(b->h_scan);
    if (b->st) cudaStreamDestroy(b->st);
    delete b;
}

// Call after anything that may have moved a member's buffers or changed its list capacity outside the batch
// (sphb_ctx_upload of positions, sphb_ctx_sort_by_cell, sphb_ctx_run_until's capacity recovery, ...).
SPHB_API int sphb_batch_refresh(sphb_batch *b)
{
    clear_error();
    if (!b) return set_error("SPHB:arg", "null batch");
    SPHB_CUDA(cudaSetDevice(b->device));
    return batch_refresh(b);
}

// n_outer outer steps of every channel; each = density + KGC, n_inner acoustic substeps (the first fused with the
// viscous / transport pass), [cell sort], list rebuild, wall-shear monitor -- sphb_ctx_run_dual's schedule.
SPHB_API int sphb_batch_run_dual(sphb_batch *b, int n_outer, int n_inner)
{
    clear_error();
    if (!b) return set_error("SPHB:arg", "null batch");
    if (n_inner < 1 || n_outer < 0) return set_error("SPHB:arg", "n_inner >= 1 and n_outer >= 0 required");
    SPHB_CUDA(cudaSetDevice(b->device));
    for (sphb_ctx *c : b->ctx) {
        if (!c->have_list) return set_error("SPHB:state", "a channel of the batch has no neighbour list");
        SPHB_TRY(ctx_fresh_mirrors(c, false));   // rows uploaded behind a standing list (eager, on the batch's stream)
    }
    if (b->graph_inner != n_inner) batch_drop_graph(b);
    const int si = b->sort_interval;
    for (int o = 0; o < n_outer; ++o) {
        b->outer_count += 1;
        const bool with_sort = si > 0 && b->outer_count % si == 0 && b->outer_count != 1;   // m:272
        if (with_sort) {
            SPHB_TRY(enqueue_physics(b, n_inner));
            for (sphb_ctx *c : b->ctx) SPHB_TRY(ctx_do_sort(c));   // swaps the channel's buffers
            SPHB_TRY(batch_refresh(b));
            SPHB_TRY(enqueue_list_and_shear(b));
            after_outer(b, n_inner);
            continue;
        }
        if (!b->graph) {
            const long long l0 = b->launches;
            cudaGraph_t g = nullptr;
            SPHB_CUDA(cudaStreamBeginCapture(b->st, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue_physics(b, n_inner);
            if (!rc) rc = enqueue_list_and_shear(b);
            cudaError_t e = cudaStreamEndCapture(b->st, &g);
            if (rc) {
                if (g) cudaGraphDestroy(g);
                b->launches = l0;
                return rc;
            }
            if (e != cudaSuccess) return set_error("SPHB:cuda", "batch graph capture failed: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&b->graph, g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return set_error("SPHB:cuda", "batch graph instantiate failed: %s", cudaGetErrorString(e));
            b->graph_inner = n_inner;
            b->graph_launches = b->launches - l0;
            b->launches = l0;
        }
        SPHB_CUDA(cudaGraphLaunch(b->graph, b->st));
        b->launches += b->graph_launches;
        after_outer(b, n_inner);
    }
    return 0;
}

SPHB_API int sphb_batch_sync(sphb_batch *b)
{
    clear_error();
    if (!b) return set_error("SPHB:arg", "null batch");
    SPHB_CUDA(cudaSetDevice(b->device));
    SPHB_CUDA(cudaStreamSynchronize(b->st));
    SPHB_CUDA(cudaGetLastError());
    // a member that raised a device flag froze at its last consistent step (step_poisoned): report it
    int all = 0;
    SPHB_TRY(sphb_batch_error_flags(b, &all));
    if (all) return set_error(all & kErrNonFinite ? "SPHB:diverged" : "SPHB:capacity",
                              "a channel of the batch raised device error flags (OR over the channels: %d); it stands at its last "
                              "consistent step -- sphb_ctx_status of the members tells which", all);
    return 0;
}

// Kernels launched by the batch so far (graph replays counted by their captured kernels).
SPHB_API long long sphb_batch_launches(const sphb_batch *b) { return b ? b->launches : 0; }

// OR of the members' device error flags (sphb_status.error_flags); synchronises the batch's stream.
SPHB_API int sphb_batch_error_flags(sphb_batch *b, int *flags_or)
{
    clear_error();
    if (!b || !flags_or) return set_error("SPHB:arg", "null argument");
    SPHB_CUDA(cudaSetDevice(b->device));
    std::vector<int> f(b->ctx.size(), 0);
    for (size_t k = 0; k < b->ctx.size(); ++k)
        SPHB_CUDA(cudaMemcpyAsync(&f[k], b->ctx[k]->nw.flags.p, sizeof(int), cudaMemcpyDeviceToHost, b->st));
    SPHB_CUDA(cudaStreamSynchronize(b->st));
    int all = 0;
    for (int v : f) all |= v;
    *flags_or = all;
    return 0;
}

}  // extern "C"
