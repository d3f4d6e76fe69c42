// slab_group.cu -- the host of a sharded channel inside ONE process: N x-slabs over the GPUs of a node, driven from C
// (include/sph_b200.h section 3b).  No Python, no torch, no NCCL: the slabs talk through the
// peer-to-peer mailboxes of slab.cu (plain device pointers + peer access inside one process), so a step is N streams of
// this library's kernels and nothing else.  Host code only -- it sits on top of the C ABI like any other caller:
//
//     cut        fluid rows by wrapped cell column (nb:253-270), wall rows of the slab and its halo columns unwrapped
//                around the slab (x -+ DL across the periodic seam, nb:284,291) -- the C++ form of slab.py's cut_slab
//     step       every slab: sphb_ctx_slab_step_begin, sphb_ctx_slab_p2p_send | sphb_ctx_slab_p2p_recv,
//                sphb_ctx_slab_step_end; with one slab per device the four calls of a slab are captured once as a CUDA
//                graph per slab and replayed (one graph launch per GPU and step from one host thread)
//     run_until  the reference loop m:250-283: the steps the target cannot cut short (context.cu, sphb_ctx_run_until) as
//                replayed graphs without reading the clock, the last few one by one with dt = min(dt, remain) (m:259)
//
// Slabs that share a device are stepped eagerly with a host wait between the sends and the receives (a waiting unpack
// kernel must not keep the sender of the same GPU off its SMs): the parity-test and small-box form.
#include "common.cuh"
#include "../../include/sph_b200.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

using namespace sphb;

namespace {

constexpr int kHaloCols = 4;   // list -> density -> KGC -> forces -> continuity: four columns of validity per step
constexpr int kHdrD = 16, kRecD = 16;   // doubles of buffer header / per record (slab.cu kHdr, kRec)

struct OneSlab {
    sphb_ctx *ctx = nullptr;
    int device = 0;
    int col_lo = 0, col_hi = 0;
    int nf_cap = 0, n_wall = 0;
    double *send_left = nullptr, *send_right = nullptr;   // device
    void *mailbox = nullptr;
    cudaGraphExec_t graph = nullptr;
    std::vector<int> wall_ids;   // global row of every wall row of the slab
};

}  // namespace

struct sphb_slabs {
    sphb_params prm{};   // the whole channel
    int n = 0, ncx = 0, rec_cap = 0;
    std::vector<OneSlab> s;
    bool one_per_device = false;
    bool graphs_ok = true;
    int compact_every = 256;
    long long k = 0;     // exchanges so far
    std::vector<double> wall_pos, wall_vel, wall_mass;   // the static wall rows of the channel (column-major, n_wall)
};

namespace {

int wrapped_column(double x, double DL, double cell, int ncx)
{
    const double xw = x - std::floor(x / DL) * DL;   // nb:270
    int c = (int)std::floor(xw / cell);
    return std::min(std::max(c, 0), ncx - 1);
}

int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
int floor_mod(int a, int b) { return a - floor_div(a, b) * b; }

void drop_graphs(sphb_slabs *g)
{
    for (OneSlab &o : g->s) {
        if (o.graph) cudaGraphExecDestroy(o.graph);
        o.graph = nullptr;
    }
}

int exchange(sphb_slabs *g)
{
    g->k += 1;
    const bool squeeze = g->compact_every > 0 && g->k % g->compact_every == 0;
    for (OneSlab &o : g->s) {
        SPHB_TRY(sphb_ctx_slab_p2p_send(o.ctx, o.send_left, o.send_right));
        if (squeeze) SPHB_TRY(sphb_ctx_slab_compact(o.ctx));
    }
    if (!g->one_per_device)
        for (OneSlab &o : g->s) SPHB_TRY(sphb_ctx_sync(o.ctx));
    for (OneSlab &o : g->s) SPHB_TRY(sphb_ctx_slab_p2p_recv(o.ctx, o.send_left, o.send_right));
    return 0;
}

int step_eager(sphb_slabs *g, double remain)
{
    for (OneSlab &o : g->s) SPHB_TRY(sphb_ctx_slab_step_begin(o.ctx, remain));
    SPHB_TRY(exchange(g));
    for (OneSlab &o : g->s) SPHB_TRY(sphb_ctx_slab_step_end(o.ctx));
    return 0;
}

// one step with remain = inf as one graph launch per slab (captured on first use; every slab on its own device)
int step_graph(sphb_slabs *g)
{
    const double inf = std::numeric_limits<double>::infinity();
    const bool squeeze_now = g->compact_every > 0 && (g->k + 1) % g->compact_every == 0;
    if (!g->one_per_device || !g->graphs_ok || squeeze_now) {
        if (squeeze_now) drop_graphs(g);   // the compaction swaps the state buffers
        return step_eager(g, inf);
    }
    if (!g->s[0].graph) {
        // capture the four calls of every slab on its own stream; nothing runs during the capture
        for (OneSlab &o : g->s) {
            cudaStream_t st = (cudaStream_t)sphb_ctx_stream(o.ctx);
            SPHB_CUDA(cudaSetDevice(o.device));
            SPHB_CUDA(cudaStreamSynchronize(st));
            cudaGraph_t graph = nullptr;
            SPHB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            int rc = sphb_ctx_slab_step_begin(o.ctx, inf);
            if (!rc) rc = sphb_ctx_slab_p2p_send(o.ctx, o.send_left, o.send_right);
            if (!rc) rc = sphb_ctx_slab_p2p_recv(o.ctx, o.send_left, o.send_right);
            if (!rc) rc = sphb_ctx_slab_step_end(o.ctx);
            const cudaError_t e = cudaStreamEndCapture(st, &graph);
            if (rc || e != cudaSuccess || !graph ||
                cudaGraphInstantiate(&o.graph, graph, 0) != cudaSuccess) {   // not capturable here: stay eager
                if (graph) cudaGraphDestroy(graph);
                (void)cudaGetLastError();
                drop_graphs(g);
                g->graphs_ok = false;
                clear_error();
                return step_eager(g, inf);
            }
            cudaGraphDestroy(graph);
        }
    }
    g->k += 1;
    for (OneSlab &o : g->s) {
        SPHB_CUDA(cudaSetDevice(o.device));
        SPHB_CUDA(cudaGraphLaunch(o.graph, (cudaStream_t)sphb_ctx_stream(o.ctx)));
    }
    return 0;
}

}  // namespace

extern "C" {

SPHB_API int sphb_slabs_create(const sphb_params *prm, int n_slabs, const int *devices, int n_devices, const double *pos,
                               const double *vel, const double *rho, const double *drho_dt, const double *mass,
                               const double *wall_vel, sphb_slabs **out)
{
    clear_error();
    if (!prm || !out || !pos || !vel || !rho || !drho_dt || !mass) return set_error("SPHB:arg", "null argument");
    *out = nullptr;
    const int nf = prm->n_fluid, nt = prm->n_total, nw_all = nt - nf;
    if (nf < 1 || nw_all < 0 || !(prm->h > 0.0) || !(prm->DL > 0.0)) return set_error("SPHB:arg", "bad channel description");
    const double cell = 2.0 * prm->h;
    const int ncx = std::max(1, (int)std::ceil(prm->DL / cell));   // nb:254
    if (n_slabs < 2 || ncx / n_slabs < kHaloCols || ncx < ncx / n_slabs + 1 + 2 * kHaloCols + 2)
        return set_error("SPHB:arg", "%d slabs of %d halo columns do not fit a channel of %d cell columns", n_slabs, kHaloCols, ncx);
    int visible = 0;
    SPHB_CUDA(cudaGetDeviceCount(&visible));
    if (visible < 1) return set_error("SPHB:cuda", "no CUDA device");
    std::vector<int> devs;
    if (devices && n_devices > 0) devs.assign(devices, devices + n_devices);
    else
        for (int d = 0; d < visible; ++d) devs.push_back(d);
    for (int d : devs)
        if (d < 0 || d >= visible) return set_error("SPHB:arg", "device %d is not visible (%d devices)", d, visible);

    sphb_slabs *g = new sphb_slabs();
    g->prm = *prm;
    g->n = n_slabs;
    g->ncx = ncx;
    g->s.resize((size_t)n_slabs);
    g->one_per_device = (int)devs.size() >= n_slabs;
    for (int r = 0; r < n_slabs && g->one_per_device; ++r)
        for (int q = 0; q < r; ++q)
            if (devs[(size_t)r] == devs[(size_t)q]) g->one_per_device = false;
    const double per_col = std::max(1.0, (double)nf / ncx);
    g->rec_cap = (int)((kHaloCols + 2) * per_col * 1.5) + 256;
    // static wall rows of the channel (download returns them as given)
    g->wall_pos.resize(2 * (size_t)nw_all);
    g->wall_vel.assign(2 * (size_t)nw_all, 0.0);
    g->wall_mass.resize((size_t)nw_all);
    for (int w = 0; w < nw_all; ++w) {
        g->wall_pos[(size_t)w] = pos[nf + w];
        g->wall_pos[(size_t)nw_all + w] = pos[(size_t)nt + nf + w];
        g->wall_mass[(size_t)w] = mass[nf + w];
        if (wall_vel) {
            g->wall_vel[(size_t)w] = wall_vel[nf + w];
            g->wall_vel[(size_t)nw_all + w] = wall_vel[(size_t)nt + nf + w];
        }
    }
    // columns of every row
    std::vector<int> col((size_t)nt);
    for (int i = 0; i < nt; ++i) col[(size_t)i] = wrapped_column(pos[i], prm->DL, cell, ncx);

    int rc = 0;
    const int base = ncx / n_slabs, extra = ncx % n_slabs;
    int lo = 0;
    for (int r = 0; r < n_slabs && !rc; ++r) {
        OneSlab &o = g->s[(size_t)r];
        const int hi = lo + base + (r < extra ? 1 : 0);
        o.col_lo = lo;
        o.col_hi = hi;
        o.device = devs[(size_t)r % devs.size()];
        const int width = hi - lo;
        // owned fluid rows, in row order
        std::vector<int> own;
        for (int i = 0; i < nf; ++i)
            if (col[(size_t)i] >= lo && col[(size_t)i] < hi) own.push_back(i);
        // wall rows of the slab and its halo columns, unwrapped around the slab
        std::vector<int> widx;
        std::vector<double> wx;
        for (int w = 0; w < nw_all; ++w) {
            const int cw = col[(size_t)nf + w];
            int rel = floor_mod(cw - lo + kHaloCols, ncx) - kHaloCols;
            bool keep = rel < width + kHaloCols;
            if (!keep && lo == 0) {
                // towards the periodic seam the halo is kHaloCols CELLS of x, not columns: the channel's last column is
                // a sliver unless DL is a whole number of cells (slab.cu slab_class has the same rule for the copies)
                const double x = pos[nf + w];
                const double xw = x - std::floor(x / prm->DL) * prm->DL;
                if (xw >= prm->DL - kHaloCols * cell) {
                    rel = cw - ncx;
                    keep = true;
                }
            }
            if (!keep) continue;
            const int kshift = floor_div(lo + rel - cw, ncx);   // -1, 0 or +1 periods
            widx.push_back(w);
            wx.push_back(pos[nf + w] + kshift * prm->DL);
        }
        const int n_owned = (int)own.size(), n_wall = (int)widx.size();
        o.n_wall = n_wall;
        o.nf_cap = (int)((n_owned + 2.0 * (kHaloCols + 1) * per_col) * 1.25) + 1024;
        o.wall_ids.resize((size_t)n_wall);
        // column-major blocks for sphb_ctx_slab_init
        std::vector<double> p2(2 * (size_t)n_owned), v2(2 * (size_t)n_owned), r1((size_t)n_owned), d1((size_t)n_owned),
            m1((size_t)n_owned), wp(2 * (size_t)n_wall), wm((size_t)n_wall), wv(2 * (size_t)n_wall, 0.0);
        std::vector<int> ids((size_t)n_owned);
        for (int a = 0; a < n_owned; ++a) {
            const int i = own[(size_t)a];
            ids[(size_t)a] = i;
            p2[(size_t)a] = pos[i];
            p2[(size_t)n_owned + a] = pos[(size_t)nt + i];
            v2[(size_t)a] = vel[i];
            v2[(size_t)n_owned + a] = vel[(size_t)nt + i];
            r1[(size_t)a] = rho[i];
            d1[(size_t)a] = drho_dt[i];
            m1[(size_t)a] = mass[i];
        }
        for (int a = 0; a < n_wall; ++a) {
            const int w = widx[(size_t)a];
            o.wall_ids[(size_t)a] = nf + w;
            wp[(size_t)a] = wx[(size_t)a];
            wp[(size_t)n_wall + a] = pos[(size_t)nt + nf + w];
            wm[(size_t)a] = mass[nf + w];
            if (wall_vel) {
                wv[(size_t)a] = wall_vel[nf + w];
                wv[(size_t)n_wall + a] = wall_vel[(size_t)nt + nf + w];
            }
        }
        sphb_params sp = *prm;
        sp.n_fluid = o.nf_cap;
        sp.n_total = o.nf_cap + n_wall;
        sp.sort_interval = 0;   // the now-and-then compaction is the slab form of the cell sort (slab.cu)
        sp.device = o.device;
        rc = sphb_ctx_create(&sp, &o.ctx);
        if (!rc)
            rc = sphb_ctx_slab_init(o.ctx, lo, hi, kHaloCols, n_owned, p2.data(), v2.data(), r1.data(), d1.data(), m1.data(),
                                    ids.data(), n_wall, wp.data(), wm.data(), wv.data());
        if (!rc) {
            const size_t nb = ((size_t)kHdrD + (size_t)g->rec_cap * kRecD) * sizeof(double);
            if (cudaSetDevice(o.device) != cudaSuccess || cudaMalloc((void **)&o.send_left, nb) != cudaSuccess ||
                cudaMalloc((void **)&o.send_right, nb) != cudaSuccess || cudaMemset(o.send_left, 0, nb) != cudaSuccess ||
                cudaMemset(o.send_right, 0, nb) != cudaSuccess)
                rc = set_error("SPHB:cuda", "exchange buffers of slab %d: %s", r, cudaGetErrorString(cudaGetLastError()));
        }
        if (!rc) {
            size_t bytes = 0;
            unsigned char handle[64];
            rc = sphb_ctx_slab_p2p_create(o.ctx, r, n_slabs, g->rec_cap, &o.mailbox, &bytes, handle);
        }
        lo = hi;
    }
    // every slab learns every mailbox (scalars go to everybody, records to the two neighbours)
    for (int r = 0; r < n_slabs && !rc; ++r)
        for (int q = 0; q < n_slabs && !rc; ++q)
            rc = sphb_ctx_slab_p2p_connect(g->s[(size_t)r].ctx, q, nullptr, g->s[(size_t)q].mailbox, g->s[(size_t)q].device);
    // the first exchange brings the halo copies; then the first list and the shear monitor of the initial state
    if (!rc) rc = exchange(g);
    for (int r = 0; r < n_slabs && !rc; ++r) rc = sphb_ctx_slab_step_end(g->s[(size_t)r].ctx);
    if (rc) {
        // keep the callee's error text across the clean-up (the calls below clear it)
        const std::string id = sphb_last_error_id() ? sphb_last_error_id() : "SPHB:error";
        const std::string msg = sphb_last_error_msg() ? sphb_last_error_msg() : "";
        sphb_slabs_destroy(g);
        set_error(id.c_str(), "%s", msg.c_str());
        return rc;
    }
    *out = g;
    return 0;
}

SPHB_API void sphb_slabs_destroy(sphb_slabs *g)
{
    if (!g) return;
    drop_graphs(g);
    for (OneSlab &o : g->s) {
        if (o.ctx) (void)sphb_ctx_sync(o.ctx);
    }
    for (OneSlab &o : g->s) {
        if (o.send_left || o.send_right) {
            (void)cudaSetDevice(o.device);
            if (o.send_left) cudaFree(o.send_left);
            if (o.send_right) cudaFree(o.send_right);
        }
        if (o.ctx) sphb_ctx_destroy(o.ctx);
    }
    delete g;
}

SPHB_API int sphb_slabs_count(const sphb_slabs *g) { return g ? g->n : 0; }

SPHB_API int sphb_slabs_set_compaction(sphb_slabs *g, int every_n_steps)
{
    clear_error();
    if (!g) return set_error("SPHB:arg", "null group");
    g->compact_every = every_n_steps > 0 ? every_n_steps : 0;
    return 0;
}

SPHB_API sphb_ctx *sphb_slabs_member(sphb_slabs *g, int r, int *col_lo, int *col_hi, int *device)
{
    if (!g || r < 0 || r >= g->n) return nullptr;
    if (col_lo) *col_lo = g->s[(size_t)r].col_lo;
    if (col_hi) *col_hi = g->s[(size_t)r].col_hi;
    if (device) *device = g->s[(size_t)r].device;
    return g->s[(size_t)r].ctx;
}

SPHB_API int sphb_slabs_step(sphb_slabs *g, double remain)
{
    clear_error();
    if (!g) return set_error("SPHB:arg", "null group");
    if (remain < std::numeric_limits<double>::infinity()) return step_eager(g, remain);
    return step_graph(g);
}

SPHB_API int sphb_slabs_run(sphb_slabs *g, long long n_steps)
{
    clear_error();
    if (!g) return set_error("SPHB:arg", "null group");
    for (long long k = 0; k < n_steps; ++k) SPHB_TRY(step_graph(g));
    return 0;
}

SPHB_API int sphb_slabs_sync(sphb_slabs *g)
{
    clear_error();
    if (!g) return set_error("SPHB:arg", "null group");
    for (OneSlab &o : g->s) SPHB_TRY(sphb_ctx_sync(o.ctx));
    return 0;
}

// Clock of slab 0 (all slabs carry the same one: the time step follows the channel-wide max|v|), monitors of the whole
// channel reduced on the host in slab order (MAX / SUM: deterministic), error flags OR-ed, pair counts summed (they
// include the pairs of the halo copies).
SPHB_API int sphb_slabs_status(sphb_slabs *g, sphb_status *out)
{
    clear_error();
    if (!g || !out) return set_error("SPHB:arg", "null argument");
    sphb_status acc{};
    for (int r = 0; r < g->n; ++r) {
        sphb_status s{};
        SPHB_TRY(sphb_ctx_status(g->s[(size_t)r].ctx, &s));
        if (r == 0) acc = s;
        else {
            acc.vmax = std::max(acc.vmax, s.vmax);
            acc.tau_bottom += s.tau_bottom;
            acc.tau_top += s.tau_top;
            acc.n_half_pairs += s.n_half_pairs;
            acc.n_csr_entries += s.n_csr_entries;
            acc.error_flags |= s.error_flags;
        }
    }
    *out = acc;
    return 0;
}

SPHB_API int sphb_slabs_run_until(sphb_slabs *g, double t_target, double t_end, long long max_steps, long long *steps_done)
{
    clear_error();
    if (!g) return set_error("SPHB:arg", "null group");
    const sphb_params &p = g->prm;
    const double nu = p.mu / p.rho0;
    const double dt_max = std::fmin(std::fmin(0.25 * p.h / std::fmax(p.c_f, 1e-12), 0.125 * p.h * p.h / std::fmax(nu, 1e-12)),
                                    0.25 * std::sqrt(p.h / std::fmax(std::fabs(p.gravity_g), 1e-12)));
    long long n = 0;
    sphb_status s{};
    SPHB_TRY(sphb_ctx_status(g->s[0].ctx, &s));
    while (s.t < t_target - 1e-12 && (max_steps <= 0 || n < max_steps)) {   // m:250
        double safe = std::floor((std::fmin(t_target, t_end) - s.t) / dt_max) - 1.0;
        if (max_steps > 0) safe = std::fmin(safe, (double)(max_steps - n));
        const long long step0 = s.step;
        if (safe >= 2.0) {
            SPHB_TRY(sphb_slabs_run(g, (long long)std::fmin(safe, 4096.0)));
        } else {
            SPHB_TRY(step_eager(g, std::fmin(t_target - s.t, t_end - s.t)));   // m:252, 259
        }
        SPHB_TRY(sphb_slabs_status(g, &s));
        if (s.error_flags)
            return set_error("SPHB:capacity", "device error flags %d at step %lld (list, tile or exchange-buffer capacity of a slab)",
                             s.error_flags, s.step);
        if (s.dt < 1e-14) return set_error("SPHB:diverged", "time step collapsed (dt=%g) at t=%g step=%lld", s.dt, s.t, s.step);
        if (s.step == step0) return set_error("SPHB:state", "the sharded channel did not advance at step %lld", s.step);
        n += s.step - step0;
    }
    if (steps_done) *steps_done = n;
    return 0;
}

// Fluid rows by GLOBAL row (the order of the arrays given to sphb_slabs_create), the static wall rows behind them:
// column-major n_total arrays like sphb_ctx_download; any pointer may be NULL.
SPHB_API int sphb_slabs_download(sphb_slabs *g, double *pos, double *vel, double *rho, double *p, double *drho_dt)
{
    clear_error();
    if (!g) return set_error("SPHB:arg", "null group");
    const int nf = g->prm.n_fluid, nt = g->prm.n_total, nw_all = nt - nf;
    std::vector<char> seen((size_t)nf, 0);
    for (OneSlab &o : g->s) {
        int n_owned = 0, n_halo = 0, n_wall = 0;
        SPHB_TRY(sphb_ctx_slab_counts(o.ctx, &n_owned, &n_halo, &n_wall));
        const size_t cap = (size_t)o.nf_cap + (size_t)o.n_wall;
        std::vector<double> bp(pos ? 2 * cap : 0), bv(vel ? 2 * cap : 0), br(rho ? cap : 0), bq(p ? cap : 0), bd(drho_dt ? cap : 0);
        SPHB_TRY(sphb_ctx_download(o.ctx, pos ? bp.data() : nullptr, vel ? bv.data() : nullptr, rho ? br.data() : nullptr,
                                   p ? bq.data() : nullptr, drho_dt ? bd.data() : nullptr, nullptr, nullptr, nullptr, nullptr));
        std::vector<int> ids((size_t)(n_owned + n_halo) + 1);
        SPHB_TRY(sphb_ctx_slab_ids(o.ctx, ids.data()));
        for (int a = 0; a < n_owned; ++a) {
            const int i = ids[(size_t)a];
            if (i < 0) continue;   // a row that migrated away and awaits compaction
            if (i >= nf || seen[(size_t)i]) return set_error("SPHB:state", "row %d is owned twice or out of range", i);
            seen[(size_t)i] = 1;
            if (pos) {
                pos[i] = bp[(size_t)a];
                pos[(size_t)nt + i] = bp[cap + a];
            }
            if (vel) {
                vel[i] = bv[(size_t)a];
                vel[(size_t)nt + i] = bv[cap + a];
            }
            if (rho) rho[i] = br[(size_t)a];
            if (p) p[i] = bq[(size_t)a];
            if (drho_dt) drho_dt[i] = bd[(size_t)a];
        }
    }
    for (int i = 0; i < nf; ++i)
        if (!seen[(size_t)i]) return set_error("SPHB:state", "row %d is owned by no slab", i);
    for (int w = 0; w < nw_all; ++w) {
        if (pos) {
            pos[nf + w] = g->wall_pos[(size_t)w];
            pos[(size_t)nt + nf + w] = g->wall_pos[(size_t)nw_all + w];
        }
        if (vel) vel[nf + w] = vel[(size_t)nt + nf + w] = 0.0;
        if (rho) rho[nf + w] = g->prm.rho0;
        if (p) p[nf + w] = 0.0;
        if (drho_dt) drho_dt[nf + w] = 0.0;
    }
    return 0;
}

// Binned u(y) monitor of the whole channel (m:579-605; sphb_ctx_profile for one context): the fluid rows are gathered
// and binned on the host -- an output-point operation, not part of the step.
SPHB_API int sphb_slabs_profile(sphb_slabs *g, double x0, double half_width, int n_bins, double *u_mean, double *count)
{
    clear_error();
    if (!g || n_bins < 1 || !u_mean) return set_error("SPHB:arg", "bad profile arguments");
    const int nf = g->prm.n_fluid, nt = g->prm.n_total;
    const double DL = g->prm.DL, DH = g->prm.DH;
    std::vector<double> pos(2 * (size_t)nt), vel(2 * (size_t)nt), edges((size_t)n_bins + 1), sum((size_t)n_bins, 0.0),
        cnt((size_t)n_bins, 0.0);
    SPHB_TRY(sphb_slabs_download(g, pos.data(), vel.data(), nullptr, nullptr, nullptr));
    for (int k = 0; k <= n_bins; ++k) edges[(size_t)k] = (k == n_bins) ? DH : (double)k * (DH / (double)n_bins);   // m:581
    for (int i = 0; i < nf; ++i) {
        if (x0 >= 0.0) {
            double xw = pos[(size_t)i];
            if (!(xw >= 0.0 && xw < DL)) {   // MATLAB mod(x, DL), m:573
                xw = std::fmod(xw, DL);
                if (xw < 0.0) xw += DL;
            }
            double d = std::fabs(xw - x0);
            d = std::fmin(d, DL - d);
            if (!(d <= half_width)) continue;
        }
        const double y = pos[(size_t)nt + i];
        if (!(y >= edges[0] && y <= edges[(size_t)n_bins])) continue;   // discretize(): out of range dropped
        int b = (int)std::floor((y - edges[0]) / (edges[(size_t)n_bins] - edges[0]) * n_bins);
        b = std::min(std::max(b, 0), n_bins - 1);
        while (b > 0 && y < edges[(size_t)b]) --b;
        while (b < n_bins - 1 && y >= edges[(size_t)b + 1]) ++b;
        sum[(size_t)b] += vel[(size_t)i];
        cnt[(size_t)b] += 1.0;
    }
    for (int k = 0; k < n_bins; ++k) {
        u_mean[k] = cnt[(size_t)k] > 0.0 ? sum[(size_t)k] / cnt[(size_t)k] : std::numeric_limits<double>::quiet_NaN();   // m:588-589
        if (count) count[k] = cnt[(size_t)k];
    }
    return 0;
}

SPHB_API int sphb_slabs_set_time(sphb_slabs *g, double t, long long step)
{
    clear_error();
    if (!g) return set_error("SPHB:arg", "null group");
    for (OneSlab &o : g->s) SPHB_TRY(sphb_ctx_set_time(o.ctx, t, step));
    return 0;
}

}  // extern "C"
