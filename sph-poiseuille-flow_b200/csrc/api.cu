// api.cu -- the STATELESS C-ABI entry points (family 1 of include/sph_b200.h): one function per MEX
// gateway/mode of the reference, host pointers in, host pointers out, nothing persists.
//
// Each function (a) performs the reference's argument checks in the reference's order and reports the
// reference's MATLAB error identifier, (b) uploads the inputs, (c) runs the
// hand-written kernels of ops_pairs.cu / neighbor.cu on a private stream, (d) copies the outputs
// straight into the caller's arrays.  Device workspaces are grow-only and shared by all calls under a
// mutex (MATLAB calls MEX functions from one interpreter thread).  No CPU fallback exists.
#include <climits>
#include <mutex>

#include "common.cuh"
#include "ops_pairs.cuh"

using namespace sphb;

struct sphb_pairs {
    size_t n_pairs = 0;
    unsigned long long generation = 0;   // the search this handle belongs to (Stateless::generation)
};

namespace {

struct Stateless {
    std::mutex mu;
    cudaStream_t st = nullptr;
    DevBuf arena;         // inputs + outputs + scratch of one call
    DevBuf pos2;          // double2 positions for the neighbour search
    NeighborWork nw;
    bool pairs_pending = false;
    unsigned long long generation = 0;   // bumped by every search: a handle fetches only its own list
};

// sphb_validate_only(1): the operator entry points return after the reference's argument checks, before any device
// work (and before they look at their output pointers).  Per calling thread.
thread_local int g_validate_only = 0;

Stateless &G()
{
    static Stateless g;
    return g;
}

int ensure_stream(Stateless &g)
{
    SPHB_TRY(ensure_device());
    if (!g.st) SPHB_CUDA(cudaStreamCreateWithFlags(&g.st, cudaStreamNonBlocking));
    return 0;
}

// bump allocator over the arena, in doubles
struct Bump {
    double *base = nullptr;
    size_t used = 0, cap = 0;
    double *take(size_t n)
    {
        double *p = base + used;
        used += (n + 3) & ~(size_t)3;   // keep 32-byte alignment
        return p;
    }
};

int open_arena(Stateless &g, size_t doubles, Bump &b)
{
    SPHB_TRY(g.arena.reserve((doubles + 64) * sizeof(double)));
    b.base = g.arena.as<double>();
    b.cap = doubles + 64;
    b.used = 0;
    return 0;
}

inline size_t pad4(size_t n) { return (n + 3) & ~(size_t)3; }

int upload(Stateless &g, Bump &b, const double *src, size_t n, const double **dev)
{
    double *d = b.take(n);
    if (n) SPHB_CUDA(cudaMemcpyAsync(d, src, n * sizeof(double), cudaMemcpyHostToDevice, g.st));
    *dev = d;
    return 0;
}

int download(Stateless &g, double *dst, const double *dev, size_t n)
{
    if (n && dst) SPHB_CUDA(cudaMemcpyAsync(dst, dev, n * sizeof(double), cudaMemcpyDeviceToHost, g.st));
    return 0;
}

int finish(Stateless &g)
{
    SPHB_CUDA(cudaStreamSynchronize(g.st));
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

inline size_t numel(const sphb_mat &a) { return a.m * a.n; }
inline bool is_mat(const sphb_mat &a, size_t m, size_t n) { return a.m == m && a.n == n; }

// get_pair_data, ph:49-74
bool pairs_consistent(const sphb_mat &pi, const sphb_mat &pj, const sphb_mat &dx, const sphb_mat &dy,
                      const sphb_mat &r, const sphb_mat &dW)
{
    const size_t n = numel(pi);
    return numel(pj) == n && numel(dx) == n && numel(dy) == n && numel(r) == n && numel(dW) == n;
}

struct DevPairs {
    PairView pv;
};

int upload_pairs(Stateless &g, Bump &b, const sphb_mat &pi, const sphb_mat &pj, const sphb_mat &dx,
                 const sphb_mat &dy, const sphb_mat &r, const sphb_mat *W, const sphb_mat &dW, PairView &pv)
{
    const size_t n = numel(pi);
    pv.n = n;
    SPHB_TRY(upload(g, b, pi.p, n, &pv.pi));
    SPHB_TRY(upload(g, b, pj.p, n, &pv.pj));
    SPHB_TRY(upload(g, b, dx.p, n, &pv.dx));
    SPHB_TRY(upload(g, b, dy.p, n, &pv.dy));
    SPHB_TRY(upload(g, b, r.p, n, &pv.r));
    pv.W = nullptr;
    if (W) SPHB_TRY(upload(g, b, W->p, n, &pv.W));
    SPHB_TRY(upload(g, b, dW.p, n, &pv.dW));
    return 0;
}

#define ENTER()                                   \
    clear_error();                                \
    Stateless &g = G();                           \
    std::lock_guard<std::mutex> lock__(g.mu)

inline size_t scratch_doubles(int nt) { return pad4((size_t)kScratchCols * (size_t)nt + 4096); }

}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------
// neighbour search (nb:185-421)
// ------------------------------------------------------------------------------------------------
SPHB_API int sphb_neighbor_search(sphb_mat pos, int n_fluid, int n_total, double h, double DL, sphb_pairs **handle,
                                  size_t *n_pairs)
{
    ENTER();
    if (!handle || !n_pairs) return set_error("SPHB:arg", "null output argument");
    *handle = nullptr;
    *n_pairs = 0;
    SPHB_REQUIRE(pos.p != nullptr && pos.n == 2, "SPH:Neighbor:pos", "pos must be a double matrix of size [n_total x 2].");
    SPHB_REQUIRE(!(n_total <= 0 || n_fluid <= 0 || n_fluid > n_total || (size_t)n_total != pos.m),
                 "SPH:Neighbor:count", "Invalid n_fluid/n_total or inconsistent pos size.");
    SPHB_REQUIRE(!(h <= 0.0 || DL <= 0.0), "SPH:Neighbor:param", "h and DL must be positive.");
    if (g_validate_only) return 0;
    SPHB_TRY(ensure_stream(g));
    const size_t n = (size_t)n_total;
    double lo = pos.p[n], hi = pos.p[n];
    for (size_t i = 1; i < n; ++i) {   // sizes the cell arrays; the grid itself is derived on the device
        lo = pos.p[n + i] < lo ? pos.p[n + i] : lo;
        hi = pos.p[n + i] > hi ? pos.p[n + i] : hi;
    }
    SPHB_TRY(neighbor_plan(g.nw, n_fluid, n_total, h, DL, lo, hi, 0, 0));
    SPHB_CUDA(cudaMemsetAsync(g.nw.flags.p, 0, sizeof(int), g.st));
    Bump b;
    SPHB_TRY(open_arena(g, pad4(2 * n), b));
    const double *d_pos = nullptr;
    SPHB_TRY(upload(g, b, pos.p, 2 * n, &d_pos));
    SPHB_TRY(g.pos2.reserve(n * sizeof(double2)));
    launch_pack2(d_pos, g.pos2.as<double2>(), n_total, g.st);
    // half list only (the reference's own), ELL grown to the longest row: any input is legal here
    SPHB_TRY(neighbor_build(g.nw, g.pos2.as<double2>(), true, true, g.st, nullptr));
    int host[3] = {0, 0, 0};
    SPHB_CUDA(cudaMemcpyAsync(host, g.nw.totals.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, g.st));
    SPHB_CUDA(cudaMemcpyAsync(host + 2, g.nw.flags.p, sizeof(int), cudaMemcpyDeviceToHost, g.st));
    SPHB_TRY(finish(g));
    if (host[2] & (kErrCellCap | kErrCsrCap | kErrSeenCap))
        return set_error("SPH:Neighbor:entryCapacity", "Cell entry capacity exceeded.");
    sphb_pairs *hnd = new sphb_pairs();
    hnd->n_pairs = (size_t)host[0];
    hnd->generation = ++g.generation;
    g.pairs_pending = true;
    *handle = hnd;
    *n_pairs = hnd->n_pairs;
    return 0;
}

SPHB_API int sphb_neighbor_fetch(sphb_pairs *handle, double *pair_i, double *pair_j, double *dx, double *dy,
                                 double *r, double *W, double *dW)
{
    ENTER();
    if (!handle) return set_error("SPHB:arg", "null handle");
    if (!g.pairs_pending || handle->generation != g.generation) {
        delete handle;
        return set_error("SPHB:state", "the pair list of this handle has been overwritten by a later search");
    }
    const size_t n = handle->n_pairs;
    int rc = 0;
    if (n) {
        Bump b;
        rc = open_arena(g, 7 * pad4(n), b);
        double *d[7];
        for (int k = 0; k < 7 && !rc; ++k) d[k] = b.take(n);
        if (!rc) rc = neighbor_export_pairs(g.nw, n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], g.st);
        double *dst[7] = {pair_i, pair_j, dx, dy, r, W, dW};
        for (int k = 0; k < 7 && !rc; ++k) rc = download(g, dst[k], d[k], n);
        if (!rc) rc = finish(g);
    }
    g.pairs_pending = false;
    delete handle;
    return rc;
}

SPHB_API void sphb_neighbor_release(sphb_pairs *handle)
{
    Stateless &g = G();
    std::lock_guard<std::mutex> lock(g.mu);
    if (handle && handle->generation == g.generation) g.pairs_pending = false;   // (an older handle owns nothing any more)
    delete handle;
}

SPHB_API void sphb_validate_only(int on) { g_validate_only = on ? 1 : 0; }

// ------------------------------------------------------------------------------------------------
// density_correction (ph:95-374; checks ph:139-162)
// ------------------------------------------------------------------------------------------------
SPHB_API int sphb_density_correction(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                     sphb_mat W, sphb_mat dW, sphb_mat mass, int n_fluid, int n_total, double rho0,
                                     double h, double inv_sigma0, double *rho, double *Vol, double *B)
{
    ENTER();
    const size_t np = numel(pair_i);
    SPHB_REQUIRE(numel(pair_j) == np && numel(dx) == np && numel(dy) == np && numel(r) == np && numel(W) == np &&
                     numel(dW) == np,
                 "SPH:Physics:density:pairs", "Pair arrays mismatch.");
    SPHB_REQUIRE(np <= (size_t)INT_MAX, "SPH:Physics:density:pairsize", "Pair count exceeds INT_MAX.");
    SPHB_REQUIRE(n_fluid > 0 && n_total >= n_fluid, "SPH:Physics:density:count", "Invalid n_fluid/n_total.");
    SPHB_REQUIRE((long long)numel(mass) == (long long)n_total, "SPH:Physics:density:mass", "mass size mismatch.");
    SPHB_REQUIRE(rho0 > 0.0 && h > 0.0, "SPH:Physics:density:param", "rho0 and h must be positive.");
    if (g_validate_only) return 0;
    SPHB_TRY(ensure_stream(g));
    const size_t nt = (size_t)n_total;
    Bump b;
    SPHB_TRY(open_arena(g, 7 * pad4(np) + 7 * pad4(nt) + scratch_doubles(n_total), b));
    PairView pv;
    SPHB_TRY(upload_pairs(g, b, pair_i, pair_j, dx, dy, r, &W, dW, pv));
    const double *d_mass;
    SPHB_TRY(upload(g, b, mass.p, nt, &d_mass));
    double *d_rho = b.take(nt), *d_Vol = b.take(nt), *d_B = b.take(4 * nt), *scr = b.take(scratch_doubles(n_total));
    SPHB_TRY(pairs_density(pv, d_mass, n_fluid, n_total, rho0, h, inv_sigma0, d_rho, d_Vol, d_B, scr, g.st));
    SPHB_TRY(download(g, rho, d_rho, nt));
    SPHB_TRY(download(g, Vol, d_Vol, nt));
    SPHB_TRY(download(g, B, d_B, 4 * nt));
    return finish(g);
}

// ------------------------------------------------------------------------------------------------
// viscous_force (ph:396-550; checks ph:430-451)
// ------------------------------------------------------------------------------------------------
SPHB_API int sphb_viscous_force(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r, sphb_mat dW,
                                sphb_mat vel, sphb_mat Vol, sphb_mat B, double mu, double h, int n_fluid, int n_total,
                                sphb_mat mass, sphb_mat wall_vel, double *force)
{
    ENTER();
    SPHB_REQUIRE(pairs_consistent(pair_i, pair_j, dx, dy, r, dW), "SPH:Physics:pairs", "Pair arrays must have same length.");
    SPHB_REQUIRE(n_total > 0, "SPHB:arg", "n_total must be positive.");
    const size_t nt = (size_t)n_total;
    SPHB_REQUIRE(is_mat(vel, nt, 2), "SPH:Physics:viscous:vel", "vel size mismatch.");
    SPHB_REQUIRE(numel(Vol) == nt, "SPH:Physics:viscous:Vol", "Vol size mismatch.");
    SPHB_REQUIRE(is_mat(B, nt, 4), "SPH:Physics:viscous:B", "B size mismatch.");
    SPHB_REQUIRE(numel(mass) == nt, "SPH:Physics:viscous:mass", "mass size mismatch.");
    SPHB_REQUIRE(is_mat(wall_vel, nt, 2), "SPH:Physics:viscous:wallvel", "wall_vel size mismatch.");
    if (g_validate_only) return 0;
    SPHB_TRY(ensure_stream(g));
    const size_t np = numel(pair_i);
    Bump b;
    SPHB_TRY(open_arena(g, 6 * pad4(np) + 12 * pad4(nt) + scratch_doubles(n_total), b));
    PairView pv;
    SPHB_TRY(upload_pairs(g, b, pair_i, pair_j, dx, dy, r, nullptr, dW, pv));
    const double *d_vel, *d_Vol, *d_B, *d_wv;
    SPHB_TRY(upload(g, b, vel.p, 2 * nt, &d_vel));
    SPHB_TRY(upload(g, b, Vol.p, nt, &d_Vol));
    SPHB_TRY(upload(g, b, B.p, 4 * nt, &d_B));
    SPHB_TRY(upload(g, b, wall_vel.p, 2 * nt, &d_wv));
    double *d_force = b.take(2 * nt), *scr = b.take(scratch_doubles(n_total));
    SPHB_TRY(pairs_viscous(pv, d_vel, d_Vol, d_B, mu, h, n_fluid, n_total, d_wv, d_force, scr, g.st));
    SPHB_TRY(download(g, force, d_force, 2 * nt));
    return finish(g);
}

// ------------------------------------------------------------------------------------------------
// transport_correction (ph:569-714; checks ph:599-618)
// ------------------------------------------------------------------------------------------------
SPHB_API int sphb_transport_correction(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                       sphb_mat dW, sphb_mat Vol, sphb_mat B, sphb_mat pos, double h, int n_fluid,
                                       int n_total, int has_coeff, double transport_coeff, double *pos_out)
{
    ENTER();
    SPHB_REQUIRE(pairs_consistent(pair_i, pair_j, dx, dy, r, dW), "SPH:Physics:pairs", "Pair arrays must have same length.");
    const double coeff = has_coeff ? transport_coeff : 0.2;   // ph:584, 607-609
    SPHB_REQUIRE(coeff >= 0.0, "SPH:Physics:transport:coeff", "transport_coeff must be non-negative.");
    SPHB_REQUIRE(n_total > 0, "SPHB:arg", "n_total must be positive.");
    const size_t nt = (size_t)n_total;
    SPHB_REQUIRE(numel(Vol) == nt, "SPH:Physics:transport:Vol", "Vol size mismatch.");
    SPHB_REQUIRE(is_mat(B, nt, 4), "SPH:Physics:transport:B", "B size mismatch.");
    SPHB_REQUIRE(is_mat(pos, nt, 2), "SPH:Physics:transport:pos", "pos size mismatch.");
    if (g_validate_only) return 0;
    SPHB_TRY(ensure_stream(g));
    const size_t np = numel(pair_i);
    Bump b;
    SPHB_TRY(open_arena(g, 6 * pad4(np) + 10 * pad4(nt) + scratch_doubles(n_total), b));
    PairView pv;
    SPHB_TRY(upload_pairs(g, b, pair_i, pair_j, dx, dy, r, nullptr, dW, pv));
    const double *d_Vol, *d_B, *d_pos;
    SPHB_TRY(upload(g, b, Vol.p, nt, &d_Vol));
    SPHB_TRY(upload(g, b, B.p, 4 * nt, &d_B));
    SPHB_TRY(upload(g, b, pos.p, 2 * nt, &d_pos));
    double *d_out = b.take(2 * nt), *scr = b.take(scratch_doubles(n_total));
    SPHB_TRY(pairs_transport(pv, d_Vol, d_B, d_pos, h, n_fluid, n_total, coeff, d_out, scr, g.st));
    SPHB_TRY(download(g, pos_out, d_out, 2 * nt));
    return finish(g);
}

namespace {

// the nine size checks shared by integration_1st (tag "int1", ph:804-821) and integration_verlet
// (tag "verlet", ph:1344-1361); the two modes test them in different orders.
int check_int_inputs(bool verlet, size_t nt, const sphb_mat &Vol, const sphb_mat &B, const sphb_mat &rho,
                     const sphb_mat &mass, const sphb_mat &pos, const sphb_mat &vel, const sphb_mat &drho,
                     const sphb_mat &fp, const sphb_mat &wv)
{
    if (verlet) {
        SPHB_REQUIRE(numel(Vol) == nt, "SPH:Physics:verlet:Vol", "Vol size mismatch.");
        SPHB_REQUIRE(is_mat(B, nt, 4), "SPH:Physics:verlet:B", "B size mismatch.");
        SPHB_REQUIRE(numel(rho) == nt, "SPH:Physics:verlet:rho", "rho size mismatch.");
        SPHB_REQUIRE(numel(mass) == nt, "SPH:Physics:verlet:mass", "mass size mismatch.");
        SPHB_REQUIRE(is_mat(pos, nt, 2), "SPH:Physics:verlet:pos", "pos size mismatch.");
        SPHB_REQUIRE(is_mat(vel, nt, 2), "SPH:Physics:verlet:vel", "vel size mismatch.");
        SPHB_REQUIRE(numel(drho) == nt, "SPH:Physics:verlet:drho", "drho size mismatch.");
        SPHB_REQUIRE(is_mat(fp, nt, 2), "SPH:Physics:verlet:force_prior", "force_prior size mismatch.");
        SPHB_REQUIRE(is_mat(wv, nt, 2), "SPH:Physics:verlet:wall_vel", "wall_vel size mismatch.");
    } else {
        SPHB_REQUIRE(is_mat(B, nt, 4), "SPH:Physics:int1:B", "B size mismatch.");
        SPHB_REQUIRE(numel(Vol) == nt, "SPH:Physics:int1:Vol", "Vol size mismatch.");
        SPHB_REQUIRE(numel(rho) == nt, "SPH:Physics:int1:rho", "rho size mismatch.");
        SPHB_REQUIRE(numel(mass) == nt, "SPH:Physics:int1:mass", "mass size mismatch.");
        SPHB_REQUIRE(is_mat(pos, nt, 2), "SPH:Physics:int1:pos", "pos size mismatch.");
        SPHB_REQUIRE(is_mat(vel, nt, 2), "SPH:Physics:int1:vel", "vel size mismatch.");
        SPHB_REQUIRE(numel(drho) == nt, "SPH:Physics:int1:drho", "drho size mismatch.");
        SPHB_REQUIRE(is_mat(fp, nt, 2), "SPH:Physics:int1:force_prior", "force_prior size mismatch.");
        SPHB_REQUIRE(is_mat(wv, nt, 2), "SPH:Physics:int1:wall_vel", "wall_vel size mismatch.");
    }
    return 0;
}

struct VerletDev {
    const double *Vol, *B, *rho, *mass, *pos, *vel, *drho, *fp, *wv;
};

// integration_verlet on device arrays (ph:1386-1450): 1st stage, kick, 2nd stage, finish.
int verlet_on_device(Stateless &g, Bump &b, const PairView &pv, const VerletDev &in, double dt, int nf, int nt_i,
                     double rho0, double p0, double c_f, double **o_rho, double **o_p, double **o_pos,
                     double **o_vel, double **o_drho, double **o_force)
{
    const size_t nt = (size_t)nt_i;
    double *rho_h = b.take(nt), *p_h = b.take(nt), *pos_h = b.take(2 * nt), *force = b.take(2 * nt);
    double *diss = b.take(nt), *vel_new = b.take(2 * nt), *pos_new = b.take(2 * nt), *drho = b.take(nt);
    double *rho_new = b.take(nt), *p_new = b.take(nt), *scr = b.take(scratch_doubles(nt_i));
    SPHB_TRY(pairs_integration_1st(pv, in.Vol, in.B, in.rho, in.mass, in.pos, in.vel, in.drho, in.fp, dt, nf, nt_i,
                                   rho0, p0, c_f, rho_h, p_h, pos_h, force, diss, scr, g.st));
    SPHB_TRY(pairs_verlet_kick(in.vel, in.fp, force, in.mass, dt, nf, nt_i, vel_new, g.st));
    SPHB_TRY(pairs_integration_2nd(pv, in.Vol, rho_h, pos_h, vel_new, dt, nf, nt_i, in.wv, pos_new, drho, scr, g.st));
    SPHB_TRY(pairs_verlet_finish(rho_h, drho, dt, nf, nt_i, rho0, p0, rho_new, p_new, g.st));
    *o_rho = rho_new;
    *o_p = p_new;
    *o_pos = pos_new;
    *o_vel = vel_new;
    *o_drho = drho;
    *o_force = force;
    return 0;
}

inline size_t verlet_work_doubles(int nt) { return 14 * pad4((size_t)nt) + scratch_doubles(nt); }

}  // namespace

// ------------------------------------------------------------------------------------------------
// integration_1st (ph:736-967)
// ------------------------------------------------------------------------------------------------
SPHB_API int sphb_integration_1st(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r, sphb_mat dW,
                                  sphb_mat Vol, sphb_mat B, sphb_mat rho, sphb_mat mass, sphb_mat pos, sphb_mat vel,
                                  sphb_mat drho_dt, sphb_mat force_prior, double dt, int n_fluid, int n_total,
                                  double rho0, double p0, double c_f, sphb_mat wall_vel, double *rho_out,
                                  double *p_out, double *pos_out, double *force_out, double *drho_out)
{
    ENTER();
    SPHB_REQUIRE(pairs_consistent(pair_i, pair_j, dx, dy, r, dW), "SPH:Physics:pairs", "Pair arrays must have same length.");
    SPHB_REQUIRE(n_total > 0, "SPHB:arg", "n_total must be positive.");
    const size_t nt = (size_t)n_total;
    SPHB_TRY(check_int_inputs(false, nt, Vol, B, rho, mass, pos, vel, drho_dt, force_prior, wall_vel));
    if (g_validate_only) return 0;
    SPHB_TRY(ensure_stream(g));
    const size_t np = numel(pair_i);
    Bump b;
    SPHB_TRY(open_arena(g, 6 * pad4(np) + 16 * pad4(nt) + 8 * pad4(nt) + scratch_doubles(n_total), b));
    PairView pv;
    SPHB_TRY(upload_pairs(g, b, pair_i, pair_j, dx, dy, r, nullptr, dW, pv));
    VerletDev in;
    SPHB_TRY(upload(g, b, Vol.p, nt, &in.Vol));
    SPHB_TRY(upload(g, b, B.p, 4 * nt, &in.B));
    SPHB_TRY(upload(g, b, rho.p, nt, &in.rho));
    SPHB_TRY(upload(g, b, mass.p, nt, &in.mass));
    SPHB_TRY(upload(g, b, pos.p, 2 * nt, &in.pos));
    SPHB_TRY(upload(g, b, vel.p, 2 * nt, &in.vel));
    SPHB_TRY(upload(g, b, drho_dt.p, nt, &in.drho));
    SPHB_TRY(upload(g, b, force_prior.p, 2 * nt, &in.fp));
    double *d_rho = b.take(nt), *d_p = b.take(nt), *d_pos = b.take(2 * nt), *d_f = b.take(2 * nt), *d_dr = b.take(nt);
    double *scr = b.take(scratch_doubles(n_total));
    SPHB_TRY(pairs_integration_1st(pv, in.Vol, in.B, in.rho, in.mass, in.pos, in.vel, in.drho, in.fp, dt, n_fluid,
                                   n_total, rho0, p0, c_f, d_rho, d_p, d_pos, d_f, d_dr, scr, g.st));
    SPHB_TRY(download(g, rho_out, d_rho, nt));
    SPHB_TRY(download(g, p_out, d_p, nt));
    SPHB_TRY(download(g, pos_out, d_pos, 2 * nt));
    SPHB_TRY(download(g, force_out, d_f, 2 * nt));
    SPHB_TRY(download(g, drho_out, d_dr, nt));
    return finish(g);
}

// ------------------------------------------------------------------------------------------------
// integration_2nd (ph:987-1119; checks ph:1023-1043)
// ------------------------------------------------------------------------------------------------
SPHB_API int sphb_integration_2nd(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r, sphb_mat dW,
                                  sphb_mat Vol, sphb_mat rho, sphb_mat pos, sphb_mat vel, double dt, int n_fluid,
                                  int n_total, sphb_mat wall_vel, double *pos_out, double *drho_out,
                                  double *zeros_out)
{
    ENTER();
    SPHB_REQUIRE(pairs_consistent(pair_i, pair_j, dx, dy, r, dW), "SPH:Physics:pairs", "Pair arrays must have same length.");
    SPHB_REQUIRE(n_total > 0, "SPHB:arg", "n_total must be positive.");
    const size_t nt = (size_t)n_total;
    SPHB_REQUIRE(numel(Vol) == nt, "SPH:Physics:int2:Vol", "Vol size mismatch.");
    SPHB_REQUIRE(numel(rho) == nt, "SPH:Physics:int2:rho", "rho size mismatch.");
    SPHB_REQUIRE(is_mat(pos, nt, 2), "SPH:Physics:int2:pos", "pos size mismatch.");
    SPHB_REQUIRE(is_mat(vel, nt, 2), "SPH:Physics:int2:vel", "vel size mismatch.");
    SPHB_REQUIRE(is_mat(wall_vel, nt, 2), "SPH:Physics:int2:wall_vel", "wall_vel size mismatch.");
    if (g_validate_only) return 0;
    SPHB_TRY(ensure_stream(g));
    const size_t np = numel(pair_i);
    Bump b;
    SPHB_TRY(open_arena(g, 6 * pad4(np) + 12 * pad4(nt) + scratch_doubles(n_total), b));
    PairView pv;
    SPHB_TRY(upload_pairs(g, b, pair_i, pair_j, dx, dy, r, nullptr, dW, pv));
    const double *d_Vol, *d_rho, *d_pos, *d_vel, *d_wv;
    SPHB_TRY(upload(g, b, Vol.p, nt, &d_Vol));
    SPHB_TRY(upload(g, b, rho.p, nt, &d_rho));
    SPHB_TRY(upload(g, b, pos.p, 2 * nt, &d_pos));
    SPHB_TRY(upload(g, b, vel.p, 2 * nt, &d_vel));
    SPHB_TRY(upload(g, b, wall_vel.p, 2 * nt, &d_wv));
    double *d_po = b.take(2 * nt), *d_dr = b.take(nt), *scr = b.take(scratch_doubles(n_total));
    SPHB_TRY(pairs_integration_2nd(pv, d_Vol, d_rho, d_pos, d_vel, dt, n_fluid, n_total, d_wv, d_po, d_dr, scr, g.st));
    SPHB_TRY(download(g, pos_out, d_po, 2 * nt));
    SPHB_TRY(download(g, drho_out, d_dr, nt));
    if (zeros_out) memset(zeros_out, 0, 2 * nt * sizeof(double));   // third output is all zeros, ph:1047,1064
    return finish(g);
}

// ------------------------------------------------------------------------------------------------
// integration_verlet (ph:1316-1469)
// ------------------------------------------------------------------------------------------------
SPHB_API int sphb_integration_verlet(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                     sphb_mat dW, sphb_mat Vol, sphb_mat B, sphb_mat rho, sphb_mat mass, sphb_mat pos,
                                     sphb_mat vel, sphb_mat drho_dt, sphb_mat force_prior, double dt, int n_fluid,
                                     int n_total, double rho0, double p0, double c_f, sphb_mat wall_vel,
                                     double *rho_out, double *p_out, double *pos_out, double *vel_out,
                                     double *drho_out, double *force_out)
{
    ENTER();
    SPHB_REQUIRE(n_total > 0, "SPHB:arg", "n_total must be positive.");
    const size_t nt = (size_t)n_total;
    SPHB_TRY(check_int_inputs(true, nt, Vol, B, rho, mass, pos, vel, drho_dt, force_prior, wall_vel));
    SPHB_REQUIRE(pairs_consistent(pair_i, pair_j, dx, dy, r, dW), "SPH:Physics:pairs", "Pair arrays must have same length.");
    if (g_validate_only) return 0;
    SPHB_TRY(ensure_stream(g));
    const size_t np = numel(pair_i);
    Bump b;
    SPHB_TRY(open_arena(g, 6 * pad4(np) + 16 * pad4(nt) + verlet_work_doubles(n_total), b));
    PairView pv;
    SPHB_TRY(upload_pairs(g, b, pair_i, pair_j, dx, dy, r, nullptr, dW, pv));
    VerletDev in;
    SPHB_TRY(upload(g, b, Vol.p, nt, &in.Vol));
    SPHB_TRY(upload(g, b, B.p, 4 * nt, &in.B));
    SPHB_TRY(upload(g, b, rho.p, nt, &in.rho));
    SPHB_TRY(upload(g, b, mass.p, nt, &in.mass));
    SPHB_TRY(upload(g, b, pos.p, 2 * nt, &in.pos));
    SPHB_TRY(upload(g, b, vel.p, 2 * nt, &in.vel));
    SPHB_TRY(upload(g, b, drho_dt.p, nt, &in.drho));
    SPHB_TRY(upload(g, b, force_prior.p, 2 * nt, &in.fp));
    SPHB_TRY(upload(g, b, wall_vel.p, 2 * nt, &in.wv));
    double *o_rho, *o_p, *o_pos, *o_vel, *o_drho, *o_force;
    SPHB_TRY(verlet_on_device(g, b, pv, in, dt, n_fluid, n_total, rho0, p0, c_f, &o_rho, &o_p, &o_pos, &o_vel, &o_drho,
                              &o_force));
    SPHB_TRY(download(g, rho_out, o_rho, nt));
    SPHB_TRY(download(g, p_out, o_p, nt));
    SPHB_TRY(download(g, pos_out, o_pos, 2 * nt));
    SPHB_TRY(download(g, vel_out, o_vel, 2 * nt));
    SPHB_TRY(download(g, drho_out, o_drho, nt));
    SPHB_TRY(download(g, force_out, o_force, 2 * nt));
    return finish(g);
}

// ------------------------------------------------------------------------------------------------
// advance_shell_step (ph:1490-1639): density -> viscous -> + m g -> transport(0.2) -> verlet
// ------------------------------------------------------------------------------------------------
SPHB_API int sphb_advance_shell_step(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                     sphb_mat W, sphb_mat dW, sphb_mat mass, sphb_mat pos, sphb_mat vel,
                                     sphb_mat wall_vel, sphb_mat rho, sphb_mat drho_dt, double dt, int n_fluid,
                                     int n_total, double rho0, double p0, double c_f, double mu, double h,
                                     double inv_sigma0, double gravity_g, double *rho_out, double *p_out,
                                     double *pos_out, double *vel_out, double *drho_out, double *force_out,
                                     double *force_prior_out, double *Vol_out, double *B_out)
{
    ENTER();
    const size_t nt = n_total > 0 ? (size_t)n_total : 0;
    SPHB_REQUIRE(n_total > 0 && numel(mass) == nt, "SPH:Physics:advance:mass", "mass size mismatch.");
    SPHB_REQUIRE(is_mat(pos, nt, 2), "SPH:Physics:advance:pos", "pos size mismatch.");
    SPHB_REQUIRE(is_mat(vel, nt, 2), "SPH:Physics:advance:vel", "vel size mismatch.");
    SPHB_REQUIRE(is_mat(wall_vel, nt, 2), "SPH:Physics:advance:wall_vel", "wall_vel size mismatch.");
    SPHB_REQUIRE(numel(rho) == nt, "SPH:Physics:advance:rho", "rho size mismatch.");
    SPHB_REQUIRE(numel(drho_dt) == nt, "SPH:Physics:advance:drho_dt", "drho_dt size mismatch.");
    SPHB_REQUIRE(n_fluid > 0 && n_total >= n_fluid, "SPH:Physics:advance:count", "Invalid n_fluid/n_total.");
    // the inner density_correction call re-checks the pair arrays (ph:139-145) and the parameters
    const size_t np = numel(pair_i);
    SPHB_REQUIRE(numel(pair_j) == np && numel(dx) == np && numel(dy) == np && numel(r) == np && numel(W) == np &&
                     numel(dW) == np,
                 "SPH:Physics:density:pairs", "Pair arrays mismatch.");
    SPHB_REQUIRE(np <= (size_t)INT_MAX, "SPH:Physics:density:pairsize", "Pair count exceeds INT_MAX.");
    SPHB_REQUIRE(rho0 > 0.0 && h > 0.0, "SPH:Physics:density:param", "rho0 and h must be positive.");
    if (g_validate_only) return 0;
    SPHB_TRY(ensure_stream(g));
    Bump b;
    SPHB_TRY(open_arena(g, 7 * pad4(np) + 24 * pad4(nt) + verlet_work_doubles(n_total) + scratch_doubles(n_total), b));
    PairView pv;
    SPHB_TRY(upload_pairs(g, b, pair_i, pair_j, dx, dy, r, &W, dW, pv));
    const double *d_mass, *d_pos, *d_vel, *d_wv, *d_drho;
    SPHB_TRY(upload(g, b, mass.p, nt, &d_mass));
    SPHB_TRY(upload(g, b, pos.p, 2 * nt, &d_pos));
    SPHB_TRY(upload(g, b, vel.p, 2 * nt, &d_vel));
    SPHB_TRY(upload(g, b, wall_vel.p, 2 * nt, &d_wv));
    SPHB_TRY(upload(g, b, drho_dt.p, nt, &d_drho));
    double *d_rho = b.take(nt), *d_Vol = b.take(nt), *d_B = b.take(4 * nt), *d_fp = b.take(2 * nt);
    double *d_post = b.take(2 * nt), *scr = b.take(scratch_doubles(n_total));
    SPHB_TRY(pairs_density(pv, d_mass, n_fluid, n_total, rho0, h, inv_sigma0, d_rho, d_Vol, d_B, scr, g.st));
    SPHB_TRY(pairs_viscous(pv, d_vel, d_Vol, d_B, mu, h, n_fluid, n_total, d_wv, d_fp, scr, g.st));
    SPHB_TRY(pairs_add_gravity(d_fp, d_mass, gravity_g, n_fluid, g.st));                       // ph:1575-1580
    SPHB_TRY(pairs_transport(pv, d_Vol, d_B, d_pos, h, n_fluid, n_total, 0.2, d_post, scr, g.st));   // ph:1596 + 584
    VerletDev in{d_Vol, d_B, d_rho, d_mass, d_post, d_vel, d_drho, d_fp, d_wv};
    double *o_rho, *o_p, *o_pos, *o_vel, *o_drho, *o_force;
    SPHB_TRY(verlet_on_device(g, b, pv, in, dt, n_fluid, n_total, rho0, p0, c_f, &o_rho, &o_p, &o_pos, &o_vel, &o_drho,
                              &o_force));
    SPHB_TRY(download(g, rho_out, o_rho, nt));
    SPHB_TRY(download(g, p_out, o_p, nt));
    SPHB_TRY(download(g, pos_out, o_pos, 2 * nt));
    SPHB_TRY(download(g, vel_out, o_vel, 2 * nt));
    SPHB_TRY(download(g, drho_out, o_drho, nt));
    SPHB_TRY(download(g, force_out, o_force, 2 * nt));
    SPHB_TRY(download(g, force_prior_out, d_fp, 2 * nt));
    SPHB_TRY(download(g, Vol_out, d_Vol, nt));
    SPHB_TRY(download(g, B_out, d_B, 4 * nt));
    return finish(g);
}

// ------------------------------------------------------------------------------------------------
// wall_shear_monitor (ph:1653-1743; checks ph:1684-1707; n_total = numel(Vol), ph:1692)
// ------------------------------------------------------------------------------------------------
SPHB_API int sphb_wall_shear_monitor(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                     sphb_mat dW, sphb_mat pos, sphb_mat vel, sphb_mat wall_vel, sphb_mat Vol,
                                     sphb_mat B, int n_fluid, double DL, double DH, double mu, double h,
                                     double *tau_bottom, double *tau_top)
{
    ENTER();
    SPHB_REQUIRE(pairs_consistent(pair_i, pair_j, dx, dy, r, dW), "SPH:Physics:pairs", "Pair arrays must have same length.");
    const size_t nt = numel(Vol);
    SPHB_REQUIRE(DL > 0.0 && h > 0.0, "SPH:Physics:wallshear:param", "DL and h must be positive.");
    SPHB_REQUIRE(is_mat(pos, nt, 2), "SPH:Physics:wallshear:pos", "pos size mismatch.");
    SPHB_REQUIRE(is_mat(vel, nt, 2), "SPH:Physics:wallshear:vel", "vel size mismatch.");
    SPHB_REQUIRE(is_mat(wall_vel, nt, 2), "SPH:Physics:wallshear:wall_vel", "wall_vel size mismatch.");
    SPHB_REQUIRE(is_mat(B, nt, 4), "SPH:Physics:wallshear:B", "B size mismatch.");
    SPHB_REQUIRE(nt > 0 && nt <= (size_t)INT_MAX, "SPHB:arg", "Vol must be non-empty.");
    if (g_validate_only) return 0;
    SPHB_TRY(ensure_stream(g));
    const size_t np = numel(pair_i);
    const int nti = (int)nt;
    Bump b;
    SPHB_TRY(open_arena(g, 6 * pad4(np) + 12 * pad4(nt) + scratch_doubles(nti) + 8, b));
    PairView pv;
    SPHB_TRY(upload_pairs(g, b, pair_i, pair_j, dx, dy, r, nullptr, dW, pv));
    const double *d_pos, *d_vel, *d_wv, *d_Vol, *d_B;
    SPHB_TRY(upload(g, b, pos.p, 2 * nt, &d_pos));
    SPHB_TRY(upload(g, b, vel.p, 2 * nt, &d_vel));
    SPHB_TRY(upload(g, b, wall_vel.p, 2 * nt, &d_wv));
    SPHB_TRY(upload(g, b, Vol.p, nt, &d_Vol));
    SPHB_TRY(upload(g, b, B.p, 4 * nt, &d_B));
    double *d_tau = b.take(4), *scr = b.take(scratch_doubles(nti));
    SPHB_TRY(pairs_wall_shear(pv, d_pos, d_vel, d_wv, d_Vol, d_B, n_fluid, nti, DL, DH, mu, h, d_tau, scr, g.st));
    double host[2] = {0.0, 0.0};
    SPHB_CUDA(cudaMemcpyAsync(host, d_tau, 2 * sizeof(double), cudaMemcpyDeviceToHost, g.st));
    SPHB_TRY(finish(g));
    if (tau_bottom) *tau_bottom = host[0];
    if (tau_top) *tau_top = host[1];
    return 0;
}

}  // extern "C"
