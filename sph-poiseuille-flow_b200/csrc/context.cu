// context.cu -- the device-resident channel: state in HBM, neighbour CSR, one-call time steps.
//
// Restates the hot slice of SPH_Poiseuille.m (row S1 of SURVEY.md 8(a)): the main-loop body m:250-283
// with verlet_time_step m:519-527, periodic_bounding m:570-577, sort_particles_by_cell m:529-568 and
// the S5 initialisation m:165-173, as a fixed sequence of kernel launches on one stream; the time
// step, max|v|, clock and monitors stay on the device (DevScalars), so a step needs no host round trip.
#include "context.cuh"

#include <limits>
#include <string>
#include <thread>
#include <vector>

namespace {

__global__ void k_fill(double *p, double v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_identity(Mat2 *B, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        Mat2 m;
        m.b11 = 1.0;
        m.b12 = 0.0;
        m.b21 = 0.0;
        m.b22 = 1.0;
        B[i] = m;
    }
}
__global__ void k_vol_from_mass(const double *mass, const double *rho, double *Vol, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) Vol[i] = mass[i] / rho[i];   // m:120
}
__global__ void k_eos(const double *rho, double *p, int nf, int nt, double rho0, double p0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nt) p[i] = i < nf ? p0 * (rho[i] / rho0 - 1.0) : 0.0;   // m:172-173
}
__global__ void k_set_clock(DevScalars *sc, double t, long long step)
{
    sc->t = t;
    sc->step = step;
}
__global__ void k_zero_wall_vel(double2 *vel, int nf, int nt)
{
    const int i = nf + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nt) vel[i] = make_double2(0.0, 0.0);   // m:270
}

int bind_device(sphb_ctx *c)
{
    SPHB_CUDA(cudaSetDevice(c->device));
    return 0;
}

void drop_graphs(sphb_ctx *c)
{
    for (int k = 0; k < 2; ++k) {
        if (c->graph[k]) cudaGraphExecDestroy(c->graph[k]);
        c->graph[k] = nullptr;
    }
    c->graph_inner = -1;
}

int plan_lists(sphb_ctx *c, double y_lo, double y_hi)
{
    const sphb_params &p = c->prm;
    SPHB_TRY(neighbor_plan(c->nw, p.n_fluid, p.n_total, p.h, p.DL, y_lo, y_hi, p.max_neighbors_hint,
                           csr_max_tile_slots()));
    SPHB_TRY(csr_configure(c->nw.tile_cap));
    SPHB_TRY(c->shear_partials.reserve(2 * (size_t)c->nw.n_tiles * sizeof(double)));
    for (DevBuf *b : {&c->e_vel, &c->e_w, &c->e_pV, &c->e_b1, &c->e_b2, &c->e_vn, &c->e_rm})   // entry-ordered mirrors
        if (b->bytes < c->nw.ell_stride * sizeof(double2)) {
            SPHB_TRY(b->reserve(c->nw.ell_stride * sizeof(double2)));
            SPHB_CUDA(cudaMemsetAsync(b->p, 0, c->nw.ell_stride * sizeof(double2), c->st));
        }
    SPHB_TRY(sort_plan(c->sw, p.n_fluid, p.DL, p.h, y_hi - y_lo));
    c->planned = true;
    c->y_lo = y_lo;
    c->y_hi = y_hi;
    return 0;
}

// ---- pieces of one step ------------------------------------------------------------------------
int fresh_mirrors(sphb_ctx *c, bool need_pv)
{
    if (!c->have_list || !(c->mirrors_stale || (need_pv && !c->pv_valid))) return 0;
    SPHB_TRY(csr_refresh_mirrors(c->P, c->view(), c->st));
    c->launches += 1;
    c->mirrors_stale = false;
    c->pv_valid = true;
    return 0;
}

// publish_dt: the density launch opens a reference step (do_publish_dt folded into it)
int do_density(sphb_ctx *c, bool publish_dt = false, double remain = 0.0)
{
    SPHB_TRY(fresh_mirrors(c, false));
    const StateView S = c->view();
    {
        ProfScope ps(&c->kprof, kPfDensity, c->st);
        SPHB_TRY(csr_density(c->P, S, c->sc(), c->st, publish_dt, remain));
    }
    {
        ProfScope ps(&c->kprof, kPfKgc, c->st);
        SPHB_TRY(csr_kgc(c->P, S, c->st));
    }
    c->launches += 2;
    c->pv_valid = true;
    return 0;
}

int do_visc_transport(sphb_ctx *c, bool visc, bool gravity, bool transport)
{
    ProfScope ps(&c->kprof, kPfViscTransp, c->st);
    SPHB_TRY(csr_visc_transport(c->P, c->view(), visc, gravity, transport, c->st));
    c->launches += 1;
    return 0;
}

int do_sort(sphb_ctx *c)
{
    const int nf = c->prm.n_fluid, nt = c->prm.n_total;
    if (nf <= 0 || nf >= nt) return 0;   // m:535-537
    drop_graphs(c);   // the permutation swaps buffer pointers: a cached graph of the step holds the old ones
    ProfScope ps(&c->kprof, kPfSort, c->st);
    SPHB_TRY(sort_permutation(c->sw, c->pos.as<double2>(), nf, c->prm.DL, c->prm.h, c->sc(), c->nw.flags.as<int>(),
                              c->st, &c->launches));
    PermuteArrays A;
    DevBuf *in2[] = {&c->pos, &c->vel, &c->wall_vel, &c->force_prior, &c->force};
    DevBuf *out2[] = {&c->a_pos, &c->a_vel, &c->a_wall_vel, &c->a_force_prior, &c->a_force};
    DevBuf *in1[] = {&c->rho, &c->mass, &c->drho_dt, &c->p, &c->Vol};
    DevBuf *out1[] = {&c->a_rho, &c->a_mass, &c->a_drho_dt, &c->a_p, &c->a_Vol};
    for (int k = 0; k < PermuteArrays::kN2; ++k) {
        A.in2[k] = in2[k]->as<double2>();
        A.out2[k] = out2[k]->as<double2>();
    }
    for (int k = 0; k < PermuteArrays::kN1; ++k) {
        A.in1[k] = in1[k]->as<double>();
        A.out1[k] = out1[k]->as<double>();
    }
    A.inB = c->B.as<Mat2>();
    A.outB = c->a_B.as<Mat2>();
    SPHB_TRY(sort_apply(c->sw, nf, nt, A, c->st, &c->launches));
    for (int k = 0; k < PermuteArrays::kN2; ++k) std::swap(*in2[k], *out2[k]);
    for (int k = 0; k < PermuteArrays::kN1; ++k) std::swap(*in1[k], *out1[k]);
    std::swap(c->B, c->a_B);
    c->have_list = false;
    // the list build reads the rows through these (BuildMirrors): they follow the swap
    c->nw.mirrors.vel = c->vel.as<double2>();
    c->nw.mirrors.wall_vel = c->wall_vel.as<double2>();
    c->nw.mirrors.mass = c->mass.as<double>();
    return 0;
}

int do_neighbor(sphb_ctx *c)
{
    // the first build sizes the ELL exactly (one host round trip); later builds only raise the flag
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    (void)cudaStreamIsCapturing(c->st, &cap);
    const bool exact = !c->list_sized && cap == cudaStreamCaptureStatusNone;
    // entry-ordered mirrors the build writes for the gather kernels (the state buffers swap at a cell sort: set anew)
    c->nw.mirrors.vel = c->vel.as<double2>();
    c->nw.mirrors.wall_vel = c->wall_vel.as<double2>();
    c->nw.mirrors.mass = c->mass.as<double>();
    c->nw.mirrors.inv_rho0 = c->P.inv_rho0;
    c->nw.mirrors.e_vel = c->e_vel.as<double2>();
    c->nw.mirrors.e_w = c->e_w.as<double2>();
    c->nw.mirrors.ghost_ent = c->nw.ghost_ent.as<int>();
    SPHB_TRY(neighbor_build(c->nw, c->pos.as<double2>(), false, exact, c->st, &c->launches, &c->kprof));
    if (exact) {
        c->list_sized = true;
        drop_graphs(c);   // the list, the tiles or their slot count may have changed
        SPHB_TRY(csr_configure(c->nw.tile_cap));
        SPHB_TRY(c->shear_partials.reserve(2 * (size_t)c->nw.n_tiles * sizeof(double)));
    }
    c->have_list = true;
    c->mirrors_stale = false;   // the build wrote e_vel / e_w for the new entry order
    c->pv_valid = false;        // (p, V) and B mirrors belong to the old one until the next density pass
    return 0;
}

int do_shear(sphb_ctx *c)
{
    SPHB_TRY(fresh_mirrors(c, false));   // the monitor stages the wall entries' weights and velocities from the mirrors
    ProfScope ps(&c->kprof, kPfShear, c->st);
    SPHB_TRY(csr_wall_shear(c->P, c->view(), c->sc(), c->shear_partials.as<double>(), c->st));
    c->launches += 1;
    return 0;
}

// GridFuse record when the list build follows this continuity pass at once (sphb_ctx_step, run_dual)
GridFuse fuse_for_next_build(sphb_ctx *c, bool build_follows)
{
    if (!build_follows) return GridFuse();
    const GridFuse F = neighbor_grid_fuse(c->nw);
    if (F.g) c->nw.grid_ready = true;   // consumed by the next neighbor_build
    return F;
}

int do_substep(sphb_ctx *c, double remain, double fixed_dt, bool wrap, bool build_follows = false)
{
    const StateView S = c->view();
    {
        ProfScope ps(&c->kprof, kPfTimeStep, c->st);
        SPHB_TRY(csr_time_step(c->P, S, c->sc(), remain, fixed_dt, c->st));
    }
    {
        ProfScope ps(&c->kprof, kPfPressure, c->st);
        SPHB_TRY(csr_pressure(c->P, S, c->sc(), c->st));
    }
    {
        ProfScope ps(&c->kprof, kPfContinuity, c->st);
        SPHB_TRY(csr_continuity(c->P, S, c->sc(), wrap, true, fuse_for_next_build(c, build_follows), c->st));
    }
    c->launches += 3;
    return 0;
}

// first substep of an outer step (after do_density, which published the step's dt): viscous + transport + pressure in one
// gather, then continuity
int do_fused_first_substep(sphb_ctx *c, bool wrap, bool build_follows = false)
{
    const StateView S = c->view();
    {
        ProfScope ps(&c->kprof, kPfForceFused, c->st);
        SPHB_TRY(csr_force_fused(c->P, S, c->sc(), true, c->st));
    }
    {
        ProfScope ps(&c->kprof, kPfContinuity, c->st);
        SPHB_TRY(csr_continuity(c->P, S, c->sc(), wrap, true, fuse_for_next_build(c, build_follows), c->st));
    }
    c->launches += 2;
    return 0;
}

void tick(sphb_ctx *c, int k)
{
    if (c->timing) cudaEventRecord(c->ev[k], c->st);
}

}  // namespace

int ctx_bind_device(sphb_ctx *c) { return bind_device(c); }
int ctx_plan_lists(sphb_ctx *c, double y_lo, double y_hi) { return plan_lists(c, y_lo, y_hi); }
void ctx_drop_graphs(sphb_ctx *c) { drop_graphs(c); }
int ctx_do_neighbor(sphb_ctx *c) { return do_neighbor(c); }
int ctx_fresh_mirrors(sphb_ctx *c, bool need_pv) { return fresh_mirrors(c, need_pv); }
int ctx_do_shear(sphb_ctx *c) { return do_shear(c); }
int ctx_do_sort(sphb_ctx *c) { return do_sort(c); }
int ctx_volumes_from_mass(sphb_ctx *c)
{
    const int nt = c->prm.n_total;
    k_vol_from_mass<<<grid_for((size_t)nt), kBlock, 0, c->st>>>(c->mass.as<double>(), c->rho.as<double>(),
                                                                c->Vol.as<double>(), nt);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int ctx_step_physics(sphb_ctx *c, double remain)
{
    c->h_step += 1;
    c->last_inner = 1;
    c->last_remain = remain;
    SPHB_TRY(do_density(c, true, remain));                    // m:259 (dt), m:254-255
    // a slab's list build follows after the exchange: the continuity pass starts its grid set-up (GridFuse, derive = 0)
    SPHB_TRY(do_fused_first_substep(c, true, c->slab.on));    // m:256-270
    return 0;
}

extern "C" {

SPHB_API int sphb_ctx_create(const sphb_params *prm, sphb_ctx **out)
{
    clear_error();
    if (!prm || !out) return set_error("SPHB:arg", "null argument");
    *out = nullptr;
    SPHB_REQUIRE(prm->n_total > 0 && prm->n_fluid > 0 && prm->n_fluid <= prm->n_total, "SPH:Neighbor:count",
                 "Invalid n_fluid/n_total or inconsistent pos size.");
    SPHB_REQUIRE(prm->h > 0.0 && prm->DL > 0.0, "SPH:Neighbor:param", "h and DL must be positive.");
    SPHB_REQUIRE(prm->rho0 > 0.0, "SPH:Physics:density:param", "rho0 and h must be positive.");
    SPHB_TRY(ensure_device());
    sphb_ctx *c = new sphb_ctx();
    c->prm = *prm;
    if (prm->device >= 0) c->device = prm->device;
    else if (cudaGetDevice(&c->device) != cudaSuccess) c->device = 0;
    int rc = [&]() -> int {
        SPHB_TRY(bind_device(c));
        SPHB_CUDA(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
        const size_t nt = (size_t)prm->n_total;
        PhysParams &P = c->P;
        P.nf = prm->n_fluid;
        P.nfo = prm->n_fluid;
        P.nt = prm->n_total;
        P.h = prm->h;
        P.DL = prm->DL;
        P.DH = prm->DH;
        P.rho0 = prm->rho0;
        P.p0 = prm->p0;
        P.c_f = prm->c_f;
        P.mu = prm->mu;
        P.nu = prm->mu / prm->rho0;   // m:80
        P.g = prm->gravity_g;
        P.inv_sigma0 = prm->inv_sigma0;
        P.transport_coeff = prm->transport_coeff;
        P.sigma = 10.0 / (7.0 * 3.14159265358979323846 * prm->h * prm->h);
        P.inv_h = 1.0 / prm->h;
        P.sigma_h = P.sigma / prm->h;
        P.inv_rho0 = 1.0 / prm->rho0;
        DevBuf *v2[] = {&c->pos, &c->vel, &c->vel_new, &c->wall_vel, &c->force, &c->force_prior,
                        &c->a_pos, &c->a_vel, &c->a_wall_vel, &c->a_force, &c->a_force_prior};
        for (DevBuf *b : v2) {
            SPHB_TRY(b->reserve(nt * sizeof(double2)));
            SPHB_CUDA(cudaMemsetAsync(b->p, 0, nt * sizeof(double2), c->st));
        }
        DevBuf *v1[] = {&c->rho, &c->p, &c->drho_dt, &c->mass, &c->Vol, &c->a_rho, &c->a_p, &c->a_drho_dt,
                        &c->a_mass, &c->a_Vol};
        for (DevBuf *b : v1) {
            SPHB_TRY(b->reserve(nt * sizeof(double)));
            SPHB_CUDA(cudaMemsetAsync(b->p, 0, nt * sizeof(double), c->st));
        }
        SPHB_TRY(c->rp_half.reserve(nt * sizeof(double2)));
        SPHB_CUDA(cudaMemsetAsync(c->rp_half.p, 0, nt * sizeof(double2), c->st));
        SPHB_TRY(c->B.reserve(nt * sizeof(Mat2)));
        SPHB_TRY(c->a_B.reserve(nt * sizeof(Mat2)));
        SPHB_TRY(c->scalars.reserve(sizeof(DevScalars)));
        SPHB_CUDA(cudaMemsetAsync(c->scalars.p, 0, sizeof(DevScalars), c->st));
        SPHB_TRY(c->stage.reserve(6 * nt * sizeof(double)));   // three 2-column (or one 4- + one 2-column) arrays in flight
        k_fill<<<grid_for(nt), kBlock, 0, c->st>>>(c->rho.as<double>(), prm->rho0, (int)nt);
        k_fill_identity<<<grid_for(nt), kBlock, 0, c->st>>>(c->B.as<Mat2>(), (int)nt);
        SPHB_CUDA(cudaMallocHost((void **)&c->h_scalars, 2 * sizeof(DevScalars)));   // [1]: the copy of flags_watch
        SPHB_CUDA(cudaMallocHost((void **)&c->h_ints, 8 * sizeof(int)));
        for (auto &e : c->ev) SPHB_CUDA(cudaEventCreate(&e));
        SPHB_CUDA(cudaStreamSynchronize(c->st));
        return 0;
    }();
    if (rc) {
        sphb_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return 0;
}

SPHB_API void sphb_ctx_destroy(sphb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    drop_graphs(c);
    if (c->own_st) c->st = c->own_st;   // never destroy a stream of the caller
    DevBuf *all[] = {&c->pos, &c->vel, &c->vel_new, &c->wall_vel, &c->force, &c->force_prior, &c->rho, &c->p,
                     &c->drho_dt, &c->mass, &c->Vol, &c->B, &c->a_pos, &c->a_vel, &c->a_wall_vel, &c->a_force,
                     &c->a_force_prior, &c->a_rho, &c->a_p, &c->a_drho_dt, &c->a_mass, &c->a_Vol, &c->a_B,
                     &c->scalars, &c->shear_partials, &c->stage, &c->prof, &c->rp_half,
                     &c->e_vel, &c->e_w, &c->e_pV, &c->e_b1, &c->e_b2, &c->e_vn, &c->e_rm};
    for (DevBuf *b : all) b->release();
    for (DevBuf *b : {&c->slab.w_pos, &c->slab.w_mass, &c->slab.w_wall_vel, &c->slab.counts, &c->slab.block_tot,
                      &c->slab.pid, &c->slab.a_pid, &c->slab.sort_flag, &c->slab.sort_ptr})
        b->release();
    if (c->slab.h_counts) cudaFreeHost(c->slab.h_counts);
    for (int r = 0; r < SlabState::kMaxWorld; ++r)
        if (c->slab.peer_opened[r] && c->slab.peer[r]) cudaIpcCloseMemHandle(c->slab.peer[r]);
    if (c->slab.mail) cudaFree(c->slab.mail);
    c->slab.p2p_state.release();
    c->nw.release_all();
    c->sw.release_all();
    c->kprof.release();
    if (c->flags_ev) cudaEventDestroy(c->flags_ev);
    if (c->aux_ev) cudaEventDestroy(c->aux_ev);
    if (c->aux_st) cudaStreamDestroy(c->aux_st);
    if (c->h_scalars) cudaFreeHost(c->h_scalars);
    if (c->h_ints) cudaFreeHost(c->h_ints);
    for (auto &e : c->ev)
        if (e) cudaEventDestroy(e);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}

namespace {
void host_y_range(const double *y, size_t n, double &lo, double &hi)
{
    double lo8[8], hi8[8];
    for (int k = 0; k < 8; ++k) lo8[k] = hi8[k] = y[0];
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
        for (int k = 0; k < 8; ++k) {
            const double v = y[i + k];
            lo8[k] = v < lo8[k] ? v : lo8[k];
            hi8[k] = v > hi8[k] ? v : hi8[k];
        }
    for (; i < n; ++i) {
        lo8[0] = y[i] < lo8[0] ? y[i] : lo8[0];
        hi8[0] = y[i] > hi8[0] ? y[i] : hi8[0];
    }
    lo = lo8[0];
    hi = hi8[0];
    for (int k = 1; k < 8; ++k) {
        lo = lo8[k] < lo ? lo8[k] : lo;
        hi = hi8[k] > hi ? hi8[k] : hi;
    }
}
}  // namespace

SPHB_API int sphb_ctx_upload(sphb_ctx *c, const double *pos, const double *vel, const double *rho,
                             const double *drho_dt, const double *mass, const double *wall_vel)
{
    clear_error();
    if (!c) return set_error("SPHB:arg", "null context");
    SPHB_TRY(bind_device(c));
    const int nt = c->prm.n_total;
    const size_t n = (size_t)nt;
    // every 2-column array has its own staging region: the copies and the pack kernels queue back to back
    double *stage = c->stage.as<double>();
    int region = 0;
    auto up2 = [&](const double *src, DevBuf &dst) -> int {
        double *st2 = stage + (size_t)(region++) * 2 * n;
        SPHB_CUDA(cudaMemcpyAsync(st2, src, 2 * n * sizeof(double), cudaMemcpyHostToDevice, c->st));
        launch_pack2(st2, dst.as<double2>(), nt, c->st);
        return 0;
    };
    if (pos) {
        SPHB_TRY(up2(pos, c->pos));
        c->have_list = false;
        c->nw.grid_ready = false;
    }
    if (vel) SPHB_TRY(up2(vel, c->vel));
    if (wall_vel) SPHB_TRY(up2(wall_vel, c->wall_vel));
    if (rho) SPHB_CUDA(cudaMemcpyAsync(c->rho.p, rho, n * sizeof(double), cudaMemcpyHostToDevice, c->st));
    if (drho_dt) SPHB_CUDA(cudaMemcpyAsync(c->drho_dt.p, drho_dt, n * sizeof(double), cudaMemcpyHostToDevice, c->st));
    if (mass) SPHB_CUDA(cudaMemcpyAsync(c->mass.p, mass, n * sizeof(double), cudaMemcpyHostToDevice, c->st));
    // Vol = m / rho is the initial condition (m:120) and comes with the masses.  A caller that feeds the state of
    // the last step back (rho without mass) keeps the volumes of that step's density pass, like B: they are what
    // the reference hands to the wall-shear monitor (m:281-283), and every step recomputes them first thing.
    if (mass)
        k_vol_from_mass<<<grid_for(n), kBlock, 0, c->st>>>(c->mass.as<double>(), c->rho.as<double>(),
                                                           c->Vol.as<double>(), nt);
    if (vel) SPHB_TRY(csr_vmax(c->P, c->vel.as<double2>(), c->sc(), c->st));
    if (vel || wall_vel || mass || rho) c->mirrors_stale = true;   // (a standing list keeps its entry order)
    if (pos) {
        // y range of the uploaded rows: capacity planning only (the grid itself is derived on the device).  Done while
        // the copies above are in flight -- this pass over the host array used to sit between the first copy and the
        // rest and cost more than the whole upload (0.8 ms for 489 600 rows as a dependent min / max chain; eight
        // independent chains take a third of that and now overlap the transfer).
        double lo, hi;
        host_y_range(pos + n, n, lo, hi);
        if (!c->planned || lo < c->y_lo || hi > c->y_hi) {
            SPHB_CUDA(cudaStreamSynchronize(c->st));   // (the staging regions are in use; planning may reallocate)
            drop_graphs(c);
            SPHB_TRY(plan_lists(c, lo, hi));
        }
    }
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

namespace {
int enqueue_download(sphb_ctx *c, double *pos, double *vel, double *rho, double *p, double *drho_dt, double *force,
                     double *force_prior, double *Vol, double *B);
}

SPHB_API int sphb_ctx_download(sphb_ctx *c, double *pos, double *vel, double *rho, double *p, double *drho_dt,
                               double *force, double *force_prior, double *Vol, double *B)
{
    clear_error();
    if (!c) return set_error("SPHB:arg", "null context");
    SPHB_TRY(bind_device(c));
    if (c->have_list && !c->flags_pending) SPHB_TRY(ctx_flags_watch(c));
    SPHB_TRY(ctx_flags_poll(c, true));   // an asynchronous run that raised a device flag is recovered (or reported) first
    SPHB_TRY(enqueue_download(c, pos, vel, rho, p, drho_dt, force, force_prior, Vol, B));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

namespace {
// the copies of sphb_ctx_download, enqueued on the context's stream (no wait at the end)
int enqueue_download(sphb_ctx *c, double *pos, double *vel, double *rho, double *p, double *drho_dt, double *force,
                     double *force_prior, double *Vol, double *B)
{
    const int nt = c->prm.n_total;
    const size_t n = (size_t)nt;
    double *stage = c->stage.as<double>();
    int region = 0;
    auto down2 = [&](const DevBuf &src, double *dst) -> int {
        if (region >= 3) {   // the three staging regions are in flight: drain before reusing them
            SPHB_CUDA(cudaStreamSynchronize(c->st));
            region = 0;
        }
        double *st2 = stage + (size_t)(region++) * 2 * n;
        launch_unpack2(src.as<double2>(), st2, nt, c->st);
        SPHB_CUDA(cudaMemcpyAsync(dst, st2, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        return 0;
    };
    if (pos) SPHB_TRY(down2(c->pos, pos));
    if (vel) SPHB_TRY(down2(c->vel, vel));
    if (force) SPHB_TRY(down2(c->force, force));
    if (force_prior) SPHB_TRY(down2(c->force_prior, force_prior));
    if (B) {
        SPHB_CUDA(cudaStreamSynchronize(c->st));
        launch_unpack4(c->B.as<Mat2>(), stage, nt, c->st);
        SPHB_CUDA(cudaMemcpyAsync(B, stage, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    }
    if (rho) SPHB_CUDA(cudaMemcpyAsync(rho, c->rho.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    if (p) SPHB_CUDA(cudaMemcpyAsync(p, c->p.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    if (drho_dt) SPHB_CUDA(cudaMemcpyAsync(drho_dt, c->drho_dt.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    if (Vol) SPHB_CUDA(cudaMemcpyAsync(Vol, c->Vol.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    return 0;
}
}  // namespace

SPHB_API int sphb_ctx_set_time(sphb_ctx *c, double t, long long step)
{
    clear_error();
    if (!c) return set_error("SPHB:arg", "null context");
    SPHB_TRY(bind_device(c));
    k_set_clock<<<1, 1, 0, c->st>>>(c->sc(), t, step);
    c->h_step = step;
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

SPHB_API int sphb_ctx_neighbor_build(sphb_ctx *c)
{
    clear_error();
    if (!c) return set_error("SPHB:arg", "null context");
    SPHB_TRY(bind_device(c));
    if (!c->planned) return set_error("SPHB:state", "upload positions before building the neighbour list");
    return do_neighbor(c);
}

SPHB_API int sphb_ctx_neighbor_count(sphb_ctx *c, size_t *n_half, size_t *n_csr)
{
    clear_error();
    if (!c) return set_error("SPHB:arg", "null context");
    SPHB_TRY(bind_device(c));
    if (!c->have_list) return set_error("SPHB:state", "no neighbour list");
    SPHB_TRY(neighbor_half_total(c->nw, c->st));
    SPHB_CUDA(cudaMemcpyAsync(c->h_ints, c->nw.totals.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaMemcpyAsync(c->h_ints + 2, c->nw.flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    if (n_half) *n_half = (size_t)c->h_ints[0];
    if (n_csr) *n_csr = (size_t)c->h_ints[1];
    if (c->h_ints[2] & (kErrCsrCap | kErrCellCap | kErrSeenCap))
        return set_error("SPH:Neighbor:entryCapacity", "neighbour list capacity exceeded (flags=%d)", c->h_ints[2]);
    return 0;
}

SPHB_API int sphb_ctx_neighbor_export_pairs(sphb_ctx *c, double *pair_i, double *pair_j, double *dx, double *dy,
                                            double *r, double *W, double *dW)
{
    clear_error();
    size_t nh = 0;
    SPHB_TRY(sphb_ctx_neighbor_count(c, &nh, nullptr));
    if (nh == 0) return 0;
    DevBuf tmp;
    SPHB_TRY(tmp.reserve(7 * nh * sizeof(double)));
    double *d = tmp.as<double>();
    int rc = neighbor_export_pairs(c->nw, nh, d, d + nh, d + 2 * nh, d + 3 * nh, d + 4 * nh, d + 5 * nh, d + 6 * nh,
                                   c->st);
    double *dst[7] = {pair_i, pair_j, dx, dy, r, W, dW};
    for (int k = 0; k < 7 && !rc; ++k) {
        if (!dst[k]) continue;
        if (cudaMemcpyAsync(dst[k], d + k * nh, nh * sizeof(double), cudaMemcpyDeviceToHost, c->st) != cudaSuccess)
            rc = set_error("SPHB:cuda", "pair export copy failed");
    }
    if (cudaStreamSynchronize(c->st) != cudaSuccess && !rc) rc = set_error("SPHB:cuda", "pair export failed");
    c->launches += 1;
    tmp.release();
    return rc;
}

SPHB_API int sphb_ctx_neighbor_export_csr(sphb_ctx *c, int *row_ptr, int *col)
{
    clear_error();
    size_t nc = 0;
    SPHB_TRY(sphb_ctx_neighbor_count(c, nullptr, &nc));
    const size_t nf = (size_t)c->prm.n_fluid;
    DevBuf tmp;
    SPHB_TRY(tmp.reserve((nf + 1 + nc) * sizeof(int)));
    int *d_ptr = tmp.as<int>(), *d_col = d_ptr + nf + 1;
    std::vector<int> h_ptr(nf + 1);
    int rc = neighbor_export_csr(c->nw, d_ptr, col ? d_col : nullptr, c->st);
    if (!rc && cudaMemcpyAsync(h_ptr.data(), d_ptr, (nf + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->st) != cudaSuccess)
        rc = set_error("SPHB:cuda", "CSR export copy failed");
    if (!rc && col && nc &&
        cudaMemcpyAsync(col, d_col, nc * sizeof(int), cudaMemcpyDeviceToHost, c->st) != cudaSuccess)
        rc = set_error("SPHB:cuda", "CSR export copy failed");
    if (cudaStreamSynchronize(c->st) != cudaSuccess && !rc) rc = set_error("SPHB:cuda", "CSR export failed");
    tmp.release();
    if (rc) return rc;
    // export path (not hot): rows come in the reference's walk order; the API promises ascending j
    if (col)
        for (size_t i = 0; i < nf; ++i) std::sort(col + h_ptr[i], col + h_ptr[i + 1]);
    if (row_ptr) std::copy(h_ptr.begin(), h_ptr.end(), row_ptr);
    return 0;
}

#define CTX_ENTER(need_list)                                                                   \
    clear_error();                                                                             \
    if (!c) return set_error("SPHB:arg", "null context");                                      \
    SPHB_TRY(bind_device(c));                                                                  \
    if ((need_list) && !c->have_list) return set_error("SPHB:state", "no neighbour list: call sphb_ctx_neighbor_build")

SPHB_API int sphb_ctx_density_correction(sphb_ctx *c)
{
    CTX_ENTER(true);
    return do_density(c);
}

SPHB_API int sphb_ctx_eos(sphb_ctx *c)
{
    CTX_ENTER(false);
    k_eos<<<grid_for(c->prm.n_total), kBlock, 0, c->st>>>(c->rho.as<double>(), c->p.as<double>(), c->prm.n_fluid,
                                                          c->prm.n_total, c->prm.rho0, c->prm.p0);
    c->launches += 1;
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

SPHB_API int sphb_ctx_viscous_force(sphb_ctx *c, int add_gravity)
{
    CTX_ENTER(true);
    return do_visc_transport(c, true, add_gravity != 0, false);
}

SPHB_API int sphb_ctx_transport_correction(sphb_ctx *c)
{
    CTX_ENTER(true);
    return do_visc_transport(c, false, false, true);
}

SPHB_API int sphb_ctx_integration_verlet(sphb_ctx *c, double dt)
{
    CTX_ENTER(true);
    SPHB_TRY(fresh_mirrors(c, true));
    const StateView S = c->view();
    SPHB_TRY(csr_time_step(c->P, S, c->sc(), std::numeric_limits<double>::infinity(), dt, c->st));
    SPHB_TRY(csr_pressure(c->P, S, c->sc(), c->st));
    SPHB_TRY(csr_continuity(c->P, S, c->sc(), false, false, GridFuse(), c->st));
    c->launches += 3;
    return 0;
}

SPHB_API int sphb_ctx_wall_shear(sphb_ctx *c, double *tau_bottom, double *tau_top)
{
    CTX_ENTER(true);
    SPHB_TRY(do_shear(c));
    SPHB_CUDA(cudaMemcpyAsync(c->h_scalars, c->sc(), sizeof(DevScalars), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    if (tau_bottom) *tau_bottom = c->h_scalars->tau_bottom;
    if (tau_top) *tau_top = c->h_scalars->tau_top;
    return 0;
}

SPHB_API int sphb_ctx_periodic_wrap(sphb_ctx *c)
{
    CTX_ENTER(false);
    c->launches += 1;
    return csr_wrap(c->P, c->pos.as<double2>(), c->st);
}

SPHB_API int sphb_ctx_sort_by_cell(sphb_ctx *c)
{
    CTX_ENTER(false);
    if (!c->planned) return set_error("SPHB:state", "upload positions first");
    drop_graphs(c);
    return do_sort(c);
}

SPHB_API int sphb_ctx_status(sphb_ctx *c, sphb_status *out)
{
    CTX_ENTER(false);
    if (!out) return set_error("SPHB:arg", "null output");
    SPHB_CUDA(cudaMemcpyAsync(c->h_scalars, c->sc(), sizeof(DevScalars), cudaMemcpyDeviceToHost, c->st));
    c->h_ints[0] = c->h_ints[1] = c->h_ints[2] = 0;
    if (c->planned) {
        // (a slab's step may be replayed by a graph of its caller, which rebuilds the list behind this library's back)
        if (c->slab.on) c->nw.half_valid = false;
        if (c->have_list) SPHB_TRY(neighbor_half_total(c->nw, c->st));
        SPHB_CUDA(cudaMemcpyAsync(c->h_ints, c->nw.totals.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        SPHB_CUDA(cudaMemcpyAsync(c->h_ints + 2, c->nw.flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    }
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    SPHB_CUDA(cudaGetLastError());
    out->t = c->h_scalars->t;
    out->dt = c->h_scalars->dt;
    out->vmax = c->h_scalars->vmax;
    out->tau_bottom = c->h_scalars->tau_bottom;
    out->tau_top = c->h_scalars->tau_top;
    out->step = c->h_scalars->step;
    out->n_half_pairs = c->have_list ? c->h_ints[0] : 0;
    out->n_csr_entries = c->have_list ? c->h_ints[1] : 0;
    out->error_flags = c->h_ints[2];
    return 0;
}

// One pass of the reference main-loop body, m:251-283.
SPHB_API int sphb_ctx_step(sphb_ctx *c, double remain)
{
    CTX_ENTER(true);
    c->h_step += 1;
    c->last_inner = 1;
    c->last_remain = remain;
    tick(c, 0);
    SPHB_TRY(do_density(c, true, remain));                                            // m:259 (dt), m:254-255
    tick(c, 1);
    tick(c, 2);
    SPHB_TRY(do_fused_first_substep(c, true, true));                                  // m:256-269
    tick(c, 3);
    const int si = c->prm.sort_interval;
    if (si > 0 && c->h_step % si == 0 && c->h_step != 1) SPHB_TRY(do_sort(c));        // m:272-278
    tick(c, 4);
    SPHB_TRY(do_neighbor(c));                                                         // m:280
    tick(c, 5);
    SPHB_TRY(do_shear(c));                                                            // m:281-283
    tick(c, 6);
    return 0;
}

// One reference step for a caller that keeps the state on the HOST between steps (upload -> this -> download, the
// calling pattern of the MEX drop-in): the list for the uploaded positions (m:280, the search that closed the
// caller's previous step), the wall-shear monitor on it (m:281-283), then the physics m:254-270.  Unlike
// sphb_ctx_step it does not rebuild the list for the NEW positions: the next upload would discard it.
SPHB_API int sphb_ctx_step_uploaded(sphb_ctx *c, double remain)
{
    CTX_ENTER(false);
    if (!c->planned) return set_error("SPHB:state", "upload positions first");
    if (!c->have_list) SPHB_TRY(do_neighbor(c));
    SPHB_TRY(do_shear(c));
    return ctx_step_physics(c, remain);
}

// The same step as ONE call on host arrays (what the reference's script does per step with its two MEX calls, m:280 +
// the modes of m:254-270): upload -> list for the uploaded positions -> shear monitor -> physics -> download.  The
// transfers overlap the work where the data flow allows: the list build needs positions and velocities only and runs
// while rho and drho_dt are still on the wire (a second stream); everything else is sphb_ctx_upload /
// sphb_ctx_step_uploaded / sphb_ctx_download, with the same results bit for bit.
SPHB_API int sphb_ctx_step_host(sphb_ctx *c, double remain, const double *pos, const double *vel, const double *rho,
                                const double *drho_dt, double *pos_out, double *vel_out, double *rho_out, double *p_out,
                                double *drho_dt_out, double *force_out)
{
    clear_error();
    if (!c || !pos || !vel || !rho || !drho_dt) return set_error("SPHB:arg", "null argument");
    SPHB_TRY(bind_device(c));
    if (c->slab.on) return set_error("SPHB:state", "not available on a slab of a sharded channel");
    SPHB_TRY(ctx_flags_poll(c, true));   // what an earlier asynchronous run left behind
    if (!c->aux_st) SPHB_CUDA(cudaStreamCreateWithFlags(&c->aux_st, cudaStreamNonBlocking));
    if (!c->aux_ev) SPHB_CUDA(cudaEventCreateWithFlags(&c->aux_ev, cudaEventDisableTiming));
    const int nt = c->prm.n_total;
    const size_t n = (size_t)nt;
    double *stage = c->stage.as<double>();
    // (the previous call returned after a synchronisation: nothing on the main stream still reads rho / drho_dt)
    SPHB_CUDA(cudaMemcpyAsync(c->rho.p, rho, n * sizeof(double), cudaMemcpyHostToDevice, c->aux_st));
    SPHB_CUDA(cudaMemcpyAsync(c->drho_dt.p, drho_dt, n * sizeof(double), cudaMemcpyHostToDevice, c->aux_st));
    SPHB_CUDA(cudaEventRecord(c->aux_ev, c->aux_st));
    SPHB_CUDA(cudaMemcpyAsync(stage, pos, 2 * n * sizeof(double), cudaMemcpyHostToDevice, c->st));
    launch_pack2(stage, c->pos.as<double2>(), nt, c->st);
    SPHB_CUDA(cudaMemcpyAsync(stage + 2 * n, vel, 2 * n * sizeof(double), cudaMemcpyHostToDevice, c->st));
    launch_pack2(stage + 2 * n, c->vel.as<double2>(), nt, c->st);
    double lo = 0.0, hi = 0.0;
    if (!c->planned) {   // first call: the lists are planned for the y range of these rows
        host_y_range(pos + n, n, lo, hi);
        SPHB_CUDA(cudaStreamSynchronize(c->st));
        SPHB_CUDA(cudaStreamSynchronize(c->aux_st));
        drop_graphs(c);
        SPHB_TRY(plan_lists(c, lo, hi));
    }
    // The whole step is enqueued without looking at the rows on the host: the grid of the list is derived on the
    // device, and a build that outgrows what was planned (cells, rows, tile slots) raises a flag that stops the
    // kernels behind it, so the state stays what was uploaded.  Only then is the host's y range needed -- to plan again
    // and repeat the step from the same, untouched input arrays.  (A pass over the host array per call, as
    // sphb_ctx_upload makes, costs as much as the upload itself.)
    for (int attempt = 0;; ++attempt) {
        c->have_list = false;
        c->nw.grid_ready = false;
        c->mirrors_stale = true;
        if (attempt == 0) SPHB_TRY(csr_vmax(c->P, c->vel.as<double2>(), c->sc(), c->st));
        SPHB_TRY(do_neighbor(c));                                   // m:280 of the caller's previous step
        if (attempt == 0) SPHB_CUDA(cudaStreamWaitEvent(c->st, c->aux_ev, 0));
        SPHB_TRY(do_shear(c));                                      // m:281-283
        SPHB_TRY(ctx_step_physics(c, remain));                      // m:254-270
        SPHB_TRY(enqueue_download(c, pos_out, vel_out, rho_out, p_out, drho_dt_out, force_out, nullptr, nullptr, nullptr));
        SPHB_CUDA(cudaMemcpyAsync(c->h_ints + 5, c->nw.flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
        SPHB_CUDA(cudaStreamSynchronize(c->st));
        SPHB_CUDA(cudaGetLastError());
        int f = c->h_ints[5];
        if (!f) return 0;
        if ((f & ~(kErrCellCap | kErrCsrCap | kErrTileCap)) || attempt > 0) {
            if (f & kErrNonFinite)
                return set_error("SPHB:diverged", "non-finite density or velocity in the uploaded step (flags %d)", f);
            return set_error("SPHB:capacity", "device error flags %d: the step was not taken", f);
        }
        // capacity: clear the flags, plan for the rows as they are now, build exactly (grows the list / the tile slots)
        f = 0;
        SPHB_CUDA(cudaMemcpy(c->nw.flags.p, &f, sizeof(int), cudaMemcpyHostToDevice));
        host_y_range(pos + n, n, lo, hi);
        drop_graphs(c);
        if (lo < c->y_lo || hi > c->y_hi) SPHB_TRY(plan_lists(c, std::min(lo, c->y_lo), std::max(hi, c->y_hi)));
        c->list_sized = false;
        c->h_step -= 1;   // (the stopped attempt counted a step the device did not take)
    }
}

namespace {

// clock, dt and flags only (one small D2H, no kernel): the per-step read-back of sphb_ctx_run_until
int status_light(sphb_ctx *c, sphb_status *out)
{
    SPHB_CUDA(cudaMemcpyAsync(c->h_scalars, c->sc(), sizeof(DevScalars), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaMemcpyAsync(c->h_ints + 2, c->nw.flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    out->t = c->h_scalars->t;
    out->dt = c->h_scalars->dt;
    out->vmax = c->h_scalars->vmax;
    out->step = c->h_scalars->step;
    out->error_flags = c->h_ints[2];
    return 0;
}

// A list that outgrew its row or tile capacity during a run (the state itself is intact: the build only
// reads it): clear the flags, rebuild with the exact, growing build and redo the monitor of the step.
int recover_list_capacity(sphb_ctx *c)
{
    int f = 0;
    SPHB_CUDA(cudaMemcpyAsync(&f, c->nw.flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    f &= ~(kErrCsrCap | kErrTileCap);
    SPHB_CUDA(cudaMemcpyAsync(c->nw.flags.p, &f, sizeof(int), cudaMemcpyHostToDevice, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    c->list_sized = false;
    SPHB_TRY(do_neighbor(c));
    SPHB_TRY(do_shear(c));
    return 0;
}

}  // namespace

SPHB_API int sphb_ctx_run_until(sphb_ctx *c, double t_target, double t_end, long long max_steps,
                                long long *steps_done)
{
    CTX_ENTER(true);
    long long n = 0;
    sphb_status s{};
    SPHB_TRY(status_light(c, &s));
    // No criterion of m:519-527 allows a step longer than dt_max (max|v| = 0), so the target cannot cut a step short
    // (m:259) during the next floor((target - t) / dt_max) - 1 steps: those run as replayed graphs with remain = inf --
    // the same arithmetic -- without the per-step read-back of the clock; the last few steps before the target go one
    // by one as the reference loop does.
    const PhysParams &Pp = c->P;
    const double dt_max = fmin(fmin(0.25 * Pp.h / fmax(Pp.c_f, 1e-12), 0.125 * Pp.h * Pp.h / fmax(Pp.nu, 1e-12)),
                               0.25 * sqrt(Pp.h / fmax(fabs(Pp.g), 1e-12)));
    while (s.t < t_target - 1e-12 && (max_steps <= 0 || n < max_steps)) {   // m:250
        if (!c->slab.on && !c->kprof.on && !c->timing) {
            double safe_d = floor((fmin(t_target, t_end) - s.t) / dt_max) - 1.0;
            if (max_steps > 0) safe_d = fmin(safe_d, (double)(max_steps - n));
            if (safe_d >= 2.0) {
                const long long step0 = s.step;
                SPHB_TRY(sphb_ctx_run_dual(c, (int)fmin(safe_d, 4096.0), 1));
                SPHB_TRY(ctx_flags_watch(c));
                SPHB_TRY(ctx_flags_poll(c, true));   // a list that outgrew its capacity: rebuilt, the skipped steps redone
                SPHB_TRY(status_light(c, &s));
                if (s.error_flags) return set_error("SPHB:capacity", "device error flags %d at step %lld", s.error_flags, s.step);
                if (s.dt < 1e-14) return set_error("SPHB:diverged", "time step collapsed (dt=%g) at t=%g step=%lld", s.dt, s.t, s.step);
                n += s.step - step0;
                continue;
            }
        }
        const double remain = fmin(t_target - s.t, t_end - s.t);           // m:252
        SPHB_TRY(sphb_ctx_step(c, remain));
        SPHB_TRY(status_light(c, &s));
        ++n;
        if ((s.error_flags & (kErrCsrCap | kErrTileCap)) && !(s.error_flags & ~(kErrCsrCap | kErrTileCap))) {
            SPHB_TRY(recover_list_capacity(c));
            SPHB_TRY(status_light(c, &s));
        }
        if (s.error_flags) return set_error("SPHB:capacity", "device error flags %d at step %lld", s.error_flags, s.step);
        if (s.dt < 1e-14) return set_error("SPHB:diverged", "time step collapsed (dt=%g) at t=%g step=%lld", s.dt, s.t, s.step);   // m:260-263
    }
    if (steps_done) *steps_done = n;
    return 0;
}

namespace {

int enqueue_outer(sphb_ctx *c, int n_inner, bool with_sort, double first_remain = std::numeric_limits<double>::infinity())
{
    const double inf = std::numeric_limits<double>::infinity();
    SPHB_TRY(do_density(c, true, first_remain));   // (publishes the step's dt: m:259)
    SPHB_TRY(do_fused_first_substep(c, true, n_inner == 1));
    for (int s = 1; s < n_inner; ++s) SPHB_TRY(do_substep(c, inf, 0.0, true, s == n_inner - 1));
    if (with_sort) SPHB_TRY(do_sort(c));
    SPHB_TRY(do_neighbor(c));
    SPHB_TRY(do_shear(c));
    return 0;
}

}  // namespace

}  // extern "C"

namespace {

// One outer step of the DEVICE-SIDE dual-criteria scheduler: Dt and every dt are computed on the device, the number
// of acoustic substeps is data dependent -- `max_inner` substeps are enqueued and those beyond ceil(Dt / dt) return at
// once (DevScalars::sub_on), so the whole outer step is one fixed launch sequence (capturable as a CUDA graph) with
// no host round trip.
int enqueue_outer_auto(sphb_ctx *c, int max_inner, double remain, bool with_sort)
{
    const double inf = std::numeric_limits<double>::infinity();
    SPHB_TRY(csr_dual_begin(c->P, c->sc(), remain, c->st));   // Dt, first dt (the density pass evaluates p(half) with it)
    c->launches += 1;
    SPHB_TRY(do_density(c));
    SPHB_TRY(do_fused_first_substep(c, true, false));
    for (int s = 1; s < max_inner; ++s) SPHB_TRY(do_substep(c, inf, 0.0, true, false));
    SPHB_TRY(csr_dual_end(c->sc(), c->st));
    c->launches += 1;
    if (with_sort) SPHB_TRY(do_sort(c));
    SPHB_TRY(do_neighbor(c));
    SPHB_TRY(do_shear(c));
    return 0;
}

// Error flags behind asynchronous runs.  The kernels of a step return at once while a flag is up (step_poisoned), so
// the state of a replayed run stays at the last consistent step; the host learns of it from a 4-byte copy into pinned
// memory enqueued behind every run (flags_watch) and looked at when the next call comes (flags_poll, non-blocking) or
// when the caller waits (sphb_ctx_sync / status / download, blocking).  A list that outgrew its fitted capacity is
// rebuilt with the exact, growing build and the outer steps the poisoned run skipped are taken again; anything else
// (non-finite state, cell / slab capacity) is returned as an error.
int flags_watch(sphb_ctx *c)
{
    if (!c->planned) return 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    (void)cudaStreamIsCapturing(c->st, &cap);
    if (cap != cudaStreamCaptureStatusNone) return 0;
    if (!c->flags_ev) SPHB_CUDA(cudaEventCreateWithFlags(&c->flags_ev, cudaEventDisableTiming));
    SPHB_CUDA(cudaMemcpyAsync(c->h_ints + 4, c->nw.flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaEventRecord(c->flags_ev, c->st));
    c->flags_pending = true;
    return 0;
}

int flags_handle(sphb_ctx *c)   // the copy of flags_watch has landed
{
    c->flags_pending = false;
    const int f = c->h_ints[4];
    if (!f) return 0;
    // (rare path) where the device stands: the clock stopped at the last consistent step
    SPHB_CUDA(cudaMemcpyAsync(&c->h_scalars[1], c->sc(), sizeof(DevScalars), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    const long long dev_step = c->h_scalars[1].step;
    if (f & kErrNonFinite)
        return set_error("SPHB:diverged", "non-finite density or velocity at step %lld (t = %g): the run is frozen there",
                         dev_step, c->h_scalars[1].t);
    if (f & ~(kErrCsrCap | kErrTileCap))
        return set_error("SPHB:capacity", "device error flags %d at step %lld: the run is frozen there", f, dev_step);
    if (c->slab.on) return set_error("SPHB:capacity", "neighbour list capacity exceeded in a slab (flags %d) at step %lld", f, dev_step);
    // list capacity: rebuild exactly (grows the list / the tile slots), then redo what the poisoned run skipped
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    SPHB_TRY(recover_list_capacity(c));
    drop_graphs(c);
    const long long missing = c->h_step - dev_step;
    const int inner = std::max(1, c->last_inner);
    // (a single step that was cut short by the caller's target, m:259, is taken again with the same cut)
    for (long long o = 0; o < missing / inner; ++o)
        SPHB_TRY(enqueue_outer(c, inner, false, (missing == 1 && inner == 1) ? c->last_remain : std::numeric_limits<double>::infinity()));
    return 0;
}

int flags_poll(sphb_ctx *c, bool block)
{
    if (!c->flags_pending) return 0;
    if (block) SPHB_CUDA(cudaEventSynchronize(c->flags_ev));
    else if (cudaEventQuery(c->flags_ev) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return flags_handle(c);
}

}  // namespace

int ctx_flags_poll(sphb_ctx *c, bool block) { return flags_poll(c, block); }
int ctx_flags_watch(sphb_ctx *c) { return flags_watch(c); }

extern "C" {

// Dual-criteria schedule as a replayed CUDA graph (one graph per outer step).
SPHB_API int sphb_ctx_run_dual(sphb_ctx *c, int n_outer, int n_inner)
{
    CTX_ENTER(true);
    if (n_inner < 1 || n_outer < 0) return set_error("SPHB:arg", "n_inner >= 1 and n_outer >= 0 required");
    SPHB_TRY(flags_poll(c, false));   // what an earlier asynchronous run left behind
    c->last_inner = n_inner;
    c->last_remain = std::numeric_limits<double>::infinity();
    const int si = c->prm.sort_interval;
    if (c->graph_inner != n_inner) drop_graphs(c);
    for (int o = 0; o < n_outer; ++o) {
        if ((o & 63) == 63) {   // long runs: look now and then without stalling the queue
            SPHB_TRY(flags_watch(c));
            SPHB_TRY(flags_poll(c, false));
        }
        c->outer_count += 1;
        c->h_step += n_inner;
        // m:272: mod(step, sort_interval) == 0 && step ~= 1 on the GLOBAL step counter (it survives restarts:
        // sphb_ctx_set_time seeds h_step); with n_inner > 1 an outer step is n_inner counts of it
        const long long kstep = c->h_step / n_inner;
        const bool with_sort = si > 0 && kstep % si == 0 && kstep != 1;
        if (with_sort) {
            // the permutation swaps buffer pointers: run eagerly, then re-capture on the new pointers
            drop_graphs(c);
            SPHB_TRY(enqueue_outer(c, n_inner, true));
            if (c->kprof.on) SPHB_TRY(c->kprof.collect(c->st));
            continue;
        }
        if (c->kprof.on) {   // per-kernel events cannot live inside a replayed graph: run this outer step eagerly
            SPHB_TRY(enqueue_outer(c, n_inner, false));
            SPHB_TRY(c->kprof.collect(c->st));
            continue;
        }
        if (!c->graph[0]) {
            const long long l0 = c->launches;
            cudaGraph_t g = nullptr;
            SPHB_CUDA(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue_outer(c, n_inner, false);
            cudaError_t e = cudaStreamEndCapture(c->st, &g);
            if (rc) {
                if (g) cudaGraphDestroy(g);
                return rc;
            }
            if (e != cudaSuccess) return set_error("SPHB:cuda", "graph capture failed: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&c->graph[0], g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return set_error("SPHB:cuda", "graph instantiate failed: %s", cudaGetErrorString(e));
            c->graph_inner = n_inner;
            c->graph_launches = c->launches - l0;
            c->launches = l0;
        }
        SPHB_CUDA(cudaGraphLaunch(c->graph[0], c->st));
        c->nw.half_valid = false;   // the replayed build rewrote the list
        c->launches += c->graph_launches;
    }
    // a 4-byte copy of the flags behind the run, not behind every call of a step-by-step caller (a copy is a stream
    // operation of its own, ~3 us): every 32nd call, and whenever the caller waits (sphb_ctx_sync / download)
    if (n_outer > 1 || (++c->watch_skip & 31) == 0) SPHB_TRY(flags_watch(c));
    return 0;
}

// Dual-criteria schedule with the criteria evaluated ON THE DEVICE (BASELINE.json north_star (3); the reference
// describes the scheme in README.md:61-65 and has no code for it at HEAD -- parity unpinned):
//   outer step  Dt = min(0.25 h / max|v|, 0.125 h^2 / nu, 0.25 sqrt(h / |g|), remain)      advection / viscous / body
//   substeps    dt = min(0.25 h / (c_f + max|v|), Dt_left) until Dt is covered, at most max_inner of them
// density + KGC, viscous force and transport shift once per outer step, the list rebuilt after it.  max|v| comes from
// the warp-shuffle + atomicMax reduction of the continuity kernel; Dt, dt and the substep count never leave the device.
// Replayed as one CUDA graph per outer step when remain is infinite.  The counts actually taken are in
// sphb_ctx_dual_counts.
SPHB_API int sphb_ctx_run_dual_auto(sphb_ctx *c, int n_outer, int max_inner, double remain)
{
    CTX_ENTER(true);
    if (max_inner < 1 || n_outer < 0) return set_error("SPHB:arg", "max_inner >= 1 and n_outer >= 0 required");
    if (c->slab.on) return set_error("SPHB:state", "the device-side scheduler is not available on a slab of a sharded channel");
    SPHB_TRY(flags_poll(c, true));
    const int si = c->prm.sort_interval;
    const bool graphable = !(remain < std::numeric_limits<double>::infinity()) && !c->kprof.on;
    if (c->graph_inner != -max_inner) drop_graphs(c);   // (negative: the auto schedule with this budget)
    for (int o = 0; o < n_outer; ++o) {
        c->outer_count += 1;
        const bool with_sort = si > 0 && c->outer_count % si == 0 && c->outer_count != 1;
        if (with_sort || !graphable) {
            if (with_sort) drop_graphs(c);   // the permutation swaps buffer pointers
            SPHB_TRY(enqueue_outer_auto(c, max_inner, remain, with_sort));
            if (c->kprof.on) SPHB_TRY(c->kprof.collect(c->st));
            continue;
        }
        if (!c->graph[0]) {
            const long long l0 = c->launches;
            cudaGraph_t g = nullptr;
            SPHB_CUDA(cudaStreamBeginCapture(c->st, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue_outer_auto(c, max_inner, remain, false);
            cudaError_t e = cudaStreamEndCapture(c->st, &g);
            if (rc) {
                if (g) cudaGraphDestroy(g);
                return rc;
            }
            if (e != cudaSuccess) return set_error("SPHB:cuda", "graph capture failed: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&c->graph[0], g, 0);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return set_error("SPHB:cuda", "graph instantiate failed: %s", cudaGetErrorString(e));
            c->graph_inner = -max_inner;
            c->graph_launches = c->launches - l0;
            c->launches = l0;
        }
        SPHB_CUDA(cudaGraphLaunch(c->graph[0], c->st));
        c->nw.half_valid = false;
        c->launches += c->graph_launches;
    }
    // the host mirror of the step counter follows the device (the substep count is the device's decision)
    SPHB_CUDA(cudaMemcpyAsync(c->h_scalars, c->sc(), sizeof(DevScalars), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaMemcpyAsync(c->h_ints + 2, c->nw.flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    c->h_step = c->h_scalars->step;
    if (c->h_ints[2])
        return set_error(c->h_ints[2] & kErrNonFinite ? "SPHB:diverged" : "SPHB:capacity",
                         "device error flags %d at step %lld: the run stands at its last consistent step", c->h_ints[2],
                         c->h_scalars->step);
    return 0;
}

// outer steps and acoustic substeps taken so far by sphb_ctx_run_dual_auto, and the last outer step Dt
SPHB_API int sphb_ctx_dual_counts(sphb_ctx *c, long long *outer_steps, long long *sub_steps, double *last_Dt)
{
    CTX_ENTER(false);
    SPHB_CUDA(cudaMemcpyAsync(c->h_scalars, c->sc(), sizeof(DevScalars), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    if (outer_steps) *outer_steps = c->h_scalars->outer_steps;
    if (sub_steps) *sub_steps = c->h_scalars->sub_steps;
    if (last_Dt) *last_Dt = c->h_scalars->Dt;
    return 0;
}

// Ensembles of independent channels (BASELINE.json configs[3]; SURVEY.md 8(e) "replicas only"): every
// channel is a context of its own with its own stream and its own cached graph; this call enqueues
// `n_outer` outer steps of each of them round-robin from one host loop, so the small kernels of many
// channels overlap on the device.  No data-path communication between channels.
SPHB_API int sphb_ensemble_run_dual(sphb_ctx **ctxs, int n_ctx, int n_outer, int n_inner)
{
    clear_error();
    if (!ctxs || n_ctx < 0) return set_error("SPHB:arg", "bad ensemble");
    // A graph launch costs the host ~10 us, a small channel's step the device about as much once several
    // are in flight: the launches are spread over a few host threads (one slice of the channels each; a
    // context is only ever touched by its own thread, error state is thread-local).
    const int n_thr = std::max(1, std::min({8, n_ctx / 16, (int)std::thread::hardware_concurrency() / 2}));
    if (n_thr == 1) {
        for (int o = 0; o < n_outer; ++o)
            for (int k = 0; k < n_ctx; ++k) SPHB_TRY(sphb_ctx_run_dual(ctxs[k], 1, n_inner));
        return 0;
    }
    std::vector<int> rc((size_t)n_thr, 0);
    std::vector<std::string> ids((size_t)n_thr), msgs((size_t)n_thr);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_thr; ++t)
        pool.emplace_back([&, t]() {
            const int k0 = (int)((long long)n_ctx * t / n_thr), k1 = (int)((long long)n_ctx * (t + 1) / n_thr);
            for (int o = 0; o < n_outer && !rc[t]; ++o)
                for (int k = k0; k < k1 && !rc[t]; ++k) rc[t] = sphb_ctx_run_dual(ctxs[k], 1, n_inner);
            if (rc[t]) {
                ids[t] = last_error_id();
                msgs[t] = last_error_msg();
            }
        });
    for (auto &th : pool) th.join();
    for (int t = 0; t < n_thr; ++t)
        if (rc[t]) return set_error(ids[t].c_str(), "%s", msgs[t].c_str());
    return 0;
}

SPHB_API int sphb_ensemble_sync(sphb_ctx **ctxs, int n_ctx)
{
    clear_error();
    if (!ctxs || n_ctx < 0) return set_error("SPHB:arg", "bad ensemble");
    for (int k = 0; k < n_ctx; ++k) SPHB_TRY(sphb_ctx_sync(ctxs[k]));
    return 0;
}

SPHB_API int sphb_ctx_profile(sphb_ctx *c, double x0, double half_width, int n_bins, double *u_mean, double *count)
{
    CTX_ENTER(false);
    if (n_bins < 1 || !u_mean) return set_error("SPHB:arg", "bad profile arguments");
    std::vector<double> edges((size_t)n_bins + 1), host(2 * (size_t)n_bins);
    for (int k = 0; k <= n_bins; ++k)   // linspace(0, DH, n_bins + 1), m:581
        edges[k] = (k == n_bins) ? c->prm.DH : 0.0 + (double)k * ((c->prm.DH - 0.0) / (double)n_bins);
    SPHB_TRY(c->prof.reserve((3 * (size_t)n_bins + 1) * sizeof(double)));
    double *d_edges = c->prof.as<double>(), *d_sum = d_edges + n_bins + 1, *d_cnt = d_sum + n_bins;
    SPHB_CUDA(cudaMemcpyAsync(d_edges, edges.data(), edges.size() * sizeof(double), cudaMemcpyHostToDevice, c->st));
    SPHB_CUDA(cudaMemsetAsync(d_sum, 0, 2 * (size_t)n_bins * sizeof(double), c->st));
    SPHB_TRY(csr_profile(c->P, c->view(), x0, half_width, n_bins, d_edges, d_sum, d_cnt, c->st));
    c->launches += 1;
    SPHB_CUDA(cudaMemcpyAsync(host.data(), d_sum, host.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    for (int k = 0; k < n_bins; ++k) {
        const double cnt = host[(size_t)n_bins + k];
        u_mean[k] = cnt > 0.0 ? host[k] / cnt : std::numeric_limits<double>::quiet_NaN();   // m:588-589
        if (count) count[k] = cnt;
    }
    return 0;
}

SPHB_API int sphb_ctx_enable_timing(sphb_ctx *c, int on)
{
    CTX_ENTER(false);
    c->timing = on != 0;
    return 0;
}

SPHB_API int sphb_ctx_timing(sphb_ctx *c, sphb_timing *out)
{
    CTX_ENTER(false);
    if (!out) return set_error("SPHB:arg", "null output");
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    sphb_timing t{};
    if (c->timing) {
        float ms[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 6; ++k)
            if (cudaEventElapsedTime(&ms[k], c->ev[k], c->ev[k + 1]) != cudaSuccess) {
                ms[k] = 0.f;
                (void)cudaGetLastError();
            }
        t.density_ms = ms[0];
        t.visc_transport_ms = ms[1];
        t.verlet_ms = ms[2];
        t.sort_ms = ms[3];
        t.neighbor_ms = ms[4];
        t.shear_ms = ms[5];
        t.total_ms = ms[0] + ms[1] + ms[2] + ms[3] + ms[4] + ms[5];
    }
    t.launches = c->launches;
    *out = t;
    return 0;
}

SPHB_API int sphb_ctx_kernel_profile(sphb_ctx *c, int enable, int reset, sphb_kprof *out)
{
    CTX_ENTER(false);
    SPHB_TRY(c->kprof.collect(c->st));
    if (out) {
        out->n = kPfCount;
        for (int k = 0; k < SPHB_KPROF_SLOTS; ++k) {
            out->name[k] = k < kPfCount ? prof_name(k) : "";
            out->ms[k] = k < kPfCount ? c->kprof.ms[k] : 0.0;
            out->calls[k] = k < kPfCount ? c->kprof.n[k] : 0;
        }
    }
    if (reset) c->kprof.reset();
    if (enable >= 0) c->kprof.on = enable != 0;
    return 0;
}

SPHB_API void *sphb_ctx_stream(sphb_ctx *c) { return c ? (void *)c->st : nullptr; }

// Launch on a stream of the caller (e.g. torch's current stream, so that NCCL collectives issued through
// torch.distributed are ordered with the library's kernels without host synchronisation).  NULL restores
// the context's own stream.  Cached graphs are dropped.
SPHB_API int sphb_ctx_set_stream(sphb_ctx *c, void *stream)
{
    CTX_ENTER(false);
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    drop_graphs(c);
    if (!c->own_st) c->own_st = c->st;
    c->st = stream ? (cudaStream_t)stream : c->own_st;
    return 0;
}

SPHB_API int sphb_ctx_sync(sphb_ctx *c)
{
    CTX_ENTER(false);
    if (c->have_list && !c->flags_pending) SPHB_TRY(flags_watch(c));
    SPHB_TRY(flags_poll(c, true));   // an asynchronous run that raised a device flag is recovered (or reported) here
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
