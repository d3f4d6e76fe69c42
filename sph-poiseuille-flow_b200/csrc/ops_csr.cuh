// ops_csr.cuh -- device-resident state views and launchers of the CSR gather kernels (ops_csr.cu).
#pragma once
#include "common.cuh"
#include "tile.cuh"

namespace sphb {

// Scalars that live in device memory so that no step needs a host round trip.
struct DevScalars {
    double t;
    double dt;
    double vmax;                    // max |v| over fluid rows, as of the last reduction
    double tau_bottom, tau_top;
    unsigned long long vmax_bits;   // running atomicMax target (bit image of a non-negative double)
    long long step;
    double y0_fluid;                // min fluid y (cell sort key origin, m:565)
    long long y0_key;
    unsigned int ticket;            // CTAs of the continuity kernel that have finished (clock advance)
    unsigned int shear_ticket;      // CTAs of the wall-shear kernel that have finished (final sum by the last one)
    // dual-criteria schedule on the device (sphb_ctx_run_dual_auto; README.md:61-65 of the reference): the outer step
    // Dt from the advection / viscous / body criteria, consumed by acoustic substeps dt = min(acoustic, Dt_left)
    int dual_on;                    // 1 while an outer step of the device-side scheduler is in flight
    int sub_on;                     // 0: Dt is used up -- the kernels of the remaining (budgeted) substeps return at once
    double Dt;                      // outer step of the current outer step
    double Dt_left;                 // what the substeps still have to cover
    long long outer_steps;          // outer steps taken by the device-side scheduler
    long long sub_steps;            // acoustic substeps taken by it
};

struct PhysParams {
    int nf, nt;
    int nfo;           // fluid rows OWNED by this context (= nf unless it is a slab with halo copies)
    double h, DL, DH, rho0, p0, c_f, mu, nu, g, inv_sigma0, transport_coeff;
    double sigma;      // 10/(7 pi h^2)
    double inv_h;      // 1/h
    double sigma_h;    // sigma/h
    double inv_rho0;   // 1/rho0
    const int *dev_counts;   // slab mode: {nf, nfo, nt} live in device memory (they change every step and the
                             // host never waits for them); NULL = the values above
};
#ifdef __CUDACC__
// every thread reads the same word; callers branch on it CTA-uniformly
__device__ __forceinline__ bool step_poisoned(const int *flags) { return *reinterpret_cast<const volatile int *>(flags) != 0; }
__device__ __forceinline__ void load_counts(PhysParams &P)
{
    if (P.dev_counts) {
        P.nf = P.dev_counts[0];
        P.nfo = P.dev_counts[1];
        P.nt = P.dev_counts[2];
    }
}
#endif

// SoA state in HBM.  2-vectors are double2 (one 128-bit load per particle), B is one 32-byte record.
struct StateView {
    double2 *pos, *vel, *vel_new, *wall_vel, *force, *force_prior;
    double *rho, *p, *drho_dt, *mass, *Vol;
    Mat2 *B;
    double2 *rp_half;   // (rho, p) at the half step of the current substep (ph:857-862), written by k_time_step
    // neighbour structures (frozen geometry = entry positions at list-build time), see neighbor.cu / tile.cuh
    const double2 *snap;      // wrapped position of particle i at build time
    const GridDesc *grid;     // device-side grid (n_cell_y)
    const int *ent_key;       // particle << 3 | multi << 2 | kind, cell-sorted
    const double2 *ent_pos;   // entry position (ghost x already shifted)
    const int *ent_of;        // real entry of particle i
    const TileDesc *tiles;    // one descriptor per cell tile
    const int *ell_cnt;       // neighbours of the walker whose real entry is e
    const uint4 *ell_q;       // ell_q[(k / 8) * ell_stride + e]: eight 16-bit (slot | shift) entries
    // entry-ordered mirrors (index = entry; ghost entries carry their particle's values): what the gather kernels stage with bulk
    // asynchronous copies instead of gathering it through ent_key (async_copy.cuh, ops_csr.cu)
    double2 *e_vel;           // velocity at list-build time (wall entry: wall velocity)      written by the list build
    double2 *e_w;             // (wall weight m / rho0 or -1 for fluid, 0)                     written by the list build
    double2 *e_pV;            // (half-step pressure, volume; volume < 0 marks a wall entry)  written by k_density
    double2 *e_b1, *e_b2;     // rows of B                                                     written by k_kgc
    double2 *e_vn;            // new velocity (wall entry: wall velocity)                      written by the force kernels
    double2 *e_rm;            // (half-step density, mass) of a fluid walker                   written by the density kernel
    const int *ghost_ent;     // [2 i + kind - 1]: entry of particle i's ghost of that kind (validated against ent_key): the
                              // producers above write the mirrors of a row's ghost entries as well (ghost_store)
    // device error flags (sphb_status.error_flags).  Once a flag is up the state-changing kernels of the step return
    // at once ("poisoned"): inside a replayed graph the state then stays where the last consistent step left it
    // until the host looks (sphb_ctx_run_dual, sphb_ctx_sync), instead of advancing on a truncated list.
    int *flags;
    size_t ell_stride;
    int ell_cap;
    int tile_cap;             // shared-memory slots per tile
    int tile_w, ncx;          // tile width in cells, cell columns
    unsigned n_tiles;
};

// publish_dt: the launch opens a reference step -- it derives dt from max|v| (m:519-527, cut by `remain`, m:259),
// publishes it and re-arms the max|v| reduction (csr_publish_dt folded in)
int csr_density(const PhysParams &P, const StateView &S, DevScalars *sc, cudaStream_t st, bool publish_dt = false,
                double remain = 0.0);   // P1 pass 1 (+ half-step p)
int csr_kgc(const PhysParams &P, const StateView &S, cudaStream_t st);                           // P1 pass 2
int csr_visc_transport(const PhysParams &P, const StateView &S, bool do_visc, bool add_gravity, bool do_transport,
                       cudaStream_t st);                                                         // P2 + P3
// dt = min(acoustic, viscous, body, remain) from the device vmax (m:519-527); fixed_dt > 0 overrides
int csr_time_step(const PhysParams &P, const StateView &S, DevScalars *sc, double remain, double fixed_dt,
                  cudaStream_t st);
int csr_pressure(const PhysParams &P, const StateView &S, const DevScalars *sc, cudaStream_t st);    // P4 + kick
// P5 + finish (+ clock); F.g != NULL: also the y range / grid / zeroed cell counters of the list build that follows
int csr_continuity(const PhysParams &P, const StateView &S, DevScalars *sc, bool wrap, bool advance_clock,
                   const GridFuse &F, cudaStream_t st);
// Device-side dual-criteria scheduler: opens an outer step (Dt = min(0.25 h / max|v|, 0.125 h^2 / nu, 0.25 sqrt(h/|g|),
// remain), first substep dt = min(acoustic, Dt)) / closes it (dual_on = 0).  Between the two the substep kernels
// (csr_time_step, csr_pressure, csr_continuity) take dt = min(acoustic, Dt_left) and return at once when Dt is used up.
int csr_dual_begin(const PhysParams &P, DevScalars *sc, double remain, cudaStream_t st);
int csr_dual_end(DevScalars *sc, cudaStream_t st);
int csr_force_fused(const PhysParams &P, const StateView &S, const DevScalars *sc, bool add_gravity,
                    cudaStream_t st);                                                            // P2 + P3 + P4 + kick
int csr_vmax(const PhysParams &P, const double2 *vel, DevScalars *sc, cudaStream_t st);          // standalone reduction
int csr_wrap(const PhysParams &P, double2 *pos, cudaStream_t st);                                // m:570-577
int csr_wall_shear(const PhysParams &P, const StateView &S, DevScalars *sc, double *partials, cudaStream_t st);  // P8
int csr_profile(const PhysParams &P, const StateView &S, double x0, double half_width, int n_bins,
                const double *edges_dev, double *sum_dev, double *cnt_dev, cudaStream_t st);
int csr_refresh_mirrors(const PhysParams &P, const StateView &S, cudaStream_t st);   // StateView::e_* from the rows
int csr_configure(int tile_cap);     // grants the gather kernels the shared memory `tile_cap` slots need
int csr_max_tile_slots();            // what the widest gather kernel (pressure) can stage

// ---- batched launches over an ensemble of independent channels (batch.cu; BASELINE configs[3]) ----------
// Every channel keeps its own state, list and scalars; a batched launch runs the SAME kernel body for all of
// them in one grid: blockIdx.x = the channel's own block index (tile), blockIdx.y = channel.  The per-channel
// arguments live in a device array of BatchChan records (stable across steps; refreshed after a cell sort
// has swapped buffers).  Launch geometry (n_tiles, rows, shared memory) must be uniform over the channels.
struct BatchChan {
    PhysParams P;
    StateView S;
    DevScalars *sc;
    double *shear_partials;
    GridFuse F;   // of the channel's next list build (used by the step's last continuity pass)
};
struct BatchLaunch {
    const BatchChan *chan;   // device array [n_chan]
    int n_chan;
    unsigned n_tiles;        // tiles of every channel
    int n_fluid;             // fluid rows of every channel
    int tile_cap;            // largest slot count over the channels (sizes the dynamic shared memory)
    int ncx, tile_w;         // cell columns and tile width of every channel
    cudaStream_t st;
};
int csr_batch_configure(int tile_cap);
int csr_batch_density(const BatchLaunch &L);
int csr_batch_kgc(const BatchLaunch &L);
int csr_batch_publish_dt(const BatchLaunch &L, double remain, double fixed_dt);
int csr_batch_force_fused(const BatchLaunch &L, bool add_gravity);
int csr_batch_time_step(const BatchLaunch &L, double remain, double fixed_dt);
int csr_batch_pressure(const BatchLaunch &L);
int csr_batch_continuity(const BatchLaunch &L, bool wrap, bool advance_clock, bool fuse_grid);
int csr_batch_wall_shear(const BatchLaunch &L);

}  // namespace sphb
