/*
 * sph_b200.h -- C ABI of libsph_b200, the B200-native (sm_100a) replacement for the particle hot
 * path of KIYOYOZU/SPH-Poiseuille-Flow.  Plain C, plain pointers and sizes, no torch / CUDA types.
 *
 * What it replaces (reference file:line):
 *   - the gateway of mex/sph_neighbor_search_mex.c:185-421          -> sphb_neighbor_search*
 *   - the eight string-dispatched modes of mex/sph_physics_mex.c:1745-1772
 *       density_correction  :95-374    -> sphb_density_correction
 *       viscous_force       :396-550   -> sphb_viscous_force
 *       transport_correction:569-714   -> sphb_transport_correction
 *       integration_1st     :736-967   -> sphb_integration_1st
 *       integration_2nd     :987-1119  -> sphb_integration_2nd
 *       integration_verlet  :1316-1469 -> sphb_integration_verlet
 *       advance_shell_step  :1490-1639 -> sphb_advance_shell_step
 *       wall_shear_monitor  :1653-1743 -> sphb_wall_shear_monitor
 *   - the main-loop body of SPH_Poiseuille.m:250-283 (+ :519-577)   -> sphb_ctx_* (device-resident)
 *
 * Two families of entry points:
 *   1. STATELESS ("MEX path").  Host pointers in, host pointers out, nothing persists: exactly the
 *      contract of the reference's MEX files.  Array arguments are passed as
 *      sphb_mat views (pointer + MATLAB dims) so that every size check the reference performs -- and
 *      the MATLAB error identifier it raises -- is reproduced inside the library; the MEX shims in
 *      sph-poiseuille-flow_b200/mex/ only unpack prhs[] and pack plhs[].
 *   2. DEVICE-RESIDENT context.  State lives in HBM across steps; the neighbour list is a CSR built
 *      on the device; one call advances whole reference steps (or dual-criteria outer/inner steps).
 *      This is what the standalone driver and bench.py run.
 *
 * Error model: every function returns 0 on success, non-zero on failure, never exits/aborts.  After a
 * failure sphb_last_error_id() holds the MATLAB identifier the reference would have raised
 * ("SPH:Neighbor:count", "SPH:Physics:viscous:vel", ...) or "SPHB:*" for
 * conditions the reference cannot have (CUDA failures, capacity), and sphb_last_error_msg() a text.
 * Both are thread-local.  There is NO CPU fallback: without a usable CUDA device every compute entry
 * point fails with SPHB:cuda.
 *
 * Conventions: matrices are column-major [m x n] doubles as MATLAB stores them; particle indices in
 * pair lists are 1-based doubles; fluid particles are rows 0..n_fluid-1, wall particles follow.
 */
#ifndef SPH_B200_H
#define SPH_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SPHB_API __attribute__((visibility("default")))
#else
#define SPHB_API
#endif

#define SPHB_VERSION 100

/* A borrowed, read-only view of a MATLAB double matrix (mxGetDoubles / mxGetM / mxGetN). */
typedef struct sphb_mat {
    const double *p;
    size_t m;
    size_t n;
} sphb_mat;

/* ----------------------------------------- errors ----------------------------------------- */
SPHB_API const char *sphb_last_error_id(void);
SPHB_API const char *sphb_last_error_msg(void);
SPHB_API int sphb_version(void);
/* 0 when a CUDA device is usable, else non-zero (and the last error says why). */
SPHB_API int sphb_device_check(void);
/* Measured DFMA throughput of the current device in TFLOP/s (2 flops per DFMA; eight independent chains per thread,
 * 8 warps per SM sub-partition), best of `reps` launches of iters x 32 DFMAs per thread, CUDA events.  The fp64
 * roofline denominator of bench.py; the reference has no counterpart. */
SPHB_API int sphb_fp64_peak(int iters, int reps, double *tflops);

/* ============================== 1. stateless MEX-path entry points ============================== */

/*
 * sphb_validate_only(1) makes the operator entry points of this section return right after the reference's argument
 * checks (same order, same identifiers), before any device work and before they touch their output pointers;
 * sphb_validate_only(0) restores normal operation.  Per calling thread.  The MEX shims use it to run the checks of
 * ph:125-162 etc. BEFORE they allocate outputs from a caller-supplied n_total, as the reference does.
 */
SPHB_API void sphb_validate_only(int on);

/*
 * Neighbour search, replaces mexFunction of mex/sph_neighbor_search_mex.c:185-421.
 * Call sphb_neighbor_search() -> *n_pairs, allocate 7 arrays of *n_pairs doubles (the MEX shim uses
 * mxCreateDoubleMatrix), then sphb_neighbor_fetch() copies device results straight into them and
 * releases the handle.  sphb_neighbor_release() drops a handle without fetching.
 * Output order and values are those of the reference (i ascending, its 3x3 walk order, its
 * real/ghost-entry arithmetic): pair_i/pair_j 1-based, dx = x_i - x_j minimum-image.
 * Errors: SPH:Neighbor:pos | count | param | entryCapacity (nb:223-242, 94).
 */
typedef struct sphb_pairs sphb_pairs;
SPHB_API int sphb_neighbor_search(sphb_mat pos, int n_fluid, int n_total, double h, double DL,
                                  sphb_pairs **handle, size_t *n_pairs);
SPHB_API int sphb_neighbor_fetch(sphb_pairs *handle, double *pair_i, double *pair_j, double *dx, double *dy,
                                 double *r, double *W, double *dW);
SPHB_API void sphb_neighbor_release(sphb_pairs *handle);

/* density_correction (ph:95-374).  Out: rho[n_total], Vol[n_total], B[n_total x 4]. */
SPHB_API int sphb_density_correction(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                     sphb_mat W, sphb_mat dW, sphb_mat mass, int n_fluid, int n_total,
                                     double rho0, double h, double inv_sigma0, double *rho, double *Vol,
                                     double *B);

/* viscous_force (ph:396-550).  Out: force[n_total x 2].  (mass is size-checked, unused: ph:547.) */
SPHB_API int sphb_viscous_force(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                sphb_mat dW, sphb_mat vel, sphb_mat Vol, sphb_mat B, double mu, double h,
                                int n_fluid, int n_total, sphb_mat mass, sphb_mat wall_vel, double *force);

/* transport_correction (ph:569-714).  has_coeff = 0 selects the MEX default 0.2 (ph:584). */
SPHB_API int sphb_transport_correction(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                       sphb_mat dW, sphb_mat Vol, sphb_mat B, sphb_mat pos, double h,
                                       int n_fluid, int n_total, int has_coeff, double transport_coeff,
                                       double *pos_out);

/* integration_1st (ph:736-967).  Out: rho, p [n_total]; pos, force [n_total x 2]; drho_dt [n_total]. */
SPHB_API int sphb_integration_1st(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                  sphb_mat dW, sphb_mat Vol, sphb_mat B, sphb_mat rho, sphb_mat mass,
                                  sphb_mat pos, sphb_mat vel, sphb_mat drho_dt, sphb_mat force_prior, double dt,
                                  int n_fluid, int n_total, double rho0, double p0, double c_f,
                                  sphb_mat wall_vel, double *rho_out, double *p_out, double *pos_out,
                                  double *force_out, double *drho_out);

/* integration_2nd (ph:987-1119).  Out: pos [n_total x 2], drho_dt [n_total], zeros [n_total x 2]. */
SPHB_API int sphb_integration_2nd(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                  sphb_mat dW, sphb_mat Vol, sphb_mat rho, sphb_mat pos, sphb_mat vel, double dt,
                                  int n_fluid, int n_total, sphb_mat wall_vel, double *pos_out,
                                  double *drho_out, double *zeros_out);

/* integration_verlet (ph:1316-1469).  Out: rho, p, pos, vel, drho_dt, force. */
SPHB_API int sphb_integration_verlet(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                     sphb_mat dW, sphb_mat Vol, sphb_mat B, sphb_mat rho, sphb_mat mass,
                                     sphb_mat pos, sphb_mat vel, sphb_mat drho_dt, sphb_mat force_prior,
                                     double dt, int n_fluid, int n_total, double rho0, double p0, double c_f,
                                     sphb_mat wall_vel, double *rho_out, double *p_out, double *pos_out,
                                     double *vel_out, double *drho_out, double *force_out);

/* advance_shell_step (ph:1490-1639).  Out: rho, p, pos, vel, drho_dt, force, force_prior, Vol, B. */
SPHB_API int sphb_advance_shell_step(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                     sphb_mat W, sphb_mat dW, sphb_mat mass, sphb_mat pos, sphb_mat vel,
                                     sphb_mat wall_vel, sphb_mat rho, sphb_mat drho_dt, double dt, int n_fluid,
                                     int n_total, double rho0, double p0, double c_f, double mu, double h,
                                     double inv_sigma0, double gravity_g, double *rho_out, double *p_out,
                                     double *pos_out, double *vel_out, double *drho_out, double *force_out,
                                     double *force_prior_out, double *Vol_out, double *B_out);

/* wall_shear_monitor (ph:1653-1743).  n_total is inferred from Vol (ph:1692).  Out: two scalars. */
SPHB_API int sphb_wall_shear_monitor(sphb_mat pair_i, sphb_mat pair_j, sphb_mat dx, sphb_mat dy, sphb_mat r,
                                     sphb_mat dW, sphb_mat pos, sphb_mat vel, sphb_mat wall_vel, sphb_mat Vol,
                                     sphb_mat B, int n_fluid, double DL, double DH, double mu, double h,
                                     double *tau_bottom, double *tau_top);

/* ================================ 2. device-resident context ================================ */

typedef struct sphb_ctx sphb_ctx;

/* Physical / numerical parameters of one channel (SPH_Poiseuille.m:47-80 after derivation). */
typedef struct sphb_params {
    int n_fluid;
    int n_total;
    double DL, DH, dp, h;
    double rho0, p0, c_f, mu, gravity_g, inv_sigma0;
    double transport_coeff;   /* 0.30 in the script (m:77) */
    int sort_interval;        /* cell sort every sort_interval steps, not at step 1 (m:272); <=0: never */
    int device;               /* CUDA device ordinal; -1 = current device */
    int max_neighbors_hint;   /* 0 = default; sizes the CSR when the exact count is not known yet */
} sphb_params;

SPHB_API int sphb_ctx_create(const sphb_params *prm, sphb_ctx **ctx);
SPHB_API void sphb_ctx_destroy(sphb_ctx *ctx);

/* Host (pageable or pinned) column-major arrays; any pointer may be NULL = leave untouched.  Uploading `mass` sets
 * Vol = mass / rho (the initial condition, m:120); without it the volumes of the last density pass are kept. */
SPHB_API int sphb_ctx_upload(sphb_ctx *ctx, const double *pos, const double *vel, const double *rho,
                             const double *drho_dt, const double *mass, const double *wall_vel);
SPHB_API int sphb_ctx_download(sphb_ctx *ctx, double *pos, double *vel, double *rho, double *p, double *drho_dt,
                               double *force, double *force_prior, double *Vol, double *B);
/* Override/inspect the clock (restart). */
SPHB_API int sphb_ctx_set_time(sphb_ctx *ctx, double t, long long step);

/* Neighbour list (N1-N6): counting sort by 2h cell, ghost entries, dedup, CSR.  */
SPHB_API int sphb_ctx_neighbor_build(sphb_ctx *ctx);
SPHB_API int sphb_ctx_neighbor_count(sphb_ctx *ctx, size_t *n_half_pairs, size_t *n_csr_entries);
/* The reference's half-pair COO list (same order/values as sphb_neighbor_search), 7 x n_half_pairs. */
SPHB_API int sphb_ctx_neighbor_export_pairs(sphb_ctx *ctx, double *pair_i, double *pair_j, double *dx,
                                            double *dy, double *r, double *W, double *dW);
/* The full (gather) CSR: row_ptr[n_fluid+1], col[n_csr_entries] (0-based j, ascending per row).  It is the symmetric
 * closure of the reference's half list with ONE documented exception: a fluid-fluid pair that one end reaches only
 * through a rounding accident at the periodic seam (a ghost entry in an unusual column, or x_w exactly on a
 * ghost-insertion threshold, nb:282-294) can be missing from the OTHER end's row.  Such pairs lie within rounding of
 * r = 2h (measured: |r/2h - 1| < 1e-12 on thousands of adversarial clouds), where W ~ (2 - q)^3 and dW ~ (2 - q)^2
 * vanish, so no sum of any operator changes; there is never an entry the reference lacks, and the half list
 * (sphb_ctx_neighbor_export_pairs, sphb_neighbor_search) is exact.  DESIGN.md section 4. */
SPHB_API int sphb_ctx_neighbor_export_csr(sphb_ctx *ctx, int *row_ptr, int *col);

/* Single operators on the resident state (for operator-level parity tests; P1-P8). */
SPHB_API int sphb_ctx_density_correction(sphb_ctx *ctx);              /* -> rho, Vol, B */
SPHB_API int sphb_ctx_eos(sphb_ctx *ctx);                             /* p = p0 (rho/rho0 - 1), wall 0 (m:172-173) */
SPHB_API int sphb_ctx_viscous_force(sphb_ctx *ctx, int add_gravity);  /* -> force_prior (+ m*g on x) */
SPHB_API int sphb_ctx_transport_correction(sphb_ctx *ctx);            /* pos += shift */
SPHB_API int sphb_ctx_integration_verlet(sphb_ctx *ctx, double dt);   /* dt <= 0: use the device dt */
SPHB_API int sphb_ctx_wall_shear(sphb_ctx *ctx, double *tau_bottom, double *tau_top);
SPHB_API int sphb_ctx_periodic_wrap(sphb_ctx *ctx);
SPHB_API int sphb_ctx_sort_by_cell(sphb_ctx *ctx);                    /* m:529-568, unconditional */

/* Status of the resident run, read back from the device (one small D2H). */
typedef struct sphb_status {
    double t;
    double dt;            /* last step's dt */
    double vmax;          /* max |v| over fluid rows after the last step */
    double tau_bottom, tau_top;
    long long step;
    long long n_half_pairs;
    long long n_csr_entries;
    int error_flags;      /* device-side flags: 1 list rows, 2 cell capacity, 4 dedup capacity, 8 non-finite, 16 tile slots,
                             32 slab rows */
} sphb_status;
SPHB_API int sphb_ctx_status(sphb_ctx *ctx, sphb_status *out);

/*
 * One pass of the reference main-loop body (SPH_Poiseuille.m:251-283): density+KGC, viscous(+g),
 * transport, dt = min(acoustic, viscous, body, remain), Verlet, wrap, wall velocity zeroing,
 * cell sort when due, neighbour rebuild, wall-shear monitor.  `remain` as in m:252 (use a huge value
 * for "no output point pending").  Requires a neighbour list (sphb_ctx_neighbor_build) and is
 * asynchronous: call sphb_ctx_status()/download to synchronise.
 */
SPHB_API int sphb_ctx_step(sphb_ctx *ctx, double remain);
/*
 * One reference step for a caller that keeps the state on the host between steps (sphb_ctx_upload -> this ->
 * sphb_ctx_download; what the MEX drop-in does per pass of m:251-283): list build for the uploaded positions
 * (skipped when the context still holds a valid list), wall-shear monitor, then the physics m:254-270.  The list is
 * not rebuilt for the new positions (the next upload would discard it); asynchronous like sphb_ctx_step.
 */
SPHB_API int sphb_ctx_step_uploaded(sphb_ctx *ctx, double remain);
/* The same step as ONE call on host arrays (column-major n_total-row arrays as in sphb_ctx_upload / _download; the
 * inputs are required, any output may be NULL): upload -> list -> shear monitor -> physics -> download with the
 * transfers overlapped where the data flow allows (the list build runs while rho / drho_dt are on the wire).  Same
 * results, bit for bit, as sphb_ctx_upload + sphb_ctx_step_uploaded + sphb_ctx_download.  Returns when the outputs
 * are in place. */
SPHB_API int sphb_ctx_step_host(sphb_ctx *ctx, double remain, const double *pos, const double *vel, const double *rho,
                                const double *drho_dt, double *pos_out, double *vel_out, double *rho_out, double *p_out,
                                double *drho_dt_out, double *force_out);
/* Reference schedule until t >= t_target - 1e-12 (m:250) or max_steps; returns steps done. */
SPHB_API int sphb_ctx_run_until(sphb_ctx *ctx, double t_target, double t_end, long long max_steps,
                                long long *steps_done);
/*
 * Dual-criteria schedule (README.md:61-65; no reference code at HEAD -- "parity unpinned"): n_outer
 * outer steps, each = density+KGC, viscous(+g), transport on the current list, then n_inner acoustic
 * Verlet substeps (dt = min(acoustic, viscous, body)) with immediate wrap, then [sort], list rebuild,
 * shear monitor.  n_inner = 1 reproduces sphb_ctx_step with remain = inf.  Runs as a CUDA graph.
 */
SPHB_API int sphb_ctx_run_dual(sphb_ctx *ctx, int n_outer, int n_inner);

/*
 * Dual-criteria schedule with both criteria evaluated on the device (BASELINE.json north_star (3); scheme of the
 * reference's README.md:61-65 -- no code at HEAD, so this schedule is not pinned by the reference, only its operators
 * are).  Per outer step:  Dt = min(0.25 h / max|v|, 0.125 h^2 / nu, 0.25 sqrt(h / |g|), remain);  density + KGC,
 * viscous force (+ m g) and transport shift once;  acoustic substeps dt = min(0.25 h / (c_f + max|v|), Dt_left) until
 * Dt is covered -- at most max_inner of them: the launch sequence has a fixed length and the substeps beyond
 * ceil(Dt / dt) return at once on the device;  [cell sort], list rebuild, wall-shear monitor.  max|v|, Dt, dt and the
 * substep count never leave the device; one CUDA graph per outer step when remain is infinite.  Returns after the
 * last outer step has finished (it reads the step counter back).  Not available on slabs of a sharded channel.
 */
SPHB_API int sphb_ctx_run_dual_auto(sphb_ctx *ctx, int n_outer, int max_inner, double remain);
/* outer steps / acoustic substeps taken so far by sphb_ctx_run_dual_auto, and the Dt of the last outer step */
SPHB_API int sphb_ctx_dual_counts(sphb_ctx *ctx, long long *outer_steps, long long *sub_steps, double *last_Dt);

/*
 * Ensembles of independent channels (swept mu / g, ...): one context per channel, each with its own stream and
 * cached graph; `n_outer` outer steps of every channel are enqueued round-robin from one host loop so that the
 * small kernels of many channels overlap on the device.  Pure data parallelism: no communication.
 */
SPHB_API int sphb_ensemble_run_dual(sphb_ctx **ctxs, int n_ctx, int n_outer, int n_inner);
SPHB_API int sphb_ensemble_sync(sphb_ctx **ctxs, int n_ctx);

/*
 * The same ensembles stepped by BATCHED launches: every kernel of the step (m:251-283) is launched once for
 * all channels (grid = the channel's own grid x channels; the per-channel pointers, mu / g / ... and device
 * scalars come from records in device memory), so the cost of a step no longer grows with the number of kernel
 * launches per channel.  Every channel advances with its own dt; results are those of sphb_ctx_run_dual on each
 * channel (same kernel bodies, same per-channel data).  Requirements, checked at creation (SPHB:batch
 * otherwise): one device; equal n_fluid, n_total, h, DL, sort_interval and cell / tile geometry; every channel
 * initialised (first neighbour build done) and at the same step count.  While a context is a member of a batch
 * it launches on the batch's stream: upload / download / status / profile keep working; after anything that moves
 * its buffers or changes its list capacity (position upload, sphb_ctx_sort_by_cell, sphb_ctx_run_until) call
 * sphb_batch_refresh.  Destroy the batch before its member contexts.
 */
typedef struct sphb_batch sphb_batch;
SPHB_API int sphb_batch_create(sphb_ctx **ctxs, int n_ctx, sphb_batch **out);
SPHB_API void sphb_batch_destroy(sphb_batch *batch);
SPHB_API int sphb_batch_refresh(sphb_batch *batch);
SPHB_API int sphb_batch_run_dual(sphb_batch *batch, int n_outer, int n_inner);
SPHB_API int sphb_batch_sync(sphb_batch *batch);
/* kernels launched by the batch so far (graph replays counted by their captured kernels) */
SPHB_API long long sphb_batch_launches(const sphb_batch *batch);
/* OR of the members' sphb_status.error_flags (synchronises) */
SPHB_API int sphb_batch_error_flags(sphb_batch *batch, int *flags_or);

/* x-/bin-averaged u(y) of the fluid rows (m:579-605).  x0 < 0: all fluid rows (full-length profile). */
SPHB_API int sphb_ctx_profile(sphb_ctx *ctx, double x0, double half_width, int n_bins, double *u_mean,
                              double *count);

/* Device timing of the phases of the last sphb_ctx_step / run (ms, CUDA events). */
typedef struct sphb_timing {
    float neighbor_ms, density_ms, visc_transport_ms, verlet_ms, shear_ms, sort_ms, total_ms;
    long long launches;   /* kernels launched by this context since creation */
} sphb_timing;
SPHB_API int sphb_ctx_timing(sphb_ctx *ctx, sphb_timing *out);
SPHB_API int sphb_ctx_enable_timing(sphb_ctx *ctx, int on);
/*
 * Per-kernel device time of the resident path, measured with CUDA events recorded on the context's
 * stream around every kernel (group) while enabled.  While enabled, sphb_ctx_run_dual launches its
 * kernels eagerly instead of replaying a CUDA graph.  enable: 1 on, 0 off, -1 unchanged; reset != 0
 * zeroes the accumulators after reporting.  `calls` counts timed scopes, not individual launches.
 */
#define SPHB_KPROF_SLOTS 24
typedef struct sphb_kprof {
    int n;                               /* slots in use */
    const char *name[SPHB_KPROF_SLOTS];  /* static strings: "pressure", "continuity", "nbr_walk_fill", ... */
    double ms[SPHB_KPROF_SLOTS];         /* accumulated device time */
    long long calls[SPHB_KPROF_SLOTS];
} sphb_kprof;
SPHB_API int sphb_ctx_kernel_profile(sphb_ctx *ctx, int enable, int reset, sphb_kprof *out);
/* The CUDA stream the context launches on (as a void* cudaStream_t), for event timing by callers. */
SPHB_API void *sphb_ctx_stream(sphb_ctx *ctx);
/* Launch on a stream of the caller (cudaStream_t as void*); NULL = back to the context's own stream. */
SPHB_API int sphb_ctx_set_stream(sphb_ctx *ctx, void *stream);
SPHB_API int sphb_ctx_sync(sphb_ctx *ctx);

/* ================== 3. x-slab sharding of one long periodic channel ==================
 * The reference has no domain decomposition; this is the multi-GPU form of its main loop (m:250-283),
 * checked against the single-domain run.  One context per rank owns the fluid rows whose wrapped x lies in
 * the cell columns [col_lo, col_hi) (cell = 2h, nb:253) plus `halo_cols` columns of copies on either
 * side; halo_cols = 4 lets a rank compute a whole reference step without a mid-step exchange (list ->
 * density -> KGC -> forces -> continuity each consume one column of validity).  Per step:
 *     sphb_ctx_slab_step_begin   density + KGC, forces, continuity, wrap            (m:254-270)
 *     sphb_ctx_slab_scalars_export -> caller all-reduces max|v| (MAX) and the shear partials (SUM)
 *     sphb_ctx_slab_pack         migrants + edge copies into two device buffers of the caller
 *     (caller: NCCL send/recv of the two buffers to the left / right neighbour rank)
 *     sphb_ctx_slab_compact      the rows that stay, compacted (runs while the buffers are on the wire)
 *     sphb_ctx_slab_unpack       new owned rows, new halo copies (shifted by -+DL across the seam, nb:284,291)
 *     sphb_ctx_slab_vmax_import  the all-reduced max|v| (time step of the next step, m:519-527)
 *     sphb_ctx_slab_step_end     list rebuild + wall-shear partials                 (m:280-283)
 * Buffers: 16 header doubles + rec_cap records of 16 doubles, DEVICE pointers owned by the caller.
 * The context is created by sphb_ctx_create with n_fluid = fluid-row CAPACITY (owned + copies, with
 * head-room) and n_total = capacity + n_wall; rows are [owned fluid | halo fluid | wall].
 */
SPHB_API int sphb_ctx_slab_init(sphb_ctx *ctx, int col_lo, int col_hi, int halo_cols, int n_owned, const double *pos,
                                const double *vel, const double *rho, const double *drho_dt, const double *mass,
                                const int *ids, int n_wall, const double *wall_pos, const double *wall_mass,
                                const double *wall_wall_vel);
/* global identity (the `ids` given at init, carried through migrations) of the fluid rows, owned first */
SPHB_API int sphb_ctx_slab_ids(sphb_ctx *ctx, int *ids);
SPHB_API int sphb_ctx_slab_pack(sphb_ctx *ctx, double *send_left_dev, double *send_right_dev, int rec_cap);
/* second half of the pack: the stayers are compacted (overlaps the transfer when called after the send) */
SPHB_API int sphb_ctx_slab_compact(sphb_ctx *ctx);
SPHB_API int sphb_ctx_slab_unpack(sphb_ctx *ctx, const double *recv_left_dev, const double *recv_right_dev,
                                  const double *sent_left_dev, const double *sent_right_dev, int rec_cap);
SPHB_API int sphb_ctx_slab_step_begin(sphb_ctx *ctx, double remain);
SPHB_API int sphb_ctx_slab_step_end(sphb_ctx *ctx);
SPHB_API int sphb_ctx_slab_scalars_export(sphb_ctx *ctx, double *out_dev3);   /* vmax, tau_bottom, tau_top partials */
SPHB_API int sphb_ctx_slab_vmax_import(sphb_ctx *ctx, const double *in_dev1);
/* the scalars of all ranks after an all-gather ([world][3], device): max|v| (MAX) into the context; the
 * reduced (max|v|, sum tau_bottom, sum tau_top) to out_dev3 (may be NULL) */
SPHB_API int sphb_ctx_slab_scalars_reduce(sphb_ctx *ctx, const double *gathered_dev, int world, double *out_dev3);
/* same with the triple of rank r at gathered_dev + offsets_dev[r] (in doubles; device array of `world` int64) */
SPHB_API int sphb_ctx_slab_scalars_reduce_at(sphb_ctx *ctx, const double *gathered_dev, const long long *offsets_dev,
                                             int world, double *out_dev3);
SPHB_API int sphb_ctx_slab_counts(sphb_ctx *ctx, int *n_owned, int *n_halo, int *n_wall);

/*
 * Peer-to-peer exchange ("mailboxes"): the B200-native form of pack -> all-to-all -> reduce -> unpack.  Every rank
 * owns a mailbox in its device memory; the pack kernel of a neighbour stores its records straight into it over
 * NVLink (peer pointer) and publishes a sequence number behind a system-wide fence; the unpack kernel waits for that
 * number.  max|v| and the wall-shear partials travel the same way to every rank.  No collective and no host
 * synchronisation: a sharded step is one stream of this library's kernels and can be captured as ONE CUDA graph.
 *   sphb_ctx_slab_p2p_create   allocates the mailbox (rec_cap records per neighbour), returns its device pointer and a
 *                              64-byte CUDA IPC handle for ranks in other processes
 *   sphb_ctx_slab_p2p_connect  makes rank r's mailbox known: its IPC handle (other process) or its device pointer
 *                              (same process; peer access is enabled when peer_device differs from the context's)
 *   sphb_ctx_slab_p2p_send     scalars to everybody + pack into the neighbours' mailboxes (and the local send buffers)
 *   sphb_ctx_slab_p2p_recv     waits for the scalars (imports max|v|) and the neighbours' records, appends them
 *                              (= sphb_ctx_slab_unpack); sphb_ctx_slab_compact may run between send and recv
 *   sphb_ctx_slab_p2p_monitors max|v|, tau_bottom, tau_top of the whole channel as of the last exchange
 * A peer that never arrives raises error flag 32 after ~2 s instead of hanging the device.
 */
SPHB_API int sphb_ctx_slab_p2p_create(sphb_ctx *ctx, int rank, int world, int rec_cap, void **mailbox_dev, size_t *bytes,
                                      unsigned char ipc_handle[64]);
SPHB_API int sphb_ctx_slab_p2p_connect(sphb_ctx *ctx, int rank, const unsigned char ipc_handle[64], void *same_process_ptr,
                                       int peer_device);
SPHB_API int sphb_ctx_slab_p2p_send(sphb_ctx *ctx, double *send_left_dev, double *send_right_dev);
SPHB_API int sphb_ctx_slab_p2p_recv(sphb_ctx *ctx, const double *sent_left_dev, const double *sent_right_dev);
SPHB_API int sphb_ctx_slab_p2p_monitors(sphb_ctx *ctx, double *out3);
/* pos, vel, rho, drho_dt of the rows [0, n_rows) of the owned region to (to_host != 0) or from host arrays (column-major,
 * ideally pinned); asynchronous on the context's stream.  What a host-resident caller carries from step to step. */
SPHB_API int sphb_ctx_slab_rows_io(sphb_ctx *ctx, int to_host, int n_rows, double *pos, double *vel, double *rho,
                                   double *drho_dt);

/* ================== 3b. the sharded channel driven from C: all slabs of a node in one process ==================
 * The C host of section 3 (csrc/slab_group.cu): cuts the channel into n_slabs x-slabs on cell-column boundaries,
 * creates one context per slab on devices[r % n_devices] (NULL / 0: all visible devices), connects their peer-to-peer
 * mailboxes (plain device pointers + peer access) and steps them from one host thread.  With one slab per device a
 * step is one CUDA graph launch per GPU; slabs that share a device are stepped eagerly (tests, small boxes).  No
 * Python, torch or NCCL involved -- what `sph_poiseuille_b200 --slabs N` runs (driver/sph_poiseuille_main.c).
 * Arrays are the column-major n_total-row host arrays of sphb_ctx_upload; prm describes the WHOLE channel
 * (prm->device and prm->sort_interval are ignored: the slabs' now-and-then compaction is their cell sort).
 *   sphb_slabs_step        one reference step m:251-283 of the whole channel (remain as in sphb_ctx_step)
 *   sphb_slabs_run         n steps with remain = inf, no host synchronisation
 *   sphb_slabs_run_until   the loop m:250-283 up to t_target, as sphb_ctx_run_until
 *   sphb_slabs_status      clock of the channel, max|v| / wall shear reduced over the slabs (slab order), flags OR-ed
 *   sphb_slabs_download    rows in the GLOBAL order of the arrays given at creation (fluid rows carry their identity
 *                          through migrations), static wall rows behind them
 *   sphb_slabs_member      context of slab r (monitors, per-kernel profile, tests); owned by the group
 */
typedef struct sphb_slabs sphb_slabs;
SPHB_API int sphb_slabs_create(const sphb_params *prm, int n_slabs, const int *devices, int n_devices, const double *pos,
                               const double *vel, const double *rho, const double *drho_dt, const double *mass,
                               const double *wall_vel, sphb_slabs **out);
SPHB_API void sphb_slabs_destroy(sphb_slabs *group);
SPHB_API int sphb_slabs_count(const sphb_slabs *group);
SPHB_API int sphb_slabs_set_compaction(sphb_slabs *group, int every_n_steps);   /* default 256; 0 = never */
SPHB_API sphb_ctx *sphb_slabs_member(sphb_slabs *group, int r, int *col_lo, int *col_hi, int *device);
SPHB_API int sphb_slabs_step(sphb_slabs *group, double remain);
SPHB_API int sphb_slabs_run(sphb_slabs *group, long long n_steps);
SPHB_API int sphb_slabs_run_until(sphb_slabs *group, double t_target, double t_end, long long max_steps,
                                  long long *steps_done);
SPHB_API int sphb_slabs_sync(sphb_slabs *group);
SPHB_API int sphb_slabs_status(sphb_slabs *group, sphb_status *out);
SPHB_API int sphb_slabs_download(sphb_slabs *group, double *pos, double *vel, double *rho, double *p, double *drho_dt);
/* binned u(y) of the whole channel (sphb_ctx_profile's arguments; gathers the rows: an output-point operation) */
SPHB_API int sphb_slabs_profile(sphb_slabs *group, double x0, double half_width, int n_bins, double *u_mean, double *count);
SPHB_API int sphb_slabs_set_time(sphb_slabs *group, double t, long long step);

#ifdef __cplusplus
}
#endif
#endif /* SPH_B200_H */
