/*
 * oracle/sph_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C, serial, deterministic restatement of the reference's SPH particle loop
 * (mex/sph_neighbor_search_mex.c, mex/sph_physics_mex.c, and the step order of SPH_Poiseuille.m).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may link or call it.
 * It is pinned against oracle/_ref/libsph_ref.so (the reference sources compiled unmodified):
 * tests/test_oracle_pin.py demands BIT-IDENTICAL outputs at OMP_NUM_THREADS=1.
 *
 * Conventions (same as the MEX boundary): matrices are column-major [n x k]; pair indices are
 * 1-based and stored as doubles; fluid particles are rows 0..n_fluid-1, wall rows follow.
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    size_t n_pairs;
    double *pair_i, *pair_j, *dx, *dy, *r, *W, *dW; /* each n_pairs long, malloc'ed */
} orc_pairs;

/* returns 0, or a negative code on invalid arguments (same conditions as SPH:Neighbor:count/param) */
int orc_neighbor_search(const double *pos, int n_fluid, int n_total, double h, double DL, orc_pairs *out);
void orc_pairs_free(orc_pairs *p);

void orc_density_correction(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                            const double *dy, const double *r, const double *W, const double *dW,
                            const double *mass, int n_fluid, int n_total, double rho0, double h,
                            double inv_sigma0, double *rho, double *Vol, double *B);

void orc_viscous_force(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                       const double *dy, const double *r, const double *dW, const double *vel,
                       const double *Vol, const double *B, double mu, double h, int n_fluid, int n_total,
                       const double *wall_vel, double *force);

void orc_transport_correction(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                              const double *dy, const double *r, const double *dW, const double *Vol,
                              const double *B, const double *pos, double h, int n_fluid, int n_total,
                              double transport_coeff, double *pos_out);

void orc_integration_1st(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                         const double *dy, const double *r, const double *dW, const double *Vol,
                         const double *B, const double *rho, const double *mass, const double *pos,
                         const double *vel, const double *drho_dt, const double *force_prior, double dt,
                         int n_fluid, int n_total, double rho0, double p0, double c_f,
                         double *rho_out, double *p_out, double *pos_out, double *force_out, double *drho_out);

void orc_integration_2nd(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                         const double *dy, const double *r, const double *dW, const double *Vol,
                         const double *rho, const double *pos, const double *vel, double dt, int n_fluid,
                         int n_total, const double *wall_vel, double *pos_out, double *drho_out);

void orc_integration_verlet(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                            const double *dy, const double *r, const double *dW, const double *Vol,
                            const double *B, const double *rho, const double *mass, const double *pos,
                            const double *vel, const double *drho_dt, const double *force_prior, double dt,
                            int n_fluid, int n_total, double rho0, double p0, double c_f,
                            const double *wall_vel, double *rho_out, double *p_out, double *pos_out,
                            double *vel_out, double *drho_out, double *force_out);

void orc_advance_shell_step(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                            const double *dy, const double *r, const double *W, const double *dW,
                            const double *mass, const double *pos, const double *vel, const double *wall_vel,
                            const double *drho_dt, double dt, int n_fluid, int n_total, double rho0, double p0,
                            double c_f, double mu, double h, double inv_sigma0, double gravity_g,
                            double *rho_out, double *p_out, double *pos_out, double *vel_out, double *drho_out,
                            double *force_out, double *force_prior_out, double *Vol_out, double *B_out);

void orc_wall_shear_monitor(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                            const double *dy, const double *r, const double *dW, const double *pos,
                            const double *vel, const double *wall_vel, const double *Vol, const double *B,
                            int n_fluid, int n_total, double DL, double DH, double mu, double h,
                            double *tau_bottom, double *tau_top);

/* ---- the slice of SPH_Poiseuille.m that sits on the hot path ---- */
double orc_verlet_time_step(const double *vel, int n_fluid, int n_total, double c_max, double h, double nu,
                            double gravity_g, double remain);
void orc_periodic_wrap(double *pos, int n_fluid, int n_total, double DL);
/* perm[k] = old row that moves to new row k (fluid rows sorted by (cy,cx), stable; wall rows fixed) */
void orc_sort_permutation(const double *pos, int n_fluid, int n_total, double DL, double h, int *perm);

#ifdef __cplusplus
}
#endif
#endif
