// ops_pairs.cuh -- launchers of the pair-parallel (COO) operator kernels used by the stateless MEX path.
// All pointers are DEVICE pointers to column-major arrays exactly as the MEX boundary lays them out.
#pragma once
#include "common.cuh"

namespace sphb {

struct PairView {
    const double *pi, *pj, *dx, *dy, *r, *W, *dW;   // W may be null for modes that do not take it
    size_t n;
};

// scratch doubles needed by each mode (per n_total), see ops_pairs.cu
constexpr int kScratchCols = 8;

int pairs_density(const PairView &pv, const double *mass, int nf, int nt, double rho0, double h, double inv_sigma0,
                  double *rho, double *Vol, double *B, double *scratch, cudaStream_t st);
int pairs_viscous(const PairView &pv, const double *vel, const double *Vol, const double *B, double mu, double h,
                  int nf, int nt, const double *wall_vel, double *force, double *scratch, cudaStream_t st);
int pairs_add_gravity(double *force_prior, const double *mass, double g, int nf, cudaStream_t st);
int pairs_transport(const PairView &pv, const double *Vol, const double *B, const double *pos, double h, int nf,
                    int nt, double coeff, double *pos_out, double *scratch, cudaStream_t st);
int pairs_integration_1st(const PairView &pv, const double *Vol, const double *B, const double *rho,
                          const double *mass, const double *pos, const double *vel, const double *drho_dt,
                          const double *force_prior, double dt, int nf, int nt, double rho0, double p0, double c_f,
                          double *rho_out, double *p_out, double *pos_out, double *force_out, double *drho_out,
                          double *scratch, cudaStream_t st);
int pairs_integration_2nd(const PairView &pv, const double *Vol, const double *rho, const double *pos,
                          const double *vel, double dt, int nf, int nt, const double *wall_vel, double *pos_out,
                          double *drho_out, double *scratch, cudaStream_t st);
// v_new = v + (F_prior + F_p)/m * dt on fluid rows, 0 on wall rows (ph:1400-1408)
int pairs_verlet_kick(const double *vel, const double *force_prior, const double *force, const double *mass,
                      double dt, int nf, int nt, double *vel_out, cudaStream_t st);
// rho_new = rho_half + 0.5 dt drho (clamp), p = EOS; wall rho kept, wall p = 0 (ph:1440-1450)
int pairs_verlet_finish(const double *rho_half, const double *drho, double dt, int nf, int nt, double rho0,
                        double p0, double *rho_out, double *p_out, cudaStream_t st);
int pairs_wall_shear(const PairView &pv, const double *pos, const double *vel, const double *wall_vel,
                     const double *Vol, const double *B, int nf, int nt, double DL, double DH, double mu, double h,
                     double *tau2_dev, double *scratch, cudaStream_t st);

}  // namespace sphb
