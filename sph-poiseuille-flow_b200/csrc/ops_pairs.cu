// ops_pairs.cu -- pair-parallel fp64 operator kernels for the STATELESS MEX path (rows P1-P9).
//
// The MEX contract hands every mode a caller-supplied half-pair COO list (pair_i, pair_j, dx, dy, r, W,
// dW) that must be consumed as given, so these kernels keep the reference's scatter
// formulation: one thread per pair, fp64 `red.global.add.f64` into per-particle accumulators (the
// reference uses `#pragma omp atomic` the same way, e.g. ph:198-210), followed by a per-particle
// finalisation kernel.  Pairs with an out-of-range index or r <= 1e-12 are silently skipped exactly
// like the reference (ph:193, 246).  The device-resident path (ops_csr.cu) uses a gather over the CSR
// instead; this file is the drop-in-correctness path, not the throughput path.
#include "ops_pairs.cuh"
#include "physmath.cuh"

namespace sphb {

namespace {

constexpr double kPi = 3.14159265358979323846;

__device__ __forceinline__ bool pair_ids(const PairView &pv, size_t k, int nf, int nt, int &ii, int &jj)
{
    ii = (int)pv.pi[k] - 1;
    jj = (int)pv.pj[k] - 1;
    return !(ii < 0 || ii >= nf || jj < 0 || jj >= nt);
}

#define PAIR_LOOP(k) \
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < pv.n; k += (size_t)gridDim.x * blockDim.x)

inline unsigned pair_grid(size_t n)
{
    size_t b = (n + kBlock - 1) / kBlock;
    const size_t cap = 148u * 64u;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// ---------------------------------------------------------------- density_correction (ph:95-374)
__global__ void k_dens_init(double *s_in, double *s_ct, double *A, int nf, double W0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    s_in[i] = W0;
    s_ct[i] = 0.0;
    A[i] = 0.0;
    A[i + nf] = 0.0;
    A[i + 2 * (size_t)nf] = 0.0;
    A[i + 3 * (size_t)nf] = 0.0;
}

__global__ void __launch_bounds__(kBlock) k_dens_sigma(PairView pv, const double *__restrict__ mass, int nf, int nt,
                                                      double rho0, double *s_in, double *s_ct)
{
    PAIR_LOOP(k)
    {
        int ii, jj;
        if (!pair_ids(pv, k, nf, nt, ii, jj)) continue;
        const double w = pv.W[k];
        if (jj < nf) {
            atomicAdd(&s_in[ii], w);
            atomicAdd(&s_in[jj], w);
        } else {
            atomicAdd(&s_ct[ii], w * (mass[jj] / rho0));
        }
    }
}

__global__ void k_dens_rho(const double *s_in, const double *s_ct, const double *__restrict__ mass, int nf, int nt,
                           double rho0, double inv_sigma0, double *rho, double *Vol)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    double v = rho0;
    if (i < nf) {
        v = s_in[i] * rho0 * inv_sigma0;
        v += s_ct[i] * rho0 * rho0 * inv_sigma0 / mass[i];
        if (v <= 1e-12) v = rho0;
    }
    rho[i] = v;
    Vol[i] = mass[i] / v;
}

__global__ void __launch_bounds__(kBlock) k_dens_A(PairView pv, const double *__restrict__ Vol, int nf, int nt,
                                                  double *A)
{
    double *A11 = A, *A12 = A + nf, *A21 = A + 2 * (size_t)nf, *A22 = A + 3 * (size_t)nf;
    PAIR_LOOP(k)
    {
        int ii, jj;
        const double rk = pv.r[k];
        if (!pair_ids(pv, k, nf, nt, ii, jj) || rk <= 1e-12) continue;
        const double dx = pv.dx[k], dy = pv.dy[k], dW = pv.dW[k];
        const double ex = dx / rk, ey = dy / rk;
        const double fj = dW * Vol[jj];
        atomicAdd(&A11[ii], -(dx * (fj * ex)));
        atomicAdd(&A12[ii], -(dx * (fj * ey)));
        atomicAdd(&A21[ii], -(dy * (fj * ex)));
        atomicAdd(&A22[ii], -(dy * (fj * ey)));
        if (jj < nf) {
            const double fi = dW * Vol[ii];
            atomicAdd(&A11[jj], -(dx * (fi * ex)));
            atomicAdd(&A12[jj], -(dx * (fi * ey)));
            atomicAdd(&A21[jj], -(dy * (fi * ex)));
            atomicAdd(&A22[jj], -(dy * (fi * ey)));
        }
    }
}

__global__ void k_dens_B(const double *A, int nf, int nt, double *B)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    double b11 = 1.0, b12 = 0.0, b21 = 0.0, b22 = 1.0;
    if (i < nf) kgc_from_A(A[i], A[i + nf], A[i + 2 * (size_t)nf], A[i + 3 * (size_t)nf], b11, b12, b21, b22);
    B[i] = b11;
    B[i + nt] = b12;
    B[i + 2 * (size_t)nt] = b21;
    B[i + 3 * (size_t)nt] = b22;
}

// ---------------------------------------------------------------- viscous_force (ph:396-550)
__global__ void __launch_bounds__(kBlock) k_visc_pairs(PairView pv, const double *__restrict__ vel,
                                                      const double *__restrict__ Vol, const double *__restrict__ B,
                                                      double mu, double h, int nf, int nt,
                                                      const double *__restrict__ wall_vel, double *acc)
{
    const double *vx = vel, *vy = vel + nt, *wx = wall_vel, *wy = wall_vel + nt;
    const double *B11 = B, *B12 = B + nt, *B21 = B + 2 * (size_t)nt, *B22 = B + 3 * (size_t)nt;
    double *ax = acc, *ay = acc + nt;
    PAIR_LOOP(k)
    {
        int ii, jj;
        const double rk = pv.r[k];
        if (!pair_ids(pv, k, nf, nt, ii, jj) || rk <= 1e-12) continue;
        const double ex = pv.dx[k] / rk, ey = pv.dy[k] / rk, dW = pv.dW[k];
        const double denom = rk + 0.01 * h;
        if (jj < nf) {
            const double s11 = B11[ii] + B11[jj], s12 = B12[ii] + B12[jj];
            const double s21 = B21[ii] + B21[jj], s22 = B22[ii] + B22[jj];
            const double eBe = ex * (s11 * ex + s12 * ey) + ey * (s21 * ex + s22 * ey);
            const double dvx = vx[ii] - vx[jj], dvy = vy[ii] - vy[jj];
            const double ci = eBe * mu * dW * Vol[jj] / denom;
            const double cj = eBe * mu * dW * Vol[ii] / denom;
            atomicAdd(&ax[ii], ci * dvx);
            atomicAdd(&ay[ii], ci * dvy);
            atomicAdd(&ax[jj], -(cj * dvx));
            atomicAdd(&ay[jj], -(cj * dvy));
        } else {
            const double eBe = ex * (B11[ii] * ex + B12[ii] * ey) + ey * (B21[ii] * ex + B22[ii] * ey);
            const double dvx = vx[ii] - wx[jj], dvy = vy[ii] - wy[jj];
            const double c = 4.0 * eBe * mu * dW * Vol[jj] / denom;
            atomicAdd(&ax[ii], c * dvx);
            atomicAdd(&ay[ii], c * dvy);
        }
    }
}

__global__ void k_visc_fin(const double *acc, const double *__restrict__ Vol, int nf, int nt, double *force)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const bool fl = i < nf;
    force[i] = fl ? acc[i] * Vol[i] : 0.0;
    force[i + nt] = fl ? acc[i + nt] * Vol[i] : 0.0;
}

__global__ void k_add_gravity(double *force_prior, const double *__restrict__ mass, double g, int nf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nf) force_prior[i] += mass[i] * g;
}

// ---------------------------------------------------------------- transport_correction (ph:569-714)
__global__ void __launch_bounds__(kBlock) k_transp_pairs(PairView pv, const double *__restrict__ Vol,
                                                        const double *__restrict__ B, int nf, int nt, double *inc)
{
    const double *B11 = B, *B12 = B + nt, *B21 = B + 2 * (size_t)nt, *B22 = B + 3 * (size_t)nt;
    double *ix = inc, *iy = inc + nt;
    PAIR_LOOP(k)
    {
        int ii, jj;
        const double rk = pv.r[k];
        if (!pair_ids(pv, k, nf, nt, ii, jj) || rk <= 1e-12) continue;
        const double ex = pv.dx[k] / rk, ey = pv.dy[k] / rk, dW = pv.dW[k];
        if (jj < nf) {
            const double s11 = B11[ii] + B11[jj], s12 = B12[ii] + B12[jj];
            const double s21 = B21[ii] + B21[jj], s22 = B22[ii] + B22[jj];
            const double tx = s11 * ex + s12 * ey, ty = s21 * ex + s22 * ey;
            const double ci = -dW * Vol[jj], cj = dW * Vol[ii];
            atomicAdd(&ix[ii], ci * tx);
            atomicAdd(&iy[ii], ci * ty);
            atomicAdd(&ix[jj], cj * tx);
            atomicAdd(&iy[jj], cj * ty);
        } else {
            const double tx = B11[ii] * ex + B12[ii] * ey, ty = B21[ii] * ex + B22[ii] * ey;
            const double c = -2.0 * dW * Vol[jj];
            atomicAdd(&ix[ii], c * tx);
            atomicAdd(&iy[ii], c * ty);
        }
    }
}

__global__ void k_transp_fin(const double *inc, const double *__restrict__ pos, double h, double coeff, int nf,
                             int nt, double *pos_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    double x = pos[i], y = pos[i + nt];
    if (i < nf) {
        const double ix = inc[i], iy = inc[i + nt];
        const double n2 = ix * ix + iy * iy;
        double lim = 100.0 * n2 / (h * h);
        const double scale = coeff * h * h;
        if (lim > 1.0) lim = 1.0;
        if (lim < 0.0) lim = 0.0;
        x += scale * lim * ix;
        y += scale * lim * iy;
    }
    pos_out[i] = x;
    pos_out[i + nt] = y;
}

// ---------------------------------------------------------------- integration_1st (ph:736-967)
__global__ void k_int1_pre(const double *__restrict__ rho, const double *__restrict__ drho,
                           const double *__restrict__ pos, const double *__restrict__ vel, double dt, int nf, int nt,
                           double rho0, double p0, double *rho_out, double *p_out, double *pos_out, double *force,
                           double *diss)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    double r = rho[i], p = 0.0, x = pos[i], y = pos[i + nt];
    if (i < nf) {
        r = r + 0.5 * dt * drho[i];
        if (r < 1e-10) r = rho0;
        p = p0 * (r / rho0 - 1.0);
        x += 0.5 * dt * vel[i];
        y += 0.5 * dt * vel[i + nt];
    }
    rho_out[i] = r;
    p_out[i] = p;
    pos_out[i] = x;
    pos_out[i + nt] = y;
    force[i] = 0.0;
    force[i + nt] = 0.0;
    diss[i] = 0.0;
}

__global__ void __launch_bounds__(kBlock) k_int1_pairs(PairView pv, const double *__restrict__ Vol,
                                                      const double *__restrict__ B, const double *__restrict__ rho_h,
                                                      const double *__restrict__ p_h, const double *__restrict__ mass,
                                                      const double *__restrict__ vel,
                                                      const double *__restrict__ force_prior, int nf, int nt,
                                                      double rho0, double c_f, double *force, double *diss)
{
    const double *vx = vel, *vy = vel + nt, *fpx = force_prior, *fpy = force_prior + nt;
    const double *B11 = B, *B12 = B + nt, *B21 = B + 2 * (size_t)nt, *B22 = B + 3 * (size_t)nt;
    double *fx = force, *fy = force + nt;
    PAIR_LOOP(k)
    {
        int ii, jj;
        const double rk = pv.r[k];
        if (!pair_ids(pv, k, nf, nt, ii, jj) || rk <= 1e-12) continue;
        const double ex = pv.dx[k] / rk, ey = pv.dy[k] / rk, dW = pv.dW[k];
        if (jj < nf) {
            const double pi_ = p_h[ii], pj_ = p_h[jj];
            const double rho_bar = 0.5 * (rho_h[ii] + rho_h[jj]);
            const double ul = vx[ii] * ex + vy[ii] * ey;
            const double ur = vx[jj] * ex + vy[jj] * ey;
            double comp = ul - ur;
            if (comp < 0.0) comp = 0.0;
            const double beta = fmin(3.0 * comp, c_f);   // riemann_beta, ph:1121-1129
            const double p_star = 0.5 * (pi_ + pj_) + 0.5 * beta * rho_bar * (ul - ur);
            const double p_face = 0.5 * (0.5 * (pi_ + pj_) + p_star);
            const double tx = p_face * ((B11[ii] + B11[jj]) * ex + (B12[ii] + B12[jj]) * ey);
            const double ty = p_face * ((B21[ii] + B21[jj]) * ex + (B22[ii] + B22[jj]) * ey);
            const double gj = dW * Vol[jj], gi = dW * Vol[ii];
            const double pd = pi_ - pj_;
            atomicAdd(&fx[ii], -(tx * gj));
            atomicAdd(&fy[ii], -(ty * gj));
            atomicAdd(&fx[jj], tx * gi);
            atomicAdd(&fy[jj], ty * gi);
            atomicAdd(&diss[ii], (pd / (rho0 * c_f)) * gj);
            atomicAdd(&diss[jj], (-pd / (rho0 * c_f)) * gi);
        } else {
            const double pi_ = p_h[ii], rho_i = rho_h[ii];
            const double gj = dW * Vol[jj];
            const double ax = fpx[ii] / mass[ii], ay = fpy[ii] / mass[ii];
            const double ext = -(ax * ex + ay * ey);
            const double pw = pi_ + rho_i * rk * fmax(0.0, ext);   // virtual wall pressure, ph:934
            const double tx = B11[ii] * ex + B12[ii] * ey, ty = B21[ii] * ex + B22[ii] * ey;
            atomicAdd(&fx[ii], -((pi_ + pw) * gj * tx));
            atomicAdd(&fy[ii], -((pi_ + pw) * gj * ty));
            atomicAdd(&diss[ii], ((pi_ - pw) / (rho0 * c_f)) * gj);
        }
    }
}

__global__ void k_int1_fin(const double *__restrict__ Vol, const double *__restrict__ rho_h, const double *diss,
                           int nf, int nt, double *force, double *drho_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    if (i < nf) {
        force[i] *= Vol[i];
        force[i + nt] *= Vol[i];
        drho_out[i] = diss[i] * rho_h[i];
    } else {
        force[i] = 0.0;
        force[i + nt] = 0.0;
        drho_out[i] = 0.0;
    }
}

// ---------------------------------------------------------------- integration_2nd (ph:987-1119)
__global__ void k_int2_pre(const double *__restrict__ pos, const double *__restrict__ vel, double dt, int nf, int nt,
                           double *pos_out, double *rate)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    double x = pos[i], y = pos[i + nt];
    if (i < nf) {
        x += 0.5 * dt * vel[i];
        y += 0.5 * dt * vel[i + nt];
    }
    pos_out[i] = x;
    pos_out[i + nt] = y;
    rate[i] = 0.0;
}

__global__ void __launch_bounds__(kBlock) k_int2_pairs(PairView pv, const double *__restrict__ Vol,
                                                      const double *__restrict__ vel,
                                                      const double *__restrict__ wall_vel, int nf, int nt,
                                                      double *rate)
{
    const double *vx = vel, *vy = vel + nt, *wx = wall_vel, *wy = wall_vel + nt;
    PAIR_LOOP(k)
    {
        int ii, jj;
        const double rk = pv.r[k];
        if (!pair_ids(pv, k, nf, nt, ii, jj) || rk <= 1e-12) continue;
        const double ex = pv.dx[k] / rk, ey = pv.dy[k] / rk, dW = pv.dW[k];
        if (jj < nf) {
            const double u = (vx[ii] - vx[jj]) * ex + (vy[ii] - vy[jj]) * ey;
            atomicAdd(&rate[ii], u * dW * Vol[jj]);
            atomicAdd(&rate[jj], u * dW * Vol[ii]);
        } else {
            const double mx = 2.0 * wx[jj] - vx[ii], my = 2.0 * wy[jj] - vy[ii];   // mirrored wall velocity
            const double u = (vx[ii] - mx) * ex + (vy[ii] - my) * ey;
            atomicAdd(&rate[ii], u * dW * Vol[jj]);
        }
    }
}

__global__ void k_int2_fin(const double *rate, const double *__restrict__ rho, int nf, int nt, double *drho_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    drho_out[i] = (i < nf) ? rate[i] * rho[i] : 0.0;
}

// ---------------------------------------------------------------- integration_verlet glue (ph:1388-1450)
__global__ void k_verlet_kick(const double *__restrict__ vel, const double *__restrict__ fp,
                              const double *__restrict__ f, const double *__restrict__ mass, double dt, int nf, int nt,
                              double *vel_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    double vx = 0.0, vy = 0.0;
    if (i < nf) {
        const double inv_m = 1.0 / mass[i];
        vx = vel[i] + (fp[i] + f[i]) * inv_m * dt;
        vy = vel[i + nt] + (fp[i + nt] + f[i + nt]) * inv_m * dt;
    }
    vel_out[i] = vx;
    vel_out[i + nt] = vy;
}

__global__ void k_verlet_finish(const double *__restrict__ rho_h, const double *__restrict__ drho, double dt, int nf,
                                int nt, double rho0, double p0, double *rho_out, double *p_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    double r = rho_h[i], p = 0.0;
    if (i < nf) {
        r += drho[i] * (0.5 * dt);
        if (r < 1e-10) r = rho0;
        p = p0 * (r / rho0 - 1.0);
    }
    rho_out[i] = r;
    p_out[i] = p;
}

// ---------------------------------------------------------------- wall_shear_monitor (ph:1653-1743)
constexpr int kShearBlocks = 148 * 4;

__global__ void __launch_bounds__(256) k_shear_pairs(PairView pv, const double *__restrict__ pos,
                                                    const double *__restrict__ vel,
                                                    const double *__restrict__ wall_vel,
                                                    const double *__restrict__ Vol, const double *__restrict__ B,
                                                    int nf, int nt, double DH, double mu, double h, double *partial)
{
    const double *py = pos + nt, *vx = vel, *wx = wall_vel;
    const double *B11 = B, *B12 = B + nt, *B21 = B + 2 * (size_t)nt, *B22 = B + 3 * (size_t)nt;
    double sb = 0.0, st = 0.0;
    PAIR_LOOP(k)
    {
        const int ii = (int)pv.pi[k] - 1, jj = (int)pv.pj[k] - 1;
        const double rk = pv.r[k];
        if (ii < 0 || ii >= nf || jj < nf || jj >= nt || rk <= 1e-12) continue;   // ph:1722 (+ bounds guard)
        const double ex = pv.dx[k] / rk, ey = pv.dy[k] / rk;
        const double eBe = ex * (B11[ii] * ex + B12[ii] * ey) + ey * (B21[ii] * ex + B22[ii] * ey);
        const double dvx = vx[ii] - wx[jj];
        const double f = 4.0 * mu * eBe * pv.dW[k] * Vol[jj] * dvx / (rk + 0.01 * h) * Vol[ii];
        const double yj = py[jj];
        if (yj <= 0.0) sb += f;
        else if (yj >= DH) st += f;
    }
    __shared__ double shb[256], sht[256];
    shb[threadIdx.x] = sb;
    sht[threadIdx.x] = st;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            shb[threadIdx.x] += shb[threadIdx.x + d];
            sht[threadIdx.x] += sht[threadIdx.x + d];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = shb[0];
        partial[gridDim.x + blockIdx.x] = sht[0];
    }
}

__global__ void __launch_bounds__(256) k_shear_final(const double *partial, int nblocks, double DL, double *tau2)
{
    __shared__ double shb[256], sht[256];
    double sb = 0.0, st = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) {
        sb += partial[b];
        st += partial[nblocks + b];
    }
    shb[threadIdx.x] = sb;
    sht[threadIdx.x] = st;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            shb[threadIdx.x] += shb[threadIdx.x + d];
            sht[threadIdx.x] += sht[threadIdx.x + d];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tau2[0] = -shb[0] / DL;
        tau2[1] = -sht[0] / DL;
    }
}

}  // namespace

// ================================================================================================
// launchers
// ================================================================================================
int pairs_density(const PairView &pv, const double *mass, int nf, int nt, double rho0, double h, double inv_sigma0,
                  double *rho, double *Vol, double *B, double *scratch, cudaStream_t st)
{
    double *s_in = scratch, *s_ct = scratch + nf, *A = scratch + 2 * (size_t)nf;
    const double W0 = 10.0 / (7.0 * kPi * h * h);   // ph:33-38
    k_dens_init<<<grid_for(nf), kBlock, 0, st>>>(s_in, s_ct, A, nf, W0);
    if (pv.n) k_dens_sigma<<<pair_grid(pv.n), kBlock, 0, st>>>(pv, mass, nf, nt, rho0, s_in, s_ct);
    k_dens_rho<<<grid_for(nt), kBlock, 0, st>>>(s_in, s_ct, mass, nf, nt, rho0, inv_sigma0, rho, Vol);
    if (pv.n) k_dens_A<<<pair_grid(pv.n), kBlock, 0, st>>>(pv, Vol, nf, nt, A);
    k_dens_B<<<grid_for(nt), kBlock, 0, st>>>(A, nf, nt, B);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

int pairs_viscous(const PairView &pv, const double *vel, const double *Vol, const double *B, double mu, double h,
                  int nf, int nt, const double *wall_vel, double *force, double *scratch, cudaStream_t st)
{
    SPHB_CUDA(cudaMemsetAsync(scratch, 0, 2 * (size_t)nt * sizeof(double), st));
    if (pv.n) k_visc_pairs<<<pair_grid(pv.n), kBlock, 0, st>>>(pv, vel, Vol, B, mu, h, nf, nt, wall_vel, scratch);
    k_visc_fin<<<grid_for(nt), kBlock, 0, st>>>(scratch, Vol, nf, nt, force);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

int pairs_add_gravity(double *force_prior, const double *mass, double g, int nf, cudaStream_t st)
{
    k_add_gravity<<<grid_for(nf), kBlock, 0, st>>>(force_prior, mass, g, nf);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

int pairs_transport(const PairView &pv, const double *Vol, const double *B, const double *pos, double h, int nf,
                    int nt, double coeff, double *pos_out, double *scratch, cudaStream_t st)
{
    SPHB_CUDA(cudaMemsetAsync(scratch, 0, 2 * (size_t)nt * sizeof(double), st));
    if (pv.n) k_transp_pairs<<<pair_grid(pv.n), kBlock, 0, st>>>(pv, Vol, B, nf, nt, scratch);
    k_transp_fin<<<grid_for(nt), kBlock, 0, st>>>(scratch, pos, h, coeff, nf, nt, pos_out);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

int pairs_integration_1st(const PairView &pv, const double *Vol, const double *B, const double *rho,
                          const double *mass, const double *pos, const double *vel, const double *drho_dt,
                          const double *force_prior, double dt, int nf, int nt, double rho0, double p0, double c_f,
                          double *rho_out, double *p_out, double *pos_out, double *force_out, double *drho_out,
                          double *scratch, cudaStream_t st)
{
    double *diss = scratch;
    k_int1_pre<<<grid_for(nt), kBlock, 0, st>>>(rho, drho_dt, pos, vel, dt, nf, nt, rho0, p0, rho_out, p_out, pos_out,
                                                force_out, diss);
    if (pv.n)
        k_int1_pairs<<<pair_grid(pv.n), kBlock, 0, st>>>(pv, Vol, B, rho_out, p_out, mass, vel, force_prior, nf, nt,
                                                         rho0, c_f, force_out, diss);
    k_int1_fin<<<grid_for(nt), kBlock, 0, st>>>(Vol, rho_out, diss, nf, nt, force_out, drho_out);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

int pairs_integration_2nd(const PairView &pv, const double *Vol, const double *rho, const double *pos,
                          const double *vel, double dt, int nf, int nt, const double *wall_vel, double *pos_out,
                          double *drho_out, double *scratch, cudaStream_t st)
{
    double *rate = scratch;
    k_int2_pre<<<grid_for(nt), kBlock, 0, st>>>(pos, vel, dt, nf, nt, pos_out, rate);
    if (pv.n) k_int2_pairs<<<pair_grid(pv.n), kBlock, 0, st>>>(pv, Vol, vel, wall_vel, nf, nt, rate);
    k_int2_fin<<<grid_for(nt), kBlock, 0, st>>>(rate, rho, nf, nt, drho_out);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

int pairs_verlet_kick(const double *vel, const double *force_prior, const double *force, const double *mass,
                      double dt, int nf, int nt, double *vel_out, cudaStream_t st)
{
    k_verlet_kick<<<grid_for(nt), kBlock, 0, st>>>(vel, force_prior, force, mass, dt, nf, nt, vel_out);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

int pairs_verlet_finish(const double *rho_half, const double *drho, double dt, int nf, int nt, double rho0,
                        double p0, double *rho_out, double *p_out, cudaStream_t st)
{
    k_verlet_finish<<<grid_for(nt), kBlock, 0, st>>>(rho_half, drho, dt, nf, nt, rho0, p0, rho_out, p_out);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

int pairs_wall_shear(const PairView &pv, const double *pos, const double *vel, const double *wall_vel,
                     const double *Vol, const double *B, int nf, int nt, double DL, double DH, double mu, double h,
                     double *tau2_dev, double *scratch, cudaStream_t st)
{
    int blocks = (int)((pv.n + 255) / 256);
    if (blocks > kShearBlocks) blocks = kShearBlocks;
    if (blocks < 1) blocks = 1;
    k_shear_pairs<<<blocks, 256, 0, st>>>(pv, pos, vel, wall_vel, Vol, B, nf, nt, DH, mu, h, scratch);
    k_shear_final<<<1, 256, 0, st>>>(scratch, blocks, DL, tau2_dev);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sphb
