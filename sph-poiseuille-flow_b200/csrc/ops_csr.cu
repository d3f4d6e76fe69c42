// ops_csr.cu -- fp64 physics operators of the DEVICE-RESIDENT path as per-particle gathers over the
// ELL neighbour list (column-major, so a warp's k-th neighbour indices are one coalesced load) (rows P1-P9 and the epilogues of S1 in SURVEY.md 8(a)).
//
// One thread owns one particle and walks its CSR row; every pair is therefore evaluated from both
// ends, there are no atomics and the sums are deterministic.  All the reference's pair formulas are
// symmetric under (i <-> j, e -> -e) (ph:262-291, 502-520, 667-685, 891-925, 1091-1099), so the
// gather uses one formula per entry with e pointing from j to i.
//
// Geometry is FROZEN at list-build time, as in the reference, where every mode consumes the pair
// cache computed by the neighbour search: dx = (x_i - x_j) + shift*DL from the
// position snapshot `snap` and the image code stored in the top four bits of each list entry.
//
// Fusion: viscous + transport share one row walk (same inputs B_j, Vol_j); the pressure kernel
// recomputes the neighbour's half-step density/pressure inline (ph:857-862) and its epilogue does the
// velocity kick and the full position update (ph:1400-1404, 863-864, 1066-1069); the continuity
// kernel's epilogue finishes density/EOS (ph:1440-1446), wraps x (m:570-577) and feeds the grid-wide
// max|v| reduction (warp shuffle + one atomicMax per warp) that the next dt is computed from on the
// device (m:519-527).
#include "ops_csr.cuh"
#include "physmath.cuh"
#include "tile.cuh"

namespace sphb {

namespace {

constexpr int kPhysCap = 512;   // staged entries per tile (3 x 18 cells x ~6.8 = 365 on the lattice)

// Geometry of one list entry, recomputed from build-time entry positions (16 B staged per neighbour
// instead of the reference's 40 B cached per pair): e points from j to i.
struct Geo {
    double dx, dy, r, ex, ey;
};

__device__ __forceinline__ Geo pair_geometry(double2 pi, double xj, double yj, int m, const PhysParams &P)
{
    Geo g;
    const double shift = m == 0 ? 0.0 : (m > 0 ? P.DL : -P.DL);   // minimum image (nb:359-363)
    g.dx = (pi.x - xj) + shift;
    g.dy = pi.y - yj;
    const double r2 = g.dx * g.dx + g.dy * g.dy;
    const double inv_r = fast_rsqrt(r2);
    g.r = r2 * inv_r;
    g.ex = g.dx * inv_r;
    g.ey = g.dy * inv_r;
    return g;
}

// cubic spline, nb:116-133 (r < 2h is guaranteed by the list)
__device__ __forceinline__ double spline_W(double r, const PhysParams &P)
{
    const double q = r * P.inv_h;
    if (q < 1.0) return P.sigma * (1.0 - q * q * (1.5 - 0.75 * q));
    const double t = fmax(2.0 - q, 0.0);
    return 0.25 * P.sigma * t * t * t;
}
__device__ __forceinline__ double spline_dW(double r, const PhysParams &P)
{
    const double q = r * P.inv_h;
    if (q < 1.0) return P.sigma_h * q * (2.25 * q - 3.0);
    const double t = fmax(2.0 - q, 0.0);
    return -0.75 * P.sigma_h * t * t;
}

// ---- the tile frame shared by every gather kernel --------------------------------------------------
// TILE_BEGIN: descriptor + early exit.  TILE_WALKERS: one thread per REAL entry of the tile's own
// cells (ew = entry, i = particle); ghost entries are skipped.  tile_for_row: the walker's ELL row; for
// each neighbour `ea` = its entry, `sl` = its shared-memory slot (or -1: load from global), `m` = shift.
#define SPHB_TILE_BEGIN()                                                                         \
    __shared__ TileDesc T;                                                                        \
    tile_setup(T, blockIdx.x, S.ncx, S.grid->ncy, S.cell_start, kPhysCap);                        \
    if (!T.valid || T.own_b == T.own_e) return

#define SPHB_TILE_WALKERS(ew, i)                                                                  \
    for (int ew = T.own_b + (int)threadIdx.x; ew < T.own_e; ew += kTileThreads)                    \
        if (const int key__ = S.ent_key[ew]; (key__ & 3) == 0)                                    \
            if (const int i = key__ >> 3; true)

// The walker's ELL row.  The column indices are fetched eight at a time (independent coalesced loads,
// next chunk in flight while the current one is processed) so a row costs ~3 global latencies, not 20.
template <class Body>
__device__ __forceinline__ void tile_for_row(const StateView &S, const TileDesc &T, int ew, Body &&body)
{
    constexpr int U = 8;
    const int n = (size_t)ew < S.ell_stride ? min(S.ell_cnt[ew], S.ell_cap) : 0;
    const unsigned *col = S.ell_col + ew;
    unsigned cur[U], nxt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = u < n ? col[(size_t)u * S.ell_stride] : 0u;
    for (int k0 = 0; k0 < n; k0 += U) {
#pragma unroll
        for (int u = 0; u < U; ++u) nxt[u] = (k0 + U + u) < n ? col[(size_t)(k0 + U + u) * S.ell_stride] : 0u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (k0 + u < n) {
                const unsigned c = cur[u];
                const int ea = ell_entry(c);
                body(ea, tile_lookup(T, ea, ell_row(c)), ell_shift(c));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = nxt[u];
    }
}

// ---------------------------------------------------------------- P1 pass 1: sigma sums -> rho, Vol
__global__ void __launch_bounds__(kTileThreads) k_density(PhysParams P, StateView S)
{
    __shared__ double s_x[kPhysCap], s_y[kPhysCap];
    __shared__ int s_id[kPhysCap];
    SPHB_TILE_BEGIN();
    SPHB_TILE_FOR_STAGED(T, e, slot)
    {
        const double2 p = S.ent_pos[e];
        s_x[slot] = p.x;
        s_y[slot] = p.y;
        s_id[slot] = S.ent_key[e] >> 3;
    }
    __syncthreads();
    SPHB_TILE_WALKERS(ew, i)
    {
        const double m_i = S.mass[i];
        if (i >= P.nf) {   // wall rows: rho = rho0 (ph:214-216), Vol = m/rho0
            S.rho[i] = P.rho0;
            S.Vol[i] = m_i / P.rho0;
            continue;
        }
        const double2 pi = S.ent_pos[ew];
        double s_in = P.sigma;   // W(0), ph:178-181
        double s_ct = 0.0;
        tile_for_row(S, T, ew, [&](const int ea, const int sl, const int m) {
            double xj, yj;
            int j;
            if (sl >= 0) {
                xj = s_x[sl];
                yj = s_y[sl];
                j = s_id[sl];
            } else {
                const double2 p = S.ent_pos[ea];
                xj = p.x;
                yj = p.y;
                j = S.ent_key[ea] >> 3;
            }
            const Geo g = pair_geometry(pi, xj, yj, m, P);
            const double W = spline_W(g.r, P);
            if (j < P.nf) s_in += W;
            else s_ct += W * (S.mass[j] * P.inv_rho0);
        });
        double rho = s_in * P.rho0 * P.inv_sigma0;
        rho += s_ct * P.rho0 * P.rho0 * P.inv_sigma0 / m_i;
        if (rho <= 1e-12) rho = P.rho0;
        S.rho[i] = rho;
        S.Vol[i] = m_i / rho;
    }
}

// ---------------------------------------------------------------- P1 pass 2: A -> B (ph:239-366)
__global__ void __launch_bounds__(kTileThreads) k_kgc(PhysParams P, StateView S)
{
    __shared__ double s_x[kPhysCap], s_y[kPhysCap], s_V[kPhysCap];
    SPHB_TILE_BEGIN();
    SPHB_TILE_FOR_STAGED(T, e, slot)
    {
        const double2 p = S.ent_pos[e];
        s_x[slot] = p.x;
        s_y[slot] = p.y;
        s_V[slot] = S.Vol[S.ent_key[e] >> 3];
    }
    __syncthreads();
    SPHB_TILE_WALKERS(ew, i)
    {
        Mat2 B;
        B.b11 = 1.0;
        B.b12 = 0.0;
        B.b21 = 0.0;
        B.b22 = 1.0;
        if (i < P.nf) {
            const double2 pi = S.ent_pos[ew];
            double a11 = 0.0, a12 = 0.0, a21 = 0.0, a22 = 0.0;
            tile_for_row(S, T, ew, [&](const int ea, const int sl, const int m) {
                double xj, yj, Vj;
                if (sl >= 0) {
                    xj = s_x[sl];
                    yj = s_y[sl];
                    Vj = s_V[sl];
                } else {
                    const double2 p = S.ent_pos[ea];
                    xj = p.x;
                    yj = p.y;
                    Vj = S.Vol[S.ent_key[ea] >> 3];
                }
                const Geo g = pair_geometry(pi, xj, yj, m, P);
                if (!(g.r > 1e-12)) return;
                const double f = spline_dW(g.r, P) * Vj;
                const double fx = f * g.ex, fy = f * g.ey;
                a11 -= g.dx * fx;
                a12 -= g.dx * fy;
                a21 -= g.dy * fx;
                a22 -= g.dy * fy;
            });
            kgc_from_A(a11, a12, a21, a22, B.b11, B.b12, B.b21, B.b22);
        }
        S.B[i] = B;
    }
}

// ---------------------------------------------------------------- P2 + P3 fused (ph:469-545, 636-710)
template <bool VISC, bool TRANSP>
__global__ void __launch_bounds__(kTileThreads) k_visc_transport(PhysParams P, StateView S, int add_gravity)
{
    __shared__ double s_x[kPhysCap], s_y[kPhysCap], s_V[kPhysCap], s_vx[kPhysCap], s_vy[kPhysCap];
    __shared__ double s_b11[kPhysCap], s_b12[kPhysCap], s_b21[kPhysCap], s_b22[kPhysCap];
    __shared__ int s_id[kPhysCap];
    SPHB_TILE_BEGIN();
    SPHB_TILE_FOR_STAGED(T, e, slot)
    {
        const int j = S.ent_key[e] >> 3;
        const double2 p = S.ent_pos[e];
        s_x[slot] = p.x;
        s_y[slot] = p.y;
        s_id[slot] = j;
        s_V[slot] = S.Vol[j];
        const double2 v = j < P.nf ? S.vel[j] : S.wall_vel[j];   // a wall neighbour contributes its wall velocity
        s_vx[slot] = v.x;
        s_vy[slot] = v.y;
        const Mat2 B = S.B[j];
        s_b11[slot] = B.b11;
        s_b12[slot] = B.b12;
        s_b21[slot] = B.b21;
        s_b22[slot] = B.b22;
    }
    __syncthreads();
    const double h001 = 0.01 * P.h;
    SPHB_TILE_WALKERS(ew, i)
    {
        if (i >= P.nf) {
            if (VISC) S.force_prior[i] = make_double2(0.0, 0.0);
            continue;
        }
        const double2 pi = S.ent_pos[ew];
        const double2 vi = S.vel[i];
        const Mat2 Bi = S.B[i];
        const double Vi = S.Vol[i];
        double ax = 0.0, ay = 0.0, ix = 0.0, iy = 0.0;
        tile_for_row(S, T, ew, [&](const int ea, const int sl, const int m) {
            double xj, yj, Vj, vjx, vjy;
            Mat2 Bj;
            int j;
            if (sl >= 0) {
                xj = s_x[sl];
                yj = s_y[sl];
                Vj = s_V[sl];
                vjx = s_vx[sl];
                vjy = s_vy[sl];
                Bj.b11 = s_b11[sl];
                Bj.b12 = s_b12[sl];
                Bj.b21 = s_b21[sl];
                Bj.b22 = s_b22[sl];
                j = s_id[sl];
            } else {
                j = S.ent_key[ea] >> 3;
                const double2 p = S.ent_pos[ea];
                xj = p.x;
                yj = p.y;
                Vj = S.Vol[j];
                const double2 v = j < P.nf ? S.vel[j] : S.wall_vel[j];
                vjx = v.x;
                vjy = v.y;
                Bj = S.B[j];
            }
            const Geo g = pair_geometry(pi, xj, yj, m, P);
            if (!(g.r > 1e-12)) return;
            const double ex = g.ex, ey = g.ey;
            const double gj = spline_dW(g.r, P) * Vj;
            double tx, ty, wv, wt;
            if (j < P.nf) {
                const double s11 = Bi.b11 + Bj.b11, s12 = Bi.b12 + Bj.b12;
                const double s21 = Bi.b21 + Bj.b21, s22 = Bi.b22 + Bj.b22;
                tx = s11 * ex + s12 * ey;
                ty = s21 * ex + s22 * ey;
                wv = 1.0;
                wt = 1.0;
            } else {
                tx = Bi.b11 * ex + Bi.b12 * ey;
                ty = Bi.b21 * ex + Bi.b22 * ey;
                wv = 4.0;   // no-slip mirror, ph:526
                wt = 2.0;   // ph:690
            }
            if (VISC) {
                const double eBe = ex * tx + ey * ty;
                const double cf = wv * eBe * P.mu * gj * fast_rcp(g.r + h001);
                ax += cf * (vi.x - vjx);
                ay += cf * (vi.y - vjy);
            }
            if (TRANSP) {
                const double ci = -wt * gj;
                ix += ci * tx;
                iy += ci * ty;
            }
        });
        if (VISC) {
            double fx = ax * Vi, fy = ay * Vi;
            if (add_gravity) fx += S.mass[i] * P.g;   // m:392
            S.force_prior[i] = make_double2(fx, fy);
        }
        if (TRANSP) {
            const double n2 = ix * ix + iy * iy;
            double lim = 100.0 * n2 / (P.h * P.h);
            const double scale = P.transport_coeff * P.h * P.h;
            if (lim > 1.0) lim = 1.0;
            if (lim < 0.0) lim = 0.0;
            double2 x = S.pos[i];
            x.x += scale * lim * ix;
            x.y += scale * lim * iy;
            S.pos[i] = x;
        }
    }
}

__device__ __forceinline__ double half_step_rho(double rho, double drho, double dt, double rho0)
{
    double r = rho + 0.5 * dt * drho;   // ph:858
    if (r < 1e-10) r = rho0;            // ph:859-861
    return r;
}

// ---------------------------------------------------------------- dt on the device (m:519-527) and the
// half-step density / pressure of every particle (ph:857-862), which the pressure gather reads per
// neighbour.  Every thread derives the same dt from the device vmax; thread 0 publishes it.
__global__ void __launch_bounds__(kBlock) k_time_step(PhysParams P, StateView S, DevScalars *sc, double remain,
                                                     double fixed_dt)
{
    double dt;
    if (fixed_dt > 0.0) {
        dt = fixed_dt;
    } else {
        const double vmax = sc->vmax;
        const double dt_a = 0.25 * P.h / fmax(P.c_f + vmax, 1e-12);
        const double dt_v = 0.125 * P.h * P.h / fmax(P.nu, 1e-12);
        const double dt_b = 0.25 * sqrt(P.h / fmax(fabs(P.g), 1e-12));
        dt = fmin(fmin(dt_a, dt_v), fmin(dt_b, remain));
        dt = fmax(dt, 1e-12);
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        sc->dt = dt;
        sc->vmax_bits = 0ull;
    }
    if (i < P.nf) {
        const double rho_h = half_step_rho(S.rho[i], S.drho_dt[i], dt, P.rho0);
        S.rp_half[i] = make_double2(rho_h, P.p0 * (rho_h / P.rho0 - 1.0));   // ph:862
    }
}

__global__ void k_advance_clock(DevScalars *sc)
{
    sc->t += sc->dt;
    sc->step += 1;
    sc->vmax = __longlong_as_double((long long)sc->vmax_bits);
}

// ---------------------------------------------------------------- P4 + velocity kick + position update
__global__ void __launch_bounds__(kTileThreads) k_pressure(PhysParams P, StateView S, const DevScalars *__restrict__ sc)
{
    __shared__ double s_x[kPhysCap], s_y[kPhysCap], s_V[kPhysCap], s_vx[kPhysCap], s_vy[kPhysCap];
    __shared__ double s_rho[kPhysCap], s_p[kPhysCap];
    __shared__ double s_b11[kPhysCap], s_b12[kPhysCap], s_b21[kPhysCap], s_b22[kPhysCap];
    __shared__ int s_id[kPhysCap];
    SPHB_TILE_BEGIN();
    SPHB_TILE_FOR_STAGED(T, e, slot)
    {
        const int j = S.ent_key[e] >> 3;
        const double2 p = S.ent_pos[e];
        s_x[slot] = p.x;
        s_y[slot] = p.y;
        s_id[slot] = j;
        s_V[slot] = S.Vol[j];
        if (j < P.nf) {   // wall neighbours need only position and volume (ph:927-947)
            const double2 v = S.vel[j];
            s_vx[slot] = v.x;
            s_vy[slot] = v.y;
            const double2 rp = S.rp_half[j];
            s_rho[slot] = rp.x;
            s_p[slot] = rp.y;
            const Mat2 B = S.B[j];
            s_b11[slot] = B.b11;
            s_b12[slot] = B.b12;
            s_b21[slot] = B.b21;
            s_b22[slot] = B.b22;
        }
    }
    __syncthreads();
    const double dt = sc->dt;
    SPHB_TILE_WALKERS(ew, i)
    {
        if (i >= P.nf) {   // wall rows: force 0, velocity 0 (ph:958-962, 1405-1408)
            S.force[i] = make_double2(0.0, 0.0);
            S.vel_new[i] = make_double2(0.0, 0.0);
            continue;
        }
        const double2 pi = S.ent_pos[ew];
        const double2 vi = S.vel[i];
        const Mat2 Bi = S.B[i];
        const double Vi = S.Vol[i], m_i = S.mass[i];
        const double2 fp = S.force_prior[i];
        const double2 rpi = S.rp_half[i];
        const double rho_i = rpi.x, p_i = rpi.y;
        const double inv_m = 1.0 / m_i;
        const double wax = fp.x * inv_m, way = fp.y * inv_m;   // a = F_prior / m, ph:929-930
        double fx = 0.0, fy = 0.0;
        tile_for_row(S, T, ew, [&](const int ea, const int sl, const int m) {
            double xj, yj, Vj;
            int j;
            if (sl >= 0) {
                xj = s_x[sl];
                yj = s_y[sl];
                Vj = s_V[sl];
                j = s_id[sl];
            } else {
                j = S.ent_key[ea] >> 3;
                const double2 p = S.ent_pos[ea];
                xj = p.x;
                yj = p.y;
                Vj = S.Vol[j];
            }
            const Geo g = pair_geometry(pi, xj, yj, m, P);
            if (!(g.r > 1e-12)) return;
            const double ex = g.ex, ey = g.ey;
            const double gj = spline_dW(g.r, P) * Vj;
            if (j < P.nf) {
                double vjx, vjy, rho_j, p_j;
                Mat2 Bj;
                if (sl >= 0) {
                    vjx = s_vx[sl];
                    vjy = s_vy[sl];
                    rho_j = s_rho[sl];
                    p_j = s_p[sl];
                    Bj.b11 = s_b11[sl];
                    Bj.b12 = s_b12[sl];
                    Bj.b21 = s_b21[sl];
                    Bj.b22 = s_b22[sl];
                } else {
                    const double2 v = S.vel[j];
                    vjx = v.x;
                    vjy = v.y;
                    const double2 rp = S.rp_half[j];
                    rho_j = rp.x;
                    p_j = rp.y;
                    Bj = S.B[j];
                }
                const double rho_bar = 0.5 * (rho_i + rho_j);
                const double du = (vi.x - vjx) * ex + (vi.y - vjy) * ey;     // u_l - u_r
                const double beta = fmin(3.0 * fmax(du, 0.0), P.c_f);        // ph:1121-1129
                const double p_avg = 0.5 * (p_i + p_j);
                const double p_star = p_avg + 0.5 * beta * rho_bar * du;
                const double pg = 0.5 * (p_avg + p_star) * gj;
                fx -= pg * ((Bi.b11 + Bj.b11) * ex + (Bi.b12 + Bj.b12) * ey);
                fy -= pg * ((Bi.b21 + Bj.b21) * ex + (Bi.b22 + Bj.b22) * ey);
            } else {
                const double ext = -(wax * ex + way * ey);
                const double pw = p_i + rho_i * g.r * fmax(0.0, ext);   // virtual wall pressure, ph:934
                const double pg = (p_i + pw) * gj;
                fx -= pg * (Bi.b11 * ex + Bi.b12 * ey);
                fy -= pg * (Bi.b21 * ex + Bi.b22 * ey);
            }
        });
        fx *= Vi;
        fy *= Vi;
        S.force[i] = make_double2(fx, fy);
        double2 vn;
        vn.x = vi.x + (fp.x + fx) / m_i * dt;   // ph:1402-1403
        vn.y = vi.y + (fp.y + fy) / m_i * dt;
        S.vel_new[i] = vn;
        double2 x = S.pos[i];
        x.x += 0.5 * dt * vi.x;   // ph:863-864
        x.y += 0.5 * dt * vi.y;
        x.x += 0.5 * dt * vn.x;   // ph:1067-1068
        x.y += 0.5 * dt * vn.y;
        S.pos[i] = x;
    }
}

__device__ __forceinline__ double wrap_x(double x, double DL)   // MATLAB mod(x, DL), m:573
{
    if (x >= 0.0 && x < DL) return x;
    double m = fmod(x, DL);
    if (m < 0.0) m += DL;
    return m;
}

// ---------------------------------------------------------------- P5 + density/EOS finish + wrap + max|v|
__global__ void __launch_bounds__(kTileThreads) k_continuity(PhysParams P, StateView S, DevScalars *sc, int do_wrap)
{
    __shared__ double s_x[kPhysCap], s_y[kPhysCap], s_V[kPhysCap], s_vx[kPhysCap], s_vy[kPhysCap];
    __shared__ int s_id[kPhysCap];
    SPHB_TILE_BEGIN();
    SPHB_TILE_FOR_STAGED(T, e, slot)
    {
        const int j = S.ent_key[e] >> 3;
        const double2 p = S.ent_pos[e];
        s_x[slot] = p.x;
        s_y[slot] = p.y;
        s_id[slot] = j;
        s_V[slot] = S.Vol[j];
        const double2 v = j < P.nf ? S.vel_new[j] : S.wall_vel[j];
        s_vx[slot] = v.x;
        s_vy[slot] = v.y;
    }
    __syncthreads();
    const double dt = sc->dt;
    double speed = 0.0;
    SPHB_TILE_WALKERS(ew, i)
    {
        if (i >= P.nf) {
            S.drho_dt[i] = 0.0;   // ph:1114-1116
            S.p[i] = 0.0;         // ph:1449
            S.vel[i] = make_double2(0.0, 0.0);   // ph:1405-1408, m:270
            continue;
        }
        const double2 pi = S.ent_pos[ew];
        const double2 vi = S.vel_new[i];
        double rate = 0.0;
        tile_for_row(S, T, ew, [&](const int ea, const int sl, const int m) {
            double xj, yj, Vj, vjx, vjy;
            int j;
            if (sl >= 0) {
                xj = s_x[sl];
                yj = s_y[sl];
                Vj = s_V[sl];
                vjx = s_vx[sl];
                vjy = s_vy[sl];
                j = s_id[sl];
            } else {
                j = S.ent_key[ea] >> 3;
                const double2 p = S.ent_pos[ea];
                xj = p.x;
                yj = p.y;
                Vj = S.Vol[j];
                const double2 v = j < P.nf ? S.vel_new[j] : S.wall_vel[j];
                vjx = v.x;
                vjy = v.y;
            }
            const Geo g = pair_geometry(pi, xj, yj, m, P);
            if (!(g.r > 1e-12)) return;
            const double gj = spline_dW(g.r, P) * Vj;
            double ux, uy;
            if (j < P.nf) {
                ux = vi.x - vjx;
                uy = vi.y - vjy;
            } else {
                ux = vi.x - (2.0 * vjx - vi.x);   // mirrored velocity, ph:1101-1102
                uy = vi.y - (2.0 * vjy - vi.y);
            }
            rate += (ux * g.ex + uy * g.ey) * gj;
        });
        const double rho_h = S.rp_half[i].x;
        const double drho = rate * rho_h;          // ph:1112
        double rho = rho_h + drho * (0.5 * dt);    // ph:1441
        if (rho < 1e-10) rho = P.rho0;
        S.drho_dt[i] = drho;
        S.rho[i] = rho;
        S.p[i] = P.p0 * (rho / P.rho0 - 1.0);
        if (do_wrap) {
            double2 x = S.pos[i];
            x.x = wrap_x(x.x, P.DL);
            S.pos[i] = x;
        }
        S.vel[i] = vi;   // nobody reads `vel` in this kernel (neighbours read vel_new): commit in place
        speed = fmax(speed, sqrt(vi.x * vi.x + vi.y * vi.y));
    }
    // grid-wide max |v| over fluid rows: warp shuffle, then one atomicMax per warp
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) speed = fmax(speed, __shfl_xor_sync(0xffffffffu, speed, d));
    if ((threadIdx.x & 31) == 0 && speed > 0.0)
        atomicMax(&sc->vmax_bits, (unsigned long long)__double_as_longlong(speed));
}

__global__ void __launch_bounds__(256) k_vmax(int nf, const double2 *__restrict__ vel, DevScalars *sc)
{
    double speed = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) {
        const double2 v = vel[i];
        speed = fmax(speed, sqrt(v.x * v.x + v.y * v.y));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) speed = fmax(speed, __shfl_xor_sync(0xffffffffu, speed, d));
    if ((threadIdx.x & 31) == 0 && speed > 0.0)
        atomicMax(&sc->vmax_bits, (unsigned long long)__double_as_longlong(speed));
}
__global__ void k_vmax_reset(DevScalars *sc) { sc->vmax_bits = 0ull; }
__global__ void k_vmax_commit(DevScalars *sc) { sc->vmax = __longlong_as_double((long long)sc->vmax_bits); }

__global__ void __launch_bounds__(kBlock) k_wrap(int nf, double DL, double2 *pos)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    double2 x = pos[i];
    x.x = wrap_x(x.x, DL);
    pos[i] = x;
}

// ---------------------------------------------------------------- P8 wall shear (ph:1713-1742)
// Only wall particles at y <= 0 or y >= DH count (ph:1731-1736) and the pair distance is < 2h, so a
// fluid particle with 2h <= y <= DH - 2h cannot contribute: those rows are skipped unread.
__global__ void __launch_bounds__(256) k_wall_shear(PhysParams P, StateView S, double *partials)
{
    double sb = 0.0, st = 0.0;
    const double band = 2.0 * P.h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.nf; i += gridDim.x * blockDim.x) {
        const double2 pi = S.snap[i];
        if (pi.y >= band && pi.y <= P.DH - band) continue;
        const int ew = S.ent_of[i];
        const double vx = S.vel[i].x;
        const Mat2 Bi = S.B[i];
        const double Vi = S.Vol[i];
        const int n = (size_t)ew < S.ell_stride ? min(S.ell_cnt[ew], S.ell_cap) : 0;
        for (int k = 0; k < n; ++k) {
            const unsigned c = S.ell_col[(size_t)k * S.ell_stride + ew];
            const int ea = ell_entry(c);
            const int j = S.ent_key[ea] >> 3;
            if (j < P.nf) continue;
            const double2 pj = S.ent_pos[ea];
            const Geo g = pair_geometry(pi, pj.x, pj.y, ell_shift(c), P);
            if (!(g.r > 1e-12)) continue;
            const double ex = g.ex, ey = g.ey;
            const double eBe = ex * (Bi.b11 * ex + Bi.b12 * ey) + ey * (Bi.b21 * ex + Bi.b22 * ey);
            const double dvx = vx - S.wall_vel[j].x;
            const double f = 4.0 * P.mu * eBe * spline_dW(g.r, P) * S.Vol[j] * dvx / (g.r + 0.01 * P.h) * Vi;
            if (pj.y <= 0.0) sb += f;
            else if (pj.y >= P.DH) st += f;
        }
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
        partials[blockIdx.x] = shb[0];
        partials[gridDim.x + blockIdx.x] = sht[0];
    }
}

__global__ void __launch_bounds__(256) k_wall_shear_final(const double *partials, int nblocks, double DL,
                                                         DevScalars *sc)
{
    __shared__ double shb[256], sht[256];
    double sb = 0.0, st = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) {
        sb += partials[b];
        st += partials[nblocks + b];
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
        sc->tau_bottom = -shb[0] / DL;
        sc->tau_top = -sht[0] / DL;
    }
}

// ---------------------------------------------------------------- binned u(y) monitor (m:579-605)
__global__ void __launch_bounds__(kBlock) k_profile(PhysParams P, StateView S, double x0, double half_width,
                                                   int n_bins, const double *__restrict__ edges, double *sum,
                                                   double *cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.nf) return;
    const double2 x = S.pos[i];
    if (x0 >= 0.0) {
        const double xw = wrap_x(x.x, P.DL);
        double d = fabs(xw - x0);
        d = fmin(d, P.DL - d);
        if (!(d <= half_width)) return;
    }
    const double y = x.y;
    if (!(y >= edges[0] && y <= edges[n_bins])) return;   // discretize(): out of range dropped
    int b = (int)floor((y - edges[0]) / (edges[n_bins] - edges[0]) * n_bins);
    if (b < 0) b = 0;
    if (b > n_bins - 1) b = n_bins - 1;
    while (b > 0 && y < edges[b]) --b;
    while (b < n_bins - 1 && y >= edges[b + 1]) ++b;
    atomicAdd(&sum[b], S.vel[i].x);
    atomicAdd(&cnt[b], 1.0);
}

}  // namespace

// ================================================================================================
int csr_density(const PhysParams &P, const StateView &S, cudaStream_t st)
{
    k_density<<<S.n_tiles, kTileThreads, 0, st>>>(P, S);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_kgc(const PhysParams &P, const StateView &S, cudaStream_t st)
{
    k_kgc<<<S.n_tiles, kTileThreads, 0, st>>>(P, S);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_visc_transport(const PhysParams &P, const StateView &S, bool do_visc, bool add_gravity, bool do_transport,
                       cudaStream_t st)
{
    const unsigned g = S.n_tiles;
    if (do_visc && do_transport) k_visc_transport<true, true><<<g, kTileThreads, 0, st>>>(P, S, add_gravity ? 1 : 0);
    else if (do_visc) k_visc_transport<true, false><<<g, kTileThreads, 0, st>>>(P, S, add_gravity ? 1 : 0);
    else if (do_transport) k_visc_transport<false, true><<<g, kTileThreads, 0, st>>>(P, S, 0);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_time_step(const PhysParams &P, const StateView &S, DevScalars *sc, double remain, double fixed_dt,
                  cudaStream_t st)
{
    k_time_step<<<grid_for(P.nf), kBlock, 0, st>>>(P, S, sc, remain, fixed_dt);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_pressure(const PhysParams &P, const StateView &S, const DevScalars *sc, cudaStream_t st)
{
    k_pressure<<<S.n_tiles, kTileThreads, 0, st>>>(P, S, sc);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_continuity(const PhysParams &P, const StateView &S, DevScalars *sc, bool wrap, cudaStream_t st)
{
    k_continuity<<<S.n_tiles, kTileThreads, 0, st>>>(P, S, sc, wrap ? 1 : 0);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_advance_clock(DevScalars *sc, cudaStream_t st)
{
    k_advance_clock<<<1, 1, 0, st>>>(sc);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_vmax(const PhysParams &P, const double2 *vel, DevScalars *sc, cudaStream_t st)
{
    int blocks = (P.nf + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    k_vmax_reset<<<1, 1, 0, st>>>(sc);
    k_vmax<<<blocks, 256, 0, st>>>(P.nf, vel, sc);
    k_vmax_commit<<<1, 1, 0, st>>>(sc);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_wrap(const PhysParams &P, double2 *pos, cudaStream_t st)
{
    k_wrap<<<grid_for(P.nf), kBlock, 0, st>>>(P.nf, P.DL, pos);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_wall_shear(const PhysParams &P, const StateView &S, DevScalars *sc, double *partials, cudaStream_t st)
{
    int blocks = (P.nf + 255) / 256;
    if (blocks > kShearPartialBlocks) blocks = kShearPartialBlocks;
    if (blocks < 1) blocks = 1;
    k_wall_shear<<<blocks, 256, 0, st>>>(P, S, partials);
    k_wall_shear_final<<<1, 256, 0, st>>>(partials, blocks, P.DL, sc);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_profile(const PhysParams &P, const StateView &S, double x0, double half_width, int n_bins,
                const double *edges_dev, double *sum_dev, double *cnt_dev, cudaStream_t st)
{
    k_profile<<<grid_for(P.nf), kBlock, 0, st>>>(P, S, x0, half_width, n_bins, edges_dev, sum_dev, cnt_dev);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sphb
