/*
 * oracle/sph_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Serial plain-C restatement of the reference's particle loop.  Every function cites the reference
 * lines it follows (nb = mex/sph_neighbor_search_mex.c, ph = mex/sph_physics_mex.c,
 * m = SPH_Poiseuille.m).  Arithmetic is written with the same expression grouping as the reference
 * so that, compiled with the same flags (-O2 -ffp-contract=off, no FMA), the outputs are
 * bit-identical to oracle/_ref/libsph_ref.so run with OMP_NUM_THREADS=1 -- that identity is what
 * tests/test_oracle_pin.py checks ("pinned" oracle).  The data structures are deliberately not the
 * reference's: the cell list is a counting-sorted entry array walked backwards (the reference uses
 * head/next linked lists, whose walk order is "last inserted first"), and the per-mode temporaries
 * are kept in one scratch block.
 */
#include "sph_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846

/* ---------------------------------------------------------------------------------------------
 * Cubic-spline kernel, 2-D.  Follows nb:116-133 (cubic_kernel_2d) and ph:33-38 (W(0)).
 * ------------------------------------------------------------------------------------------- */
static void spline2d(double r, double h, double *W, double *dW)
{
    const double sigma = 10.0 / (7.0 * ORC_PI * h * h);
    const double q = r / h;
    if (q < 1.0) {
        *W = sigma * (1.0 - 1.5 * q * q + 0.75 * q * q * q);
        *dW = sigma * (-3.0 * q + 2.25 * q * q) / h;
    } else if (q < 2.0) {
        const double t = 2.0 - q;
        *W = sigma * 0.25 * t * t * t;
        *dW = -sigma * 0.75 * t * t / h;
    } else {
        *W = 0.0;
        *dW = 0.0;
    }
}

static int wrap_col(int c, int ncx) /* nb:57-67 */
{
    if (ncx <= 0) return 0;
    c %= ncx;
    if (c < 0) c += ncx;
    return c;
}

typedef struct {
    int particle;
    int cell;
    double x, y;
} cell_entry;

static void pairs_push(orc_pairs *p, size_t *cap, int i, int j, double dx, double dy, double r, double W,
                       double dW)
{
    if (p->n_pairs == *cap) {
        size_t ncap = *cap ? *cap * 2 : 1024;
        p->pair_i = (double *)realloc(p->pair_i, ncap * sizeof(double));
        p->pair_j = (double *)realloc(p->pair_j, ncap * sizeof(double));
        p->dx = (double *)realloc(p->dx, ncap * sizeof(double));
        p->dy = (double *)realloc(p->dy, ncap * sizeof(double));
        p->r = (double *)realloc(p->r, ncap * sizeof(double));
        p->W = (double *)realloc(p->W, ncap * sizeof(double));
        p->dW = (double *)realloc(p->dW, ncap * sizeof(double));
        *cap = ncap;
    }
    p->pair_i[p->n_pairs] = (double)(i + 1);
    p->pair_j[p->n_pairs] = (double)(j + 1);
    p->dx[p->n_pairs] = dx;
    p->dy[p->n_pairs] = dy;
    p->r[p->n_pairs] = r;
    p->W[p->n_pairs] = W;
    p->dW[p->n_pairs] = dW;
    p->n_pairs++;
}

/*
 * Neighbour search.  Follows nb:185-421:
 *   y range over ALL particles nb:245-250; grid nb:253-258; wrap + binning nb:270-279;
 *   ghost entries nb:282-295; stencil walk nb:312-392 (skip j==i, seen_neighbor, fluid j<i;
 *   minimum image nb:359-363; accept nb:368,372; seen mark only on accept nb:383).
 * Entry order inside a cell: the reference pushes at the list head, so it visits entries in
 * DEcreasing insertion order; here entries are counting-sorted by cell (stable) and each cell's
 * range is walked from its end to its start, which is the same order.
 */
int orc_neighbor_search(const double *pos, int n_fluid, int n_total, double h, double DL, orc_pairs *out)
{
    const double *x = pos, *y = pos + n_total;
    double y_min, y_max, cell, cutoff, rc2;
    int ncx, ncy, ncells, i, n_entries = 0;
    cell_entry *ent, *sorted;
    int *cell_start, *cx_of, *cy_of, *seen;
    size_t cap = 0;

    memset(out, 0, sizeof(*out));
    if (n_total <= 0 || n_fluid <= 0 || n_fluid > n_total) return -1; /* SPH:Neighbor:count nb:237 */
    if (h <= 0.0 || DL <= 0.0) return -2;                              /* SPH:Neighbor:param nb:240 */

    y_min = y_max = y[0];
    for (i = 1; i < n_total; ++i) {
        if (y[i] < y_min) y_min = y[i];
        if (y[i] > y_max) y_max = y[i];
    }
    cell = 2.0 * h;
    ncx = (int)ceil(DL / cell);
    if (ncx < 1) ncx = 1;
    ncy = (int)ceil((y_max - y_min + 1e-12) / cell) + 1;
    if (ncy < 1) ncy = 1;
    ncells = ncx * ncy;
    cutoff = 2.0 * h;

    ent = (cell_entry *)malloc((size_t)n_total * 3 * sizeof(cell_entry));
    sorted = (cell_entry *)malloc((size_t)n_total * 3 * sizeof(cell_entry));
    cell_start = (int *)calloc((size_t)ncells + 1, sizeof(int));
    cx_of = (int *)malloc((size_t)n_total * sizeof(int));
    cy_of = (int *)malloc((size_t)n_total * sizeof(int));
    seen = (int *)malloc((size_t)n_total * sizeof(int));

    for (i = 0; i < n_total; ++i) {
        double xw = x[i] - floor(x[i] / DL) * DL;
        int cxi = wrap_col((int)floor(xw / cell), ncx);
        int cyi = (int)floor((y[i] - y_min) / cell);
        if (cyi < 0) cyi = 0;
        if (cyi >= ncy) cyi = ncy - 1;
        cx_of[i] = cxi;
        cy_of[i] = cyi;
        ent[n_entries].particle = i;
        ent[n_entries].cell = cyi * ncx + cxi;
        ent[n_entries].x = xw;
        ent[n_entries].y = y[i];
        n_entries++;
        if (xw > DL - cutoff) {
            double gx = xw - DL;
            int gcx = wrap_col((int)floor(gx / cell), ncx);
            if (gcx != cxi) {
                ent[n_entries].particle = i;
                ent[n_entries].cell = cyi * ncx + gcx;
                ent[n_entries].x = gx;
                ent[n_entries].y = y[i];
                n_entries++;
            }
        }
        if (xw < cutoff) {
            double gx = xw + DL;
            int gcx = wrap_col((int)floor(gx / cell), ncx);
            if (gcx != cxi) {
                ent[n_entries].particle = i;
                ent[n_entries].cell = cyi * ncx + gcx;
                ent[n_entries].x = gx;
                ent[n_entries].y = y[i];
                n_entries++;
            }
        }
        seen[i] = -1;
    }
    /* stable counting sort of the entries by cell */
    for (i = 0; i < n_entries; ++i) cell_start[ent[i].cell + 1]++;
    for (i = 0; i < ncells; ++i) cell_start[i + 1] += cell_start[i];
    {
        int *fill = (int *)malloc((size_t)ncells * sizeof(int));
        memcpy(fill, cell_start, (size_t)ncells * sizeof(int));
        for (i = 0; i < n_entries; ++i) sorted[fill[ent[i].cell]++] = ent[i];
        free(fill);
    }

    rc2 = (2.0 * h) * (2.0 * h);
    for (i = 0; i < n_fluid; ++i) {
        int oy, ox;
        for (oy = -1; oy <= 1; ++oy) {
            int cy = cy_of[i] + oy;
            if (cy < 0 || cy >= ncy) continue;
            for (ox = -1; ox <= 1; ++ox) {
                int cx = cx_of[i] + ox, cid, e;
                if (cx < 0) cx += ncx;
                else if (cx >= ncx) cx -= ncx;
                if (cx < 0 || cx >= ncx) continue;
                cid = cy * ncx + cx;
                for (e = cell_start[cid + 1] - 1; e >= cell_start[cid]; --e) {
                    int j = sorted[e].particle;
                    double xw, ddx, ddy, r2, rr, W, dW;
                    if (j == i || seen[j] == i) continue;
                    if (j < n_fluid && j < i) continue;
                    xw = x[i] - floor(x[i] / DL) * DL;
                    ddx = xw - sorted[e].x;
                    if (ddx > 0.5 * DL) ddx -= DL;
                    else if (ddx < -0.5 * DL) ddx += DL;
                    ddy = y[i] - sorted[e].y;
                    r2 = ddx * ddx + ddy * ddy;
                    if (!(r2 > 1e-24 && r2 < rc2)) continue;
                    rr = sqrt(r2);
                    spline2d(rr, h, &W, &dW);
                    if (W > 0.0 || fabs(dW) > 0.0) {
                        pairs_push(out, &cap, i, j, ddx, ddy, rr, W, dW);
                        seen[j] = i;
                    }
                }
            }
        }
    }
    free(ent);
    free(sorted);
    free(cell_start);
    free(cx_of);
    free(cy_of);
    free(seen);
    return 0;
}

void orc_pairs_free(orc_pairs *p)
{
    free(p->pair_i);
    free(p->pair_j);
    free(p->dx);
    free(p->dy);
    free(p->r);
    free(p->W);
    free(p->dW);
    memset(p, 0, sizeof(*p));
}

/* pair k -> 0-based (ii, jj); returns 0 when the reference would silently skip it (e.g. ph:193) */
static int pair_ids(const double *pair_i, const double *pair_j, size_t k, int n_fluid, int n_total, int *ii,
                    int *jj)
{
    *ii = (int)pair_i[k] - 1;
    *jj = (int)pair_j[k] - 1;
    return !(*ii < 0 || *ii >= n_fluid || *jj < 0 || *jj >= n_total);
}

/*
 * density_correction.  Follows ph:95-374: sigma sums ph:188-212 (no r test in this first loop),
 * rho/Vol ph:214-234, A matrix ph:239-312, regularised pseudo-inverse + blend ph:321-366.
 */
void orc_density_correction(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                            const double *dy, const double *r, const double *W, const double *dW,
                            const double *mass, int n_fluid, int n_total, double rho0, double h,
                            double inv_sigma0, double *rho, double *Vol, double *B)
{
    const double eps_reg = 1e-8; /* ph:30 */
    const double W0 = 10.0 / (7.0 * ORC_PI * h * h);
    double *scr = (double *)calloc((size_t)n_fluid * 6, sizeof(double));
    double *s_in = scr, *s_ct = scr + n_fluid;
    double *A11 = scr + 2 * (size_t)n_fluid, *A12 = scr + 3 * (size_t)n_fluid;
    double *A21 = scr + 4 * (size_t)n_fluid, *A22 = scr + 5 * (size_t)n_fluid;
    size_t k;
    int i, ii, jj;

    for (i = 0; i < n_fluid; ++i) s_in[i] = W0;
    for (k = 0; k < n_pairs; ++k) {
        if (!pair_ids(pair_i, pair_j, k, n_fluid, n_total, &ii, &jj)) continue;
        if (jj < n_fluid) {
            s_in[ii] += W[k];
            s_in[jj] += W[k];
        } else {
            s_ct[ii] += W[k] * (mass[jj] / rho0);
        }
    }
    for (i = 0; i < n_total; ++i) rho[i] = rho0;
    for (i = 0; i < n_fluid; ++i) {
        double v = s_in[i] * rho0 * inv_sigma0;
        v += s_ct[i] * rho0 * rho0 * inv_sigma0 / mass[i];
        if (v <= 1e-12) v = rho0;
        rho[i] = v;
    }
    for (i = 0; i < n_total; ++i) {
        if (rho[i] <= 1e-12) rho[i] = rho0;
        Vol[i] = mass[i] / rho[i];
    }
    for (k = 0; k < n_pairs; ++k) {
        double ex, ey, fj;
        if (!pair_ids(pair_i, pair_j, k, n_fluid, n_total, &ii, &jj) || r[k] <= 1e-12) continue;
        ex = dx[k] / r[k];
        ey = dy[k] / r[k];
        fj = dW[k] * Vol[jj];
        A11[ii] -= dx[k] * (fj * ex);
        A12[ii] -= dx[k] * (fj * ey);
        A21[ii] -= dy[k] * (fj * ex);
        A22[ii] -= dy[k] * (fj * ey);
        if (jj < n_fluid) {
            double fi = dW[k] * Vol[ii];
            A11[jj] -= dx[k] * (fi * ex);
            A12[jj] -= dx[k] * (fi * ey);
            A21[jj] -= dy[k] * (fi * ex);
            A22[jj] -= dy[k] * (fi * ey);
        }
    }
    for (i = 0; i < n_total; ++i) {
        B[i] = 1.0;
        B[i + n_total] = 0.0;
        B[i + 2 * (size_t)n_total] = 0.0;
        B[i + 3 * (size_t)n_total] = 1.0;
    }
    for (i = 0; i < n_fluid; ++i) {
        double a11 = A11[i], a12 = A12[i], a21 = A21[i], a22 = A22[i];
        double m11 = a11 * a11 + a21 * a21 + eps_reg;
        double m12 = a11 * a12 + a21 * a22;
        double m22 = a12 * a12 + a22 * a22 + eps_reg;
        double det_m = m11 * m22 - m12 * m12;
        double p11 = 1.0, p12 = 0.0, p21 = 0.0, p22 = 1.0;
        double det_a, lack, denom, w1, w2;
        if (!(fabs(det_m) < 1e-20)) {
            double i11 = m22 / det_m, i12 = -m12 / det_m, i22 = m11 / det_m;
            p11 = i11 * a11 + i12 * a12;
            p12 = i11 * a21 + i12 * a22;
            p21 = i12 * a11 + i22 * a12;
            p22 = i12 * a21 + i22 * a22;
        }
        det_a = a11 * a22 - a12 * a21;
        lack = fmax(1.0 - det_a, 0.0);
        denom = det_a + lack;
        if (fabs(denom) < 1e-12) {
            w1 = 0.0;
            w2 = 1.0;
        } else {
            w1 = det_a / denom;
            w2 = lack / denom;
        }
        B[i] = w1 * p11 + w2;
        B[i + n_total] = w1 * p12;
        B[i + 2 * (size_t)n_total] = w1 * p21;
        B[i + 3 * (size_t)n_total] = w1 * p22 + w2;
    }
    free(scr);
}

/* viscous_force.  Follows ph:396-550 (pair loop ph:469-536, scaling ph:538-545). */
void orc_viscous_force(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                       const double *dy, const double *r, const double *dW, const double *vel,
                       const double *Vol, const double *B, double mu, double h, int n_fluid, int n_total,
                       const double *wall_vel, double *force)
{
    const double *vx = vel, *vy = vel + n_total, *wx = wall_vel, *wy = wall_vel + n_total;
    const double *B11 = B, *B12 = B + n_total, *B21 = B + 2 * (size_t)n_total, *B22 = B + 3 * (size_t)n_total;
    double *ax = (double *)calloc((size_t)n_total * 2, sizeof(double));
    double *ay = ax + n_total;
    size_t k;
    int i, ii, jj;

    for (k = 0; k < n_pairs; ++k) {
        double ex, ey, denom;
        if (!pair_ids(pair_i, pair_j, k, n_fluid, n_total, &ii, &jj) || r[k] <= 1e-12) continue;
        ex = dx[k] / r[k];
        ey = dy[k] / r[k];
        denom = r[k] + 0.01 * h;
        if (jj < n_fluid) {
            double s11 = B11[ii] + B11[jj], s12 = B12[ii] + B12[jj];
            double s21 = B21[ii] + B21[jj], s22 = B22[ii] + B22[jj];
            double eBe = ex * (s11 * ex + s12 * ey) + ey * (s21 * ex + s22 * ey);
            double dvx = vx[ii] - vx[jj], dvy = vy[ii] - vy[jj];
            double ci = eBe * mu * dW[k] * Vol[jj] / denom;
            double cj = eBe * mu * dW[k] * Vol[ii] / denom;
            ax[ii] += ci * dvx;
            ay[ii] += ci * dvy;
            ax[jj] -= cj * dvx;
            ay[jj] -= cj * dvy;
        } else {
            double eBe = ex * (B11[ii] * ex + B12[ii] * ey) + ey * (B21[ii] * ex + B22[ii] * ey);
            double dvx = vx[ii] - wx[jj], dvy = vy[ii] - wy[jj];
            double c = 4.0 * eBe * mu * dW[k] * Vol[jj] / denom;
            ax[ii] += c * dvx;
            ay[ii] += c * dvy;
        }
    }
    for (i = 0; i < n_fluid; ++i) {
        force[i] = ax[i] * Vol[i];
        force[i + n_total] = ay[i] * Vol[i];
    }
    for (i = n_fluid; i < n_total; ++i) {
        force[i] = 0.0;
        force[i + n_total] = 0.0;
    }
    free(ax);
}

/* transport_correction.  Follows ph:569-714 (pair loop ph:636-700, limiter ph:702-710). */
void orc_transport_correction(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                              const double *dy, const double *r, const double *dW, const double *Vol,
                              const double *B, const double *pos, double h, int n_fluid, int n_total,
                              double transport_coeff, double *pos_out)
{
    const double *B11 = B, *B12 = B + n_total, *B21 = B + 2 * (size_t)n_total, *B22 = B + 3 * (size_t)n_total;
    double *incx = (double *)calloc((size_t)n_total * 2, sizeof(double));
    double *incy = incx + n_total;
    size_t k;
    int i, ii, jj;

    memcpy(pos_out, pos, (size_t)n_total * 2 * sizeof(double));
    for (k = 0; k < n_pairs; ++k) {
        double ex, ey;
        if (!pair_ids(pair_i, pair_j, k, n_fluid, n_total, &ii, &jj) || r[k] <= 1e-12) continue;
        ex = dx[k] / r[k];
        ey = dy[k] / r[k];
        if (jj < n_fluid) {
            double s11 = B11[ii] + B11[jj], s12 = B12[ii] + B12[jj];
            double s21 = B21[ii] + B21[jj], s22 = B22[ii] + B22[jj];
            double tx = s11 * ex + s12 * ey, ty = s21 * ex + s22 * ey;
            double ci = -dW[k] * Vol[jj], cj = dW[k] * Vol[ii];
            incx[ii] += ci * tx;
            incy[ii] += ci * ty;
            incx[jj] += cj * tx;
            incy[jj] += cj * ty;
        } else {
            double tx = B11[ii] * ex + B12[ii] * ey, ty = B21[ii] * ex + B22[ii] * ey;
            double c = -2.0 * dW[k] * Vol[jj];
            incx[ii] += c * tx;
            incy[ii] += c * ty;
        }
    }
    for (i = 0; i < n_fluid; ++i) {
        double n2 = incx[i] * incx[i] + incy[i] * incy[i];
        double lim = 100.0 * n2 / (h * h);
        double scale = transport_coeff * h * h;
        if (lim > 1.0) lim = 1.0;
        if (lim < 0.0) lim = 0.0;
        pos_out[i] += scale * lim * incx[i];
        pos_out[i + n_total] += scale * lim * incy[i];
    }
    free(incx);
}

static double dissipation_limiter(double un_l, double un_r, double c_f) /* riemann_beta, ph:1121-1129 */
{
    double comp = un_l - un_r;
    if (comp < 0.0) comp = 0.0;
    return fmin(3.0 * comp, c_f);
}

/*
 * integration_1st.  Follows ph:736-967: half-step density / EOS / position ph:857-865, pressure
 * pair loop ph:870-951 (fluid-fluid ph:884-925, virtual wall pressure ph:926-950), scaling ph:953-962.
 */
void orc_integration_1st(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                         const double *dy, const double *r, const double *dW, const double *Vol,
                         const double *B, const double *rho, const double *mass, const double *pos,
                         const double *vel, const double *drho_dt, const double *force_prior, double dt,
                         int n_fluid, int n_total, double rho0, double p0, double c_f,
                         double *rho_out, double *p_out, double *pos_out, double *force_out, double *drho_out)
{
    const double *vx = vel, *vy = vel + n_total;
    const double *fpx = force_prior, *fpy = force_prior + n_total;
    const double *B11 = B, *B12 = B + n_total, *B21 = B + 2 * (size_t)n_total, *B22 = B + 3 * (size_t)n_total;
    double *fx = force_out, *fy = force_out + n_total;
    double *diss = (double *)calloc((size_t)n_total, sizeof(double));
    size_t k;
    int i, ii, jj;

    memcpy(rho_out, rho, (size_t)n_total * sizeof(double));
    memcpy(pos_out, pos, (size_t)n_total * 2 * sizeof(double));
    memset(p_out, 0, (size_t)n_total * sizeof(double));
    memset(force_out, 0, (size_t)n_total * 2 * sizeof(double));
    memset(drho_out, 0, (size_t)n_total * sizeof(double));

    for (i = 0; i < n_fluid; ++i) {
        rho_out[i] = rho_out[i] + 0.5 * dt * drho_dt[i];
        if (rho_out[i] < 1e-10) rho_out[i] = rho0;
        p_out[i] = p0 * (rho_out[i] / rho0 - 1.0);
        pos_out[i] += 0.5 * dt * vx[i];
        pos_out[i + n_total] += 0.5 * dt * vy[i];
    }
    for (k = 0; k < n_pairs; ++k) {
        double ex, ey;
        if (!pair_ids(pair_i, pair_j, k, n_fluid, n_total, &ii, &jj) || r[k] <= 1e-12) continue;
        ex = dx[k] / r[k];
        ey = dy[k] / r[k];
        if (jj < n_fluid) {
            double pi_ = p_out[ii], pj_ = p_out[jj];
            double rho_bar = 0.5 * (rho_out[ii] + rho_out[jj]);
            double ul = vx[ii] * ex + vy[ii] * ey;
            double ur = vx[jj] * ex + vy[jj] * ey;
            double beta = dissipation_limiter(ul, ur, c_f);
            double p_star = 0.5 * (pi_ + pj_) + 0.5 * beta * rho_bar * (ul - ur);
            double p_face = 0.5 * (0.5 * (pi_ + pj_) + p_star);
            double tx = p_face * ((B11[ii] + B11[jj]) * ex + (B12[ii] + B12[jj]) * ey);
            double ty = p_face * ((B21[ii] + B21[jj]) * ex + (B22[ii] + B22[jj]) * ey);
            double gj = dW[k] * Vol[jj], gi = dW[k] * Vol[ii];
            double pd = pi_ - pj_;
            fx[ii] -= tx * gj;
            fy[ii] -= ty * gj;
            fx[jj] += tx * gi;
            fy[jj] += ty * gi;
            diss[ii] += (pd / (rho0 * c_f)) * gj;
            diss[jj] += (-pd / (rho0 * c_f)) * gi;
        } else {
            double pi_ = p_out[ii], rho_i = rho_out[ii];
            double gj = dW[k] * Vol[jj];
            double ax = fpx[ii] / mass[ii], ay = fpy[ii] / mass[ii];
            double ext = -(ax * ex + ay * ey);
            double pw = pi_ + rho_i * r[k] * fmax(0.0, ext);
            double tx = B11[ii] * ex + B12[ii] * ey, ty = B21[ii] * ex + B22[ii] * ey;
            fx[ii] -= (pi_ + pw) * gj * tx;
            fy[ii] -= (pi_ + pw) * gj * ty;
            diss[ii] += ((pi_ - pw) / (rho0 * c_f)) * gj;
        }
    }
    for (i = 0; i < n_fluid; ++i) {
        fx[i] *= Vol[i];
        fy[i] *= Vol[i];
        drho_out[i] = diss[i] * rho_out[i];
    }
    for (i = n_fluid; i < n_total; ++i) {
        fx[i] = 0.0;
        fy[i] = 0.0;
        drho_out[i] = 0.0;
    }
    free(diss);
}

/* integration_2nd.  Follows ph:987-1119 (position ph:1066-1069, continuity loop ph:1076-1109). */
void orc_integration_2nd(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                         const double *dy, const double *r, const double *dW, const double *Vol,
                         const double *rho, const double *pos, const double *vel, double dt, int n_fluid,
                         int n_total, const double *wall_vel, double *pos_out, double *drho_out)
{
    const double *vx = vel, *vy = vel + n_total, *wx = wall_vel, *wy = wall_vel + n_total;
    double *rate = (double *)calloc((size_t)n_total, sizeof(double));
    size_t k;
    int i, ii, jj;

    memcpy(pos_out, pos, (size_t)n_total * 2 * sizeof(double));
    for (i = 0; i < n_fluid; ++i) {
        pos_out[i] += 0.5 * dt * vx[i];
        pos_out[i + n_total] += 0.5 * dt * vy[i];
    }
    for (k = 0; k < n_pairs; ++k) {
        double ex, ey;
        if (!pair_ids(pair_i, pair_j, k, n_fluid, n_total, &ii, &jj) || r[k] <= 1e-12) continue;
        ex = dx[k] / r[k];
        ey = dy[k] / r[k];
        if (jj < n_fluid) {
            double u = (vx[ii] - vx[jj]) * ex + (vy[ii] - vy[jj]) * ey;
            rate[ii] += u * dW[k] * Vol[jj];
            rate[jj] += u * dW[k] * Vol[ii];
        } else {
            double mx = 2.0 * wx[jj] - vx[ii], my = 2.0 * wy[jj] - vy[ii];
            double u = (vx[ii] - mx) * ex + (vy[ii] - my) * ey;
            rate[ii] += u * dW[k] * Vol[jj];
        }
    }
    for (i = 0; i < n_fluid; ++i) drho_out[i] = rate[i] * rho[i];
    for (i = n_fluid; i < n_total; ++i) drho_out[i] = 0.0;
    free(rate);
}

/*
 * integration_verlet.  Follows ph:1316-1469: 1st stage ph:1386, velocity kick ph:1400-1408,
 * 2nd stage on (rho_half, x_half, v_new) ph:1411-1427, density/EOS completion ph:1440-1450.
 * The dissipation rate of the 1st stage is discarded (ph:1463).
 */
void orc_integration_verlet(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                            const double *dy, const double *r, const double *dW, const double *Vol,
                            const double *B, const double *rho, const double *mass, const double *pos,
                            const double *vel, const double *drho_dt, const double *force_prior, double dt,
                            int n_fluid, int n_total, double rho0, double p0, double c_f,
                            const double *wall_vel, double *rho_out, double *p_out, double *pos_out,
                            double *vel_out, double *drho_out, double *force_out)
{
    size_t n = (size_t)n_total;
    double *rho_h = (double *)malloc(n * 7 * sizeof(double));
    double *p_h = rho_h + n, *pos_h = rho_h + 2 * n, *diss = rho_h + 4 * n;
    int i;

    orc_integration_1st(n_pairs, pair_i, pair_j, dx, dy, r, dW, Vol, B, rho, mass, pos, vel, drho_dt,
                        force_prior, dt, n_fluid, n_total, rho0, p0, c_f, rho_h, p_h, pos_h, force_out, diss);
    memcpy(vel_out, vel, n * 2 * sizeof(double));
    for (i = 0; i < n_fluid; ++i) {
        double inv_m = 1.0 / mass[i];
        vel_out[i] += (force_prior[i] + force_out[i]) * inv_m * dt;
        vel_out[i + n] += (force_prior[i + n] + force_out[i + n]) * inv_m * dt;
    }
    for (i = n_fluid; i < n_total; ++i) {
        vel_out[i] = 0.0;
        vel_out[i + n] = 0.0;
    }
    orc_integration_2nd(n_pairs, pair_i, pair_j, dx, dy, r, dW, Vol, rho_h, pos_h, vel_out, dt, n_fluid,
                        n_total, wall_vel, pos_out, drho_out);
    memcpy(rho_out, rho_h, n * sizeof(double));
    for (i = 0; i < n_fluid; ++i) {
        rho_out[i] += drho_out[i] * (0.5 * dt);
        if (rho_out[i] < 1e-10) rho_out[i] = rho0;
        p_out[i] = p0 * (rho_out[i] / rho0 - 1.0);
    }
    for (i = n_fluid; i < n_total; ++i) p_out[i] = 0.0;
    free(rho_h);
}

/*
 * advance_shell_step.  Follows ph:1490-1639: density ph:1554 -> viscous ph:1573 -> + m*g on x
 * ph:1575-1580 -> transport with the MEX default coefficient 0.2 (13-argument call ph:1596, default
 * ph:584) -> verlet ph:1621.  The incoming rho is never used (ph:1532 only size-checks it).
 */
void orc_advance_shell_step(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                            const double *dy, const double *r, const double *W, const double *dW,
                            const double *mass, const double *pos, const double *vel, const double *wall_vel,
                            const double *drho_dt, double dt, int n_fluid, int n_total, double rho0, double p0,
                            double c_f, double mu, double h, double inv_sigma0, double gravity_g,
                            double *rho_out, double *p_out, double *pos_out, double *vel_out, double *drho_out,
                            double *force_out, double *force_prior_out, double *Vol_out, double *B_out)
{
    size_t n = (size_t)n_total;
    double *rho_d = (double *)malloc(n * 3 * sizeof(double));
    double *pos_t = rho_d + n;
    int i;

    orc_density_correction(n_pairs, pair_i, pair_j, dx, dy, r, W, dW, mass, n_fluid, n_total, rho0, h,
                           inv_sigma0, rho_d, Vol_out, B_out);
    orc_viscous_force(n_pairs, pair_i, pair_j, dx, dy, r, dW, vel, Vol_out, B_out, mu, h, n_fluid, n_total,
                      wall_vel, force_prior_out);
    for (i = 0; i < n_fluid; ++i) force_prior_out[i] += mass[i] * gravity_g;
    orc_transport_correction(n_pairs, pair_i, pair_j, dx, dy, r, dW, Vol_out, B_out, pos, h, n_fluid, n_total,
                             0.2, pos_t);
    orc_integration_verlet(n_pairs, pair_i, pair_j, dx, dy, r, dW, Vol_out, B_out, rho_d, mass, pos_t, vel,
                           drho_dt, force_prior_out, dt, n_fluid, n_total, rho0, p0, c_f, wall_vel, rho_out,
                           p_out, pos_out, vel_out, drho_out, force_out);
    free(rho_d);
}

/* wall_shear_monitor.  Follows ph:1653-1743 (serial loop ph:1713-1739, result ph:1741-1742). */
void orc_wall_shear_monitor(size_t n_pairs, const double *pair_i, const double *pair_j, const double *dx,
                            const double *dy, const double *r, const double *dW, const double *pos,
                            const double *vel, const double *wall_vel, const double *Vol, const double *B,
                            int n_fluid, int n_total, double DL, double DH, double mu, double h,
                            double *tau_bottom, double *tau_top)
{
    const double *py = pos + n_total, *vx = vel, *wx = wall_vel;
    const double *B11 = B, *B12 = B + n_total, *B21 = B + 2 * (size_t)n_total, *B22 = B + 3 * (size_t)n_total;
    double sb = 0.0, st = 0.0;
    size_t k;

    for (k = 0; k < n_pairs; ++k) {
        int ii = (int)pair_i[k] - 1, jj = (int)pair_j[k] - 1;
        double ex, ey, eBe, dvx, f;
        if (ii < 0 || ii >= n_fluid || jj < n_fluid || r[k] <= 1e-12) continue; /* ph:1722 */
        ex = dx[k] / r[k];
        ey = dy[k] / r[k];
        eBe = ex * (B11[ii] * ex + B12[ii] * ey) + ey * (B21[ii] * ex + B22[ii] * ey);
        dvx = vx[ii] - wx[jj];
        f = 4.0 * mu * eBe * dW[k] * Vol[jj] * dvx / (r[k] + 0.01 * h) * Vol[ii];
        if (py[jj] <= 0.0) sb += f;
        else if (py[jj] >= DH) st += f;
    }
    *tau_bottom = -sb / DL;
    *tau_top = -st / DL;
}

/* verlet_time_step.  Follows m:519-527 (v_max over fluid rows only, m:259). */
double orc_verlet_time_step(const double *vel, int n_fluid, int n_total, double c_max, double h, double nu,
                            double gravity_g, double remain)
{
    double vmax = 0.0, dt_a, dt_v, dt_b, dt;
    int i;
    for (i = 0; i < n_fluid; ++i) {
        double s = sqrt(vel[i] * vel[i] + vel[i + n_total] * vel[i + n_total]);
        if (s > vmax) vmax = s;
    }
    dt_a = 0.25 * h / fmax(c_max + vmax, 1e-12);
    dt_v = 0.125 * h * h / fmax(nu, 1e-12);
    dt_b = 0.25 * sqrt(h / fmax(fabs(gravity_g), 1e-12));
    dt = fmin(fmin(dt_a, dt_v), fmin(dt_b, remain));
    return fmax(dt, 1e-12);
}

/* periodic_bounding.  Follows m:570-577: x <- mod(x, DL) on fluid rows (MATLAB mod: result in [0,DL)). */
void orc_periodic_wrap(double *pos, int n_fluid, int n_total, double DL)
{
    int i;
    (void)n_total;
    for (i = 0; i < n_fluid; ++i) {
        double m = fmod(pos[i], DL);
        if (m != 0.0 && m < 0.0) m += DL;
        pos[i] = m;
    }
}

/*
 * sort_particles_by_cell / sort_subset_indices.  Follows m:529-568: fluid rows only, key (cy, cx) with
 * cx = floor(mod(x,DL)/2h), cy = floor((y - min fluid y)/2h), sortrows is stable; wall rows keep place.
 */
void orc_sort_permutation(const double *pos, int n_fluid, int n_total, double DL, double h, int *perm)
{
    const double cell = 2.0 * h;
    double y0;
    int i, max_cx = 0, max_cy = 0, ncx, nkeys;
    int *cx, *cy, *start;

    for (i = 0; i < n_total; ++i) perm[i] = i;
    if (n_fluid <= 0 || n_fluid >= n_total) return; /* m:535-537 */
    cx = (int *)malloc((size_t)n_fluid * 2 * sizeof(int));
    cy = cx + n_fluid;
    y0 = pos[n_total];
    for (i = 1; i < n_fluid; ++i)
        if (pos[i + n_total] < y0) y0 = pos[i + n_total];
    for (i = 0; i < n_fluid; ++i) {
        double m = fmod(pos[i], DL);
        if (m != 0.0 && m < 0.0) m += DL;
        cx[i] = (int)floor(m / cell);
        cy[i] = (int)floor((pos[i + n_total] - y0) / cell);
        if (cx[i] > max_cx) max_cx = cx[i];
        if (cy[i] > max_cy) max_cy = cy[i];
    }
    ncx = max_cx + 1;
    nkeys = ncx * (max_cy + 1);
    start = (int *)calloc((size_t)nkeys + 1, sizeof(int));
    for (i = 0; i < n_fluid; ++i) start[cy[i] * ncx + cx[i] + 1]++;
    for (i = 0; i < nkeys; ++i) start[i + 1] += start[i];
    for (i = 0; i < n_fluid; ++i) perm[start[cy[i] * ncx + cx[i]]++] = i;
    free(start);
    free(cx);
}
