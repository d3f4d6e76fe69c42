// neighbor.cu -- the cell-linked list rebuilt as a device counting sort, emitted as CSR (rows N1-N6 of
// SURVEY.md 8(a); reference: mex/sph_neighbor_search_mex.c, cited as nb:line).
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false.  Every floating-point expression that decides
// cell membership, ghost insertion or pair acceptance is written with the reference's grouping and is
// evaluated with separate IEEE multiply/add (fp64 div, sqrt, floor, ceil are correctly rounded on the
// device), so the neighbour SETS and the exported dx,dy,r,W,dW are bit-identical to the reference's.
//
// Pipeline (all kernels hand-written; prefix sums in scan.cu):
//   k_yrange        y_min/y_max over ALL particles (nb:245-250), atomic min/max on ordered integer keys
//   k_grid_setup    n_cell_y and cell count (nb:253-258)
//   k_bin_count     wrap, bin, count real + ghost entries per cell (nb:270-296) + position snapshot
//   scan            cell_start
//   k_scatter       entry records into their cell ranges (atomic slot), then
//   k_cell_canon    each cell's range sorted by insertion key: the reference's linked list is walked
//                   "last inserted first", i.e. in DEcreasing (particle, real<ghost-<ghost+) order
//   k_half_walk<0>  per fluid particle: 3x3 walk in the reference's order with its skip rules, dedup of
//                   real/ghost duplicates (seen_neighbor, nb:342,383) -> counts for both lists
//   scan x2         half_ptr (the reference's half-pair list as CSR-by-i) and row_ptr (full CSR)
//   k_half_walk<1>  same walk, fills half_col and scatters both directions into the full CSR
//   k_row_sort      full rows sorted by j (deterministic order, monotone gathers)
//   k_export_pairs  half list -> the 7 MEX output columns (nb:375-381), same arithmetic as the walk
#include "common.cuh"

#include <cmath>
#include <climits>

namespace sphb {

namespace {

constexpr unsigned kIdMask = 0x3fffffffu;
constexpr int kSeenCap = 64;

struct NbrParams {
    int n_fluid, n_total;
    int ncx;
    int cell_cap;
    double h, DL;
    double cell;      // 2.0 * h                         nb:253
    double cutoff;    // r_cut * h = 2.0 * h             nb:264
    double rc2;       // (r_cut*h) * (r_cut*h)           nb:306
    double sigma;     // 10 / (7 pi h h)                 nb:119
    size_t half_cap, csr_cap;
};

__device__ __forceinline__ long long ordered_key(double d)
{
    long long b = __double_as_longlong(d);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double from_ordered_key(long long k)
{
    return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}

__device__ __forceinline__ int wrap_col(int c, int ncx)   // nb:57-67
{
    c %= ncx;
    if (c < 0) c += ncx;
    return c;
}

// nb:116-133, same expression grouping (this TU has FMA contraction off)
__device__ __forceinline__ void spline2d(double r, double h, double sigma, double &W, double &dW)
{
    const double q = r / h;
    if (q < 1.0) {
        W = sigma * (1.0 - 1.5 * q * q + 0.75 * q * q * q);
        dW = sigma * (-3.0 * q + 2.25 * q * q) / h;
    } else if (q < 2.0) {
        const double t = 2.0 - q;
        W = sigma * 0.25 * t * t * t;
        dW = -sigma * 0.75 * t * t / h;
    } else {
        W = 0.0;
        dW = 0.0;
    }
}

__global__ void k_grid_init(GridDesc *g)
{
    g->ymin_key = LLONG_MAX;
    g->ymax_key = LLONG_MIN;
    g->n_entries = 0;
}

__global__ void __launch_bounds__(256) k_yrange(const double2 *__restrict__ pos, int n, GridDesc *g)
{
    long long kmin = LLONG_MAX, kmax = LLONG_MIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        long long k = ordered_key(pos[i].y);
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        long long a = __shfl_xor_sync(0xffffffffu, kmin, d);
        long long b = __shfl_xor_sync(0xffffffffu, kmax, d);
        kmin = a < kmin ? a : kmin;
        kmax = b > kmax ? b : kmax;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&g->ymin_key, kmin);
        atomicMax(&g->ymax_key, kmax);
    }
}

__global__ void k_grid_setup(GridDesc *g, NbrParams P, int *flags)
{
    const double y_min = from_ordered_key(g->ymin_key);
    const double y_max = from_ordered_key(g->ymax_key);
    int ncy = (int)ceil((y_max - y_min + 1e-12) / P.cell) + 1;   // nb:256
    if (ncy < 1) ncy = 1;
    if ((long long)ncy * P.ncx > (long long)P.cell_cap) {
        atomicOr(flags, kErrCellCap);
        ncy = P.cell_cap / P.ncx;
        if (ncy < 1) ncy = 1;
    }
    g->y_min = y_min;
    g->y_max = y_max;
    g->cell = P.cell;
    g->ncx = P.ncx;
    g->ncy = ncy;
    g->ncells = ncy * P.ncx;
}

struct Bins {
    double xw, y;
    int cx, cy;
    int gcx_lo, gcx_hi;   // ghost columns, -1 when that ghost entry is not inserted
};

// nb:270-295 for one particle
__device__ __forceinline__ Bins bin_particle(double2 p, const NbrParams &P, double y_min, int ncy)
{
    Bins b;
    b.xw = p.x - floor(p.x / P.DL) * P.DL;
    b.y = p.y;
    b.cx = wrap_col((int)floor(b.xw / P.cell), P.ncx);
    int cy = (int)floor((p.y - y_min) / P.cell);
    if (cy < 0) cy = 0;
    if (cy >= ncy) cy = ncy - 1;
    b.cy = cy;
    b.gcx_lo = -1;
    b.gcx_hi = -1;
    if (b.xw > P.DL - P.cutoff) {
        const double gx = b.xw - P.DL;
        const int g = wrap_col((int)floor(gx / P.cell), P.ncx);
        if (g != b.cx) b.gcx_lo = g;
    }
    if (b.xw < P.cutoff) {
        const double gx = b.xw + P.DL;
        const int g = wrap_col((int)floor(gx / P.cell), P.ncx);
        if (g != b.cx) b.gcx_hi = g;
    }
    return b;
}

__global__ void __launch_bounds__(kBlock) k_bin_count(const double2 *__restrict__ pos, NbrParams P,
                                                     const GridDesc *__restrict__ g, double2 *__restrict__ snap,
                                                     int *__restrict__ cell_of, int *__restrict__ cell_cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_total) return;
    const Bins b = bin_particle(pos[i], P, g->y_min, g->ncy);
    snap[i] = make_double2(b.xw, b.y);
    const int row = b.cy * P.ncx;
    cell_of[i] = row + b.cx;
    atomicAdd(&cell_cnt[row + b.cx], 1);
    if (b.gcx_lo >= 0) atomicAdd(&cell_cnt[row + b.gcx_lo], 1);
    if (b.gcx_hi >= 0) atomicAdd(&cell_cnt[row + b.gcx_hi], 1);
}

__global__ void __launch_bounds__(kBlock) k_scatter(const double2 *__restrict__ pos, NbrParams P,
                                                   const GridDesc *__restrict__ g,
                                                   const int *__restrict__ cell_start, int *__restrict__ cell_fill,
                                                   int *__restrict__ ent_key, double2 *__restrict__ ent_pos)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_total) return;
    const Bins b = bin_particle(pos[i], P, g->y_min, g->ncy);
    const int row = b.cy * P.ncx;
    const int multi = (b.gcx_lo >= 0 || b.gcx_hi >= 0) ? 4 : 0;
    {
        const int c = row + b.cx;
        const int s = cell_start[c] + atomicAdd(&cell_fill[c], 1);
        ent_key[s] = (i << 3) | multi | 0;
        ent_pos[s] = make_double2(b.xw, b.y);
    }
    if (b.gcx_lo >= 0) {
        const int c = row + b.gcx_lo;
        const int s = cell_start[c] + atomicAdd(&cell_fill[c], 1);
        ent_key[s] = (i << 3) | multi | 1;
        ent_pos[s] = make_double2(b.xw - P.DL, b.y);
    }
    if (b.gcx_hi >= 0) {
        const int c = row + b.gcx_hi;
        const int s = cell_start[c] + atomicAdd(&cell_fill[c], 1);
        ent_key[s] = (i << 3) | multi | 2;
        ent_pos[s] = make_double2(b.xw + P.DL, b.y);
    }
}

// insertion sort of one cell's entries by key (ascending = the reference's insertion order)
__global__ void __launch_bounds__(kBlock) k_cell_canon(const GridDesc *__restrict__ g,
                                                      const int *__restrict__ cell_start,
                                                      int *__restrict__ ent_key, double2 *__restrict__ ent_pos)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g->ncells) return;
    const int s = cell_start[c], e = cell_start[c + 1];
    for (int a = s + 1; a < e; ++a) {
        const int key = ent_key[a];
        const double2 p = ent_pos[a];
        int b = a - 1;
        while (b >= s && ent_key[b] > key) {
            ent_key[b + 1] = ent_key[b];
            ent_pos[b + 1] = ent_pos[b];
            --b;
        }
        ent_key[b + 1] = key;
        ent_pos[b + 1] = p;
    }
}

// The 3x3 walk of nb:312-392 for fluid particle i.  FILL = false: count; FILL = true: write lists.
template <bool FILL>
__global__ void __launch_bounds__(kBlock) k_half_walk(NbrParams P, const GridDesc *__restrict__ g,
                                                     const double2 *__restrict__ snap,
                                                     const int *__restrict__ cell_of,
                                                     const int *__restrict__ cell_start,
                                                     const int *__restrict__ ent_key,
                                                     const double2 *__restrict__ ent_pos,
                                                     int *__restrict__ half_cnt, int *__restrict__ row_cnt,
                                                     const int *__restrict__ half_ptr,
                                                     const int *__restrict__ row_ptr, int *__restrict__ row_fill,
                                                     unsigned *__restrict__ half_col, unsigned *__restrict__ col,
                                                     int want_csr, int *__restrict__ flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_fluid) return;
    const int ncx = P.ncx, ncy = g->ncy;
    const double2 pi = snap[i];
    const int home = cell_of[i];
    const int cxi = home % ncx, cyi = home / ncx;
    const double half_DL = 0.5 * P.DL;
    int seen[kSeenCap];
    int nseen = 0;
    int count = 0;
    size_t hbase = 0;
    if (FILL) hbase = (size_t)half_ptr[i];

    for (int oy = -1; oy <= 1; ++oy) {
        const int cy = cyi + oy;
        if (cy < 0 || cy >= ncy) continue;
        int cx_prev[2] = {-1, -1};
        for (int ox = -1; ox <= 1; ++ox) {
            int cx = cxi + ox;
            if (cx < 0) cx += ncx;
            else if (cx >= ncx) cx -= ncx;
            if (cx < 0 || cx >= ncx) continue;
            // n_cell_x <= 2: the wrapped stencil revisits a column.  The reference re-walks it, but
            // every entry it accepted is already in seen_neighbor and every entry it rejected is
            // rejected again by the same arithmetic (nb:342-372), so the revisit emits nothing.
            if (cx == cx_prev[0] || cx == cx_prev[1]) continue;
            cx_prev[ox == -1 ? 0 : 1] = cx;
            const int cid = cy * ncx + cx;
            const int s = cell_start[cid];
            for (int e = cell_start[cid + 1] - 1; e >= s; --e) {
                const int key = ent_key[e];
                const int j = key >> 3;
                if (j == i) continue;
                if (key & 4) {   // particle j owns ghost entries: it may already have been accepted
                    bool dup = false;
                    for (int q = 0; q < nseen; ++q) dup |= (seen[q] == j);
                    if (dup) continue;
                }
                if (j < P.n_fluid && j < i) continue;   // fluid-fluid pairs stored once, i < j (nb:353)
                const double2 pe = ent_pos[e];
                double dx = pi.x - pe.x;
                int m = 0;                               // minimum image (nb:359-363)
                if (dx > half_DL) {
                    dx -= P.DL;
                    m = -1;
                } else if (dx < -half_DL) {
                    dx += P.DL;
                    m = 1;
                }
                const double dy = pi.y - pe.y;
                const double r2 = dx * dx + dy * dy;
                if (!(r2 > 1e-24 && r2 < P.rc2)) continue;
                const double r = sqrt(r2);
                double W, dW;
                spline2d(r, P.h, P.sigma, W, dW);
                if (!(W > 0.0 || fabs(dW) > 0.0)) continue;
                // ---- accepted ----
                const int kind = key & 3;
                if (key & 4) {
                    if (nseen < kSeenCap) seen[nseen++] = j;
                    else atomicOr(flags, kErrSeenCap);
                }
                if (FILL) {
                    const size_t hp = hbase + count;
                    if (hp < P.half_cap) half_col[hp] = (unsigned)j | ((unsigned)kind << 30);
                    if (want_csr) {
                        // image shift of x_j as seen from i: ghost kind 1 sits at x_j - DL, kind 2 at x_j + DL
                        const int sh = m + (kind == 1 ? 1 : (kind == 2 ? -1 : 0));
                        const size_t a = (size_t)row_ptr[i] + atomicAdd(&row_fill[i], 1);
                        if (a < P.csr_cap) col[a] = (unsigned)j | ((unsigned)(sh + 1) << 30);
                        if (j < P.n_fluid) {
                            const size_t b = (size_t)row_ptr[j] + atomicAdd(&row_fill[j], 1);
                            if (b < P.csr_cap) col[b] = (unsigned)i | ((unsigned)(1 - sh) << 30);
                        }
                    }
                } else if (want_csr && j < P.n_fluid) {
                    atomicAdd(&row_cnt[j], 1);
                }
                ++count;
            }
        }
    }
    if (!FILL) {
        half_cnt[i] = count;
        if (want_csr) atomicAdd(&row_cnt[i], count);
    }
}

__global__ void k_check_caps(NbrParams P, const int *__restrict__ half_ptr, const int *__restrict__ row_ptr,
                             int want_csr, int *__restrict__ totals, int *__restrict__ flags)
{
    const int nh = half_ptr[P.n_fluid];
    const int nc = want_csr ? row_ptr[P.n_fluid] : 0;
    totals[0] = nh;
    totals[1] = nc;
    if ((size_t)nh > P.half_cap || (size_t)nc > P.csr_cap) atomicOr(flags, kErrCsrCap);
}

__global__ void __launch_bounds__(kBlock) k_row_sort(int n_fluid, const int *__restrict__ row_ptr, size_t csr_cap,
                                                    unsigned *__restrict__ col)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_fluid) return;
    const size_t s = (size_t)row_ptr[i];
    size_t e = (size_t)row_ptr[i + 1];
    if (e > csr_cap) e = csr_cap;
    for (size_t a = s + 1; a < e; ++a) {
        const unsigned v = col[a];
        const unsigned key = v & kIdMask;
        size_t b = a;
        while (b > s && (col[b - 1] & kIdMask) > key) {
            col[b] = col[b - 1];
            --b;
        }
        col[b] = v;
    }
}

// nb:357-381 re-evaluated for the stored (j, kind): identical operations, identical bits
__global__ void __launch_bounds__(kBlock) k_export_pairs(NbrParams P, const double2 *__restrict__ snap,
                                                        const int *__restrict__ half_ptr,
                                                        const unsigned *__restrict__ half_col,
                                                        double *__restrict__ pair_i, double *__restrict__ pair_j,
                                                        double *__restrict__ odx, double *__restrict__ ody,
                                                        double *__restrict__ orr, double *__restrict__ oW,
                                                        double *__restrict__ odW)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_fluid) return;
    const double2 pi = snap[i];
    const double half_DL = 0.5 * P.DL;
    const int s = half_ptr[i], e = half_ptr[i + 1];
    for (int k = s; k < e; ++k) {
        const unsigned v = half_col[k];
        const int j = (int)(v & kIdMask);
        const int kind = (int)(v >> 30);
        const double2 pj = snap[j];
        const double ex = kind == 0 ? pj.x : (kind == 1 ? pj.x - P.DL : pj.x + P.DL);
        double dx = pi.x - ex;
        if (dx > half_DL) dx -= P.DL;
        else if (dx < -half_DL) dx += P.DL;
        const double dy = pi.y - pj.y;
        const double r = sqrt(dx * dx + dy * dy);
        double W, dW;
        spline2d(r, P.h, P.sigma, W, dW);
        pair_i[k] = (double)(i + 1);
        pair_j[k] = (double)(j + 1);
        odx[k] = dx;
        ody[k] = dy;
        orr[k] = r;
        oW[k] = W;
        odW[k] = dW;
    }
}

NbrParams make_params(const NeighborWork &w)
{
    NbrParams P;
    const double pi = 3.14159265358979323846;
    P.n_fluid = w.n_fluid;
    P.n_total = w.n_total;
    P.h = w.h;
    P.DL = w.DL;
    P.cell = 2.0 * w.h;
    P.cutoff = 2.0 * w.h;
    P.rc2 = (2.0 * w.h) * (2.0 * w.h);
    P.sigma = 10.0 / (7.0 * pi * w.h * w.h);
    int ncx = (int)std::ceil(w.DL / P.cell);   // nb:254
    if (ncx < 1) ncx = 1;
    P.ncx = ncx;
    P.cell_cap = w.cell_cap;
    P.half_cap = w.half_cap;
    P.csr_cap = w.csr_cap;
    return P;
}

}  // namespace

void NeighborWork::release_all()
{
    DevBuf *all[] = {&grid, &flags, &snap, &cell_of, &cell_cnt, &cell_start, &cell_fill, &ent_key, &ent_pos,
                     &half_cnt, &half_ptr, &half_col, &row_cnt, &row_ptr, &row_fill, &col, &scan_tmp, &totals};
    for (DevBuf *b : all) b->release();
}

// Size every buffer whose size is known a priori.  y hints bound the cell rows (a fluid particle
// escaping beyond the hinted range raises the cell-capacity flag instead of writing out of bounds).
int neighbor_plan(NeighborWork &w, int n_fluid, int n_total, double h, double DL, double y_min_hint,
                  double y_max_hint, int max_nbr_hint)
{
    w.n_fluid = n_fluid;
    w.n_total = n_total;
    w.h = h;
    w.DL = DL;
    const double cell = 2.0 * h;
    const double ncx_d = std::ceil(DL / cell);
    const double ncy_d = std::ceil((y_max_hint - y_min_hint + 1e-12) / cell) + 1.0 + 8.0;   // slack rows
    if (!(ncx_d >= 1.0) || !(ncy_d >= 1.0) || ncx_d * ncy_d > 1.0e9)
        return set_error("SPH:Neighbor:entryCapacity", "cell grid of %g x %g cells is not representable",
                         ncx_d, ncy_d);
    if ((long long)n_total >= (1LL << 28))
        return set_error("SPH:Neighbor:entryCapacity", "n_total exceeds the 2^28 particle limit of the entry key");
    w.cell_cap = (int)(ncx_d * ncy_d);
    const size_t nt = (size_t)n_total, nf = (size_t)n_fluid;
    SPHB_TRY(w.grid.reserve(sizeof(GridDesc)));
    if (!w.flags.p) {
        SPHB_TRY(w.flags.reserve(sizeof(int)));
        SPHB_CUDA(cudaMemset(w.flags.p, 0, sizeof(int)));
    }
    SPHB_TRY(w.totals.reserve(4 * sizeof(int)));
    SPHB_TRY(w.snap.reserve(nt * sizeof(double2)));
    SPHB_TRY(w.cell_of.reserve(nt * sizeof(int)));
    SPHB_TRY(w.cell_cnt.reserve(((size_t)w.cell_cap + 1) * sizeof(int)));
    SPHB_TRY(w.cell_start.reserve(((size_t)w.cell_cap + 1) * sizeof(int)));
    SPHB_TRY(w.cell_fill.reserve(((size_t)w.cell_cap + 1) * sizeof(int)));
    SPHB_TRY(w.ent_key.reserve(3 * nt * sizeof(int)));
    SPHB_TRY(w.ent_pos.reserve(3 * nt * sizeof(double2)));
    SPHB_TRY(w.half_cnt.reserve((nf + 1) * sizeof(int)));
    SPHB_TRY(w.half_ptr.reserve((nf + 1) * sizeof(int)));
    SPHB_TRY(w.row_cnt.reserve((nf + 1) * sizeof(int)));
    SPHB_TRY(w.row_ptr.reserve((nf + 1) * sizeof(int)));
    SPHB_TRY(w.row_fill.reserve((nf + 1) * sizeof(int)));
    if (max_nbr_hint > 0) {
        const size_t csr = nf * (size_t)max_nbr_hint;
        const size_t half = nf * (size_t)((max_nbr_hint + 1) / 2 + 4);
        if (csr > w.csr_cap) {
            SPHB_TRY(w.col.reserve(csr * sizeof(unsigned)));
            w.csr_cap = csr;
        }
        if (half > w.half_cap) {
            SPHB_TRY(w.half_col.reserve(half * sizeof(unsigned)));
            w.half_cap = half;
        }
    }
    return 0;
}

// Builds the lists from `pos` (device, double2[n_total]).  When a capacity is exceeded the build stops
// after the counting pass, synchronises, grows the buffer and continues (stateless path and first
// resident build); in steady state no host synchronisation happens.
int neighbor_build(NeighborWork &w, const double2 *pos, bool want_full_csr, cudaStream_t st, long long *launches,
                   Prof *prof)
{
    NbrParams P = make_params(w);
    GridDesc *g = w.grid.as<GridDesc>();
    int *flags = w.flags.as<int>();
    const int nt = w.n_total, nf = w.n_fluid;
    long long nl = 0;

    {
        ProfScope ps(prof, kPfNbrGrid, st);
        k_grid_init<<<1, 1, 0, st>>>(g);
        int blocks = (nt + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_yrange<<<blocks, 256, 0, st>>>(pos, nt, g);
        k_grid_setup<<<1, 1, 0, st>>>(g, P, flags);
        SPHB_CUDA(cudaMemsetAsync(w.cell_cnt.p, 0, ((size_t)w.cell_cap + 1) * sizeof(int), st));
        SPHB_CUDA(cudaMemsetAsync(w.cell_fill.p, 0, ((size_t)w.cell_cap + 1) * sizeof(int), st));
    }
    {
        ProfScope ps(prof, kPfNbrBin, st);
        k_bin_count<<<grid_for(nt), kBlock, 0, st>>>(pos, P, g, w.snap.as<double2>(), w.cell_of.as<int>(),
                                                     w.cell_cnt.as<int>());
    }
    nl += 4;
    {
        ProfScope ps(prof, kPfNbrScan, st);
        SPHB_TRY(exclusive_scan_i32(w.cell_cnt.as<int>(), w.cell_start.as<int>(), (size_t)w.cell_cap, w.scan_tmp,
                                    &g->n_entries, st, &nl));
    }
    {
        ProfScope ps(prof, kPfNbrScatter, st);
        k_scatter<<<grid_for(nt), kBlock, 0, st>>>(pos, P, g, w.cell_start.as<int>(), w.cell_fill.as<int>(),
                                                   w.ent_key.as<int>(), w.ent_pos.as<double2>());
    }
    {
        ProfScope ps(prof, kPfNbrCanon, st);
        k_cell_canon<<<grid_for((size_t)w.cell_cap), kBlock, 0, st>>>(g, w.cell_start.as<int>(), w.ent_key.as<int>(),
                                                                      w.ent_pos.as<double2>());
    }
    nl += 2;

    {
        ProfScope ps(prof, kPfNbrWalkCount, st);
        if (want_full_csr) SPHB_CUDA(cudaMemsetAsync(w.row_cnt.p, 0, ((size_t)nf + 1) * sizeof(int), st));
        k_half_walk<false><<<grid_for(nf), kBlock, 0, st>>>(
            P, g, w.snap.as<double2>(), w.cell_of.as<int>(), w.cell_start.as<int>(), w.ent_key.as<int>(),
            w.ent_pos.as<double2>(), w.half_cnt.as<int>(), w.row_cnt.as<int>(), nullptr, nullptr, nullptr, nullptr,
            nullptr, want_full_csr ? 1 : 0, flags);
    }
    nl += 1;
    {
        ProfScope ps(prof, kPfNbrScan, st);
        SPHB_TRY(exclusive_scan_i32(w.half_cnt.as<int>(), w.half_ptr.as<int>(), (size_t)nf, w.scan_tmp, nullptr, st,
                                    &nl));
        if (want_full_csr)
            SPHB_TRY(exclusive_scan_i32(w.row_cnt.as<int>(), w.row_ptr.as<int>(), (size_t)nf, w.scan_tmp, nullptr, st,
                                        &nl));
    }

    // capacity: known-sufficient in steady state; otherwise read the totals back once and grow
    bool need_sync = (w.half_cap == 0) || (want_full_csr && w.csr_cap == 0);
    if (!need_sync) {
        k_check_caps<<<1, 1, 0, st>>>(P, w.half_ptr.as<int>(), w.row_ptr.as<int>(), want_full_csr ? 1 : 0,
                                      w.totals.as<int>(), flags);
        nl += 1;
    } else {
        int nh = 0, nc = 0;
        SPHB_CUDA(cudaMemcpyAsync(&nh, w.half_ptr.as<int>() + nf, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (want_full_csr)
            SPHB_CUDA(cudaMemcpyAsync(&nc, w.row_ptr.as<int>() + nf, sizeof(int), cudaMemcpyDeviceToHost, st));
        SPHB_CUDA(cudaStreamSynchronize(st));
        const size_t half_need = (size_t)nh + (size_t)nh / 4 + 1024;   // 25 % head-room for later rebuilds
        const size_t csr_need = (size_t)nc + (size_t)nc / 4 + 1024;
        if (half_need > w.half_cap) {
            SPHB_TRY(w.half_col.reserve(half_need * sizeof(unsigned)));
            w.half_cap = half_need;
        }
        if (want_full_csr && csr_need > w.csr_cap) {
            SPHB_TRY(w.col.reserve(csr_need * sizeof(unsigned)));
            w.csr_cap = csr_need;
        }
        P = make_params(w);
        k_check_caps<<<1, 1, 0, st>>>(P, w.half_ptr.as<int>(), w.row_ptr.as<int>(), want_full_csr ? 1 : 0,
                                      w.totals.as<int>(), flags);
        nl += 1;
    }

    {
        ProfScope ps(prof, kPfNbrWalkFill, st);
        if (want_full_csr) SPHB_CUDA(cudaMemsetAsync(w.row_fill.p, 0, ((size_t)nf + 1) * sizeof(int), st));
        k_half_walk<true><<<grid_for(nf), kBlock, 0, st>>>(
            P, g, w.snap.as<double2>(), w.cell_of.as<int>(), w.cell_start.as<int>(), w.ent_key.as<int>(),
            w.ent_pos.as<double2>(), nullptr, nullptr, w.half_ptr.as<int>(), w.row_ptr.as<int>(),
            w.row_fill.as<int>(), w.half_col.as<unsigned>(), w.col.as<unsigned>(), want_full_csr ? 1 : 0, flags);
    }
    nl += 1;
    if (want_full_csr) {
        ProfScope ps(prof, kPfNbrRowSort, st);
        k_row_sort<<<grid_for(nf), kBlock, 0, st>>>(nf, w.row_ptr.as<int>(), w.csr_cap, w.col.as<unsigned>());
        nl += 1;
    }
    SPHB_CUDA(cudaGetLastError());
    if (launches) *launches += nl;
    return 0;
}

int neighbor_export_pairs(NeighborWork &w, size_t n_half, double *pair_i, double *pair_j, double *dx, double *dy,
                          double *r, double *W, double *dW, cudaStream_t st)
{
    (void)n_half;
    NbrParams P = make_params(w);
    k_export_pairs<<<grid_for(w.n_fluid), kBlock, 0, st>>>(P, w.snap.as<double2>(), w.half_ptr.as<int>(),
                                                           w.half_col.as<unsigned>(), pair_i, pair_j, dx, dy, r, W,
                                                           dW);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sphb
