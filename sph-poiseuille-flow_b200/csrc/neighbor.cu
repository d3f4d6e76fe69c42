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
//   k_cell_canon    each cell's range sorted by DEcreasing insertion key: the reference's linked list
//                   is walked "last inserted first", so a forward scan of the range is its walk order
//   k_walk          ONE pass per fluid particle over its 3x3 stencil in the reference's order with its
//                   skip rules and its real/ghost dedup (seen_neighbor, nb:342,383), writing the row
//                   straight into an ELL (column-major, coalesced for the gather kernels): no count
//                   pass, no scatter atomics, no row sort
//   k_half_count / scan / k_export_pairs   (MEX export only) the reference's half-pair COO list, same
//                   order, same bits (nb:375-381)
#include "common.cuh"
#include "tile.cuh"

#include <cmath>
#include <climits>

namespace sphb {

namespace {


struct NbrParams {
    int n_fluid, n_total;
    int ncx;
    int cell_cap;
    double h, DL;
    double cell;      // 2.0 * h                         nb:253
    double cutoff;    // r_cut * h = 2.0 * h             nb:264
    double rc2;       // (r_cut*h) * (r_cut*h)           nb:306
    double rc2_safe;  // rc2 (1 - 1e-12): below it the kernel-support test cannot fail
    double half_DL;   // 0.5 * DL                        nb:359
    double sigma;     // 10 / (7 pi h h)                 nb:119
    int ell_cap;      // rows of the ELL (longest storable neighbour row)
    size_t ell_stride;
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

// insertion sort of one cell's entries by DEcreasing key (= the order the reference's list is walked in);
// also records where each particle's REAL entry ended up (rows of the ELL are indexed by entry)
__global__ void __launch_bounds__(kBlock) k_cell_canon(const GridDesc *__restrict__ g,
                                                      const int *__restrict__ cell_start,
                                                      int *__restrict__ ent_key, double2 *__restrict__ ent_pos,
                                                      int *__restrict__ ent_of)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g->ncells) return;
    const int s = cell_start[c], e = cell_start[c + 1];
    for (int a = s + 1; a < e; ++a) {
        const int key = ent_key[a];
        const double2 p = ent_pos[a];
        int b = a - 1;
        while (b >= s && ent_key[b] < key) {
            ent_key[b + 1] = ent_key[b];
            ent_pos[b + 1] = ent_pos[b];
            --b;
        }
        ent_key[b + 1] = key;
        ent_pos[b + 1] = p;
    }
    for (int a = s; a < e; ++a) {
        const int key = ent_key[a];
        if ((key & 3) == 0) ent_of[key >> 3] = a;
    }
}

// nb:357-372 for "particle a (wrapped xa, ya) meets an entry at (ex, ey)": minimum image, cut-off and
// kernel-support tests with the reference's arithmetic.  m = -1/0/+1 is the minimum-image shift.
__device__ __forceinline__ bool accept_entry(double xa, double ya, double ex, double ey, const NbrParams &P, int &m)
{
    double dx = xa - ex;
    m = 0;
    if (dx > P.half_DL) {
        dx -= P.DL;
        m = -1;
    } else if (dx < -P.half_DL) {
        dx += P.DL;
        m = 1;
    }
    const double dy = ya - ey;
    const double r2 = dx * dx + dy * dy;
    if (!(r2 > 1e-24 && r2 < P.rc2)) return false;
    // r2 < rc2 (1 - 1e-12) guarantees fl(fl(sqrt(r2)) / h) < 2, i.e. W > 0 (nb:116-133, 372); only the
    // last 1e-12 below the cut-off needs the reference's sqrt / divide / spline evaluated exactly
    if (r2 < P.rc2_safe) return true;
    const double r = sqrt(r2);
    double W, dW;
    spline2d(r, P.h, P.sigma, W, dW);
    return W > 0.0 || fabs(dW) > 0.0;
}

// Entries of one particle as the reference inserts them (nb:278-295), from its wrapped position and
// home cell: kind 0 real, 1 ghost at x - DL, 2 ghost at x + DL.
struct Entries {
    double x[3];
    int cx[3];     // -1: not inserted
    int cy;
};
__device__ __forceinline__ Entries entries_of(double xw, int home, const NbrParams &P)
{
    Entries E;
    const int cx = home % P.ncx;
    E.cy = home / P.ncx;
    E.x[0] = xw;
    E.cx[0] = cx;
    E.x[1] = xw - P.DL;
    E.cx[1] = -1;
    E.x[2] = xw + P.DL;
    E.cx[2] = -1;
    if (xw > P.DL - P.cutoff) {
        const int g = wrap_col((int)floor(E.x[1] / P.cell), P.ncx);
        if (g != cx) E.cx[1] = g;
    }
    if (xw < P.cutoff) {
        const int g = wrap_col((int)floor(E.x[2] / P.cell), P.ncx);
        if (g != cx) E.cx[2] = g;
    }
    return E;
}

// The reference stores a fluid-fluid pair once, found by the LOWER index's walk (nb:353).  When ghost
// entries are involved the two ends do not round dx identically, so the higher index replays the
// lower one's walk restricted to its own entries: cells in (oy, ox) order, within a cell the
// higher-kind entry first (list order = reverse insertion), first accepted entry wins.  Returns the
// image shift of x_high as seen from the low particle, or 2 when the pair is not accepted.
__device__ int replay_low_walk(double xl, double yl, int home_l, double xh, double yh, int home_h, int ncy,
                               const NbrParams &P)
{
    const Entries E = entries_of(xh, home_h, P);
    const int cxl = home_l % P.ncx, cyl = home_l / P.ncx;
    if (E.cy < cyl - 1 || E.cy > cyl + 1 || E.cy < 0 || E.cy >= ncy) return 2;
    int cx_prev[2] = {-1, -1};
    for (int ox = -1; ox <= 1; ++ox) {
        int cx = cxl + ox;
        if (cx < 0) cx += P.ncx;
        else if (cx >= P.ncx) cx -= P.ncx;
        if (cx < 0 || cx >= P.ncx) continue;
        if (cx == cx_prev[0] || cx == cx_prev[1]) continue;
        cx_prev[ox == -1 ? 0 : 1] = cx;
#pragma unroll
        for (int kind = 2; kind >= 0; --kind) {
            if (E.cx[kind] != cx) continue;
            int m;
            if (accept_entry(xl, yl, E.x[kind], yh, P, m)) return m + (kind == 1 ? 1 : (kind == 2 ? -1 : 0));
        }
    }
    return 2;
}

constexpr int kWalkCap = 640;   // staged candidate entries per tile (3 x 18 cells x ~6.8 = 365 on the lattice)

// The 3x3 stencil walk of nb:312-392, one CTA per tile of cells, candidates staged in shared memory.
//   HALF = true : the reference's own list (fluid-fluid pairs only from the lower index) -- MEX export
//   HALF = false: the full gather list: every neighbour of i, fluid-fluid pairs decided by the lower
//                 index exactly as the reference decides them (symmetric closure of its list)
// A thread owns one walker = the real entry of a fluid particle; lanes of a warp are consecutive
// entries of the cell-sorted array, i.e. particles of the same or adjacent cells.  Output: ELL rows
// indexed by the walker's ENTRY, column-major col[k * stride + e], in the reference's walk order
// (dropping the fluid j < i entries of a full row gives the reference's emission order).
template <bool HALF>
__global__ void __launch_bounds__(kTileThreads) k_walk_tile(NbrParams P, const GridDesc *__restrict__ g,
                                                           const double2 *__restrict__ snap,
                                                           const int *__restrict__ cell_of,
                                                           const int *__restrict__ ent_of,
                                                           const int *__restrict__ cell_start,
                                                           const int *__restrict__ ent_key,
                                                           const double2 *__restrict__ ent_pos, unsigned *col,
                                                           int *__restrict__ cnt, int *__restrict__ totals,
                                                           int *__restrict__ flags)
{
    __shared__ TileDesc T;
    __shared__ int s_key[kWalkCap];
    __shared__ double s_x[kWalkCap], s_y[kWalkCap];
    const int ncx = P.ncx, ncy = g->ncy;
    tile_setup(T, blockIdx.x, ncx, ncy, cell_start, kWalkCap);
    if (!T.valid || T.own_b == T.own_e) return;
    SPHB_TILE_FOR_STAGED(T, e, slot)
    {
        s_key[slot] = ent_key[e];
        const double2 p = ent_pos[e];
        s_x[slot] = p.x;
        s_y[slot] = p.y;
    }
    __syncthreads();

    const size_t stride = P.ell_stride;
    const int cyi = T.cy;
    int t_half = 0, t_full = 0, t_max = 0;
    for (int e0 = T.own_b; e0 < T.own_e; e0 += kTileThreads) {
        const int ew = e0 + (int)threadIdx.x;
        if (ew >= T.own_e) break;
        if ((size_t)ew >= stride) {   // more entries than ELL columns: reported, never written
            atomicOr(flags, kErrCsrCap);
            continue;
        }
        const int own_slot = tile_lookup(T, ew, 1);
        const int key_i = own_slot >= 0 ? s_key[own_slot] : ent_key[ew];
        const int i = key_i >> 3;
        if ((key_i & 3) != 0 || i >= P.n_fluid) {   // ghost or wall entry: not a walker
            cnt[ew] = 0;
            continue;
        }
        const double2 pi = own_slot >= 0 ? make_double2(s_x[own_slot], s_y[own_slot]) : ent_pos[ew];
        const int home = cell_of[i];
        const int cxi = home % ncx;
        // does i itself own ghost entries?  (then a lower fluid index may have met one of them)
        const bool i_multi = (pi.x > P.DL - P.cutoff || pi.x < P.cutoff) && ncx > 1;
        int count = 0, n_half = 0;

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
                int cb, ce;
                tile_cell_range(T, oy + 1, cx, ncx, cb, ce);
                if (cb == ce) continue;
                int n_staged;
                const int sb = tile_lookup_run(T, cb, ce, oy + 1, n_staged);
                for (int e = cb; e < ce; ++e) {   // entries are stored last-inserted first
                    const int s = (e - cb) < n_staged ? sb + (e - cb) : -1;   // -1: beyond the staged part
                    const int key = s >= 0 ? s_key[s] : ent_key[e];
                    const int j = key >> 3;
                    if (j == i) continue;
                    const bool low_fluid = j < P.n_fluid && j < i;
                    if (HALF && low_fluid) continue;                       // nb:353
                    const bool seam = (key & 4) || (low_fluid && i_multi);
                    if (seam) {   // seen_neighbor (nb:342): j may already be in the row through another entry
                        bool dup = false;
                        const int lim = count < P.ell_cap ? count : P.ell_cap;
                        for (int q = 0; q < lim; ++q)
                            dup |= ((ent_key[ell_entry(col[(size_t)q * stride + ew])] >> 3) == j);
                        if (dup) continue;
                    }
                    int m, acc = e, row_off = oy;
                    if (low_fluid && seam) {
                        const double2 pj = snap[j];
                        const int home_j = cell_of[j];
                        const int sh = replay_low_walk(pj.x, pj.y, home_j, pi.x, pi.y, home, ncy, P);
                        if (sh == 2) continue;
                        acc = ent_of[j];   // the row names j's REAL entry; the image shift goes into m
                        m = -sh;
                        row_off = home_j / ncx - cyi;
                    } else {
                        const double ex = s >= 0 ? s_x[s] : ent_pos[e].x;
                        const double ey = s >= 0 ? s_y[s] : ent_pos[e].y;
                        if (!accept_entry(pi.x, pi.y, ex, ey, P, m)) continue;
                    }
                    if (count < P.ell_cap) col[(size_t)count * stride + ew] = ell_pack(acc, m, row_off);
                    ++count;
                    n_half += low_fluid ? 0 : 1;
                }
            }
        }
        cnt[ew] = count;
        t_half += n_half;
        t_full += count;
        t_max = max(t_max, count);
    }
    // totals: half pairs, full entries, longest row (capacity check)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        t_half += __shfl_xor_sync(0xffffffffu, t_half, d);
        t_full += __shfl_xor_sync(0xffffffffu, t_full, d);
        t_max = max(t_max, __shfl_xor_sync(0xffffffffu, t_max, d));
    }
    if ((threadIdx.x & 31) == 0) {
        if (t_half) atomicAdd(&totals[0], t_half);
        if (t_full) atomicAdd(&totals[1], t_full);
        if (t_max) atomicMax(&totals[2], t_max);
        if (t_max > P.ell_cap) atomicOr(flags, kErrCsrCap);
    }
}

// rows of the reference's half list inside a FULL row set: count of entries with j wall or j > i
__global__ void __launch_bounds__(kBlock) k_half_count(NbrParams P, const int *__restrict__ ent_of,
                                                      const int *__restrict__ ent_key,
                                                      const unsigned *__restrict__ col,
                                                      const int *__restrict__ cnt, int *__restrict__ half_cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_fluid) return;
    const int e = ent_of[i];
    const int n = min(cnt[e], P.ell_cap);
    int c = 0;
    for (int k = 0; k < n; ++k) {
        const int j = ent_key[ell_entry(col[(size_t)k * P.ell_stride + e])] >> 3;
        c += (j >= P.n_fluid || j > i) ? 1 : 0;
    }
    half_cnt[i] = c;
}

// nb:357-381 re-evaluated for the stored entry: identical operations, identical bits
__global__ void __launch_bounds__(kBlock) k_export_pairs(NbrParams P, const int *__restrict__ ent_of,
                                                        const int *__restrict__ ent_key,
                                                        const double2 *__restrict__ ent_pos,
                                                        const unsigned *__restrict__ col,
                                                        const int *__restrict__ cnt,
                                                        const int *__restrict__ half_ptr,
                                                        double *__restrict__ pair_i, double *__restrict__ pair_j,
                                                        double *__restrict__ odx, double *__restrict__ ody,
                                                        double *__restrict__ orr, double *__restrict__ oW,
                                                        double *__restrict__ odW)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_fluid) return;
    const int ew = ent_of[i];
    const double2 pi = ent_pos[ew];
    const int n = min(cnt[ew], P.ell_cap);
    size_t o = (size_t)half_ptr[i];
    for (int k = 0; k < n; ++k) {
        const int ea = ell_entry(col[(size_t)k * P.ell_stride + ew]);
        const int j = ent_key[ea] >> 3;
        if (j < P.n_fluid && j < i) continue;
        const double2 pe = ent_pos[ea];   // a ghost entry already holds x_j -+ DL (nb:284,291)
        double dx = pi.x - pe.x;
        if (dx > P.half_DL) dx -= P.DL;
        else if (dx < -P.half_DL) dx += P.DL;
        const double dy = pi.y - pe.y;
        const double r = sqrt(dx * dx + dy * dy);
        double W, dW;
        spline2d(r, P.h, P.sigma, W, dW);
        pair_i[o] = (double)(i + 1);
        pair_j[o] = (double)(j + 1);
        odx[o] = dx;
        ody[o] = dy;
        orr[o] = r;
        oW[o] = W;
        odW[o] = dW;
        ++o;
    }
}

// full rows by PARTICLE (export only): counts, then particle ids in walk order
__global__ void __launch_bounds__(kBlock) k_row_count(NbrParams P, const int *__restrict__ ent_of,
                                                     const int *__restrict__ cnt, int *__restrict__ row_cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.n_fluid) row_cnt[i] = min(cnt[ent_of[i]], P.ell_cap);
}
__global__ void __launch_bounds__(kBlock) k_row_fill(NbrParams P, const int *__restrict__ ent_of,
                                                    const int *__restrict__ ent_key,
                                                    const unsigned *__restrict__ col, const int *__restrict__ cnt,
                                                    const int *__restrict__ row_ptr, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_fluid) return;
    const int e = ent_of[i];
    const int n = min(cnt[e], P.ell_cap);
    const size_t o = (size_t)row_ptr[i];
    for (int k = 0; k < n; ++k) out[o + (size_t)k] = ent_key[ell_entry(col[(size_t)k * P.ell_stride + e])] >> 3;
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
    P.rc2_safe = P.rc2 * (1.0 - 1e-12);
    P.half_DL = 0.5 * w.DL;
    P.sigma = 10.0 / (7.0 * pi * w.h * w.h);
    int ncx = (int)std::ceil(w.DL / P.cell);   // nb:254
    if (ncx < 1) ncx = 1;
    P.ncx = ncx;
    P.cell_cap = w.cell_cap;
    P.ell_cap = w.ell_cap;
    P.ell_stride = w.ell_stride;
    return P;
}

}  // namespace

void NeighborWork::release_all()
{
    DevBuf *all[] = {&grid, &flags, &snap, &cell_of, &ent_of, &cell_cnt, &cell_start, &cell_fill, &ent_key, &ent_pos,
                     &ell_cnt, &ell_col, &half_cnt, &half_ptr, &scan_tmp, &totals};
    for (DevBuf *b : all) b->release();
}

// ELL rows are indexed by ENTRY (real + ghost): the stride is the entry capacity the rows may use.
// Ghost entries are a sliver of the total unless the box is only a few cells long, so large problems
// get n_total * 1.25 columns, small ones the reference's own worst case 3 n_total (nb:265).
int neighbor_reserve_ell(NeighborWork &w, int rows)
{
    const size_t nt = (size_t)w.n_total;
    size_t stride = nt < (1u << 20) ? 3 * nt : nt + nt / 4 + 4096;
    stride = (stride + 31) & ~(size_t)31;
    if (rows <= w.ell_cap && w.ell_col.p && stride == w.ell_stride) return 0;
    if (rows < w.ell_cap) rows = w.ell_cap;
    w.ell_stride = stride;
    SPHB_TRY(w.ell_col.reserve((size_t)rows * stride * sizeof(unsigned)));
    SPHB_TRY(w.ell_cnt.reserve(stride * sizeof(int)));
    w.ell_cap = rows;
    return 0;
}

// Size every buffer whose size is known a priori.  y hints bound the cell rows (a fluid particle
// escaping beyond the hinted range raises the cell-capacity flag instead of writing out of bounds).
int neighbor_plan(NeighborWork &w, int n_fluid, int n_total, double h, double DL, double y_min_hint,
                  double y_max_hint, int max_nbr_hint)
{
    const bool resized = n_total != w.n_total;
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
    if (3LL * (long long)n_total >= (1LL << 28))
        return set_error("SPH:Neighbor:entryCapacity", "3 n_total exceeds the 2^28 limit of the entry index");
    w.cell_cap = (int)(ncx_d * ncy_d);
    const size_t nt = (size_t)n_total;
    SPHB_TRY(w.grid.reserve(sizeof(GridDesc)));
    if (!w.flags.p) {
        SPHB_TRY(w.flags.reserve(sizeof(int)));
        SPHB_CUDA(cudaMemset(w.flags.p, 0, sizeof(int)));
    }
    SPHB_TRY(w.totals.reserve(4 * sizeof(int)));
    SPHB_TRY(w.snap.reserve(nt * sizeof(double2)));
    SPHB_TRY(w.cell_of.reserve(nt * sizeof(int)));
    SPHB_TRY(w.ent_of.reserve(nt * sizeof(int)));
    SPHB_TRY(w.cell_cnt.reserve(((size_t)w.cell_cap + 1) * sizeof(int)));
    SPHB_TRY(w.cell_start.reserve(((size_t)w.cell_cap + 1) * sizeof(int)));
    SPHB_TRY(w.cell_fill.reserve(((size_t)w.cell_cap + 1) * sizeof(int)));
    SPHB_TRY(w.ent_key.reserve(3 * nt * sizeof(int)));
    SPHB_TRY(w.ent_pos.reserve(3 * nt * sizeof(double2)));
    if (resized) {
        w.ell_cap = 0;   // the stride depends on n_total
        w.ell_col.release();
    }
    SPHB_TRY(neighbor_reserve_ell(w, max_nbr_hint > 0 ? max_nbr_hint : kDefaultEllRows));
    return 0;
}

// Builds the list from `pos` (device, double2[n_total]).  A row longer than the ELL raises the
// capacity flag; with `exact` the build then synchronises, grows the ELL to the longest row and walks
// again (stateless path and the first resident build); in steady state there is no host round trip.
int neighbor_build(NeighborWork &w, const double2 *pos, bool half_only, bool exact, cudaStream_t st,
                   long long *launches, Prof *prof)
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
                                                                      w.ent_pos.as<double2>(), w.ent_of.as<int>());
    }
    nl += 2;

    for (int attempt = 0; attempt < 2; ++attempt) {
        ProfScope ps(prof, kPfNbrWalk, st);
        SPHB_CUDA(cudaMemsetAsync(w.totals.p, 0, 4 * sizeof(int), st));
        // one CTA per tile of kTileCells cells; rows beyond the device-side n_cell_y exit at once
        const unsigned n_tiles = (unsigned)((w.cell_cap / P.ncx) * tiles_per_row(P.ncx));
        if (half_only)
            k_walk_tile<true><<<n_tiles, kTileThreads, 0, st>>>(
                P, g, w.snap.as<double2>(), w.cell_of.as<int>(), w.ent_of.as<int>(), w.cell_start.as<int>(),
                w.ent_key.as<int>(), w.ent_pos.as<double2>(), w.ell_col.as<unsigned>(), w.ell_cnt.as<int>(),
                w.totals.as<int>(), flags);
        else
            k_walk_tile<false><<<n_tiles, kTileThreads, 0, st>>>(
                P, g, w.snap.as<double2>(), w.cell_of.as<int>(), w.ent_of.as<int>(), w.cell_start.as<int>(),
                w.ent_key.as<int>(), w.ent_pos.as<double2>(), w.ell_col.as<unsigned>(), w.ell_cnt.as<int>(),
                w.totals.as<int>(), flags);
        nl += 1;
        if (!exact || attempt == 1) break;
        int host[4] = {0, 0, 0, 0};
        SPHB_CUDA(cudaMemcpyAsync(host, w.totals.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        SPHB_CUDA(cudaStreamSynchronize(st));
        if (host[2] <= w.ell_cap) break;
        // grow to the longest row (+ head-room for later rebuilds) and walk again
        SPHB_TRY(neighbor_reserve_ell(w, host[2] + host[2] / 4 + 4));
        int f = 0;
        SPHB_CUDA(cudaMemcpyAsync(&f, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        SPHB_CUDA(cudaStreamSynchronize(st));
        f &= ~kErrCsrCap;
        SPHB_CUDA(cudaMemcpyAsync(flags, &f, sizeof(int), cudaMemcpyHostToDevice, st));
        SPHB_CUDA(cudaStreamSynchronize(st));
        P = make_params(w);
    }
    w.half_only = half_only;
    SPHB_CUDA(cudaGetLastError());
    if (launches) *launches += nl;
    return 0;
}

// The reference's half-pair COO list from the ELL rows (device output pointers, n_half slots each).
int neighbor_export_pairs(NeighborWork &w, size_t n_half, double *pair_i, double *pair_j, double *dx, double *dy,
                          double *r, double *W, double *dW, cudaStream_t st)
{
    (void)n_half;
    NbrParams P = make_params(w);
    const size_t nf = (size_t)w.n_fluid;
    SPHB_TRY(w.half_cnt.reserve((nf + 1) * sizeof(int)));
    SPHB_TRY(w.half_ptr.reserve((nf + 1) * sizeof(int)));
    k_half_count<<<grid_for(nf), kBlock, 0, st>>>(P, w.ent_of.as<int>(), w.ent_key.as<int>(),
                                                  w.ell_col.as<unsigned>(), w.ell_cnt.as<int>(), w.half_cnt.as<int>());
    SPHB_TRY(exclusive_scan_i32(w.half_cnt.as<int>(), w.half_ptr.as<int>(), nf, w.scan_tmp, nullptr, st, nullptr));
    k_export_pairs<<<grid_for(nf), kBlock, 0, st>>>(P, w.ent_of.as<int>(), w.ent_key.as<int>(),
                                                    w.ent_pos.as<double2>(), w.ell_col.as<unsigned>(),
                                                    w.ell_cnt.as<int>(), w.half_ptr.as<int>(), pair_i, pair_j, dx, dy,
                                                    r, W, dW);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

// The full gather list as CSR by particle (device outputs: row_ptr[n_fluid+1], col[n_full] in walk order).
int neighbor_export_csr(NeighborWork &w, int *row_ptr, int *col, cudaStream_t st)
{
    NbrParams P = make_params(w);
    const size_t nf = (size_t)w.n_fluid;
    SPHB_TRY(w.half_cnt.reserve((nf + 1) * sizeof(int)));
    k_row_count<<<grid_for(nf), kBlock, 0, st>>>(P, w.ent_of.as<int>(), w.ell_cnt.as<int>(), w.half_cnt.as<int>());
    SPHB_TRY(exclusive_scan_i32(w.half_cnt.as<int>(), row_ptr, nf, w.scan_tmp, nullptr, st, nullptr));
    if (col)
        k_row_fill<<<grid_for(nf), kBlock, 0, st>>>(P, w.ent_of.as<int>(), w.ent_key.as<int>(), w.ell_col.as<unsigned>(),
                                                    w.ell_cnt.as<int>(), row_ptr, col);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sphb
