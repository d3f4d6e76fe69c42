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
//   k_scatter       entry records into their cell ranges (atomic slot: any order inside a cell)
//   k_walk_tile     one CTA per cell tile: its descriptor (tile.cuh: staged entry ranges and their slots), the
//                   canonical order of every staged cell on the way into shared memory (rank sort by DEcreasing
//                   insertion key: the reference's linked list is walked "last inserted first", so a forward scan
//                   of the range is its walk order; the owner tile writes it to the entry arrays), then
//                   ONE pass per fluid particle over its 3x3 stencil out of shared memory, in the
//                   reference's order with its skip rules and its real/ghost dedup (seen_neighbor,
//                   nb:342,383); rows of 16-bit tile-local entries, eight to a 128-bit store: no count
//                   pass, no scatter atomics, no row sort
//   k_half_count / scan / k_export_pairs   (MEX export only) the reference's half-pair COO list, same
//                   order, same bits (nb:375-381)
#include "common.cuh"
#include "tile.cuh"

#include <algorithm>
#include <cstdlib>
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
    int slab;         // x-slab of a sharded channel: no wrap, no ghost entries, columns counted from col0
    int col0;         // global cell column of local column 0 (slab mode)
    const int *dev_counts;   // slab mode: {n_fluid, owned, n_total} live in device memory; NULL = the values above
    const int *alive; // slab mode: fluid row i takes part only if alive[i] >= 0 (rows that migrated away stay
                      // in place, marked dead, until the next compaction); NULL = every row
};
__device__ __forceinline__ void load_counts(NbrParams &P)
{
    if (P.dev_counts) {
        P.n_fluid = P.dev_counts[0];
        P.n_total = P.dev_counts[2];
    }
}
__device__ __forceinline__ bool row_dead(const NbrParams &P, int i)
{
    return P.alive != nullptr && i < P.n_fluid && P.alive[i] < 0;
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

// Per-channel arguments of a BATCHED list build (ensembles of independent channels, batch.cu): the same
// kernel bodies run for every channel in one grid, blockIdx.y = channel.
struct __align__(16) NbrChan {   // (batch.cu strides its record array by the size rounded up to 16)
    NbrParams P;
    const double2 *pos;
    GridDesc *g;
    int *flags, *totals;
    int *cell_cnt;      // counts, then (at + cell_cap + 1) the fill cursors
    int *cell_start;
    double2 *snap;
    int *cell_of, *ent_of;
    int *tmp_key;
    double2 *tmp_pos;
    int *ent_key;
    double2 *ent_pos;
    TileDesc *tiles;
    uint4 *ell_q;
    int *ell_cnt;
    int tile_w, tile_cap;
    unsigned n_tiles;
    BuildMirrors M;
};

__global__ void k_grid_init(GridDesc *g)
{
    g->ymin_key = LLONG_MAX;
    g->ymax_key = LLONG_MIN;
    g->n_entries = 0;
    g->ticket = 0u;
    g->ghost_anomaly = 0u;
}

// y range over ALL particles (nb:245-250); the last CTA to finish derives the grid (nb:253-258), re-arms
// the reduction for the next build and zeroes the build's totals: one launch instead of three.
__device__ __forceinline__ void yrange_grid_body(const double2 *__restrict__ pos, int n, GridDesc *g, NbrParams P, int *flags,
                                                 int *totals)
{
    load_counts(P);
    if (P.dev_counts) n = P.n_total;
    long long kmin = LLONG_MAX, kmax = LLONG_MIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (row_dead(P, i)) continue;
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
    // one atomic pair per CTA, not per warp: thousands of 64-bit atomics on one address serialise
    __shared__ long long s_min[8], s_max[8];
    if ((threadIdx.x & 31) == 0) {
        s_min[threadIdx.x >> 5] = kmin;
        s_max[threadIdx.x >> 5] = kmax;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
        kmin = s_min[w] < kmin ? s_min[w] : kmin;
        kmax = s_max[w] > kmax ? s_max[w] : kmax;
    }
    atomicMin(&g->ymin_key, kmin);
    atomicMax(&g->ymax_key, kmax);
    __threadfence();
    if (atomicAdd(&g->ticket, 1u) != gridDim.x - 1) return;
    __threadfence();
    grid_derive(g, P.cell, P.ncx, P.cell_cap, flags, totals);
}
__global__ void __launch_bounds__(256) k_yrange_grid(const double2 *__restrict__ pos, int n, GridDesc *g, NbrParams P,
                                                    int *flags, int *totals)
{
    yrange_grid_body(pos, n, g, P, flags, totals);
}
__global__ void __launch_bounds__(256) k_yrange_grid_batch(const NbrChan *__restrict__ chan)
{
    const NbrChan &c = chan[blockIdx.y];
    yrange_grid_body(c.pos, c.P.n_total, c.g, c.P, c.flags, c.totals);
}
// zeroes the cell counters and fill cursors of every channel (the cudaMemsetAsync of the single build)
__global__ void __launch_bounds__(kBlock) k_zero_cells_batch(const NbrChan *__restrict__ chan, int n_ints)
{
    int *dst = chan[blockIdx.y].cell_cnt;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_ints; k += gridDim.x * blockDim.x) dst[k] = 0;
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
    b.y = p.y;
    if (P.slab) {
        // slab of a sharded channel: positions are kept unwrapped around the slab (halo particles from
        // across the periodic seam arrive shifted by -+DL, exactly like the reference's ghost entries,
        // nb:284,291), the column is the GLOBAL one counted from col0, and there are no ghost entries
        b.xw = p.x;
        int c = (int)floor(p.x / P.cell) - P.col0;
        b.cx = c < 0 ? 0 : (c >= P.ncx ? P.ncx - 1 : c);
        int cys = (int)floor((p.y - y_min) / P.cell);
        b.cy = cys < 0 ? 0 : (cys >= ncy ? ncy - 1 : cys);
        b.gcx_lo = -1;
        b.gcx_hi = -1;
        return b;
    }
    b.xw = p.x - floor(p.x / P.DL) * P.DL;
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

// (A build inside a poisoned step -- ops_csr.cuh step_poisoned -- runs on: the state stands still, so it rebuilds the
// list it already has; the continuity kernel of a poisoned step still zeroes the cell counters the binning counts into.)
__device__ __forceinline__ void bin_count_body(const double2 *__restrict__ pos, NbrParams P, const GridDesc *__restrict__ g,
                                               double2 *__restrict__ snap, int *__restrict__ cell_of,
                                               int *__restrict__ cell_cnt, unsigned int *ghost_anomaly, const int *flags)
{
    load_counts(P);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_total || row_dead(P, i)) return;
    const Bins b = bin_particle(pos[i], P, g->y_min, g->ncy);
    if (!P.slab) {   // see GridDesc::ghost_anomaly
        // the reference's wrap x - floor(x / DL) DL returns DL itself for a tiny negative x (and can return a
        // negative sliver): a real entry that needs the minimum image wherever it is met
        unsigned odd = (b.xw >= 0.0 && b.xw < P.DL) ? 0u : 1u;
        if (b.gcx_lo >= 0 || b.gcx_hi >= 0) {
            const int last = P.ncx - 1;
            const bool ghosts_ok = (b.gcx_lo < 0 || b.gcx_lo == last) && (b.gcx_hi < 0 || b.gcx_hi == last);
            const bool owner_ok = b.cx == 0 || b.cx >= P.ncx - 2;
            if (!(ghosts_ok && owner_ok)) odd |= 2u;
        }
        if (odd) atomicOr(ghost_anomaly, odd);
    }
    snap[i] = make_double2(b.xw, b.y);
    const int row = b.cy * P.ncx;
    cell_of[i] = row + b.cx;
    atomicAdd(&cell_cnt[row + b.cx], 1);
    if (b.gcx_lo >= 0) atomicAdd(&cell_cnt[row + b.gcx_lo], 1);
    if (b.gcx_hi >= 0) atomicAdd(&cell_cnt[row + b.gcx_hi], 1);
}
__global__ void __launch_bounds__(kBlock) k_bin_count(const double2 *__restrict__ pos, NbrParams P,
                                                     const GridDesc *__restrict__ g, double2 *__restrict__ snap,
                                                     int *__restrict__ cell_of, int *__restrict__ cell_cnt,
                                                     unsigned int *ghost_anomaly, const int *flags)
{
    pdl_wait();
    pdl_trigger_top();
    bin_count_body(pos, P, g, snap, cell_of, cell_cnt, ghost_anomaly, flags);
    pdl_trigger_end();
}
__global__ void __launch_bounds__(kBlock) k_bin_count_batch(const NbrChan *__restrict__ chan)
{
    const NbrChan &c = chan[blockIdx.y];
    bin_count_body(c.pos, c.P, c.g, c.snap, c.cell_of, c.cell_cnt, &c.g->ghost_anomaly, c.flags);
}

__device__ __forceinline__ void scatter_body(const double2 *__restrict__ pos, NbrParams P, const GridDesc *__restrict__ g,
                                             const int *__restrict__ cell_start, int *__restrict__ cell_fill,
                                             int *__restrict__ tmp_key, double2 *__restrict__ tmp_pos, const int *flags)
{
    load_counts(P);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_total || row_dead(P, i)) return;
    const Bins b = bin_particle(pos[i], P, g->y_min, g->ncy);
    const int row = b.cy * P.ncx;
    const int multi = (b.gcx_lo >= 0 || b.gcx_hi >= 0) ? 4 : 0;
    {
        const int c = row + b.cx;
        const int s = cell_start[c] + atomicAdd(&cell_fill[c], 1);
        tmp_key[s] = (i << 3) | multi | 0;
        tmp_pos[s] = make_double2(b.xw, b.y);
    }
    if (b.gcx_lo >= 0) {
        const int c = row + b.gcx_lo;
        const int s = cell_start[c] + atomicAdd(&cell_fill[c], 1);
        tmp_key[s] = (i << 3) | multi | 1;
        tmp_pos[s] = make_double2(b.xw - P.DL, b.y);
    }
    if (b.gcx_hi >= 0) {
        const int c = row + b.gcx_hi;
        const int s = cell_start[c] + atomicAdd(&cell_fill[c], 1);
        tmp_key[s] = (i << 3) | multi | 2;
        tmp_pos[s] = make_double2(b.xw + P.DL, b.y);
    }
}
__global__ void __launch_bounds__(kBlock) k_scatter(const double2 *__restrict__ pos, NbrParams P,
                                                   const GridDesc *__restrict__ g,
                                                   const int *__restrict__ cell_start, int *__restrict__ cell_fill,
                                                   int *__restrict__ tmp_key, double2 *__restrict__ tmp_pos, const int *flags)
{
    pdl_wait();
    pdl_trigger_top();
    scatter_body(pos, P, g, cell_start, cell_fill, tmp_key, tmp_pos, flags);
    pdl_trigger_end();
}
__global__ void __launch_bounds__(kBlock) k_scatter_batch(const NbrChan *__restrict__ chan)
{
    const NbrChan &c = chan[blockIdx.y];
    scatter_body(c.pos, c.P, c.g, c.cell_start, c.cell_cnt + ((size_t)c.P.cell_cap + 1), c.tmp_key, c.tmp_pos, c.flags);
}

// One descriptor per tile (tile.cuh): entry ranges of the (kTileRows + 2) x (w + 2) cell neighbourhood and
// their slots.  One warp per tile: lane (3 r + k) fetches range k of staged row r, the slot offsets are a
// warp prefix sum, lane 0 writes the rest.  Called by the first warp of the walk's CTA for its own tile (the
// descriptor is then read by every gather kernel of the step).
__device__ __forceinline__ void tile_setup_warp(unsigned t, int lane, const GridDesc *__restrict__ g, int ncx, int slab, int w,
                                                unsigned n_tiles, int cap, int ent_cap, const int *__restrict__ cell_start,
                                                TileDesc &T, TileDesc *tiles, int *__restrict__ totals, int *__restrict__ flags)
{
    if (t >= n_tiles) {   // (never: one CTA per tile)
        if (lane == 0) T.flags = 0;
        return;
    }
    const int ncy = g->ncy;
    const int tpr = tiles_per_row(ncx, w);
    const int cy0 = (int)(t / (unsigned)tpr) * kTileRows, cx0 = (int)(t % (unsigned)tpr) * w;
    const int cx1 = min(cx0 + w, ncx);
    const int a = max(cx0 - 1, 0), b = min(cx1 + 1, ncx);
    const bool valid = cy0 < ncy;
    // lane -> (staged row r, range k): 0 main columns a..b-1, 1 wrap column ncx-1, 2 wrap column 0
    int rb = 0, re = 0;
    if (lane < kStageRows * 3) {
        const int r = lane / 3, k = lane % 3;
        const int row = cy0 + r - 1;
        if (valid && row >= 0 && row < ncy) {
            const int *cs = cell_start + (size_t)row * ncx;
            if (k == 0) {
                rb = cs[a];
                re = cs[b];
            } else if (k == 1 && cx0 == 0 && b <= ncx - 1) {   // left halo column wraps to ncx-1
                rb = cs[ncx - 1];
                re = cs[ncx];
            } else if (k == 2 && cx1 == ncx && a >= 1) {       // right halo column wraps to 0
                rb = cs[0];
                re = cs[1];
            }
        }
    }
    const int len = re - rb;
    int inc = len;   // inclusive prefix over the lanes = slot offsets
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    const int used = __shfl_sync(0xffffffffu, inc, kStageRows * 3 - 1);
    const bool fits = __all_sync(0xffffffffu, re <= ent_cap);   // (entry overflow itself is reported by the walk)
    if (lane < kStageRows * 3) {
        T.seg[lane].b = rb;
        T.seg[lane].e = re;
        T.seg[lane].off = inc - len;
    }
    // own rows: lane k < kTileRows fetches its own range; it is inside the main range of staged row k + 1
    const int main_lane = 3 * (min(lane, kTileRows - 1) + 1);   // (all lanes take part in the shuffles)
    const int main_b = __shfl_sync(0xffffffffu, rb, main_lane);
    const int main_off = __shfl_sync(0xffffffffu, inc - len, main_lane);
    if (lane < kTileRows) {
        const bool row_ok = valid && cy0 + lane < ncy;
        const int *cs = cell_start + (size_t)(row_ok ? cy0 + lane : 0) * ncx;
        const int ob = row_ok ? cs[cx0] : 0, on = row_ok ? cs[cx1] - cs[cx0] : 0;
        T.own_b[lane] = ob;
        T.own_n[lane] = on;
        T.own_off[lane] = main_off + (ob - main_b);
    }
    if (lane == 0) {
        T.n_staged = used;
        T.cy0 = cy0;
        T.cx0 = cx0;
        T.cx1 = cx1;
        T.a = a;
        T.b = b;
        int f = valid && fits ? kTileValid : 0;
        // walkers of columns 0, ncx-2 and ncx-1 can meet ghost entries, both entries of a particle that owns
        // a ghost, or the periodic wrap: they take the general walk (ghosts live in column ncx-1 only and
        // their owners in columns 0, ncx-2, ncx-1; nb:282-295)
        // (a slab of a sharded channel has neither: its outermost columns are empty margins)
        if (!slab && (cx0 < 1 || cx1 > ncx - 2 || ncx < 6 || g->ghost_anomaly != 0u)) f |= kTileSeam;
        if (used + 1 > cap) {   // cannot be staged (+ the dummy slot): reported, the tile is skipped by every kernel
            atomicOr(flags, kErrTileCap);
            f &= ~kTileValid;
        }
        T.flags = f;
        if (valid) atomicMax(&totals[3], used + 1);
    }
    // the descriptor (built in the CTA's shared memory) to the global array the gather kernels read
    __syncwarp();
    if (lane < (int)(sizeof(TileDesc) / 16)) reinterpret_cast<int4 *>(tiles + t)[lane] = reinterpret_cast<const int4 *>(&T)[lane];
}
// nb:357-372 for "particle a (wrapped xa, ya) meets an entry at (ex, ey)": minimum image, cut-off and
// kernel-support tests with the reference's arithmetic.  m = -1/0/+1 is the minimum-image shift.
__device__ __forceinline__ bool support_exact(double r2, const NbrParams &P)
{
    const double r = sqrt(r2);
    double W, dW;
    spline2d(r, P.h, P.sigma, W, dW);
    return W > 0.0 || fabs(dW) > 0.0;
}
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
    return support_exact(r2, P);
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

// position of column cx in the stencil walk of a particle of column cxi (ox = -1, 0, 1, wrapped; a
// revisited column counts at its first visit), or -1 when the stencil does not reach it
__device__ __forceinline__ int stencil_rank(int cx, int cxi, int ncx)
{
    int c0 = cxi - 1, c2 = cxi + 1;
    if (c0 < 0) c0 += ncx;
    if (c2 >= ncx) c2 -= ncx;
    if (cx == c0) return 0;
    if (cx == cxi) return 1;
    if (cx == c2) return 2;
    return -1;
}

// The reference stores a fluid-fluid pair once, found by the LOWER index's walk (nb:353).  When ghost
// entries are involved the two ends do not round dx identically, so the higher index replays the
// lower one's walk restricted to its own entries: cells in (oy, ox) order, within a cell the
// higher-kind entry first (list order = reverse insertion), first accepted entry wins.  Returns the
// image shift of x_high as seen from the low particle, or kNoPair when the pair is not accepted.
constexpr int kNoPair = 99;
constexpr int kSurvCap = 96;   // survivors of the quick tests a warp buffers between the two passes of a general walk
__device__ int replay_low_walk(double xl, double yl, int home_l, const Entries &E, double yh, int ncy,
                               const NbrParams &P)
{
    const int cxl = home_l % P.ncx, cyl = home_l / P.ncx;
    if (E.cy < cyl - 1 || E.cy > cyl + 1 || E.cy < 0 || E.cy >= ncy) return kNoPair;
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
    return kNoPair;
}

// The 3x3 stencil walk of nb:312-392, one CTA per tile of cells, candidates staged in shared memory.
//   HALF = true : the reference's own list (fluid-fluid pairs only from the lower index) -- MEX export
//   HALF = false: the full gather list: every neighbour of i, fluid-fluid pairs decided by the lower
//                 index exactly as the reference decides them (symmetric closure of its list)
// A thread owns one walker = the real entry of a fluid particle; lanes of a warp are consecutive
// entries of the cell-sorted array, i.e. particles of the same or adjacent cells.  The row is emitted
// in the reference's walk order (dropping the fluid j < i entries of a full row gives the reference's
// emission order), eight 16-bit entries per 128-bit store.
//
// Interior walkers (columns 2 .. ncx-4) cannot meet a ghost entry, a particle that owns one, or the
// periodic wrap: their three stencil cells of a row are ONE contiguous slot range and the accept test
// is the reference's arithmetic with no image shift.  The others take the general walk, where the
// reference's seen_neighbor dedup (nb:342,383) is decided from the walk positions of the (at most
// three) entries of the other particle instead of from a per-particle table.
struct RowWriter {
    uint4 q;
    int count;
    uint4 *dst;        // ell_q + ew
    size_t stride;
    int qcap;          // words a row may use
    __device__ __forceinline__ void init(uint4 *d, size_t s, int cap_words)
    {
        q = make_uint4(0u, 0u, 0u, 0u);
        count = 0;
        dst = d;
        stride = s;
        qcap = cap_words;
    }
    __device__ __forceinline__ void push(int slot, int m)
    {
        ell_push(q, ell_pack(slot, m));
        ++count;
        if ((count & 7) == 0) {
            const int w = (count >> 3) - 1;
            if (w < qcap) dst[(size_t)w * stride] = q;
        }
    }
    // pads the last word with the dummy slot (tile.cuh) and stores it
    __device__ __forceinline__ void finish(int dummy_slot)
    {
        const int c = count & 7;
        if (c) {
            for (int k = c; k < 8; ++k) ell_push(q, ell_pack(dummy_slot, 0));
            const int w = count >> 3;
            if (w < qcap) dst[(size_t)w * stride] = q;
        }
    }
};

// Seam tiles (first / last of a tile row) carry the expensive general walkers: they are dealt to the
// first CTAs so that they do not form the tail of the launch.
__device__ __forceinline__ unsigned walk_tile_of_block(unsigned b, int tpr, unsigned n_tiles)
{
    if (tpr < 3) return b;
    const unsigned n_rows = n_tiles / (unsigned)tpr;
    if (b < 2 * n_rows) return (b >> 1) * (unsigned)tpr + ((b & 1u) ? (unsigned)(tpr - 1) : 0u);
    const unsigned r = b - 2 * n_rows;
    return (r / (unsigned)(tpr - 2)) * (unsigned)tpr + 1u + r % (unsigned)(tpr - 2);
}

// The general walk of ONE walker (columns 0, ncx-2, ncx-1: ghost entries, particles that own ghosts, the
// periodic wrap), executed by a whole warp: the walker's candidates, enumerated in the reference's walk
// order, are spread over the lanes; each lane decides its candidate on its own -- the seen_neighbor dedup
// (nb:342,383) is a function of the walk positions of the (at most three) entries of the other particle,
// not of a per-walker table -- and a ballot compacts the accepted entries, in walk order, into the row.
template <bool HALF>
__device__ __forceinline__ int general_walk(const NbrParams &P, const TileDesc &T, const TileCellStarts &s_cs,
                                            const double2 *s_xy,
                                            const int *s_key, const double2 *__restrict__ snap,
                                            const int *__restrict__ cell_of, int ncy, int ew, int slot_i, int orow,
                                            int cxi, unsigned short *row_sm, int *surv, uint4 *ell_q,
                                            int *__restrict__ cnt, int dummy)
{
    const int lane = threadIdx.x & 31;
    const int ncx = P.ncx;
    const int i = s_key[slot_i] >> 3;
    const double2 pi = s_xy[slot_i];
    const int cyi = T.cy0 + orow, ri = orow + 1;
    const int home = cyi * ncx + cxi;
    const Entries Ei = entries_of(pi.x, home, P);
    // does i itself own ghost entries?  (then a lower fluid index may have met one of them)
    const bool i_multi = (pi.x > P.DL - P.cutoff || pi.x < P.cutoff) && ncx > 1;
    // the nine stencil cells in walk order (oy, ox), as slot ranges
    int rb[9], rn[9], rcx[9];
    int total = 0;
#pragma unroll
    for (int oy = -1; oy <= 1; ++oy) {
        const int cy = cyi + oy;
        int cx_prev[2] = {-1, -1};
#pragma unroll
        for (int ox = -1; ox <= 1; ++ox) {
            const int k = (oy + 1) * 3 + (ox + 1);
            int cx = cxi + ox;
            if (cx < 0) cx += ncx;
            else if (cx >= ncx) cx -= ncx;
            rb[k] = 0;
            rn[k] = 0;
            rcx[k] = cx;
            if (cy < 0 || cy >= ncy || cx < 0 || cx >= ncx) continue;
            // n_cell_x <= 2: the wrapped stencil revisits a column.  The reference re-walks it, but every
            // entry it accepted is already in seen_neighbor and every entry it rejected is rejected again
            // by the same arithmetic (nb:342-372), so the revisit emits nothing.
            if (cx == cx_prev[0] || cx == cx_prev[1]) continue;
            cx_prev[ox == -1 ? 0 : 1] = cx;
            int sb, se;
            tile_cell_slots(T, s_cs, ri + oy, cx, ncx, sb, se);
            rb[k] = sb;
            rn[k] = se - sb;
            total += se - sb;
        }
    }
    // Two passes, so that the lanes of the expensive pass are all busy.  Pass A runs over every candidate of the
    // stencil: the cheap tests (own entry, lower fluid index of a half list, clearly out of reach) leave about a third,
    // which a ballot compacts -- in walk order -- into the warp's survivor list.  Pass B makes the exact decisions
    // (dedup by walk position, replay of the lower index's walk) for the survivors only and compacts the accepted
    // ones into the row.  The list is flushed through pass B whenever it could overflow.
    int count = 0, n_surv = 0;
    auto decide_survivors = [&]() {
        __syncwarp();
        for (int base = 0; base < n_surv; base += 32) {
            const int t = base + lane;
            bool acc = false;
            int s = 0, m = 0;
            if (t < n_surv) {
                const int code = surv[t];
                s = code & 0xffff;
                const int k = code >> 16;
                int cx = cxi + (k % 3) - 1;
                if (cx < 0) cx += ncx;
                else if (cx >= ncx) cx -= ncx;
                const int cy = cyi + k / 3 - 1;
                const int rank_col = stencil_rank(cx, cxi, ncx);
                const int key = s_key[s];   // entries are stored last-inserted first
                const int j = key >> 3;
                const bool low_fluid = j < P.n_fluid && j < i;
                const double2 pe = s_xy[s];
                const bool seam = (key & 4) || (low_fluid && i_multi);
                if (!seam) {
                    acc = accept_entry(pi.x, pi.y, pe.x, pe.y, P, m);
                } else {
                    // j has up to three entries; the ones the walk reaches BEFORE this one decide whether
                    // j is already in the row (seen_neighbor, nb:342).  A real entry is j's wrapped position
                    // in j's home cell: only ghosts look them up.
                    const int kind = key & 3;
                    const double2 pj = kind == 0 ? pe : snap[j];
                    const int home_j = kind == 0 ? cy * ncx + cx : cell_of[j];
                    const Entries E = entries_of(pj.x, home_j, P);
                    const int rank_cur = rank_col * 3 + (2 - kind);
                    if (low_fluid) {
                        // the pair is decided by j's walk: the first entry of j the walk meets carries it
                        bool first = true;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            if (q == kind || E.cx[q] < 0) continue;
                            const int rk = stencil_rank(E.cx[q], cxi, ncx);
                            if (rk >= 0 && rk * 3 + (2 - q) < rank_cur) first = false;
                        }
                        if (first) {
                            const int sh = replay_low_walk(pj.x, pj.y, home_j, Ei, pi.y, ncy, P);
                            if (sh != kNoPair) {
                                acc = true;
                                m = -sh + (kind == 1 ? -1 : (kind == 2 ? 1 : 0));   // relative to the entry met
                            }
                        }
                    } else if (accept_entry(pi.x, pi.y, pe.x, pe.y, P, m)) {
                        bool dup = false;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            if (q == kind || E.cx[q] < 0) continue;
                            const int rk = stencil_rank(E.cx[q], cxi, ncx);
                            if (rk >= 0 && rk * 3 + (2 - q) < rank_cur) {
                                int m2;
                                dup |= accept_entry(pi.x, pi.y, E.x[q], pj.y, P, m2);
                            }
                        }
                        acc = !dup;
                    }
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, acc);
            if (acc) {
                const int pos = count + __popc(mask & ((1u << lane) - 1u));
                if (pos < P.ell_cap) row_sm[pos] = (unsigned short)ell_pack(s, m);
            }
            count += __popc(mask);
        }
        __syncwarp();
        n_surv = 0;
    };
    for (int base = 0; base < total; base += 32) {
        const int t = base + lane;
        bool keep = false;
        int code = 0;
        if (t < total) {
            int s = 0, kc = 0, pre = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (t >= pre && t < pre + rn[k]) {
                    s = rb[k] + (t - pre);
                    kc = k;
                }
                pre += rn[k];
            }
            const int key = s_key[s];
            const int j = key >> 3;
            const bool low_fluid = j < P.n_fluid && j < i;
            // Quick reject.  In a channel of >= 6 columns every pairing of the two particles' entries (real / ghost,
            // seen from either end) has the same minimum-image separation up to rounding (~1e-16 DL), so a candidate
            // that is beyond the cut-off by more than that is in nobody's row whatever entry a walk meets first; the
            // exact decisions of pass B (dedup, replay of the lower index's walk) are kept for the ones in reach.
            bool far = false;
            if (ncx >= 6) {
                const double2 pe = s_xy[s];
                double dxq = pi.x - pe.x;
                if (dxq > P.half_DL) dxq -= P.DL;
                else if (dxq < -P.half_DL) dxq += P.DL;
                const double dyq = pi.y - pe.y;
                far = dxq * dxq + dyq * dyq > P.rc2 * (1.0 + 1e-8);
            }
            keep = j != i && !far && !(HALF && low_fluid);                         // nb:353
            code = s | (kc << 16);
        }
        if (n_surv + 32 > kSurvCap) decide_survivors();
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) surv[n_surv + __popc(mask & ((1u << lane) - 1u))] = code;
        n_surv += __popc(mask);
    }
    decide_survivors();
    __syncwarp();
    const int nw = min((count + 7) >> 3, P.ell_cap >> 3);
    for (int w = lane; w < nw; w += 32) {
        unsigned e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = (8 * w + k < count) ? (unsigned)row_sm[8 * w + k] : ell_pack(dummy, 0);
        uint4 q;
        q.w = (e[0] << 16) | e[1];
        q.z = (e[2] << 16) | e[3];
        q.y = (e[4] << 16) | e[5];
        q.x = (e[6] << 16) | e[7];
        ell_q[(size_t)w * P.ell_stride + ew] = q;
    }
    if (lane == 0) cnt[ew] = count;
    __syncwarp();
    return count;
}

// What the walk needs to do the work of the former k_cell_canon and k_tile_setup launches for its own tile.
struct WalkSetup {
    const int *tmp_key;        // entries as scattered: any order inside a cell
    const double2 *tmp_pos;
    int *ent_key;              // canonical order (written for the tile's own cells)
    double2 *ent_pos;
    int *ent_of;
    unsigned n_tiles;
    int slab;
};

// slot range [sb, se) of the staged cell that holds slot s (main segments: binary search of the entry in the
// staged cell starts of its row; a wrap segment is one cell)
__device__ __forceinline__ void staged_cell_of_slot(const TileDesc &T, const TileCellStarts &cs, int s, int &sb, int &se)
{
    int k = 0;
#pragma unroll
    for (int j = 1; j < kStageRows * 3; ++j) k += (T.seg[j].off <= s) ? 1 : 0;
    const TileSeg sg = T.seg[k];
    if (k % 3 != 0) {
        sb = sg.off;
        se = sg.off + (sg.e - sg.b);
        return;
    }
    const int r = k / 3;
    const int e = sg.b + (s - sg.off);
    int lo = 0, hi = T.b - T.a;   // cs[r][lo] <= e < cs[r][hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cs[r][mid] <= e) lo = mid;
        else hi = mid;
    }
    sb = sg.off + (cs[r][lo] - sg.b);
    se = sg.off + (cs[r][lo + 1] - sg.b);
}

template <bool HALF>
__device__ __forceinline__ void walk_tile_body(NbrParams P, const GridDesc *__restrict__ g, TileDesc *tiles, BuildMirrors M,
                                               WalkSetup W, int tile_w, const int *__restrict__ cell_start,
                                               const double2 *__restrict__ snap, const int *__restrict__ cell_of,
                                               uint4 *ell_q, int *__restrict__ cnt, int *__restrict__ totals,
                                               int *__restrict__ flags, int cap)
{
    load_counts(P);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ TileDesc T;
    double2 *s_xy = reinterpret_cast<double2 *>(smem_raw);           // layout: s_xy, s_rawx, s_key, s_rawk, general-walk rows
    int *s_key = reinterpret_cast<int *>(s_xy + 2 * (size_t)cap);
    const int ncx = P.ncx, ncy = g->ncy;
    __shared__ TileCellStarts s_cs;
    const int tpr = tiles_per_row(ncx, tile_w);
    const unsigned tile_id = walk_tile_of_block(blockIdx.x, tpr, gridDim.x);
    // The first warp builds the tile's descriptor (what k_tile_setup used to write for every tile in a launch of
    // its own) in shared memory and copies it to the global array; the other warps fetch the cell starts of the
    // staged rows meanwhile (both only need cell_start).
    if (threadIdx.x < 32) {
        tile_setup_warp(tile_id, (int)threadIdx.x, g, ncx, W.slab, tile_w, W.n_tiles, cap, (int)P.ell_stride, cell_start, T,
                        tiles, totals, flags);
    } else {
        const int cy0 = (int)(tile_id / (unsigned)tpr) * kTileRows, cx0 = (int)(tile_id % (unsigned)tpr) * tile_w;
        const int a = max(cx0 - 1, 0), b = min(min(cx0 + tile_w, ncx) + 1, ncx);
        for (int k = (int)threadIdx.x - 32; k < kStageRows * (kMaxTileW + 3); k += kTileThreads - 32) {
            const int r = k / (kMaxTileW + 3), c = k % (kMaxTileW + 3);
            const int row = cy0 + r - 1;
            s_cs[r][c] = (row >= 0 && row < ncy && c <= b - a) ? cell_start[(size_t)row * ncx + a + c] : 0;
        }
        if (blockIdx.x == 0 && threadIdx.x == 32 && g->n_entries > (int)P.ell_stride) atomicOr(flags, kErrCsrCap);
    }
    __syncthreads();
    if (!(T.flags & kTileValid)) return;
    int walls = 0;   // does the staged neighbourhood hold a wall entry?  (kTileHasWall, read by the gather kernels)
    // Staging with the canonical order made on the way (the former k_cell_canon): inside a cell the reference's
    // linked list is walked "last inserted first", i.e. by DEcreasing insertion key (keys are distinct), so a
    // forward scan of the canonical range is its walk order.  Pass A stages keys and positions as scattered; pass B
    // ranks every entry inside its cell (the number of larger keys) and moves it to the ranked slot, shared memory
    // to shared memory.  Every tile that stages a cell derives the same order; the tile that OWNS the cell writes
    // it to the entry arrays the gather kernels stage from.
    double2 *s_rawx = s_xy + cap;
    int *s_rawk = s_key + cap;
    {
        struct StageRec {
            double2 xy;
            int key;
        };
        tile_stage_flat(
            T,
            [&](int e) {
                StageRec r;
                r.key = W.tmp_key[e];
                r.xy = W.tmp_pos[e];
                return r;
            },
            [&](int slot, const StageRec &r) {
                s_rawk[slot] = r.key;
                s_rawx[slot] = r.xy;
            });
    }
    __syncthreads();
    for (int slot = (int)threadIdx.x; slot < T.n_staged; slot += kTileThreads) {
        const int key = s_rawk[slot];
        int sb, se;
        staged_cell_of_slot(T, s_cs, slot, sb, se);
        int rank = 0;
        for (int b = sb; b < se; ++b) rank += s_rawk[b] > key ? 1 : 0;
        s_key[sb + rank] = key;
        s_xy[sb + rank] = s_rawx[slot];
        walls |= (key >> 3) >= P.n_fluid ? 1 : 0;
    }
    walls = __syncthreads_or(walls);
    if (walls && threadIdx.x == 0) tiles[tile_id].flags = T.flags | kTileHasWall;

    const size_t stride = P.ell_stride;
    const int qcap = P.ell_cap >> 3;
    const bool seam_tile = (T.flags & kTileSeam) != 0;
    const bool ghost_anomaly = g->ghost_anomaly != 0u;   // a ghost entry outside the usual columns: no fast walk
    const int dummy = T.n_staged;
    const int lane = threadIdx.x & 31;
    __shared__ int gq[kTileThreads];   // own-entry indices of the queued general walkers
    __shared__ int s_surv[kTileThreads / 32][kSurvCap];   // per warp: survivors between the passes of a general walk
    __shared__ int gq_n;
    if (threadIdx.x == 0) gq_n = 0;
    __syncthreads();
    unsigned short *row_sm = reinterpret_cast<unsigned short *>(s_key + 2 * (size_t)cap) + (size_t)(threadIdx.x >> 5) * P.ell_cap;
    int t_full = 0, t_max = 0;
    const int n_own = tile_own_total(T);
    for (int base = 0; base < n_own; base += kTileThreads) {   // warp-uniform trip count (collectives below)
        const int v = base + (int)threadIdx.x;
        bool general = false;   // this lane's walker takes the warp-cooperative general walk
        int ew = 0, slot_i = 0, orow = 0, cxi = 0;
        if (v < n_own) {
            tile_own_locate(T, v, orow, ew, slot_i);
            const int key_i = s_key[slot_i];
            const int i = key_i >> 3;
            if ((size_t)ew < P.ell_stride) {   // the canonical entry arrays of the tile's own cells
                W.ent_key[ew] = key_i;
                W.ent_pos[ew] = s_xy[slot_i];
                if ((key_i & 3) == 0) W.ent_of[i] = ew;
            }
            // entry-ordered mirrors of the row's velocity and wall weight (BuildMirrors): loaded here, stored
            // after the walk so that the latency hides behind it
            // (ghost entries get the same mirror values as the real entry of their particle, so that a seam tile is
            // staged with the same bulk copies as any other; the ghost's entry index is recorded for the kernels
            // that produce the other mirrors)
            const bool mirror = M.e_vel != nullptr && (size_t)ew < stride;
            if (mirror && (key_i & 3) != 0 && M.ghost_ent != nullptr) M.ghost_ent[2 * i + (key_i & 3) - 1] = ew;
            double2 mv = make_double2(0.0, 0.0);
            double mw = -1.0;
            if (mirror) {
                const bool fluid = i < P.n_fluid;
                mv = (fluid ? M.vel : M.wall_vel)[i];
                if (!fluid) mw = M.mass[i] * M.inv_rho0;
            }
            if ((size_t)ew >= stride) {   // more entries than list columns: reported, never written
                atomicOr(flags, kErrCsrCap);
            } else if ((key_i & 3) != 0 || i >= P.n_fluid) {   // ghost or wall entry: not a walker
                cnt[ew] = 0;
            } else {
                const int ri = orow + 1;   // staged row of the walker
                {   // the walker's column: binary search of its entry in the staged cell starts of its row
                    int lo = T.cx0 - T.a, hi = T.cx1 - T.a;   // cs[ri][lo] <= ew < cs[ri][hi]
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (s_cs[ri][mid] <= ew) lo = mid;
                        else hi = mid;
                    }
                    cxi = T.a + lo;
                }
                // Column ncx-2 holds two kinds of walkers: those within the cut-off of x = DL own a ghost entry and meet
                // the ghosts of column 0 (general walk); the others are to column ncx-1 what a walker of column 1 is to
                // column 0 -- no entry of their stencil has a second entry inside it, nothing wraps -- and take the
                // interior walk (with a sliver last column that is about half of them).  The margin keeps the ghosts of
                // column 0 (at x >= DL) clearly out of their reach: within rounding of the cut-off a pair with a lower
                // index is decided by THAT particle's walk, which does not reach column ncx-2 (found by the closure fuzz).
                const bool interior = ncx >= 6 && !ghost_anomaly && cxi >= 1 &&
                                      (cxi <= ncx - 3 ||
                                       (cxi == ncx - 2 && !(s_xy[slot_i].x > P.DL - P.cutoff * (1.0 + 1e-9))));
                if (!seam_tile && !interior) {   // slab mode: a particle clamped into the empty margin columns
                    cnt[ew] = 0;
                } else if (!interior) {
                    general = true;
                } else {
                    // ---- interior walker: per stencil row one contiguous slot range; first a branch-free
                    // pass that records the accepted candidates as a bit mask, then the mask is unrolled
                    // into the row
                    const double2 pi = s_xy[slot_i];
                    const int cyi = T.cy0 + orow;
                    RowWriter row;
                    row.init(ell_q + ew, stride, qcap);
#pragma unroll 1
                    for (int r = ri - 1; r <= ri + 1; ++r) {
                        const int cy = cyi + (r - ri);
                        if (cy < 0 || cy >= ncy) continue;
                        const int k0 = cxi - 1 - T.a;
                        const int c0 = s_cs[r][k0];
                        const int sb = T.seg[3 * r].off + (c0 - T.seg[3 * r].b);
                        const int se = sb + (s_cs[r][k0 + 3] - c0);
#pragma unroll 1
                        for (int s0 = sb; s0 < se; s0 += 32) {
                            const int s1 = min(s0 + 32, se);
                            unsigned mask = 0u, edge = 0u, bit = 1u;
                            for (int s = s0; s < s1; ++s, bit <<= 1) {
                                const double2 pj = s_xy[s];
                                const double dx = pi.x - pj.x;
                                const double dy = pi.y - pj.y;
                                const double r2 = dx * dx + dy * dy;
                                const bool in = r2 > 1e-24 && r2 < P.rc2;       // also rejects the walker itself
                                mask |= in ? bit : 0u;
                                edge |= (in && !(r2 < P.rc2_safe)) ? bit : 0u;
                            }
                            while (edge) {   // last 1e-12 below the cut-off: the reference's kernel-support test
                                const int k = __ffs((int)edge) - 1;
                                edge &= edge - 1;
                                const double2 pj = s_xy[s0 + k];
                                const double dx = pi.x - pj.x;
                                const double dy = pi.y - pj.y;
                                if (!support_exact(dx * dx + dy * dy, P)) mask &= ~(1u << k);
                            }
                            while (mask) {
                                const int k = __ffs((int)mask) - 1;
                                mask &= mask - 1;
                                if (HALF) {
                                    const int j = s_key[s0 + k] >> 3;
                                    if (j < P.n_fluid && j < i) continue;   // nb:353
                                }
                                row.push(s0 + k, 0);
                            }
                        }
                    }
                    row.finish(dummy);
                    cnt[ew] = row.count;
                    t_full += row.count;
                    t_max = max(t_max, row.count);
                }
            }
            if (mirror) {
                M.e_vel[ew] = mv;
                M.e_w[ew] = make_double2(mw, 0.0);
            }
        }
        // ---- general walkers (periodic seam): queued for the whole CTA, so that the few warps that own
        // them do not walk them alone
        if (general) gq[atomicAdd(&gq_n, 1)] = v;
        if (seam_tile) {
            __syncthreads();
            const int nq = gq_n;
            for (int q = (int)(threadIdx.x >> 5); q < nq; q += kTileThreads / 32) {
                int w_orow, w_ew, w_slot;
                tile_own_locate(T, gq[q], w_orow, w_ew, w_slot);
                int lo = T.cx0 - T.a, hi = T.cx1 - T.a;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (s_cs[w_orow + 1][mid] <= w_ew) lo = mid;
                    else hi = mid;
                }
                const int n = general_walk<HALF>(P, T, s_cs, s_xy, s_key, snap, cell_of, ncy, w_ew, w_slot, w_orow, T.a + lo,
                                                 row_sm, s_surv[threadIdx.x >> 5], ell_q, cnt, dummy);
                if (lane == 0) {
                    t_full += n;
                    t_max = max(t_max, n);
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) gq_n = 0;
            __syncthreads();
        }
    }
    // totals: entries, longest row (capacity check)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        t_full += __shfl_xor_sync(0xffffffffu, t_full, d);
        t_max = max(t_max, __shfl_xor_sync(0xffffffffu, t_max, d));
    }
    if ((threadIdx.x & 31) == 0) {
        if (t_full) atomicAdd(&totals[1], t_full);
        if (HALF && t_full) atomicAdd(&totals[0], t_full);
        if (t_max) atomicMax(&totals[2], t_max);
        if (t_max > P.ell_cap) atomicOr(flags, kErrCsrCap);
    }
}
#ifndef SPHB_WALK_MINCTAS
#define SPHB_WALK_MINCTAS 8   // resident CTAs per SM the walk is compiled for (r2 call 36, walk alone: 4 -> 102.5 us (128 registers), 5 -> 93.0, 6 -> 88.5 (80 registers), 8 -> 82.3 (64 registers, 112 B of spills))
#endif
template <bool HALF>
__global__ void __launch_bounds__(kTileThreads, SPHB_WALK_MINCTAS) k_walk_tile(NbrParams P, const GridDesc *__restrict__ g, TileDesc *tiles,
                                                           BuildMirrors M, WalkSetup W, int tile_w,
                                                           const int *__restrict__ cell_start,
                                                           const double2 *__restrict__ snap,
                                                           const int *__restrict__ cell_of, uint4 *ell_q,
                                                           int *__restrict__ cnt, int *__restrict__ totals,
                                                           int *__restrict__ flags, int cap)
{
    pdl_wait();
    pdl_trigger_top();
    walk_tile_body<HALF>(P, g, tiles, M, W, tile_w, cell_start, snap, cell_of, ell_q, cnt, totals, flags, cap);
    pdl_trigger_end();
}
// full lists only (the half list is the stateless MEX export)
__global__ void __launch_bounds__(kTileThreads, SPHB_WALK_MINCTAS) k_walk_tile_batch(const NbrChan *__restrict__ chan)
{
    const NbrChan &c = chan[blockIdx.y];
    WalkSetup W;
    W.tmp_key = c.tmp_key;
    W.tmp_pos = c.tmp_pos;
    W.ent_key = c.ent_key;
    W.ent_pos = c.ent_pos;
    W.ent_of = c.ent_of;
    W.n_tiles = c.n_tiles;
    W.slab = c.P.slab;
    walk_tile_body<false>(c.P, c.g, c.tiles, c.M, W, c.tile_w, c.cell_start, c.snap, c.cell_of, c.ell_q, c.ell_cnt, c.totals,
                          c.flags, c.tile_cap);
}

// Number of reference half pairs inside a FULL row set (status / export sizing only): entries whose
// neighbour is a wall particle or a higher fluid index.
__global__ void __launch_bounds__(kTileThreads) k_half_total(NbrParams P, const TileDesc *__restrict__ tiles,
                                                            const int *__restrict__ ent_key,
                                                            const uint4 *__restrict__ ell_q,
                                                            const int *__restrict__ cnt, int *__restrict__ totals,
                                                            int cap)
{
    load_counts(P);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ TileDesc T;
    int *s_key = reinterpret_cast<int *>(smem_raw);
    tile_load(T, tiles, blockIdx.x);
    __syncthreads();
    if (!(T.flags & kTileValid)) return;
    SPHB_TILE_FOR_STAGED(T, e, slot) s_key[slot] = ent_key[e];
    __syncthreads();
    int n_half = 0;
    SPHB_TILE_FOR_OWN(T, ew, si, orow)
    {
        if ((size_t)ew >= P.ell_stride) continue;
        const int i = s_key[si] >> 3;
        const int n = min(cnt[ew], P.ell_cap);
        for (int k = 0; k < n; ++k) {
            const uint4 q = ell_q[(size_t)(k >> 3) * P.ell_stride + ew];
            const int j = s_key[ell_slot(ell_extract(q, k & 7))] >> 3;
            n_half += (j >= P.n_fluid || j > i) ? 1 : 0;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n_half += __shfl_xor_sync(0xffffffffu, n_half, d);
    if ((threadIdx.x & 31) == 0 && n_half) atomicAdd(&totals[0], n_half);
}

// ---- export paths (not hot): rows by particle out of the tile-local lists --------------------------------
__device__ __forceinline__ const TileDesc &tile_of_cell(const TileDesc *tiles, int cell, int ncx, int w)
{
    return tiles[(size_t)((cell / ncx) / kTileRows) * tiles_per_row(ncx, w) + (cell % ncx) / w];
}

// k-th entry of the row of the walker at entry ew -> global entry index and particle
__device__ __forceinline__ int row_entry(const NbrParams &P, const TileDesc &T, const uint4 *__restrict__ ell_q, int ew, int k)
{
    const uint4 q = ell_q[(size_t)(k >> 3) * P.ell_stride + ew];
    return tile_slot_entry(T, ell_slot(ell_extract(q, k & 7)));
}

// rows of the reference's half list inside a FULL row set: count of entries with j wall or j > i
__global__ void __launch_bounds__(kBlock) k_half_count(NbrParams P, int tile_w, const TileDesc *__restrict__ tiles,
                                                      const int *__restrict__ cell_of,
                                                      const int *__restrict__ ent_of,
                                                      const int *__restrict__ ent_key,
                                                      const uint4 *__restrict__ ell_q,
                                                      const int *__restrict__ cnt, int *__restrict__ half_cnt)
{
    const int n_launch = P.n_fluid;   // slab mode: the capacity; rows beyond the live count get empty rows
    load_counts(P);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_launch) return;
    if (i >= P.n_fluid || row_dead(P, i)) {
        half_cnt[i] = 0;
        return;
    }
    const int e = ent_of[i];
    const TileDesc &T = tile_of_cell(tiles, cell_of[i], P.ncx, tile_w);
    const int n = (T.flags & kTileValid) ? min(cnt[e], P.ell_cap) : 0;
    int c = 0;
    for (int k = 0; k < n; ++k) {
        const int j = ent_key[row_entry(P, T, ell_q, e, k)] >> 3;
        c += (j >= P.n_fluid || j > i) ? 1 : 0;
    }
    half_cnt[i] = c;
}

// nb:357-381 re-evaluated for the stored entry: identical operations, identical bits
__global__ void __launch_bounds__(kBlock) k_export_pairs(NbrParams P, int tile_w, const TileDesc *__restrict__ tiles,
                                                        const int *__restrict__ cell_of,
                                                        const int *__restrict__ ent_of,
                                                        const int *__restrict__ ent_key,
                                                        const double2 *__restrict__ ent_pos,
                                                        const uint4 *__restrict__ ell_q,
                                                        const int *__restrict__ cnt,
                                                        const int *__restrict__ half_ptr,
                                                        double *__restrict__ pair_i, double *__restrict__ pair_j,
                                                        double *__restrict__ odx, double *__restrict__ ody,
                                                        double *__restrict__ orr, double *__restrict__ oW,
                                                        double *__restrict__ odW)
{
    load_counts(P);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_fluid || row_dead(P, i)) return;
    const int ew = ent_of[i];
    const TileDesc &T = tile_of_cell(tiles, cell_of[i], P.ncx, tile_w);
    const double2 pi = ent_pos[ew];
    const int n = (T.flags & kTileValid) ? min(cnt[ew], P.ell_cap) : 0;
    size_t o = (size_t)half_ptr[i];
    for (int k = 0; k < n; ++k) {
        const int ea = row_entry(P, T, ell_q, ew, k);
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
__global__ void __launch_bounds__(kBlock) k_row_count(NbrParams P, int tile_w, const TileDesc *__restrict__ tiles,
                                                     const int *__restrict__ cell_of,
                                                     const int *__restrict__ ent_of,
                                                     const int *__restrict__ cnt, int *__restrict__ row_cnt)
{
    const int n_launch = P.n_fluid;   // slab mode: the capacity; rows beyond the live count get empty rows
    load_counts(P);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_launch) return;
    if (i >= P.n_fluid || row_dead(P, i)) {
        row_cnt[i] = 0;
        return;
    }
    const TileDesc &T = tile_of_cell(tiles, cell_of[i], P.ncx, tile_w);
    row_cnt[i] = (T.flags & kTileValid) ? min(cnt[ent_of[i]], P.ell_cap) : 0;
}
__global__ void __launch_bounds__(kBlock) k_row_fill(NbrParams P, int tile_w, const TileDesc *__restrict__ tiles,
                                                    const int *__restrict__ cell_of,
                                                    const int *__restrict__ ent_of,
                                                    const int *__restrict__ ent_key,
                                                    const uint4 *__restrict__ ell_q, const int *__restrict__ cnt,
                                                    const int *__restrict__ row_ptr, int *__restrict__ out)
{
    load_counts(P);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_fluid || row_dead(P, i)) return;
    const int e = ent_of[i];
    const TileDesc &T = tile_of_cell(tiles, cell_of[i], P.ncx, tile_w);
    const int n = (T.flags & kTileValid) ? min(cnt[e], P.ell_cap) : 0;
    const size_t o = (size_t)row_ptr[i];
    for (int k = 0; k < n; ++k) out[o + (size_t)k] = ent_key[row_entry(P, T, ell_q, e, k)] >> 3;
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
    int ncx = w.slab ? w.ncx_local : (int)std::ceil(w.DL / P.cell);   // nb:254
    if (ncx < 1) ncx = 1;
    P.ncx = ncx;
    P.slab = w.slab ? 1 : 0;
    P.col0 = w.col0;
    P.alive = w.slab ? w.alive : nullptr;
    P.dev_counts = w.slab ? w.dev_counts : nullptr;
    P.cell_cap = w.cell_cap;
    P.ell_cap = w.ell_cap;
    P.ell_stride = w.ell_stride;
    return P;
}

constexpr size_t kWalkSlotBytes = sizeof(double2) + sizeof(int);

int set_walk_smem(size_t bytes)
{
    static SmemGrant g[2];
    SPHB_TRY(grant_smem(g[0], k_walk_tile<true>, bytes));
    SPHB_TRY(grant_smem(g[1], k_walk_tile<false>, bytes));
    return 0;
}

}  // namespace

void NeighborWork::release_all()
{
    DevBuf *all[] = {&grid, &flags, &snap, &cell_of, &ent_of, &ghost_ent, &cell_cnt, &cell_start, &tmp_key, &tmp_pos,
                     &ent_key, &ent_pos, &tiles, &ell_cnt, &ell_q, &half_cnt, &half_ptr, &scan_tmp, &totals};
    for (DevBuf *b : all) b->release();
}

// Rows are indexed by ENTRY (real + ghost): the stride is the entry capacity.  Ghost entries are a
// sliver of the total unless the box is only a few cells long, so large problems get n_total * 1.25
// columns, small ones the reference's own worst case 3 n_total (nb:265).
static size_t entry_capacity(size_t nt)
{
    size_t cap = nt < (1u << 20) ? 3 * nt : nt + nt / 4 + 4096;
    return (cap + 31) & ~(size_t)31;
}

int neighbor_reserve_ell(NeighborWork &w, int rows)
{
    rows = (rows + 7) & ~7;
    const size_t stride = entry_capacity((size_t)w.n_total);
    if (rows <= w.ell_cap && w.ell_q.p && stride == w.ell_stride) return 0;
    if (rows < w.ell_cap) rows = w.ell_cap;
    w.ell_stride = stride;
    SPHB_TRY(w.ell_q.reserve((size_t)(rows / 8) * stride * sizeof(uint4)));
    SPHB_TRY(w.ell_cnt.reserve(stride * sizeof(int)));
    w.ell_cap = rows;
    return 0;
}

// Tile geometry: `w` cells per tile, descriptors for every cell row the grid may have.
static int plan_tiles(NeighborWork &w, int tile_w)
{
    const int ncx = w.slab ? std::max(1, w.ncx_local) : std::max(1, (int)std::ceil(w.DL / (2.0 * w.h)));
    if (tile_w > kMaxTileW) tile_w = kMaxTileW;
    if (tile_w < 1) tile_w = 1;
    w.tile_w = tile_w;
    w.n_tiles = (unsigned)(tile_rows(w.cell_cap / ncx) * tiles_per_row(ncx, tile_w));
    SPHB_TRY(w.tiles.reserve((size_t)w.n_tiles * sizeof(TileDesc)));
    return 0;
}

// Size every buffer whose size is known a priori.  y hints bound the cell rows (a fluid particle
// escaping beyond the hinted range raises the cell-capacity flag instead of writing out of bounds).
int neighbor_plan(NeighborWork &w, int n_fluid, int n_total, double h, double DL, double y_min_hint,
                  double y_max_hint, int max_nbr_hint, int tile_cap_limit)
{
    const bool resized = n_total != w.n_total || h != w.h || DL != w.DL;
    w.n_fluid = n_fluid;
    w.n_total = n_total;
    w.h = h;
    w.DL = DL;
    const double cell = 2.0 * h;
    const double ncx_d = w.slab ? (double)w.ncx_local : std::ceil(DL / cell);
    const double ncy_d = std::ceil((y_max_hint - y_min_hint + 1e-12) / cell) + 1.0 + 8.0;   // slack rows
    if (!(ncx_d >= 1.0) || !(ncy_d >= 1.0) || ncx_d * ncy_d > 1.0e9)
        return set_error("SPH:Neighbor:entryCapacity", "cell grid of %g x %g cells is not representable",
                         ncx_d, ncy_d);
    if (3LL * (long long)n_total >= (1LL << 28))
        return set_error("SPH:Neighbor:entryCapacity", "3 n_total exceeds the 2^28 limit of the entry index");
    const int new_cell_cap = (int)(ncx_d * ncy_d);
    const bool regrid = new_cell_cap != w.cell_cap;
    w.cell_cap = new_cell_cap;
    const size_t nt = (size_t)n_total;
    const size_t ecap = entry_capacity(nt);
    if (!w.grid.p) w.grid_armed = false;
    SPHB_TRY(w.grid.reserve(sizeof(GridDesc)));
    if (!w.flags.p) {
        SPHB_TRY(w.flags.reserve(sizeof(int)));
        SPHB_CUDA(cudaMemset(w.flags.p, 0, sizeof(int)));
    }
    SPHB_TRY(w.totals.reserve(8 * sizeof(int)));
    SPHB_TRY(w.snap.reserve(nt * sizeof(double2)));
    SPHB_TRY(w.cell_of.reserve(nt * sizeof(int)));
    SPHB_TRY(w.ent_of.reserve(nt * sizeof(int)));
    if (w.ghost_ent.bytes < 2 * nt * sizeof(int)) {
        SPHB_TRY(w.ghost_ent.reserve(2 * nt * sizeof(int)));
        SPHB_CUDA(cudaMemset(w.ghost_ent.p, 0xFF, 2 * nt * sizeof(int)));   // -1: no ghost entry known
    }
    SPHB_TRY(w.cell_cnt.reserve(2 * ((size_t)w.cell_cap + 1) * sizeof(int)));
    SPHB_TRY(w.cell_start.reserve(((size_t)w.cell_cap + 1) * sizeof(int)));
    SPHB_TRY(w.tmp_key.reserve(3 * nt * sizeof(int)));
    SPHB_TRY(w.tmp_pos.reserve(3 * nt * sizeof(double2)));
    SPHB_TRY(w.ent_key.reserve(ecap * sizeof(int)));
    SPHB_TRY(w.ent_pos.reserve(ecap * sizeof(double2)));
    if (resized) {
        w.ell_cap = 0;   // the stride depends on n_total
        w.ell_q.release();
    }
    SPHB_TRY(neighbor_reserve_ell(w, max_nbr_hint > 0 ? max_nbr_hint : kDefaultEllRows));
    w.tile_cap_limit = tile_cap_limit > 0 && tile_cap_limit < kMaxSlots ? tile_cap_limit : kMaxSlots;
    {   // what the walk itself can stage: two copies of (position, key) per slot (as scattered + canonical) beside the
        // rows under assembly of the general walkers, in 227 KB of shared memory
        const size_t rows = std::max((size_t)(kTileThreads / 32) * (size_t)w.ell_cap * sizeof(unsigned short), (size_t)16 * 1024);
        const int walk_limit = (int)((227 * 1024 - 4096 - std::min(rows, (size_t)96 * 1024)) / (2 * kWalkSlotBytes));
        if (w.tile_cap_limit > walk_limit) w.tile_cap_limit = walk_limit;
    }
    if (resized || regrid || w.tile_w == 0) {
        w.tile_fitted = false;
        // ~0.9 own entries per thread for a uniformly filled grid; the first exact build narrows the
        // tiles if the densest neighbourhood does not fit the shared-memory slots
        const double per_cell = w.per_cell_hint > 0.0 ? w.per_cell_hint
                                                      : (double)n_total / std::max(1.0, ncx_d * (ncy_d - 9.0));
        int tw = (int)std::floor(0.955 * kTileThreads / std::max(per_cell * kTileRows, 1e-9));
        if (const char *env = std::getenv("SPHB_TILE_W")) {   // tuning runs only: force the tile width (cells)
            const int forced = std::atoi(env);
            if (forced > 0) tw = forced;
        }
        SPHB_TRY(plan_tiles(w, std::max(1, std::min(tw, kMaxTileW))));
        const double guess = (double)kStageRows * (w.tile_w + 2) * per_cell * 1.25 + 64.0;
        w.tile_cap = (int)std::min((double)w.tile_cap_limit, guess);
    } else {
        SPHB_TRY(plan_tiles(w, w.tile_w));
    }
    return 0;
}

GridFuse neighbor_grid_fuse(const NeighborWork &w)
{
    GridFuse F;
    if (!w.grid_armed || !w.grid.p || !w.cell_cnt.p) return F;
    const NbrParams P = make_params(w);
    F.g = w.grid.as<GridDesc>();
    F.flags = w.flags.as<int>();
    F.totals = w.totals.as<int>();
    F.cell_cnt = w.cell_cnt.as<int>();
    F.n_cell_ints = 2 * (w.cell_cap + 1);
    F.cell = P.cell;
    F.ncx = P.ncx;
    F.cell_cap = P.cell_cap;
    F.derive = w.slab ? 0 : 1;
    return F;
}

// Builds the list from `pos` (device, double2[n_total]).  A row longer than the list columns or a
// neighbourhood larger than the tile slots raises a capacity flag; with `exact` the build then
// synchronises, grows what was short (or narrows the tiles) and walks again (stateless path and the
// first resident build); in steady state there is no host round trip.
int neighbor_build(NeighborWork &w, const double2 *pos, bool half_only, bool exact, cudaStream_t st,
                   long long *launches, Prof *prof)
{
    NbrParams P = make_params(w);
    GridDesc *g = w.grid.as<GridDesc>();
    int *flags = w.flags.as<int>();
    const int nt = w.n_total;
    long long nl = 0;
    int *cell_cnt = w.cell_cnt.as<int>();
    int *cell_fill = cell_cnt + ((size_t)w.cell_cap + 1);

    {
        ProfScope ps(prof, kPfNbrGrid, st);
        if (!w.grid_armed) {   // once per plan; afterwards every build re-arms the reduction itself
            k_grid_init<<<1, 1, 0, st>>>(g);
            w.grid_armed = true;
        }
        if (!w.grid_ready) {   // (else the caller's continuity kernel has done both: GridFuse)
            int blocks = (nt + 255) / 256;
            if (blocks > 148 * 4) blocks = 148 * 4;
            k_yrange_grid<<<blocks, 256, 0, st>>>(pos, nt, g, P, flags, w.totals.as<int>());
            SPHB_CUDA(cudaMemsetAsync(cell_cnt, 0, 2 * ((size_t)w.cell_cap + 1) * sizeof(int), st));
            nl += 1;
        }
        w.grid_ready = false;
    }
    {
        ProfScope ps(prof, kPfNbrBin, st);
        SPHB_LAUNCH_PDL(k_bin_count, grid_for(nt), kBlock, 0, st, pos, P, g, w.snap.as<double2>(), w.cell_of.as<int>(), cell_cnt,
                        &g->ghost_anomaly, flags);
    }
    nl += 1;
    {
        ProfScope ps(prof, kPfNbrScan, st);
        SPHB_TRY(exclusive_scan_i32(cell_cnt, w.cell_start.as<int>(), (size_t)w.cell_cap, w.scan_tmp, &g->n_entries, st,
                                    &nl));
    }
    {
        ProfScope ps(prof, kPfNbrScatter, st);
        SPHB_LAUNCH_PDL(k_scatter, grid_for(nt), kBlock, 0, st, pos, P, g, w.cell_start.as<int>(), cell_fill, w.tmp_key.as<int>(),
                        w.tmp_pos.as<double2>(), flags);
    }
    nl += 1;
    WalkSetup W;
    W.tmp_key = w.tmp_key.as<int>();
    W.tmp_pos = w.tmp_pos.as<double2>();
    W.ent_key = w.ent_key.as<int>();
    W.ent_pos = w.ent_pos.as<double2>();
    W.ent_of = w.ent_of.as<int>();
    W.slab = P.slab;

    for (int attempt = 0; attempt < 8; ++attempt) {
        ProfScope ps(prof, kPfNbrWalk, st);
        if (attempt > 0) SPHB_CUDA(cudaMemsetAsync(w.totals.p, 0, 8 * sizeof(int), st));
        W.n_tiles = w.n_tiles;
        // staged keys + positions, the rows under assembly of the general walkers, the keys as scattered
        // shared memory of the walk: two copies of the staged (position, key), the rows under assembly of the general
        // walkers (they grow with the longest row), ~2.4 KB static.  Rows that have grown shrink what a tile may stage.
        const size_t rows_bytes = (size_t)(kTileThreads / 32) * w.ell_cap * sizeof(unsigned short);
        const size_t budget = (size_t)227 * 1024 - 4096;
        if (rows_bytes + 2 * kWalkSlotBytes * 64 > budget)
            return set_error("SPH:Neighbor:entryCapacity", "rows of up to %d neighbours exceed the shared memory of the "
                             "neighbour walk", w.ell_cap);
        const int stage_limit = (int)((budget - rows_bytes) / (2 * kWalkSlotBytes));
        if (w.tile_cap_limit > stage_limit) w.tile_cap_limit = stage_limit;
        if (w.tile_cap > w.tile_cap_limit) w.tile_cap = w.tile_cap_limit;   // (a tile that no longer fits is reported by the walk)
        const size_t smem = 2 * (size_t)w.tile_cap * kWalkSlotBytes + rows_bytes;
        SPHB_TRY(set_walk_smem(smem));
        // one CTA per tile (descriptor + canonical cell order + walk); rows beyond the device-side n_cell_y exit at once
        if (half_only)
            k_walk_tile<true><<<w.n_tiles, kTileThreads, smem, st>>>(
                P, g, w.tiles.as<TileDesc>(), BuildMirrors(), W, w.tile_w, w.cell_start.as<int>(), w.snap.as<double2>(),
                w.cell_of.as<int>(), w.ell_q.as<uint4>(), w.ell_cnt.as<int>(), w.totals.as<int>(), flags, w.tile_cap);
        else
            SPHB_LAUNCH_PDL(k_walk_tile<false>, w.n_tiles, kTileThreads, smem, st, P, g, w.tiles.as<TileDesc>(), w.mirrors, W,
                            w.tile_w, w.cell_start.as<int>(), w.snap.as<double2>(), w.cell_of.as<int>(), w.ell_q.as<uint4>(),
                            w.ell_cnt.as<int>(), w.totals.as<int>(), flags, w.tile_cap);
        nl += 1;
        if (exact) {   // (not inside a stream capture)
            const cudaError_t le = cudaGetLastError();
            if (le != cudaSuccess)
                return set_error("SPHB:cuda", "neighbour walk launch failed: %s (%u tiles of width %d, %d slots, rows of %d, "
                                 "%zu bytes of shared memory)", cudaGetErrorString(le), w.n_tiles, w.tile_w, w.tile_cap,
                                 w.ell_cap, smem);
        }
        if (!exact) break;
        int host[4] = {0, 0, 0, 0};
        SPHB_CUDA(cudaMemcpyAsync(host, w.totals.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        SPHB_CUDA(cudaStreamSynchronize(st));
        const bool rows_short = host[2] > w.ell_cap;
        const bool tile_short = host[3] > w.tile_cap;
        if (!rows_short && !tile_short) {
            // fit the slot count to what the tiles hold (shared memory per CTA decides the occupancy of
            // every gather kernel of the step); later rebuilds only raise the flag
            if (attempt == 0 && !w.tile_fitted) {
                w.tile_cap = std::min(w.tile_cap, host[3] + host[3] / 8 + 16);
                w.tile_fitted = true;
            }
            break;
        }
        if (tile_short) {
            // first give the tiles the slots they need (+ head-room for later rebuilds); if even the
            // kernels' shared memory cannot hold the densest neighbourhood, halve the tile width
            if (host[3] <= w.tile_cap_limit) {
                w.tile_cap = std::min(w.tile_cap_limit, host[3] + host[3] / 8 + 16);
            } else if (w.tile_w > 1) {
                SPHB_TRY(plan_tiles(w, w.tile_w / 2));
                w.tile_cap = w.tile_cap_limit;
            } else {
                return set_error("SPH:Neighbor:entryCapacity",
                                 "%d entries in one 3x3 cell neighbourhood exceed the %d a tile can stage", host[3],
                                 w.tile_cap_limit);
            }
        }
        if (rows_short) SPHB_TRY(neighbor_reserve_ell(w, host[2] + host[2] / 4 + 4));
        // clear the capacity flags of the failed attempt and walk again
        int f = 0;
        SPHB_CUDA(cudaMemcpyAsync(&f, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        SPHB_CUDA(cudaStreamSynchronize(st));
        f &= ~(kErrCsrCap | kErrTileCap);
        SPHB_CUDA(cudaMemcpyAsync(flags, &f, sizeof(int), cudaMemcpyHostToDevice, st));
        SPHB_CUDA(cudaStreamSynchronize(st));
        P = make_params(w);
    }
    w.half_only = half_only;
    w.half_valid = half_only;   // a full build counts its half pairs on demand (neighbor_half_total)
    SPHB_CUDA(cudaGetLastError());
    if (launches) *launches += nl;
    return 0;
}

// ---- batched build over an ensemble of channels (common.cuh: NbrBatchLaunch) --------------------------------
size_t neighbor_batch_record_bytes() { return sizeof(NbrChan); }

// Writes the channel's record (host memory, neighbor_batch_record_bytes() bytes) and its scan record.
int neighbor_batch_fill(NeighborWork &w, const double2 *pos, void *host_record, ScanChan *host_scan)
{
    if (w.slab) return set_error("SPHB:batch", "slabs of a sharded channel cannot be batched");
    if (!w.grid_armed || !w.ell_q.p || !w.tiles.p)
        return set_error("SPHB:batch", "every channel needs its first (exact) list build before it joins a batch");
    SPHB_TRY(w.scan_tmp.reserve(scan_workspace_ints((size_t)w.cell_cap) * sizeof(int)));
    NbrChan c;
    c.P = make_params(w);
    c.pos = pos;
    c.g = w.grid.as<GridDesc>();
    c.flags = w.flags.as<int>();
    c.totals = w.totals.as<int>();
    c.cell_cnt = w.cell_cnt.as<int>();
    c.cell_start = w.cell_start.as<int>();
    c.snap = w.snap.as<double2>();
    c.cell_of = w.cell_of.as<int>();
    c.ent_of = w.ent_of.as<int>();
    c.tmp_key = w.tmp_key.as<int>();
    c.tmp_pos = w.tmp_pos.as<double2>();
    c.ent_key = w.ent_key.as<int>();
    c.ent_pos = w.ent_pos.as<double2>();
    c.tiles = w.tiles.as<TileDesc>();
    c.ell_q = w.ell_q.as<uint4>();
    c.ell_cnt = w.ell_cnt.as<int>();
    c.tile_w = w.tile_w;
    c.tile_cap = w.tile_cap;
    c.n_tiles = w.n_tiles;
    c.M = w.mirrors;
    std::memcpy(host_record, &c, sizeof(NbrChan));
    host_scan->in = c.cell_cnt;
    host_scan->out = c.cell_start;
    host_scan->tile_totals = w.scan_tmp.as<int>();
    host_scan->total_out = &c.g->n_entries;
    return 0;
}

// The steady-state (non-exact, full-list) build of neighbor_build for every channel of the batch: the same
// kernel bodies, one launch each.  Capacity overflows only raise the channels' own device flags.
int neighbor_batch_build(const NbrBatchLaunch &L, long long *launches)
{
    const NbrChan *chan = static_cast<const NbrChan *>(L.chan);
    const unsigned ny = (unsigned)L.n_chan;
    const int nt = L.n_total;
    long long nl = 0;
    int blocks = (nt + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (!L.grid_ready) {
        k_yrange_grid_batch<<<dim3((unsigned)blocks, ny), 256, 0, L.st>>>(chan);
        const int n_ints = 2 * (L.cell_cap + 1);
        k_zero_cells_batch<<<dim3(std::max(1u, std::min(grid_for((size_t)n_ints), 64u)), ny), kBlock, 0, L.st>>>(chan, n_ints);
        nl += 2;
    }
    k_bin_count_batch<<<dim3(grid_for((size_t)nt), ny), kBlock, 0, L.st>>>(chan);
    nl += 1;
    SPHB_TRY(exclusive_scan_i32_batch(L.scan, L.n_chan, (size_t)L.cell_cap, L.st, &nl));
    k_scatter_batch<<<dim3(grid_for((size_t)nt), ny), kBlock, 0, L.st>>>(chan);
    const size_t smem = 2 * (size_t)L.tile_cap * kWalkSlotBytes + (size_t)(kTileThreads / 32) * L.ell_cap * sizeof(unsigned short);
    static SmemGrant grant;
    SPHB_TRY(grant_smem(grant, k_walk_tile_batch, smem));
    k_walk_tile_batch<<<dim3(L.n_tiles, ny), kTileThreads, smem, L.st>>>(chan);
    nl += 2;
    SPHB_CUDA(cudaGetLastError());
    if (launches) *launches += nl;
    return 0;
}

// totals[0] = number of reference half pairs of the current list (computed on demand after a full build)
int neighbor_half_total(NeighborWork &w, cudaStream_t st)
{
    if (w.half_valid) return 0;
    NbrParams P = make_params(w);
    const size_t smem = (size_t)w.tile_cap * sizeof(int);
    static SmemGrant grant;
    SPHB_TRY(grant_smem(grant, k_half_total, smem));
    SPHB_CUDA(cudaMemsetAsync(w.totals.p, 0, sizeof(int), st));
    k_half_total<<<w.n_tiles, kTileThreads, smem, st>>>(P, w.tiles.as<TileDesc>(), w.ent_key.as<int>(),
                                                        w.ell_q.as<uint4>(), w.ell_cnt.as<int>(), w.totals.as<int>(),
                                                        w.tile_cap);
    SPHB_CUDA(cudaGetLastError());
    w.half_valid = true;
    return 0;
}

// The reference's half-pair COO list from the rows (device output pointers, n_half slots each).
int neighbor_export_pairs(NeighborWork &w, size_t n_half, double *pair_i, double *pair_j, double *dx, double *dy,
                          double *r, double *W, double *dW, cudaStream_t st)
{
    (void)n_half;
    NbrParams P = make_params(w);
    const size_t nf = (size_t)w.n_fluid;
    SPHB_TRY(w.half_cnt.reserve((nf + 1) * sizeof(int)));
    SPHB_TRY(w.half_ptr.reserve((nf + 1) * sizeof(int)));
    k_half_count<<<grid_for(nf), kBlock, 0, st>>>(P, w.tile_w, w.tiles.as<TileDesc>(), w.cell_of.as<int>(),
                                                  w.ent_of.as<int>(), w.ent_key.as<int>(), w.ell_q.as<uint4>(),
                                                  w.ell_cnt.as<int>(), w.half_cnt.as<int>());
    SPHB_TRY(exclusive_scan_i32(w.half_cnt.as<int>(), w.half_ptr.as<int>(), nf, w.scan_tmp, nullptr, st, nullptr));
    k_export_pairs<<<grid_for(nf), kBlock, 0, st>>>(P, w.tile_w, w.tiles.as<TileDesc>(), w.cell_of.as<int>(),
                                                    w.ent_of.as<int>(), w.ent_key.as<int>(), w.ent_pos.as<double2>(),
                                                    w.ell_q.as<uint4>(), w.ell_cnt.as<int>(), w.half_ptr.as<int>(),
                                                    pair_i, pair_j, dx, dy, r, W, dW);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

// The full gather list as CSR by particle (device outputs: row_ptr[n_fluid+1], col[n_full] in walk order).
int neighbor_export_csr(NeighborWork &w, int *row_ptr, int *col, cudaStream_t st)
{
    NbrParams P = make_params(w);
    const size_t nf = (size_t)w.n_fluid;
    SPHB_TRY(w.half_cnt.reserve((nf + 1) * sizeof(int)));
    k_row_count<<<grid_for(nf), kBlock, 0, st>>>(P, w.tile_w, w.tiles.as<TileDesc>(), w.cell_of.as<int>(),
                                                 w.ent_of.as<int>(), w.ell_cnt.as<int>(), w.half_cnt.as<int>());
    SPHB_TRY(exclusive_scan_i32(w.half_cnt.as<int>(), row_ptr, nf, w.scan_tmp, nullptr, st, nullptr));
    if (col)
        k_row_fill<<<grid_for(nf), kBlock, 0, st>>>(P, w.tile_w, w.tiles.as<TileDesc>(), w.cell_of.as<int>(),
                                                    w.ent_of.as<int>(), w.ent_key.as<int>(), w.ell_q.as<uint4>(),
                                                    w.ell_cnt.as<int>(), row_ptr, col);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sphb
