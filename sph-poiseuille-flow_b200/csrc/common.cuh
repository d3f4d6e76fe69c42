// common.cuh -- shared declarations of libsph_b200 (sm_100a).  Internal header, not part of the ABI.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <climits>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sph_b200.h"

#ifndef SPHB_FLAT_STAGING
#define SPHB_FLAT_STAGING 1   // staging loops of the hot tile kernels as flat, load/store-split passes (tile.cuh); 0 = segment by segment (round 1)
#endif

#ifndef SPHB_ROW_FULLWORD
#define SPHB_ROW_FULLWORD 0
#endif

namespace sphb {

// ------------------------------------------------------------------------------------------------
// error plumbing (thread-local last error, MATLAB-style identifiers)
// ------------------------------------------------------------------------------------------------
int set_error(const char *id, const char *fmt, ...);
void clear_error();
const char *last_error_id();
const char *last_error_msg();

#define SPHB_CUDA(call)                                                                           \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return ::sphb::set_error("SPHB:cuda", "%s failed: %s (%s:%d)", #call,                 \
                                     cudaGetErrorString(e__), __FILE__, __LINE__);                \
    } while (0)

#define SPHB_REQUIRE(cond, id, msg)                      \
    do {                                                 \
        if (!(cond)) return ::sphb::set_error(id, msg);  \
    } while (0)

#define SPHB_TRY(expr)          \
    do {                        \
        int rc__ = (expr);      \
        if (rc__) return rc__;  \
    } while (0)

int ensure_device();   // 0 when a CUDA device is usable (sets SPHB:cuda otherwise)

// ------------------------------------------------------------------------------------------------
// device buffers: grow-only, reused across calls (stateless path) or owned by a context
// ------------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t need);   // grow-only; contents are NOT preserved on growth
    void release();
    template <class T> T *as() const { return static_cast<T *>(p); }
};

// 2x2 kernel-gradient-correction matrix of one particle, one 32-byte record (two 128-bit loads)
struct __align__(32) Mat2 {
    double b11, b12, b21, b22;
};

// Dynamic shared memory above 48 KB must be granted per kernel -- and per DEVICE (the attribute lives in the
// device's context).  One SmemGrant per kernel remembers what each device of the process has been granted.
struct SmemGrant {
    static constexpr int kMaxDev = 64;
    std::atomic<size_t> bytes[kMaxDev];
};
template <class Kern>
inline int grant_smem(SmemGrant &g, Kern *kern, size_t bytes)
{
    if (bytes + 8 * 1024 <= 48 * 1024) return 0;   // (the 48 KB every kernel may use include its static shared memory)
    int dev = 0;
    SPHB_CUDA(cudaGetDevice(&dev));
    const bool cached = dev >= 0 && dev < SmemGrant::kMaxDev;
    if (cached && bytes <= g.bytes[dev].load(std::memory_order_relaxed)) return 0;
    SPHB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (cached) g.bytes[dev].store(bytes, std::memory_order_relaxed);
    return 0;
}

constexpr int kBlock = 128;
inline unsigned grid_for(size_t n, int block = kBlock) { return (unsigned)((n + block - 1) / block); }

// device error flags (sphb_status.error_flags)
enum : int { kErrCsrCap = 1, kErrCellCap = 2, kErrSeenCap = 4, kErrNonFinite = 8, kErrTileCap = 16, kErrSlabCap = 32 };

// ------------------------------------------------------------------------------------------------
// cell grid descriptor, lives in device memory, written by the grid-setup kernel (nb:244-258)
// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+, CUDA graphs keep it as a programmatic edge): a kernel launched with
// SPHB_LAUNCH_PDL may be scheduled while its predecessor in the stream is still draining -- its CTAs become resident as
// the predecessor's finish -- and waits in pdl_wait() (griddepcontrol.wait) for the predecessor's completion and
// memory flush before it touches anything.  Every kernel launched this way calls pdl_wait() FIRST (then pdl_trigger()
// for its own successor); a kernel launched the ordinary way returns from both at once.  What it buys is the launch latency and the CTA ramp-up of each of the
// eleven kernels of a step.  MEASURED (r2 calls 28 / 29, 489 600 particles, replayed graph): with the trigger at the
// top of every kernel the step is SLOWER, 0.395 against 0.365 ms (the successor's waiting CTAs share the SMs with the
// tail of the running kernel); with the trigger behind the work, or without one, 0.365 against 0.365-0.367 ms: a
// replayed graph has no launch latency left to hide.  So it is OFF unless SPHB_PDL=1 is set in the environment, and
// the kernels trigger behind their work (SPHB_PDL_TRIGGER=2).
#ifdef SPHB_CPUSIM
__device__ inline void pdl_wait() {}
__device__ inline void pdl_trigger_top() {}
__device__ inline void pdl_trigger_end() {}
#define SPHB_LAUNCH_PDL(kern, grid, block, smem, st, ...) cpusim::launch(kern, grid, block, smem, st, __VA_ARGS__)
#else
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the NEXT kernel of the stream be scheduled once every CTA of this one has got here (or exited); without it the
// next kernel is launched only when this one has completed.  Issued right behind pdl_wait(): the dependent's CTAs then
// take the slots this kernel's CTAs vacate, and wait in their own pdl_wait() for its completion.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// where a kernel issues it: at its top (SPHB_PDL_TRIGGER=1: the successor's CTAs take every slot this kernel vacates) or
// behind its work (SPHB_PDL_TRIGGER=2: only the successor's launch latency is hidden); 0: never (the successor is
// launched when this kernel has completed)
#ifndef SPHB_PDL_TRIGGER
#define SPHB_PDL_TRIGGER 2
#endif
__device__ __forceinline__ void pdl_trigger_top() { if (SPHB_PDL_TRIGGER == 1) pdl_trigger(); }
__device__ __forceinline__ void pdl_trigger_end() { if (SPHB_PDL_TRIGGER == 2) pdl_trigger(); }
#endif
inline bool pdl_enabled()
{
    static int on = -1;
    if (on < 0) {
        const char *env = std::getenv("SPHB_PDL");
        on = (env && env[0] == '1') ? 1 : 0;   // off unless asked for: measured no gain (below)
    }
    return on != 0;
}
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    (void)cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // (errors surface in the caller's cudaGetLastError)
}
#define SPHB_LAUNCH_PDL(kern, grid, block, smem, st, ...) ::sphb::launch_pdl(kern, dim3(grid), dim3(block), smem, st, __VA_ARGS__)
#endif

// ------------------------------------------------------------------------------------------------
struct GridDesc {
    double y_min, y_max;
    double cell;       // 2h
    int ncx, ncy, ncells;
    int n_entries;     // real + ghost entries
    long long ymin_key, ymax_key;   // ordered-integer images used by the atomic min/max reduction
    unsigned int ticket;            // CTAs of the y-range reduction that have finished
    // Set by the binning of a build when a periodic ghost entry or its owner lies outside the columns the fast
    // walk takes for granted (ghosts in column ncx-1, owners in columns 0, ncx-2, ncx-1): exact multiples of the
    // cell size can round floor((x_w +- DL) / cell) one column further (nb:284-294) -- or when a wrapped x falls
    // outside [0, DL) (nb:270-272 returns DL itself for a tiny negative x).  Every tile of such a build is a seam
    // tile and every walker takes the general walk.  Bit 0: wrapped x out of range; bit 1: ghost / owner column.
    // The reference's HALF list (stateless search, MEX export) is exact either way.  In the FULL gather list of
    // the resident path a pair that one end finds only through such a rounding accident (a ghost in an unusual
    // column, or a ghost that exists for one end because x_w sits exactly on the insertion threshold nb:282,289)
    // can be missing from the other end's row: in exact arithmetic both ends see each other, so these pairs lie
    // within rounding of r = 2h, where W ~ (2 - q)^3 and dW ~ (2 - q)^2 vanish -- no effect on any sum.
    unsigned int ghost_anomaly;
};

#ifdef __CUDACC__
// order-preserving integer image of a double (atomic min / max of y over all particles, nb:245-250)
__device__ __forceinline__ long long ordered_key(double d)
{
    long long b = __double_as_longlong(d);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double from_ordered_key(long long k)
{
    return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}
// The grid of a build from the reduced y range (nb:253-258); re-arms the reduction and zeroes the build's totals.
// Executed by ONE thread once every contribution to ymin_key / ymax_key is visible.
__device__ __forceinline__ void grid_derive(GridDesc *g, double cell, int ncx, int cell_cap, int *flags, int *totals)
{
    const double y_min = from_ordered_key(atomicMin(&g->ymin_key, LLONG_MAX));
    const double y_max = from_ordered_key(atomicMax(&g->ymax_key, LLONG_MIN));
    int ncy = (int)ceil((y_max - y_min + 1e-12) / cell) + 1;   // nb:256
    if (ncy < 1) ncy = 1;
    if ((long long)ncy * ncx > (long long)cell_cap) {
        atomicOr(flags, kErrCellCap);
        ncy = cell_cap / ncx;
        if (ncy < 1) ncy = 1;
    }
    g->y_min = y_min;
    g->y_max = y_max;
    g->cell = cell;
    g->ncx = ncx;
    g->ncy = ncy;
    g->ncells = ncy * ncx;
    g->ymin_key = LLONG_MAX;   // re-armed for the next build
    g->ymax_key = LLONG_MIN;
    g->ticket = 0u;
    g->ghost_anomaly = 0u;     // raised again by the build's binning if need be
    for (int k = 0; k < 8; ++k) totals[k] = 0;
}
#endif

// What the continuity kernel prepares for the list build that follows it in a step of the unsharded resident
// channel (so that the build starts at the binning): the y range of ALL rows (nb:245-250), the grid derived by the
// last CTA (nb:253-258), the cell counters zeroed.  g = NULL: off.
struct GridFuse {
    GridDesc *g = nullptr;
    int *flags = nullptr, *totals = nullptr;
    int *cell_cnt = nullptr;
    int n_cell_ints = 0;
    double cell = 0.0;
    int ncx = 0, cell_cap = 0;
    int derive = 1;   // 0 (slab of a sharded channel): rows still arrive after the continuity pass -- the unpack kernels
                      // add their y and derive the grid (slab.cu)
};

// ------------------------------------------------------------------------------------------------
// neighbour-list storage shared by the stateless search and the resident context
// ------------------------------------------------------------------------------------------------
struct TileDesc;
// Entry-ordered mirrors written by the list build for the gather kernels of the resident path (ops_csr.cu): the
// walker of every real entry copies its row's velocity (a wall row: its wall velocity) and wall weight to the
// entry's index, so that a tile's neighbourhood is a few contiguous ranges.  All NULL = not wanted (stateless search).
struct BuildMirrors {
    const double2 *vel = nullptr, *wall_vel = nullptr;
    const double *mass = nullptr;
    double inv_rho0 = 0.0;
    double2 *e_vel = nullptr;   // velocity at build time
    double2 *e_w = nullptr;     // (m / rho0, 0) for a wall entry, (-1, 0) for a fluid entry (ph:205)
    int *ghost_ent = nullptr;   // [2 i + kind - 1] = entry of particle i's ghost of that kind in this build (validated
                                // against ent_key by its readers: entries of earlier builds are left standing)
};
struct NeighborWork {
    int n_fluid = 0, n_total = 0;
    double h = 0, DL = 0;
    int cell_cap = 0;            // allocated cells
    int ell_cap = 0;             // longest storable neighbour row (a multiple of 8)
    size_t ell_stride = 0;       // entry capacity (rows of the list are indexed by ENTRY)
    int tile_w = 0;              // cells per tile (one cell row)
    int tile_cap = 0;            // staged entries a tile may hold (shared-memory slots)
    int tile_cap_limit = 0;      // what the kernels of the caller can stage (set by the planner)
    unsigned n_tiles = 0;
    bool slab = false;           // x-slab of a sharded channel (no wrap, no ghosts; set before neighbor_plan)
    int col0 = 0;                // global cell column of local column 0 (slab mode)
    int ncx_local = 0;           // local cell columns (slab mode)
    const int *dev_counts = nullptr;   // slab mode: {n_fluid, owned, n_total} in device memory (see slab.cu)
    const int *alive = nullptr;  // slab mode: rows with alive[i] < 0 have migrated away and are skipped
    double per_cell_hint = 0.0;  // entries per cell the tile width is planned for (0 = n_total / cells)
    bool half_only = false;      // the last build kept only the reference's half list
    bool grid_armed = false;     // the GridDesc reduction keys have been initialised
    bool half_valid = false;     // totals[0] holds the half-pair count of the current list
    bool tile_fitted = false;    // tile_cap has been fitted to the first build
    BuildMirrors mirrors;        // set by the resident context before every build
    bool grid_ready = false;     // the caller's continuity kernel has derived the grid and zeroed the cell counters
                                 // for THIS build (GridFuse): it starts at the binning
    DevBuf grid;                 // GridDesc
    DevBuf flags;                // int error flags
    DevBuf snap;                 // double2 [n_total]   wrapped x, y at build time
    DevBuf cell_of;              // int [n_total]       home cell of the real entry
    DevBuf ent_of;               // int [n_total]       index of the real entry in the sorted entry array
    DevBuf ghost_ent;            // int [2 n_total]     entries of the particle's ghosts (BuildMirrors::ghost_ent)
    DevBuf cell_cnt;             // int [2*(cell_cap+1)] counts, then fill cursors
    DevBuf cell_start;           // int [cell_cap+1]
    DevBuf tmp_key, tmp_pos;     // entries as scattered (any order inside a cell)
    DevBuf ent_key;              // int [ent_cap]       (particle << 3) | (multi << 2) | kind, canonical order
    DevBuf ent_pos;              // double2 [ent_cap]   entry coordinates (ghosts shifted by -+DL)
    DevBuf tiles;                // TileDesc [n_tiles]
    DevBuf ell_cnt;              // int [ell_stride]    neighbours of the walker whose real entry is e
    DevBuf ell_q;                // uint4 [ell_cap/8 * ell_stride]: eight 16-bit entries per word (tile.cuh)
    DevBuf half_cnt, half_ptr;   // int [n_fluid+1]     export only (reference half-pair list)
    DevBuf scan_tmp;             // scan workspace
    DevBuf totals;               // int[8]: n_half, n_full, longest row, most staged entries of a tile
    void release_all();
};
constexpr int kDefaultEllRows = 48;

// ------------------------------------------------------------------------------------------------
// per-kernel device timing (CUDA events on the launching stream); off unless a caller enables it
// ------------------------------------------------------------------------------------------------
enum ProfId : int {
    kPfNbrGrid = 0, kPfNbrBin, kPfNbrScan, kPfNbrScatter, kPfNbrCanon, kPfNbrWalk,
    kPfDensity, kPfKgc, kPfViscTransp, kPfTimeStep, kPfPressure, kPfContinuity, kPfClock,
    kPfShear, kPfSort, kPfForceFused, kPfSlab, kPfCount
};
const char *prof_name(int id);

struct Prof {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Rec { int id; size_t e0, e1; };
    std::vector<Rec> recs;
    double ms[kPfCount] = {};
    long long n[kPfCount] = {};
    size_t grab();
    void begin(int id, cudaStream_t st);
    void end(cudaStream_t st);
    int collect(cudaStream_t st);   // synchronises the stream, folds the open records into ms[]/n[]
    void reset();
    void release();
};

struct ProfScope {
    Prof *p;
    cudaStream_t st;
    ProfScope(Prof *prof, int id, cudaStream_t s) : p(prof && prof->on ? prof : nullptr), st(s)
    {
        if (p) p->begin(id, st);
    }
    ~ProfScope()
    {
        if (p) p->end(st);
    }
};

// neighbor.cu (compiled with -fmad=false: bit-exact accept tests)
int neighbor_plan(NeighborWork &w, int n_fluid, int n_total, double h, double DL, double y_min_hint,
                  double y_max_hint, int max_nbr_hint, int tile_cap_limit);
int neighbor_reserve_ell(NeighborWork &w, int rows);
// GridFuse record of the NEXT build (off unless the y-range reduction is armed and the context is not a slab)
GridFuse neighbor_grid_fuse(const NeighborWork &w);
int neighbor_build(NeighborWork &w, const double2 *pos, bool half_only, bool exact, cudaStream_t st,
                   long long *launches, Prof *prof = nullptr);
int neighbor_export_pairs(NeighborWork &w, size_t n_half, double *pair_i, double *pair_j, double *dx, double *dy,
                          double *r, double *W, double *dW, cudaStream_t st);   // device pointers

int neighbor_export_csr(NeighborWork &w, int *row_ptr, int *col, cudaStream_t st);   // device pointers
int neighbor_half_total(NeighborWork &w, cudaStream_t st);   // makes totals[0] valid after a full build

// scan.cu
int exclusive_scan_i32(const int *in, int *out, size_t n, DevBuf &tmp, int *total_out_dev, cudaStream_t st,
                       long long *launches);
// the same scan for many independent arrays of one length in one grid (blockIdx.y = array); records in device memory
struct ScanChan {
    const int *in;
    int *out;
    int *tile_totals;   // scan_workspace_ints(n) ints
    int *total_out;     // may be NULL
};
size_t scan_workspace_ints(size_t n);
int exclusive_scan_i32_batch(const ScanChan *chan_dev, int n_chan, size_t n, cudaStream_t st, long long *launches);

// batched list build over an ensemble of channels (neighbor.cu; used by batch.cu)
struct NbrBatchLaunch {
    const void *chan;       // device array of per-channel records (neighbor_batch_record_bytes() each)
    const ScanChan *scan;   // device array [n_chan]: cell counts -> cell starts
    int n_chan;
    int n_total;            // rows of every channel
    int cell_cap;           // cells of every channel
    size_t ell_stride;      // entry capacity of every channel
    unsigned n_tiles;       // tiles of every channel
    int tile_cap, ell_cap;  // largest over the channels (size the walk's dynamic shared memory)
    bool grid_ready;        // the step's last continuity pass has derived the grids and zeroed the cell counters
    cudaStream_t st;
};
size_t neighbor_batch_record_bytes();
int neighbor_batch_fill(NeighborWork &w, const double2 *pos, void *host_record, ScanChan *host_scan);
int neighbor_batch_build(const NbrBatchLaunch &L, long long *launches);

// pack/unpack helpers (misc.cu)
void launch_pack2(const double *col_major, double2 *out, int n, cudaStream_t st);     // [n x 2] -> double2[n]
void launch_unpack2(const double2 *in, double *col_major, int n, cudaStream_t st);
void launch_pack4(const double *col_major, Mat2 *out, int n, cudaStream_t st);        // [n x 4] -> Mat2[n]
void launch_unpack4(const Mat2 *in, double *col_major, int n, cudaStream_t st);

}  // namespace sphb
