// ops_csr.cu -- fp64 physics operators of the DEVICE-RESIDENT path as per-particle gathers over the
// tile-local neighbour lists (rows P1-P9 and the epilogues of S1 in SURVEY.md 8(a)).
//
// One CTA owns one cell tile (tile.cuh): it stages what the operator needs of every entry of the tile's
// 3 x (w + 2) cell neighbourhood into shared memory with coalesced loads, then one thread per own
// particle walks its row -- 16-bit slots, eight per 128-bit load -- and reads its neighbours out of
// shared memory.  Every pair is evaluated from both ends, there are no atomics and the sums are
// deterministic.  All the reference's pair formulas are symmetric under (i <-> j, e -> -e) (ph:262-291,
// 502-520, 667-685, 891-925, 1091-1099), so the gather uses one formula per entry with e from j to i.
//
// Geometry is FROZEN at list-build time, as in the reference, where every mode consumes the pair
// cache computed by the neighbour search: dx = (x_i - x_entry) + shift*DL from the
// entry positions of the build and the image code stored with each list entry.
//
// Fusion: viscous + transport share one row walk (same inputs B_j, Vol_j); the pressure kernel reads
// the neighbours' half-step density/pressure staged by k_time_step (ph:857-862) and its epilogue does
// the velocity kick and the full position update (ph:1400-1404, 863-864, 1066-1069); the continuity
// kernel's epilogue finishes density/EOS (ph:1440-1446), wraps x (m:570-577) and feeds the grid-wide
// max|v| reduction (warp shuffle + one atomicMax per warp) that the next dt is computed from on the
// device (m:519-527).
#include <cstdlib>
#include "ops_csr.cuh"
#include "async_copy.cuh"
#include "physmath.cuh"
#include "tile.cuh"

namespace sphb {

namespace {

// Geometry of one list entry, recomputed from build-time entry positions (16 B staged per neighbour
// instead of the reference's 40 B cached per pair): e points from j to i.
struct Geo {
    double dx, dy, r, ex, ey;
};

__device__ __forceinline__ Geo pair_geometry(double2 pi, double2 pj, int m, const PhysParams &P)
{
    Geo g;
    g.dx = pi.x - pj.x;
    if (m != 0) g.dx += (double)m * P.DL;   // minimum image (nb:359-363); only seam tiles ever shift
    g.dy = pi.y - pj.y;
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
// SPHB_TILE_BEGIN: descriptor into shared memory + early exit.  SPHB_TILE_WALKERS: one thread per REAL
// entry of the tile's own cells (ew = entry, si = its slot, i = particle); ghost entries are skipped.
#define SPHB_TILE_BEGIN()                                                  \
    extern __shared__ __align__(16) unsigned char smem_raw[];              \
    __shared__ TileDesc T;                                                 \
    tile_load(T, S.tiles, blockIdx.x);                                     \
    __syncthreads();                                                       \
    if (!(T.flags & kTileValid)) return;                                   \
    const int cap = S.tile_cap;                                            \
    (void)cap

// What a thread needs of its (first) walker from global memory, fetched BEFORE the staging loads so that
// both latencies overlap: entry key, row length, first row word.
struct OwnPre {
    int key, n;
    uint4 q0;
    double2 rm;   // StateView::e_rm of the walker (only the kernels that ask for it: K::kOwnRm)
};
template <bool RM = false>
__device__ __forceinline__ OwnPre tile_prefetch_own(const StateView &S, const TileDesc &T)
{
    OwnPre p;
    p.key = 3;   // "not a real entry"
    p.n = 0;
    p.q0 = make_uint4(0u, 0u, 0u, 0u);
    p.rm = make_double2(0.0, 0.0);
    if ((int)threadIdx.x < tile_own_total(T)) {
        int orow, ew, si;
        tile_own_locate(T, (int)threadIdx.x, orow, ew, si);
        p.key = S.ent_key[ew];
        p.n = S.ell_cnt[ew];
        p.q0 = S.ell_q[ew];
        if (RM) p.rm = S.e_rm[ew];
    }
    return p;
}

#define SPHB_TILE_WALKERS(ew, si, i)                                                              \
    SPHB_TILE_FOR_OWN(T, ew, si, orow__)                                                          \
        if (const bool first__ = v__ == (int)threadIdx.x; true)                                   \
            if (const int key__ = first__ ? pre.key : S.ent_key[ew]; (key__ & 3) == 0)            \
                if (const int i = key__ >> 3; true)

// A row within the cut-off of the periodic seam owns ghost entries (key bit 2): the kernel that writes a mirror at the
// row's own entry writes it at the ghost entries too, so that seam tiles are staged by the same bulk copies as every
// other tile.  ghost_ent keeps the entry of each ghost as of the build that last had one; it counts only if that entry
// IS this particle's ghost of that kind in the current list.
__device__ __forceinline__ void ghost_store(const StateView &S, int key, int i, double2 *mirror, double2 v)
{
    if (!(key & 4) || S.ghost_ent == nullptr) return;
#pragma unroll
    for (int kind = 1; kind <= 2; ++kind) {
        const int eg = S.ghost_ent[2 * i + kind - 1];
        if (eg >= 0 && (size_t)eg < S.ell_stride && S.ent_key[eg] == ((i << 3) | 4 | kind)) mirror[eg] = v;
    }
}

// The walker's row: 128-bit words of eight 16-bit entries (padded with the dummy slot), the next word
// in flight while the current one is processed, four entries per unrolled group.  body(slot, image
// shift); SHIFT = false skips the decode of the image shift (tiles away from the periodic seam never
// shift).
template <bool SHIFT, class Body>
__device__ __forceinline__ void tile_for_row_impl(const StateView &S, int ew, int n, uint4 cur, Body &&body)
{
    n = min(n, S.ell_cap);
    const int nw = (n + 7) >> 3;
    const uint4 *q = S.ell_q + ew;
#pragma unroll 1
    for (int w = 0; w < nw; ++w) {
        const uint4 nxt = (w + 1 < nw) ? q[(size_t)(w + 1) * S.ell_stride] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned c = ell_extract(cur, u);
            body(ell_slot(c), SHIFT ? ell_shift(c) : 0);
        }
        if (w * 8 + 4 < n) {
#pragma unroll
            for (int u = 4; u < 8; ++u) {
                const unsigned c = ell_extract(cur, u);
                body(ell_slot(c), SHIFT ? ell_shift(c) : 0);
            }
        }
        cur = nxt;
    }
}
// rows of the walker loop: the first pass uses the prefetched length and word
#define SPHB_ROW(SHIFT, body_lambda)                                                               \
    tile_for_row_impl<SHIFT>(S, ew, first__ ? pre.n : S.ell_cnt[ew], first__ ? pre.q0 : S.ell_q[ew], body_lambda)

// A tile is PLAIN when its neighbourhood holds no wall entry and it is away from the periodic seam:
// every neighbour is a fluid particle at its stored position, so the row loops are branch-free.
#define SPHB_TILE_DISPATCH(walls_flag, call)                                          \
    do {                                                                              \
        const bool plain__ = !__syncthreads_or(walls_flag) && !(T.flags & kTileSeam); \
        if (plain__) {                                                                \
            call(true);                                                               \
        } else {                                                                      \
            call(false);                                                              \
        }                                                                             \
    } while (0)

__device__ __forceinline__ double half_step_rho(double rho, double drho, double dt, double rho0)
{
    double r = rho + 0.5 * dt * drho;   // ph:858
    if (r < 1e-10) r = rho0;            // ph:859-861
    return r;
}

// ================================================================================================
// The persistent tile pipeline of the four hot gather kernels (density, KGC, fused forces, continuity).
//
// A CTA takes tiles cta, cta + n_ctas, ... and keeps TWO staging buffers: while its threads walk the rows of
// one tile out of one buffer, the copy engine fills the other with the next tile's neighbourhood.  That is
// possible because everything a tile stages lives in ENTRY order (StateView::e_*, written by the kernel that
// produces the field, at the walker's own entry index): per staged cell row one contiguous range per field, moved by
// cp.async.bulk and signalled on an mbarrier (async_copy.cuh) -- no thread gathers through ent_key, and the
// dependent-load chain that used to open every tile (descriptor -> ent_key -> field -> STS -> barrier) is gone.
// The next tile's descriptor travels one tile further ahead in registers, the walkers' row heads one tile ahead.
// Tiles on the periodic seam (first and last tile of a tile row) stage their wrap columns with the same copies: ghost
// entries carry their particle's mirror values (ghost_store).
//
// A kernel K provides: kFields (16-byte fields per slot), bulk_fields(T), src(S, f), dummy(buf, cap, slot),
// rows<PLAIN>(S, T, pre, buf, cap), rows_cold(...).
// ================================================================================================
// The seam-tile staging and the row loops of tiles with wall entries run for a few percent of the tiles.  A/B switch:
// SPHB_COLD_INLINE=0 keeps them out of line (off the plain path's register allocation, at the price of a stack frame).
#ifndef SPHB_COLD_INLINE
#define SPHB_COLD_INLINE 1   // measured (r2 call 4): out of line costs a 0.5 KB stack frame for the kernel object, +20 % per kernel
#endif
#if SPHB_COLD_INLINE
#define SPHB_COLD __forceinline__
#else
#define SPHB_COLD __noinline__
#endif

template <class K>
__device__ __forceinline__ void stage_bulk(const K &k, const StateView &S, const TileDesc &T, unsigned char *buf, int cap,
                                           mbar_t *bar)
{
    if (!(T.flags & kTileValid)) return;
    if (threadIdx.x >= 32) return;
    const int lane = (int)threadIdx.x;
    const int nf = k.bulk_fields(T);
    // segments of a staged row: the main range and, for the first / last tile of a tile row, the wrap columns (their
    // entries -- real and ghost -- carry mirror values like any other: ghost_store)
    const int segs = (T.flags & kTileSeam) ? kStageRows * 3 : kStageRows;
    const int step = (T.flags & kTileSeam) ? 1 : 3;
    if (lane == 0) {
        int entries = 0;
        for (int j = 0; j < segs; ++j) entries += T.seg[j * step].e - T.seg[j * step].b;
        mbar_arrive_expect_tx(bar, (unsigned)(entries * nf) * (unsigned)sizeof(double2));
        k.dummy(buf, cap, T.n_staged);
    }
    __syncwarp();
    for (int c = lane; c < segs * nf; c += 32) {
        const int j = c / nf, f = c - j * nf;
        const TileSeg sg = T.seg[j * step];
        const int n = sg.e - sg.b;
        if (n > 0)
            bulk_copy_g2s(buf + ((size_t)f * cap + sg.off) * sizeof(double2), k.src(S, f) + sg.b,
                          (unsigned)n * (unsigned)sizeof(double2), bar);
    }
}

template <class K>
__device__ __forceinline__ void run_tiles(K &k, const StateView &S, unsigned cta, unsigned n_ctas)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // three descriptor slots: the one of tile k + 1 is written at the top of iteration k, while slower threads may
    // still be reading tile k - 1's in their row loops (no barrier in between)
    __shared__ TileDesc Ts[3];
    __shared__ __align__(8) mbar_t bars[2];
    constexpr int kDescWords = (int)(sizeof(TileDesc) / 16);
    const int cap = S.tile_cap;
    const size_t buf_bytes = (size_t)cap * K::kFields * sizeof(double2);
    const unsigned n_tiles = S.n_tiles;
    if (cta >= n_tiles) return;
    const int tid = (int)threadIdx.x;
    if (tid == 0) {
        mbar_init(&bars[0], 1u);
        mbar_init(&bars[1], 1u);
        mbar_init_fence();
    }
    int4 dnext = make_int4(0, 0, 0, 0);   // descriptor of the tile after the next one (threads 0 .. kDescWords-1)
    if (tid < kDescWords) {
        reinterpret_cast<int4 *>(&Ts[0])[tid] = reinterpret_cast<const int4 *>(S.tiles + cta)[tid];
        if (cta + n_ctas < n_tiles) dnext = reinterpret_cast<const int4 *>(S.tiles + cta + n_ctas)[tid];
    }
    __syncthreads();
    stage_bulk(k, S, Ts[0], smem_raw, cap, &bars[0]);
    OwnPre pre = tile_prefetch_own<K::kOwnRm>(S, Ts[0]);
    unsigned parity = 0u;   // bit b = phase of bars[b] to wait for next
    int cur = 0, dcur = 0;
    for (unsigned t = cta; t < n_tiles; t += n_ctas, cur ^= 1, dcur = dcur == 2 ? 0 : dcur + 1) {
        const TileDesc &T = Ts[dcur];
        TileDesc &Tn = Ts[dcur == 2 ? 0 : dcur + 1];
        unsigned char *buf = smem_raw + (size_t)cur * buf_bytes;
        unsigned char *bufn = smem_raw + (size_t)(cur ^ 1) * buf_bytes;
        const unsigned tn = t + n_ctas;
        const bool has_next = tn < n_tiles;
        if (tid < kDescWords) {
            if (has_next) reinterpret_cast<int4 *>(&Tn)[tid] = dnext;
            if (tn + n_ctas < n_tiles) dnext = reinterpret_cast<const int4 *>(S.tiles + tn + n_ctas)[tid];
        }
        async_proxy_fence();   // this thread's stores into the other buffer (a seam tile staged by hand) before the copies below
        __syncthreads();       // Tn is visible; every thread has left the rows of the previous tile: the other buffer is free
        OwnPre pre_next;
        pre_next.key = 3;
        pre_next.n = 0;
        pre_next.q0 = make_uint4(0u, 0u, 0u, 0u);
        pre_next.rm = make_double2(0.0, 0.0);
        if (has_next) {
            stage_bulk(k, S, Tn, bufn, cap, &bars[cur ^ 1]);
            pre_next = tile_prefetch_own<K::kOwnRm>(S, Tn);
        }
        if (T.flags & kTileValid) {
            mbar_wait(&bars[cur], (parity >> cur) & 1u);
            parity ^= 1u << cur;
            if (T.flags & (kTileHasWall | kTileSeam)) k.rows_cold(S, T, pre, buf, cap);
            else k.template rows<true>(S, T, pre, buf, cap);
        }
        pre = pre_next;
    }
}

// fields of a staging buffer: field f of slot s at buf + (f * cap + s) * 16
__device__ __forceinline__ double2 *field(unsigned char *buf, int cap, int f)
{
    return reinterpret_cast<double2 *>(buf) + (size_t)f * cap;
}

// the time step of a reference step (m:519-527)
__device__ __forceinline__ double device_dt(const PhysParams &P, const DevScalars *sc, double remain, double fixed_dt)
{
    if (fixed_dt > 0.0) return fixed_dt;
    const double vmax = sc->vmax;
    const double dt_a = 0.25 * P.h / fmax(P.c_f + vmax, 1e-12);
    const double dt_v = 0.125 * P.h * P.h / fmax(P.nu, 1e-12);
    const double dt_b = 0.25 * sqrt(P.h / fmax(fabs(P.g), 1e-12));
    return fmax(fmin(fmin(dt_a, dt_v), fmin(dt_b, remain)), 1e-12);
}

// ---------------------------------------------------------------- P1 pass 1: sigma sums -> rho, Vol
// staged per slot: position; tiles with wall entries also (m_j / rho0, 0) for a wall entry (ph:205), (-1, 0) for a
// fluid entry.  Also writes the entry-ordered (half-step pressure, volume) pair the later kernels stage.
struct DensityK {
    static constexpr bool kOwnRm = false;
    static constexpr int kFields = 2;
    PhysParams P;
    double dt;
    // (the wall weights are read by the row loop of wall tiles and of seam tiles, which share the cold row loop)
    __device__ __forceinline__ int bulk_fields(const TileDesc &T) const { return (T.flags & (kTileHasWall | kTileSeam)) ? 2 : 1; }
    __device__ __forceinline__ const double2 *src(const StateView &S, int f) const { return f == 0 ? S.ent_pos : S.e_w; }
    __device__ __forceinline__ void dummy(unsigned char *buf, int cap, int slot) const
    {
        field(buf, cap, 0)[slot] = make_double2(kDummyCoord, kDummyCoord);
        field(buf, cap, 1)[slot] = make_double2(-1.0, 0.0);
    }
    template <bool PLAIN>
    __device__ __forceinline__ void rows(const StateView &S, const TileDesc &T, const OwnPre &pre, unsigned char *buf, int cap)
    {
        const double2 *s_xy = field(buf, cap, 0), *s_w = field(buf, cap, 1);
        SPHB_TILE_WALKERS(ew, si, i)
        {
            const double m_i = S.mass[i];
            if (i >= P.nf) {   // wall rows: rho = rho0 (ph:214-216), Vol = m/rho0
                const double V = m_i / P.rho0;
                S.rho[i] = P.rho0;
                S.Vol[i] = V;
                S.e_pV[ew] = make_double2(0.0, -V);
                ghost_store(S, key__, i, S.e_pV, make_double2(0.0, -V));
                continue;
            }
            const double drho_i = S.drho_dt[i];
            const double2 pi = s_xy[si];
            double s_in = P.sigma;   // W(0), ph:178-181
            double s_ct = 0.0;
            SPHB_ROW(!PLAIN, ([&](const int sl, const int m) {
                const Geo g = pair_geometry(pi, s_xy[sl], m, P);
                const double W = spline_W(g.r, P);
                if (PLAIN) {
                    s_in += W;
                } else {
                    const double wj = s_w[sl].x;
                    if (wj < 0.0) s_in += W;
                    else s_ct += W * wj;
                }
            }));
            double rho = s_in * P.rho0 * P.inv_sigma0;
            rho += s_ct * P.rho0 * P.rho0 * P.inv_sigma0 / m_i;
            if (rho <= 1e-12) rho = P.rho0;
            const double V = m_i / rho;
            S.rho[i] = rho;
            S.Vol[i] = V;
            // half-step pressure of this row for the force kernel's neighbours (ph:857-862)
            const double rho_h = half_step_rho(rho, drho_i, dt, P.rho0);
            const double2 pV = make_double2(P.p0 * (rho_h / P.rho0 - 1.0), V);
            S.e_rm[ew] = make_double2(rho_h, m_i);   // the force kernel's own-row operands, with its row head
            S.e_pV[ew] = pV;
            ghost_store(S, key__, i, S.e_pV, pV);
        }
    }
    __device__ SPHB_COLD void rows_cold(const StateView &S, const TileDesc &T, const OwnPre &pre, unsigned char *buf, int cap)
    {
        rows<false>(S, T, pre, buf, cap);
    }
};

// publish != 0: this launch opens a reference step -- every thread derives the step's dt from max|v| itself (what the
// one-thread k_publish_dt launch used to do in front of this kernel: ~4 us of an otherwise idle device per step), and
// one thread publishes it for the kernels behind and re-arms the max|v| reduction.
__device__ __forceinline__ void density_tiles(PhysParams P, const StateView &S, DevScalars *sc, unsigned cta, unsigned n_ctas,
                                              int publish, double remain)
{
    double dt;
    if (publish) {
        dt = device_dt(P, sc, remain, 0.0);
        if (cta == 0 && threadIdx.x == 0) {
            sc->dt = dt;
            sc->vmax_bits = 0ull;
        }
    } else {
        dt = sc->dt;
    }
    if (step_poisoned(S.flags)) return;
    load_counts(P);
    DensityK k;
    k.P = P;
    k.dt = dt;
    run_tiles(k, S, cta, n_ctas);
}
#ifndef SPHB_DENS_MINCTAS
#define SPHB_DENS_MINCTAS 6   // (r2 calls 3-4: 8 -> 150 B of spills, 53 us; 6 -> 43 us; 5 -> 44 us)
#endif
__global__ void __launch_bounds__(kTileThreads, SPHB_DENS_MINCTAS) k_density(PhysParams P, StateView S, DevScalars *sc, int publish,
                                                                            double remain)
{
    pdl_wait();
    pdl_trigger_top();
    density_tiles(P, S, sc, blockIdx.x, gridDim.x, publish, remain);
    pdl_trigger_end();
}
__global__ void __launch_bounds__(kTileThreads, SPHB_DENS_MINCTAS) k_density_batch(const BatchChan *__restrict__ chan)
{
    const BatchChan &c = chan[blockIdx.y];
    density_tiles(c.P, c.S, c.sc, blockIdx.x, gridDim.x, 0, 0.0);
}

// ---------------------------------------------------------------- P1 pass 2: A -> B (ph:239-366)
// staged per slot: position, (-, volume).  Also writes the entry-ordered rows of B.
struct KgcK {
    static constexpr bool kOwnRm = false;
    static constexpr int kFields = 2;
    PhysParams P;
    __device__ __forceinline__ int bulk_fields(const TileDesc &) const { return 2; }
    __device__ __forceinline__ const double2 *src(const StateView &S, int f) const { return f == 0 ? S.ent_pos : S.e_pV; }
    __device__ __forceinline__ void dummy(unsigned char *buf, int cap, int slot) const
    {
        field(buf, cap, 0)[slot] = make_double2(kDummyCoord, kDummyCoord);
        field(buf, cap, 1)[slot] = make_double2(0.0, 0.0);
    }
    template <bool PLAIN>
    __device__ __forceinline__ void rows(const StateView &S, const TileDesc &T, const OwnPre &pre, unsigned char *buf, int cap)
    {
        const double2 *s_xy = field(buf, cap, 0), *s_pV = field(buf, cap, 1);
        SPHB_TILE_WALKERS(ew, si, i)
        {
            Mat2 B;
            B.b11 = 1.0;
            B.b12 = 0.0;
            B.b21 = 0.0;
            B.b22 = 1.0;
            if (i < P.nf) {
                const double2 pi = s_xy[si];
                double a11 = 0.0, a12 = 0.0, a21 = 0.0, a22 = 0.0;
                SPHB_ROW(!PLAIN, ([&](const int sl, const int m) {
                    const Geo g = pair_geometry(pi, s_xy[sl], m, P);
                    // walls take the same formula (ph:293-309); their staged volume carries the wall sign
                    const double Vs = s_pV[sl].y;
                    const double f = spline_dW(g.r, P) * (PLAIN ? Vs : fabs(Vs));
                    const double fx = f * g.ex, fy = f * g.ey;
                    a11 -= g.dx * fx;
                    a12 -= g.dx * fy;
                    a21 -= g.dy * fx;
                    a22 -= g.dy * fy;
                }));
                kgc_from_A(a11, a12, a21, a22, B.b11, B.b12, B.b21, B.b22);
                S.e_b1[ew] = make_double2(B.b11, B.b12);
                S.e_b2[ew] = make_double2(B.b21, B.b22);
                ghost_store(S, key__, i, S.e_b1, make_double2(B.b11, B.b12));
                ghost_store(S, key__, i, S.e_b2, make_double2(B.b21, B.b22));
            }
            S.B[i] = B;
        }
    }
    __device__ SPHB_COLD void rows_cold(const StateView &S, const TileDesc &T, const OwnPre &pre, unsigned char *buf, int cap)
    {
        rows<false>(S, T, pre, buf, cap);
    }
};

__device__ __forceinline__ void kgc_tiles(PhysParams P, const StateView &S, unsigned cta, unsigned n_ctas)
{
    if (step_poisoned(S.flags)) return;
    load_counts(P);
    KgcK k;
    k.P = P;
    run_tiles(k, S, cta, n_ctas);
}
#ifndef SPHB_KGC_MINCTAS
#define SPHB_KGC_MINCTAS 6    // (8 -> spills, 57 us; 6 -> 51.5 us)
#endif
__global__ void __launch_bounds__(kTileThreads, SPHB_KGC_MINCTAS) k_kgc(PhysParams P, StateView S)
{
    pdl_wait();
    pdl_trigger_top();
    kgc_tiles(P, S, blockIdx.x, gridDim.x);
    pdl_trigger_end();
}
__global__ void __launch_bounds__(kTileThreads, SPHB_KGC_MINCTAS) k_kgc_batch(const BatchChan *__restrict__ chan)
{
    const BatchChan &c = chan[blockIdx.y];
    kgc_tiles(c.P, c.S, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------- P2 + P3 fused (ph:469-545, 636-710)
// staged per slot: position, velocity (a wall entry: its wall velocity), B, and the volume with the
// sign bit set for wall entries
template <bool VISC, bool TRANSP, bool PLAIN>
__device__ __forceinline__ void visc_transport_rows(const PhysParams &P, const StateView &S, const TileDesc &T, const OwnPre &pre,
                                                    const double2 *s_xy, const double2 *s_v, const double2 *s_b1,
                                                    const double2 *s_b2, const double *s_V, int add_gravity)
{
    const double h001 = 0.01 * P.h;
    SPHB_TILE_WALKERS(ew, si, i)
    {
        if (i >= P.nf) {
            if (VISC) S.force_prior[i] = make_double2(0.0, 0.0);
            continue;
        }
        const double2 pi = s_xy[si];
        const double2 vi = VISC ? s_v[si] : make_double2(0.0, 0.0);
        const double2 bi1 = s_b1[si], bi2 = s_b2[si];
        const double Vi = s_V[si];
        double ax = 0.0, ay = 0.0, ix = 0.0, iy = 0.0;
        SPHB_ROW(!PLAIN, ([&](const int sl, const int m) {
            const Geo g = pair_geometry(pi, s_xy[sl], m, P);
            const double ex = g.ex, ey = g.ey;
            const double Vs = s_V[sl];
            const double gj = spline_dW(g.r, P) * (PLAIN ? Vs : fabs(Vs));
            double tx, ty, wv = 1.0, wt = 1.0;
            if (PLAIN || Vs >= 0.0) {
                const double2 bj1 = s_b1[sl], bj2 = s_b2[sl];
                tx = (bi1.x + bj1.x) * ex + (bi1.y + bj1.y) * ey;
                ty = (bi2.x + bj2.x) * ex + (bi2.y + bj2.y) * ey;
            } else {
                tx = bi1.x * ex + bi1.y * ey;
                ty = bi2.x * ex + bi2.y * ey;
                wv = 4.0;   // no-slip mirror, ph:526
                wt = 2.0;   // ph:690
            }
            if (VISC) {
                const double2 vj = s_v[sl];
                const double eBe = ex * tx + ey * ty;
                double cf = eBe * P.mu * gj * fast_rcp(g.r + h001);
                if (!PLAIN) cf *= wv;
                ax += cf * (vi.x - vj.x);
                ay += cf * (vi.y - vj.y);
            }
            if (TRANSP) {
                const double ci = PLAIN ? -gj : -wt * gj;
                ix += ci * tx;
                iy += ci * ty;
            }
        }));
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

template <bool VISC, bool TRANSP>
__global__ void __launch_bounds__(kTileThreads) k_visc_transport(PhysParams P, StateView S, int add_gravity)
{
    if (step_poisoned(S.flags)) return;
    load_counts(P);
    SPHB_TILE_BEGIN();
    const OwnPre pre = tile_prefetch_own(S, T);
    double2 *s_xy = reinterpret_cast<double2 *>(smem_raw);
    double2 *s_v = s_xy + cap;
    double2 *s_b1 = s_v + cap;
    double2 *s_b2 = s_b1 + cap;
    double *s_V = reinterpret_cast<double *>(s_b2 + cap);
    int walls = 0;
    SPHB_TILE_FOR_STAGED(T, e, slot)
    {
        const int j = S.ent_key[e] >> 3;
        s_xy[slot] = S.ent_pos[e];
        const double V = S.Vol[j];
        if (j < P.nf) {
            s_V[slot] = V;
            if (VISC) s_v[slot] = S.vel[j];
            const double4 B = *reinterpret_cast<const double4 *>(&S.B[j]);
            s_b1[slot] = make_double2(B.x, B.y);
            s_b2[slot] = make_double2(B.z, B.w);
        } else {   // a wall neighbour contributes its wall velocity and no B (ph:522-533, 687-697)
            s_V[slot] = -V;
            if (VISC) s_v[slot] = S.wall_vel[j];
            walls = 1;
        }
    }
    if (threadIdx.x == 0) {
        const int d = T.n_staged;
        s_xy[d] = make_double2(kDummyCoord, kDummyCoord);
        s_v[d] = make_double2(0.0, 0.0);
        s_b1[d] = make_double2(0.0, 0.0);
        s_b2[d] = make_double2(0.0, 0.0);
        s_V[d] = 0.0;
    }
#define SPHB_CALL(plain) \
    visc_transport_rows<VISC, TRANSP, plain>(P, S, T, pre, s_xy, s_v, s_b1, s_b2, s_V, add_gravity)
    SPHB_TILE_DISPATCH(walls, SPHB_CALL);
#undef SPHB_CALL
}

// ---------------------------------------------------------------- dt on the device (m:519-527) and the
// half-step density / pressure of every particle (ph:857-862), which the pressure gather reads per
// neighbour.  Every thread derives the same dt from the device vmax; thread 0 publishes it.
__device__ __forceinline__ void time_step_rows(PhysParams P, StateView S, DevScalars *sc, double remain, double fixed_dt)
{
    if (step_poisoned(S.flags)) return;
    if (sc->dual_on) {   // device-side dual-criteria scheduler: the substep covers what is left of the outer step
        if (!sc->sub_on) return;
        remain = sc->Dt_left;
    }
    load_counts(P);
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
__global__ void __launch_bounds__(kBlock) k_time_step(PhysParams P, StateView S, DevScalars *sc, double remain,
                                                     double fixed_dt)
{
    time_step_rows(P, S, sc, remain, fixed_dt);
}
__global__ void __launch_bounds__(kBlock) k_time_step_batch(const BatchChan *__restrict__ chan, double remain,
                                                           double fixed_dt)
{
    const BatchChan &c = chan[blockIdx.y];
    time_step_rows(c.P, c.S, c.sc, remain, fixed_dt);
}


// ---------------------------------------------------------------- P4 + velocity kick + position update
// staged per slot: position, volume (sign bit = wall); for fluid entries also velocity, half-step
// (rho, p) and B -- wall neighbours need only position and volume (ph:927-947)
template <bool PLAIN>
__device__ __forceinline__ void pressure_rows(const PhysParams &P, const StateView &S, const TileDesc &T, const OwnPre &pre,
                                              const double2 *s_xy, const double2 *s_v, const double2 *s_rp,
                                              const double2 *s_b1, const double2 *s_b2, const double *s_V, double dt)
{
    SPHB_TILE_WALKERS(ew, si, i)
    {
        if (i >= P.nf) {   // wall rows: force 0, velocity 0 (ph:958-962, 1405-1408)
            S.force[i] = make_double2(0.0, 0.0);
            S.vel_new[i] = make_double2(0.0, 0.0);
            const double2 wv = S.wall_vel[i];   // what the continuity pass stages of a wall neighbour
            S.e_vn[ew] = wv;
            ghost_store(S, key__, i, S.e_vn, wv);
            continue;
        }
        const double2 pi = s_xy[si];
        const double2 vi = s_v[si];
        const double2 bi1 = s_b1[si], bi2 = s_b2[si];
        const double Vi = s_V[si], m_i = S.mass[i];
        const double2 fp = S.force_prior[i];
        const double2 rpi = s_rp[si];
        const double rho_i = rpi.x, p_i = rpi.y;
        const double inv_m = 1.0 / m_i;
        const double wax = fp.x * inv_m, way = fp.y * inv_m;   // a = F_prior / m, ph:929-930
        double fx = 0.0, fy = 0.0;
        SPHB_ROW(!PLAIN, ([&](const int sl, const int m) {
            const Geo g = pair_geometry(pi, s_xy[sl], m, P);
            const double ex = g.ex, ey = g.ey;
            const double Vs = s_V[sl];
            const double gj = spline_dW(g.r, P) * (PLAIN ? Vs : fabs(Vs));
            if (PLAIN || Vs >= 0.0) {
                const double2 vj = s_v[sl];
                const double2 rpj = s_rp[sl];
                const double2 bj1 = s_b1[sl], bj2 = s_b2[sl];
                const double rho_bar = 0.5 * (rho_i + rpj.x);
                const double du = (vi.x - vj.x) * ex + (vi.y - vj.y) * ey;     // u_l - u_r
                const double beta = fmin(3.0 * fmax(du, 0.0), P.c_f);          // ph:1121-1129
                const double p_avg = 0.5 * (p_i + rpj.y);
                const double p_star = p_avg + 0.5 * beta * rho_bar * du;
                const double pg = 0.5 * (p_avg + p_star) * gj;
                fx -= pg * ((bi1.x + bj1.x) * ex + (bi1.y + bj1.y) * ey);
                fy -= pg * ((bi2.x + bj2.x) * ex + (bi2.y + bj2.y) * ey);
            } else {
                const double ext = -(wax * ex + way * ey);
                const double pw = p_i + rho_i * g.r * fmax(0.0, ext);   // virtual wall pressure, ph:934
                const double pg = (p_i + pw) * gj;
                fx -= pg * (bi1.x * ex + bi1.y * ey);
                fy -= pg * (bi2.x * ex + bi2.y * ey);
            }
        }));
        fx *= Vi;
        fy *= Vi;
        S.force[i] = make_double2(fx, fy);
        double2 vn;
        vn.x = vi.x + (fp.x + fx) / m_i * dt;   // ph:1402-1403
        vn.y = vi.y + (fp.y + fy) / m_i * dt;
        S.vel_new[i] = vn;
        S.e_vn[ew] = vn;
        ghost_store(S, key__, i, S.e_vn, vn);
        double2 x = S.pos[i];
        x.x += 0.5 * dt * vi.x;   // ph:863-864
        x.y += 0.5 * dt * vi.y;
        x.x += 0.5 * dt * vn.x;   // ph:1067-1068
        x.y += 0.5 * dt * vn.y;
        S.pos[i] = x;
    }
}

__device__ __forceinline__ void pressure_tile(PhysParams P, StateView S, const DevScalars *__restrict__ sc)
{
    if (step_poisoned(S.flags) || (sc->dual_on && !sc->sub_on)) return;
    load_counts(P);
    SPHB_TILE_BEGIN();
    const OwnPre pre = tile_prefetch_own(S, T);
    double2 *s_xy = reinterpret_cast<double2 *>(smem_raw);
    double2 *s_v = s_xy + cap;
    double2 *s_rp = s_v + cap;
    double2 *s_b1 = s_rp + cap;
    double2 *s_b2 = s_b1 + cap;
    double *s_V = reinterpret_cast<double *>(s_b2 + cap);
    int walls = 0;
    SPHB_TILE_FOR_STAGED(T, e, slot)
    {
        const int j = S.ent_key[e] >> 3;
        s_xy[slot] = S.ent_pos[e];
        const double V = S.Vol[j];
        if (j < P.nf) {
            s_V[slot] = V;
            s_v[slot] = S.vel[j];
            s_rp[slot] = S.rp_half[j];
            const double4 B = *reinterpret_cast<const double4 *>(&S.B[j]);
            s_b1[slot] = make_double2(B.x, B.y);
            s_b2[slot] = make_double2(B.z, B.w);
        } else {
            s_V[slot] = -V;
            walls = 1;
        }
    }
    if (threadIdx.x == 0) {
        const int d = T.n_staged;
        s_xy[d] = make_double2(kDummyCoord, kDummyCoord);
        s_v[d] = make_double2(0.0, 0.0);
        s_rp[d] = make_double2(0.0, 0.0);
        s_b1[d] = make_double2(0.0, 0.0);
        s_b2[d] = make_double2(0.0, 0.0);
        s_V[d] = 0.0;
    }
    const double dt = sc->dt;
#define SPHB_CALL(plain) pressure_rows<plain>(P, S, T, pre, s_xy, s_v, s_rp, s_b1, s_b2, s_V, dt)
    SPHB_TILE_DISPATCH(walls, SPHB_CALL);
#undef SPHB_CALL
}
__global__ void __launch_bounds__(kTileThreads) k_pressure(PhysParams P, StateView S, const DevScalars *__restrict__ sc)
{
    pressure_tile(P, S, sc);
}
__global__ void __launch_bounds__(kTileThreads, 5) k_pressure_batch(const BatchChan *__restrict__ chan)
{
    const BatchChan &c = chan[blockIdx.y];
    pressure_tile(c.P, c.S, c.sc);
}

// ---------------------------------------------------------------- P2 + P3 + dt + P4 + kick, one gather
// The viscous/transport pass and the pressure pass read the same neighbour fields (position, velocity,
// B, volume), and the pressure pass depends on the viscous one only through the walker's OWN F_prior
// (virtual wall pressure ph:929-934 and the kick ph:1402-1403).  One row walk therefore serves all
// three operators: 80 B gathered per neighbour once instead of 72 B + 88 B.  Walkers with wall
// neighbours take a second, short walk over their wall entries once F_prior is complete.  The time
// step (m:519-527) is published by the density pass (density_tiles), which evaluates every row's
// half-step pressure (ph:857-862) into the entry-ordered (p, V) pairs this kernel stages.
// staged per slot: position, velocity (wall entry: wall velocity), (half-step pressure, volume; volume < 0 = wall
// entry), the two rows of B.
struct ForceK {
    static constexpr bool kOwnRm = true;   // the walker's (half-step density, mass) travels with its row head
    static constexpr int kFields = 5;
    PhysParams P;
    double dt;
    int add_gravity;
    __device__ __forceinline__ int bulk_fields(const TileDesc &) const { return 5; }
    __device__ __forceinline__ const double2 *src(const StateView &S, int f) const
    {
        return f == 0 ? S.ent_pos : (f == 1 ? S.e_vel : (f == 2 ? S.e_pV : (f == 3 ? S.e_b1 : S.e_b2)));
    }
    __device__ __forceinline__ void dummy(unsigned char *buf, int cap, int slot) const
    {
        field(buf, cap, 0)[slot] = make_double2(kDummyCoord, kDummyCoord);
#pragma unroll
        for (int f = 1; f < 5; ++f) field(buf, cap, f)[slot] = make_double2(0.0, 0.0);
    }
    template <bool PLAIN>
    __device__ __forceinline__ void rows(const StateView &S, const TileDesc &T, const OwnPre &pre, unsigned char *buf, int cap)
    {
        const double2 *s_xy = field(buf, cap, 0), *s_v = field(buf, cap, 1), *s_pV = field(buf, cap, 2),
                      *s_b1 = field(buf, cap, 3), *s_b2 = field(buf, cap, 4);
        const double h001 = 0.01 * P.h;
        const double rho_per_p = P.rho0 / P.p0;
        SPHB_TILE_WALKERS(ew, si, i)
        {
            if (i >= P.nf) {   // wall rows: forces 0, velocity 0 (ph:958-962, 1405-1408)
                S.force_prior[i] = make_double2(0.0, 0.0);
                S.force[i] = make_double2(0.0, 0.0);
                S.vel_new[i] = make_double2(0.0, 0.0);
                const double2 wv = S.wall_vel[i];   // what the continuity pass reads of a wall neighbour (ph:1101-1102)
                S.e_vn[ew] = wv;
                ghost_store(S, key__, i, S.e_vn, wv);
                continue;
            }
            const double2 pi = s_xy[si];
            const double2 vi = s_v[si];
            const double2 bi1 = s_b1[si], bi2 = s_b2[si];
            const double2 pVi = s_pV[si];                      // (half-step p, volume)
            // (half-step density ph:857-861, mass): what the density pass of this step computed for this row, bit for
            // bit, fetched with the row head one tile ahead instead of three loads through the row index here
            const double2 rm = first__ ? pre.rm : S.e_rm[ew];
            const double Vi = pVi.y, m_i = rm.y;
            const double p_i = pVi.x;
            const double rho_i = rm.x;
            double ax = 0.0, ay = 0.0, ix = 0.0, iy = 0.0, fx = 0.0, fy = 0.0;
            bool has_wall = false;
            SPHB_ROW(!PLAIN, ([&](const int sl, const int m) {
                const Geo g = pair_geometry(pi, s_xy[sl], m, P);
                const double ex = g.ex, ey = g.ey;
                const double2 pVj = s_pV[sl];
                const double Vs = pVj.y;
                const double gj = spline_dW(g.r, P) * (PLAIN ? Vs : fabs(Vs));
                const double2 vj = s_v[sl];
                const double dvx = vi.x - vj.x, dvy = vi.y - vj.y;
                const double rinv = fast_rcp(g.r + h001);
                if (PLAIN || Vs >= 0.0) {
                    const double2 bj1 = s_b1[sl], bj2 = s_b2[sl];
                    const double tx = (bi1.x + bj1.x) * ex + (bi1.y + bj1.y) * ey;
                    const double ty = (bi2.x + bj2.x) * ex + (bi2.y + bj2.y) * ey;
                    // viscous (ph:502-520) and transport (ph:667-685)
                    const double cf = (ex * tx + ey * ty) * P.mu * gj * rinv;
                    ax += cf * dvx;
                    ay += cf * dvy;
                    ix -= gj * tx;
                    iy -= gj * ty;
                    // pressure with Riemann dissipation (ph:891-925)
                    // neighbour's half-step density from its staged pressure: rho = rho0 (1 + p / p0), the EOS
                    // (ph:862) inverted -- 8 B less to gather per neighbour, ~1 ulp from the stored value
                    const double rho_bar = 0.5 * (rho_i + fma(pVj.x, rho_per_p, P.rho0));
                    const double du = dvx * ex + dvy * ey;                          // u_l - u_r
                    const double beta = fmin(3.0 * fmax(du, 0.0), P.c_f);           // ph:1121-1129
                    const double p_avg = 0.5 * (p_i + pVj.x);
                    const double p_star = p_avg + 0.5 * beta * rho_bar * du;
                    const double pg = 0.5 * (p_avg + p_star) * gj;
                    fx -= pg * tx;
                    fy -= pg * ty;
                } else {
                    const double tx = bi1.x * ex + bi1.y * ey;
                    const double ty = bi2.x * ex + bi2.y * ey;
                    const double cf = 4.0 * (ex * tx + ey * ty) * P.mu * gj * rinv;   // no-slip mirror, ph:526
                    ax += cf * dvx;
                    ay += cf * dvy;
                    ix -= 2.0 * gj * tx;                                              // ph:690
                    iy -= 2.0 * gj * ty;
                    has_wall = true;
                }
            }));
            double2 fp;
            fp.x = ax * Vi;
            fp.y = ay * Vi;
            if (add_gravity) fp.x += m_i * P.g;   // m:392
            S.force_prior[i] = fp;
            if (!PLAIN && has_wall) {
                // virtual wall pressure needs the complete F_prior of the walker (ph:929-947)
                const double inv_m = 1.0 / m_i;
                const double wax = fp.x * inv_m, way = fp.y * inv_m;
                SPHB_ROW(true, ([&](const int sl, const int m) {
                    const double Vs = s_pV[sl].y;
                    if (Vs >= 0.0) return;
                    const Geo g = pair_geometry(pi, s_xy[sl], m, P);
                    const double ex = g.ex, ey = g.ey;
                    const double gj = spline_dW(g.r, P) * fabs(Vs);
                    const double ext = -(wax * ex + way * ey);
                    const double pw = p_i + rho_i * g.r * fmax(0.0, ext);   // ph:934
                    const double pg = (p_i + pw) * gj;
                    fx -= pg * (bi1.x * ex + bi1.y * ey);
                    fy -= pg * (bi2.x * ex + bi2.y * ey);
                }));
            }
            double2 x = S.pos[i];
            {   // transport correction (ph:702-710)
                const double n2 = ix * ix + iy * iy;
                double lim = 100.0 * n2 / (P.h * P.h);
                const double scale = P.transport_coeff * P.h * P.h;
                if (lim > 1.0) lim = 1.0;
                if (lim < 0.0) lim = 0.0;
                x.x += scale * lim * ix;
                x.y += scale * lim * iy;
            }
            fx *= Vi;
            fy *= Vi;
            S.force[i] = make_double2(fx, fy);
            S.rp_half[i] = make_double2(rho_i, p_i);
            double2 vn;
            vn.x = vi.x + (fp.x + fx) / m_i * dt;   // ph:1402-1403
            vn.y = vi.y + (fp.y + fy) / m_i * dt;
            S.vel_new[i] = vn;
            S.e_vn[ew] = vn;
            ghost_store(S, key__, i, S.e_vn, vn);
            x.x += 0.5 * dt * vi.x;   // ph:863-864
            x.y += 0.5 * dt * vi.y;
            x.x += 0.5 * dt * vn.x;   // ph:1067-1068
            x.y += 0.5 * dt * vn.y;
            S.pos[i] = x;
        }
    }
    __device__ SPHB_COLD void rows_cold(const StateView &S, const TileDesc &T, const OwnPre &pre, unsigned char *buf, int cap)
    {
        rows<false>(S, T, pre, buf, cap);
    }
};

__device__ __forceinline__ void force_fused_tiles(PhysParams P, const StateView &S, const DevScalars *sc, int add_gravity,
                                                  unsigned cta, unsigned n_ctas)
{
    if (step_poisoned(S.flags)) return;
    load_counts(P);
    ForceK k;
    k.P = P;
    k.dt = sc->dt;   // published by the density launch that opened the step
    k.add_gravity = add_gravity;
    run_tiles(k, S, cta, n_ctas);
}
#ifndef SPHB_FORCE_MINCTAS
// two staging buffers of 80 B per slot: three CTAs per SM fit with two-row tiles (~390 slots), four with three-row
// tiles (~320 slots)
#define SPHB_FORCE_MINCTAS (SPHB_TILE_ROWS >= 3 ? 4 : 3)
#endif
__global__ void __launch_bounds__(kTileThreads, SPHB_FORCE_MINCTAS) k_force_fused(PhysParams P, StateView S,
                                                                                 const DevScalars *sc, int add_gravity)
{
    pdl_wait();
    pdl_trigger_top();
    force_fused_tiles(P, S, sc, add_gravity, blockIdx.x, gridDim.x);
    pdl_trigger_end();
}
__global__ void __launch_bounds__(kTileThreads, SPHB_FORCE_MINCTAS) k_force_fused_batch(const BatchChan *__restrict__ chan,
                                                                                       int add_gravity)
{
    const BatchChan &c = chan[blockIdx.y];
    force_fused_tiles(c.P, c.S, c.sc, add_gravity, blockIdx.x, gridDim.x);
}

// Device-side dual-criteria scheduler (README.md:61-65 of the reference; its constants are those of m:519-527 split
// by what they bound): the outer step from the advection, viscous and body-force criteria, the first substep from
// the acoustic one.
__global__ void k_dual_begin(PhysParams P, DevScalars *sc, double remain)
{
    const double vmax = sc->vmax;
    const double Dt_a = 0.25 * P.h / fmax(vmax, 1e-12);                  // advection
    const double Dt_v = 0.125 * P.h * P.h / fmax(P.nu, 1e-12);           // viscous diffusion
    const double Dt_b = 0.25 * sqrt(P.h / fmax(fabs(P.g), 1e-12));       // body force
    const double Dt = fmax(fmin(fmin(Dt_a, Dt_v), fmin(Dt_b, remain)), 1e-12);
    const double dt_ac = 0.25 * P.h / fmax(P.c_f + vmax, 1e-12);         // acoustic
    sc->Dt = Dt;
    sc->Dt_left = Dt;
    sc->dt = fmin(dt_ac, Dt);
    sc->vmax_bits = 0ull;
    sc->dual_on = 1;
    sc->sub_on = 1;
    sc->outer_steps += 1;
}
__global__ void k_dual_end(DevScalars *sc)
{
    sc->dual_on = 0;
    sc->sub_on = 0;
}

// one thread per channel
__global__ void __launch_bounds__(kBlock) k_publish_dt_batch(const BatchChan *__restrict__ chan, int n_chan, double remain,
                                                            double fixed_dt)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_chan) return;
    DevScalars *sc = chan[k].sc;
    sc->dt = device_dt(chan[k].P, sc, remain, fixed_dt);
    sc->vmax_bits = 0ull;
}

__device__ __forceinline__ double wrap_x(double x, double DL)   // MATLAB mod(x, DL), m:573
{
    if (x >= 0.0 && x < DL) return x;
    double m = fmod(x, DL);
    if (m < 0.0) m += DL;
    return m;
}

// ---------------------------------------------------------------- P5 + density/EOS finish + wrap + max|v|
// staged per slot: position, new velocity (a wall entry: its wall velocity), (-, volume; volume < 0 = wall entry)
struct ContinuityK {
    static constexpr bool kOwnRm = false;
    static constexpr int kFields = 3;
    PhysParams P;
    double dt;
    int do_wrap;
    int track_y;                // GridFuse: reduce the y range of every row for the coming list build
    double speed;
    long long ykmin, ykmax;     // ordered keys (common.cuh)
    __device__ __forceinline__ int bulk_fields(const TileDesc &) const { return 3; }
    __device__ __forceinline__ const double2 *src(const StateView &S, int f) const
    {
        return f == 0 ? S.ent_pos : (f == 1 ? S.e_vn : S.e_pV);
    }
    __device__ __forceinline__ void see_y(double y)
    {
        const long long k = ordered_key(y);
        ykmin = k < ykmin ? k : ykmin;
        ykmax = k > ykmax ? k : ykmax;
    }
    __device__ __forceinline__ void dummy(unsigned char *buf, int cap, int slot) const
    {
        field(buf, cap, 0)[slot] = make_double2(kDummyCoord, kDummyCoord);
        field(buf, cap, 1)[slot] = make_double2(0.0, 0.0);
        field(buf, cap, 2)[slot] = make_double2(0.0, 0.0);
    }
    template <bool PLAIN>
    __device__ __forceinline__ void rows(const StateView &S, const TileDesc &T, const OwnPre &pre, unsigned char *buf, int cap)
    {
        const double2 *s_xy = field(buf, cap, 0), *s_v = field(buf, cap, 1), *s_pV = field(buf, cap, 2);
        SPHB_TILE_WALKERS(ew, si, i)
        {
            if (i >= P.nf) {
                S.drho_dt[i] = 0.0;   // ph:1114-1116
                S.p[i] = 0.0;         // ph:1449
                S.vel[i] = make_double2(0.0, 0.0);   // ph:1405-1408, m:270
                if (track_y) see_y(S.pos[i].y);
                continue;
            }
            const double2 pi = s_xy[si];
            const double2 vi = s_v[si];
            // the row's own global-memory operands are requested before the row loop, which hides their latency
            // (they were the long-scoreboard stalls of this kernel: 3.2 per issue in the r2_v2 capture)
            const double rho_h = S.rp_half[i].x;
            double2 x = S.pos[i];
            double rate = 0.0;
            SPHB_ROW(!PLAIN, ([&](const int sl, const int m) {
                const Geo g = pair_geometry(pi, s_xy[sl], m, P);
                const double Vs = s_pV[sl].y;
                const double gj = spline_dW(g.r, P) * (PLAIN ? Vs : fabs(Vs));
                const double2 vj = s_v[sl];
                double ux, uy;
                if (PLAIN || Vs >= 0.0) {
                    ux = vi.x - vj.x;
                    uy = vi.y - vj.y;
                } else {
                    ux = vi.x - (2.0 * vj.x - vi.x);   // mirrored velocity, ph:1101-1102
                    uy = vi.y - (2.0 * vj.y - vi.y);
                }
                rate += (ux * g.ex + uy * g.ey) * gj;
            }));
            const double drho = rate * rho_h;          // ph:1112
            double rho = rho_h + drho * (0.5 * dt);    // ph:1441
            if (rho < 1e-10) rho = P.rho0;
            S.drho_dt[i] = drho;
            S.rho[i] = rho;
            S.p[i] = P.p0 * (rho / P.rho0 - 1.0);
            if (do_wrap) {
                x.x = wrap_x(x.x, P.DL);
                S.pos[i] = x;
            }
            if (track_y) see_y(x.y);
            S.vel[i] = vi;   // nobody reads `vel` in this kernel (neighbours read the staged new velocity): commit in place
            if (i < P.nfo) speed = fmax(speed, sqrt(vi.x * vi.x + vi.y * vi.y));
            // a blown-up row (NaN / Inf in density or velocity; fmax above drops a NaN silently): raised once, rare
            if (!(fabs(rho) <= 1.0e300 && fabs(vi.x) + fabs(vi.y) <= 1.0e300)) atomicOr(S.flags, kErrNonFinite);
        }
    }
    __device__ SPHB_COLD void rows_cold(const StateView &S, const TileDesc &T, const OwnPre &pre, unsigned char *buf, int cap)
    {
        rows<false>(S, T, pre, buf, cap);
    }
};

__device__ __forceinline__ void continuity_tiles(PhysParams P, const StateView &S, DevScalars *sc, int do_wrap, int advance_clock,
                                                 const GridFuse &F, unsigned cta, unsigned n_ctas)
{
    if (step_poisoned(S.flags)) {
        // nobody takes a clock ticket: t and step stay at the last good step.  The list build that follows inside a
        // replayed graph still counts into the cell counters this kernel was to zero for it (GridFuse): do that much.
        if (F.g != nullptr && advance_clock)
            for (int j = (int)(cta * kTileThreads + threadIdx.x); j < F.n_cell_ints; j += (int)(n_ctas * kTileThreads)) F.cell_cnt[j] = 0;
        return;
    }
    if (sc->dual_on && !sc->sub_on) return;   // device-side scheduler: the outer step is used up
    load_counts(P);
    const bool fuse = F.g != nullptr && advance_clock;
    ContinuityK k;
    k.P = P;
    k.dt = sc->dt;
    k.do_wrap = do_wrap;
    k.track_y = fuse ? 1 : 0;
    k.speed = 0.0;
    k.ykmin = LLONG_MAX;
    k.ykmax = LLONG_MIN;
    run_tiles(k, S, cta, n_ctas);
    // grid-wide max |v| over fluid rows: warp shuffle, then one atomicMax per warp and CTA
    double speed = k.speed;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) speed = fmax(speed, __shfl_xor_sync(0xffffffffu, speed, d));
    if ((threadIdx.x & 31) == 0 && speed > 0.0)
        atomicMax(&sc->vmax_bits, (unsigned long long)__double_as_longlong(speed));
    if (fuse) {
        // for the list build that follows: y range of all rows (one atomic pair per CTA) and zeroed cell counters
        long long kmin = k.ykmin, kmax = k.ykmax;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const long long a = __shfl_xor_sync(0xffffffffu, kmin, d);
            const long long b = __shfl_xor_sync(0xffffffffu, kmax, d);
            kmin = a < kmin ? a : kmin;
            kmax = b > kmax ? b : kmax;
        }
        __shared__ long long s_min[kTileThreads / 32], s_max[kTileThreads / 32];
        if ((threadIdx.x & 31) == 0) {
            s_min[threadIdx.x >> 5] = kmin;
            s_max[threadIdx.x >> 5] = kmax;
        }
        for (int j = (int)(cta * kTileThreads + threadIdx.x); j < F.n_cell_ints; j += (int)(n_ctas * kTileThreads)) F.cell_cnt[j] = 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kTileThreads / 32; ++w) {
                kmin = s_min[w] < kmin ? s_min[w] : kmin;
                kmax = s_max[w] > kmax ? s_max[w] : kmax;
            }
            if (kmin != LLONG_MAX) atomicMin(&F.g->ymin_key, kmin);
            if (kmax != LLONG_MIN) atomicMax(&F.g->ymax_key, kmax);
        }
    }
    if (!advance_clock) return;
    // the last CTA to finish advances the clock and publishes max|v| (what k_advance_clock does as a
    // launch of its own): t += dt, step += 1, m:264-266 -- and derives the grid of the coming build (nb:253-258)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned ticket = atomicAdd(&sc->ticket, 1u);
        if (ticket == n_ctas - 1) {
            __threadfence();
            sc->t += sc->dt;
            sc->step += 1;
            sc->vmax = __longlong_as_double((long long)atomicAdd(&sc->vmax_bits, 0ull));
            sc->ticket = 0u;
            if (sc->dual_on) {   // device-side scheduler: one more substep unless Dt is covered (to rounding)
                sc->Dt_left -= sc->dt;
                sc->sub_steps += 1;
                sc->sub_on = sc->Dt_left > 1e-9 * sc->Dt ? 1 : 0;
            }
            if (fuse && F.derive) grid_derive(F.g, F.cell, F.ncx, F.cell_cap, F.flags, F.totals);
        }
    }
}
#ifndef SPHB_CONT_MINCTAS
#define SPHB_CONT_MINCTAS 5   // (6 -> 170 B of spills, 77 us; 5 -> 57.8 us; 4 -> 58.2 us)
#endif
__global__ void __launch_bounds__(kTileThreads, SPHB_CONT_MINCTAS) k_continuity(PhysParams P, StateView S, DevScalars *sc,
                                                                               int do_wrap, int advance_clock, GridFuse F)
{
    pdl_wait();
    pdl_trigger_top();
    continuity_tiles(P, S, sc, do_wrap, advance_clock, F, blockIdx.x, gridDim.x);
    pdl_trigger_end();
}
__global__ void __launch_bounds__(kTileThreads, SPHB_CONT_MINCTAS) k_continuity_batch(const BatchChan *__restrict__ chan, int do_wrap,
                                                                  int advance_clock, int fuse_grid)
{
    const BatchChan &c = chan[blockIdx.y];
    continuity_tiles(c.P, c.S, c.sc, do_wrap, advance_clock, fuse_grid ? c.F : GridFuse(), blockIdx.x, gridDim.x);
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
// Only wall particles at y <= 0 or y >= DH count (ph:1731-1736) and the pair distance is < 2h, so only the tile
// rows whose neighbourhood reaches the cell rows of y <= 0 or y >= DH can contribute: the CTAs deal exactly those
// tiles among themselves (a few hundred instead of every tile of the channel).  One partial (bottom, top) per CTA;
// the last CTA to finish sums them in CTA order -- deterministic -- and publishes tau (no second launch).
__device__ __forceinline__ void wall_shear_tiles(PhysParams P, const StateView &S, double *partials, DevScalars *sc,
                                                 unsigned cta, unsigned n_ctas)
{
    load_counts(P);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ TileDesc T;
    __shared__ double shb[kTileThreads], sht[kTileThreads];
    __shared__ int s_last;
    const int cap = S.tile_cap;
    double2 *s_xy = reinterpret_cast<double2 *>(smem_raw);
    double *s_V = reinterpret_cast<double *>(s_xy + cap);
    double *s_wvx = s_V + cap;
    // candidate tile rows: cy0 - 1 <= cyb (bottom) or cy0 + kTileRows >= cyt (top)
    const GridDesc *g = S.grid;
    const int tpr = tiles_per_row(S.ncx, S.tile_w);
    const int n_trows = (int)(S.n_tiles / (unsigned)tpr);
    int nb, t0;
    wall_tile_rows(g, P.DH, n_trows, nb, t0);
    const int ncand = (nb + max(0, n_trows - t0)) * tpr;
    double sb = 0.0, st = 0.0;
    for (int q = (int)cta; q < ncand; q += (int)n_ctas) {
        int tr = q / tpr;
        const int tc = q - tr * tpr;
        if (tr >= nb) tr = t0 + (tr - nb);
        __syncthreads();   // the previous tile's readers are done with T and the staging arrays
        tile_load(T, S.tiles, (unsigned)(tr * tpr + tc));
        __syncthreads();
        // (the walk of this build has marked the tiles whose neighbourhood holds a wall entry)
        if (!(T.flags & kTileValid) || !(T.flags & kTileHasWall)) continue;
        const OwnPre pre = tile_prefetch_own(S, T);
        // the walker's own operands are requested before the staging, whose latency hides theirs
        double own_vx = 0.0, own_V = 0.0;
        Mat2 own_B;
        own_B.b11 = own_B.b12 = own_B.b21 = own_B.b22 = 0.0;
        if ((pre.key & 3) == 0 && (pre.key >> 3) < P.nfo) {
            own_vx = S.vel[pre.key >> 3].x;
            own_B = S.B[pre.key >> 3];
            own_V = S.Vol[pre.key >> 3];
        }
        // Staged from the entry-ordered mirrors of the build (position; wall weight m / rho0 = the wall entry's volume,
        // -1 for a fluid entry; wall velocity): three independent coalesced loads per slot instead of the chain
        // ent_key -> Vol[j], wall_vel[j] segment by segment, which was a third of this latency-bound kernel's time.
        {
            struct StageRec {
                double2 xy;
                double w, wvx;
            };
            tile_stage_flat(
                T,
                [&](int e) {
                    StageRec r;
                    r.xy = S.ent_pos[e];
                    r.w = S.e_w[e].x;
                    r.wvx = S.e_vel[e].x;
                    return r;
                },
                [&](int slot, const StageRec &r) {
                    s_xy[slot] = r.xy;
                    s_V[slot] = r.w < 0.0 ? 1.0 : -r.w;   // > 0: skipped like the dummy; wall: minus its volume
                    s_wvx[slot] = r.wvx;
                });
        }
        if (threadIdx.x == 0) {
            s_xy[T.n_staged] = make_double2(kDummyCoord, kDummyCoord);
            s_V[T.n_staged] = 1.0;   // skipped like a fluid neighbour
        }
        __syncthreads();
        SPHB_TILE_WALKERS(ew, si, i)
        {
            if (i >= P.nfo) continue;   // wall rows and halo copies of a slab
            const double2 pi = s_xy[si];
            const double vx = first__ ? own_vx : S.vel[i].x;
            const Mat2 Bi = first__ ? own_B : S.B[i];
            const double Vi = first__ ? own_V : S.Vol[i];
            SPHB_ROW(true, ([&](const int sl, const int m) {
                const double Vs = s_V[sl];
                if (Vs > 0.0) return;
                const double2 pj = s_xy[sl];
                const Geo g2 = pair_geometry(pi, pj, m, P);
                if (!(g2.r > 1e-12)) return;
                const double ex = g2.ex, ey = g2.ey;
                const double eBe = ex * (Bi.b11 * ex + Bi.b12 * ey) + ey * (Bi.b21 * ex + Bi.b22 * ey);
                const double dvx = vx - s_wvx[sl];
                const double f = 4.0 * P.mu * eBe * spline_dW(g2.r, P) * (-Vs) * dvx / (g2.r + 0.01 * P.h) * Vi;
                if (pj.y <= 0.0) sb += f;
                else if (pj.y >= P.DH) st += f;
            }));
        }
    }
    shb[threadIdx.x] = sb;
    sht[threadIdx.x] = st;
    __syncthreads();
    for (int d = kTileThreads / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            shb[threadIdx.x] += shb[threadIdx.x + d];
            sht[threadIdx.x] += sht[threadIdx.x + d];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partials[cta] = shb[0];
        partials[n_ctas + cta] = sht[0];
        __threadfence();
        s_last = atomicAdd(&sc->shear_ticket, 1u) == n_ctas - 1 ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    // the last CTA: every partial is visible (fence + ticket); sum in CTA order
    __threadfence();
    sb = 0.0;
    st = 0.0;
    for (unsigned b = threadIdx.x; b < n_ctas; b += kTileThreads) {
        sb += partials[b];
        st += partials[n_ctas + b];
    }
    shb[threadIdx.x] = sb;
    sht[threadIdx.x] = st;
    __syncthreads();
    for (int d = kTileThreads / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            shb[threadIdx.x] += shb[threadIdx.x + d];
            sht[threadIdx.x] += sht[threadIdx.x + d];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sc->tau_bottom = -shb[0] / P.DL;
        sc->tau_top = -sht[0] / P.DL;
        sc->shear_ticket = 0u;
    }
}
__global__ void __launch_bounds__(kTileThreads) k_wall_shear(PhysParams P, StateView S, double *partials, DevScalars *sc)
{
    pdl_wait();
    pdl_trigger_top();
    wall_shear_tiles(P, S, partials, sc, blockIdx.x, gridDim.x);
    pdl_trigger_end();
}
__global__ void __launch_bounds__(kTileThreads, 6) k_wall_shear_batch(const BatchChan *__restrict__ chan)
{
    const BatchChan &c = chan[blockIdx.y];
    wall_shear_tiles(c.P, c.S, c.shear_partials, c.sc, blockIdx.x, gridDim.x);
}

// ---------------------------------------------------------------- entry-ordered mirrors rebuilt from the rows
// The producing kernels keep the mirrors current inside a step.  A caller that changes rows behind a standing list
// (sphb_ctx_upload without positions) or runs the Verlet operator without a density pass since the last build gets
// them from here: one thread per entry.
__global__ void __launch_bounds__(kBlock) k_refresh_mirrors(PhysParams P, StateView S)
{
    load_counts(P);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.grid->n_entries || (size_t)e >= S.ell_stride) return;
    const int j = S.ent_key[e] >> 3;
    if (j < 0 || j >= P.nt) return;
    const bool fluid = j < P.nf;
    const double V = S.Vol[j];
    S.e_vel[e] = (fluid ? S.vel : S.wall_vel)[j];
    S.e_w[e] = make_double2(fluid ? -1.0 : S.mass[j] * P.inv_rho0, 0.0);
    S.e_pV[e] = make_double2(fluid ? S.rp_half[j].y : 0.0, fluid ? V : -V);
    const Mat2 B = S.B[j];
    S.e_b1[e] = make_double2(B.b11, B.b12);
    S.e_b2[e] = make_double2(B.b21, B.b22);
    S.e_vn[e] = (fluid ? S.vel_new : S.wall_vel)[j];
    S.e_rm[e] = make_double2(fluid ? S.rp_half[j].x : 0.0, S.mass[j]);
}

// ---------------------------------------------------------------- binned u(y) monitor (m:579-605)
__global__ void __launch_bounds__(kBlock) k_profile(PhysParams P, StateView S, double x0, double half_width,
                                                   int n_bins, const double *__restrict__ edges, double *sum,
                                                   double *cnt)
{
    load_counts(P);
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
// bytes of shared memory per staged slot
// bytes of shared memory per staged slot.  The four pipelined kernels hold TWO buffers of kFields 16-byte fields.
constexpr size_t kDensitySlot = 2 * DensityK::kFields * 16, kKgcSlot = 2 * KgcK::kFields * 16, kViscSlot = 72, kPressureSlot = 88,
                 kForceSlot = 2 * ForceK::kFields * 16, kContinuitySlot = 2 * ContinuityK::kFields * 16, kShearSlot = 32;

int csr_max_tile_slots()
{
    return (int)((227 * 1024 - 4096) / kForceSlot);
}

// Persistent launches: one CTA per resident slot of the device (SMs x CTAs per SM at the kernel's register and
// shared-memory footprint), never more than there are tiles.
template <class Kern>
static unsigned persistent_ctas(Kern *kern, size_t smem, unsigned n_tiles)
{
    int dev = 0, sms = 148, per_sm = 1;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    static int sm_count[SmemGrant::kMaxDev] = {};
    if (dev >= 0 && dev < SmemGrant::kMaxDev) {
        if (sm_count[dev] == 0 && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            sm_count[dev] = sms;
        if (sm_count[dev] > 0) sms = sm_count[dev];
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, smem) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 1;
    }
    const unsigned g = (unsigned)sms * (unsigned)per_sm;
    return n_tiles < g ? (n_tiles ? n_tiles : 1u) : g;
}

// dynamic shared memory above 48 KB must be granted per kernel (and per device: SmemGrant)
int csr_configure(int tile_cap)
{
    static SmemGrant g[9];
    const size_t c = (size_t)tile_cap;
    SPHB_TRY(grant_smem(g[0], k_density, c * kDensitySlot));
    SPHB_TRY(grant_smem(g[1], k_kgc, c * kKgcSlot));
    SPHB_TRY(grant_smem(g[2], k_visc_transport<true, true>, c * kViscSlot));
    SPHB_TRY(grant_smem(g[3], k_visc_transport<true, false>, c * kViscSlot));
    SPHB_TRY(grant_smem(g[4], k_visc_transport<false, true>, c * kViscSlot));
    SPHB_TRY(grant_smem(g[5], k_pressure, c * kPressureSlot));
    SPHB_TRY(grant_smem(g[6], k_force_fused, c * kForceSlot));
    SPHB_TRY(grant_smem(g[7], k_continuity, c * kContinuitySlot));
    SPHB_TRY(grant_smem(g[8], k_wall_shear, c * kShearSlot));
    return 0;
}

int csr_density(const PhysParams &P, const StateView &S, DevScalars *sc, cudaStream_t st, bool publish_dt, double remain)
{
    const size_t sm = (size_t)S.tile_cap * kDensitySlot;
    SPHB_LAUNCH_PDL(k_density, persistent_ctas(k_density, sm, S.n_tiles), kTileThreads, sm, st, P, S, sc, publish_dt ? 1 : 0, remain);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_kgc(const PhysParams &P, const StateView &S, cudaStream_t st)
{
    const size_t sm = (size_t)S.tile_cap * kKgcSlot;
    SPHB_LAUNCH_PDL(k_kgc, persistent_ctas(k_kgc, sm, S.n_tiles), kTileThreads, sm, st, P, S);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_visc_transport(const PhysParams &P, const StateView &S, bool do_visc, bool add_gravity, bool do_transport,
                       cudaStream_t st)
{
    const unsigned g = S.n_tiles;
    const size_t sm = (size_t)S.tile_cap * kViscSlot;
    if (do_visc && do_transport) k_visc_transport<true, true><<<g, kTileThreads, sm, st>>>(P, S, add_gravity ? 1 : 0);
    else if (do_visc) k_visc_transport<true, false><<<g, kTileThreads, sm, st>>>(P, S, add_gravity ? 1 : 0);
    else if (do_transport) k_visc_transport<false, true><<<g, kTileThreads, sm, st>>>(P, S, 0);
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
    k_pressure<<<S.n_tiles, kTileThreads, (size_t)S.tile_cap * kPressureSlot, st>>>(P, S, sc);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_continuity(const PhysParams &P, const StateView &S, DevScalars *sc, bool wrap, bool advance_clock,
                   const GridFuse &F, cudaStream_t st)
{
    const size_t sm = (size_t)S.tile_cap * kContinuitySlot;
    SPHB_LAUNCH_PDL(k_continuity, persistent_ctas(k_continuity, sm, S.n_tiles), kTileThreads, sm, st, P, S, sc, wrap ? 1 : 0,
                    advance_clock ? 1 : 0, F);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_dual_begin(const PhysParams &P, DevScalars *sc, double remain, cudaStream_t st)
{
    k_dual_begin<<<1, 1, 0, st>>>(P, sc, remain);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_dual_end(DevScalars *sc, cudaStream_t st)
{
    k_dual_end<<<1, 1, 0, st>>>(sc);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
// viscous + transport + pressure + kick in one gather (the first substep of an outer step; after the csr_density that
// published the step's dt)
int csr_force_fused(const PhysParams &P, const StateView &S, const DevScalars *sc, bool add_gravity, cudaStream_t st)
{
    const size_t sm = (size_t)S.tile_cap * kForceSlot;
    SPHB_LAUNCH_PDL(k_force_fused, persistent_ctas(k_force_fused, sm, S.n_tiles), kTileThreads, sm, st, P, S, sc, add_gravity ? 1 : 0);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_vmax(const PhysParams &P, const double2 *vel, DevScalars *sc, cudaStream_t st)
{
    int blocks = (P.nf + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    k_vmax_reset<<<1, 1, 0, st>>>(sc);
    k_vmax<<<blocks, 256, 0, st>>>(P.nfo, vel, sc);
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
// CTAs of the wall-shear kernel: the tiles of six tile rows (bottom + top wall neighbourhoods of a channel whose
// walls are BW = 4 dp thick; more candidate rows are dealt round-robin), never more than there are tiles.  The
// partials buffer holds 2 x n_tiles doubles.
static unsigned shear_ctas(unsigned n_tiles, int ncx, int tile_w)
{
    const unsigned g = 6u * (unsigned)tiles_per_row(ncx, tile_w);
    return n_tiles < g ? (n_tiles ? n_tiles : 1u) : g;
}
int csr_wall_shear(const PhysParams &P, const StateView &S, DevScalars *sc, double *partials, cudaStream_t st)
{
    SPHB_LAUNCH_PDL(k_wall_shear, shear_ctas(S.n_tiles, S.ncx, S.tile_w), kTileThreads, (size_t)S.tile_cap * kShearSlot, st, P, S, partials, sc);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
// ---- batched launches (ops_csr.cuh): grid = (the channel's own grid, channels) ----------------------------
int csr_batch_configure(int tile_cap)
{
    static SmemGrant g[6];
    const size_t c = (size_t)tile_cap;
    SPHB_TRY(grant_smem(g[0], k_density_batch, c * kDensitySlot));
    SPHB_TRY(grant_smem(g[1], k_kgc_batch, c * kKgcSlot));
    SPHB_TRY(grant_smem(g[2], k_pressure_batch, c * kPressureSlot));
    SPHB_TRY(grant_smem(g[3], k_force_fused_batch, c * kForceSlot));
    SPHB_TRY(grant_smem(g[4], k_continuity_batch, c * kContinuitySlot));
    SPHB_TRY(grant_smem(g[5], k_wall_shear_batch, c * kShearSlot));
    return 0;
}
static inline dim3 batch_grid(const BatchLaunch &L, unsigned gx) { return dim3(gx, (unsigned)L.n_chan, 1u); }
// pipelined kernels: a few tiles per CTA, so that the copies of one tile overlap the rows of the one before
static unsigned batch_tiles_per_cta()
{
    static int n = -1;   // 0: by the size of the launch (batch_tile_grid)
    if (n < 0) {
        n = 0;
        if (const char *env = std::getenv("SPHB_BATCH_TILES_PER_CTA")) {   // tuning runs only
            const int v = std::atoi(env);
            if (v >= 1 && v <= 64) n = v;
        }
    }
    return (unsigned)n;
}
// Tiles per CTA of the pipelined kernels: the more, the better the copies of one tile hide behind the rows of the one
// before (measured on 512 channels, r2 call 16: 1 -> 7.45e8, 3 -> 7.66e8, 6 -> 7.81e8, 12 -> 7.96e8 particle-substeps/s),
// as long as a few thousand CTAs are left to fill the device.
static inline dim3 batch_tile_grid(const BatchLaunch &L)
{
    unsigned per = batch_tiles_per_cta();
    if (per == 0) {
        const unsigned long long total = (unsigned long long)L.n_tiles * (unsigned long long)L.n_chan;
        per = (unsigned)std::min<unsigned long long>(16ull, std::max<unsigned long long>(3ull, total / 4096ull));
    }
    return dim3((L.n_tiles + per - 1u) / per, (unsigned)L.n_chan, 1u);
}

int csr_batch_density(const BatchLaunch &L)
{
    k_density_batch<<<batch_tile_grid(L), kTileThreads, (size_t)L.tile_cap * kDensitySlot, L.st>>>(L.chan);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_batch_kgc(const BatchLaunch &L)
{
    k_kgc_batch<<<batch_tile_grid(L), kTileThreads, (size_t)L.tile_cap * kKgcSlot, L.st>>>(L.chan);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_batch_publish_dt(const BatchLaunch &L, double remain, double fixed_dt)
{
    k_publish_dt_batch<<<grid_for((size_t)L.n_chan), kBlock, 0, L.st>>>(L.chan, L.n_chan, remain, fixed_dt);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_batch_force_fused(const BatchLaunch &L, bool add_gravity)
{
    k_force_fused_batch<<<batch_tile_grid(L), kTileThreads, (size_t)L.tile_cap * kForceSlot, L.st>>>(
        L.chan, add_gravity ? 1 : 0);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_batch_time_step(const BatchLaunch &L, double remain, double fixed_dt)
{
    k_time_step_batch<<<batch_grid(L, grid_for((size_t)L.n_fluid)), kBlock, 0, L.st>>>(L.chan, remain, fixed_dt);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_batch_pressure(const BatchLaunch &L)
{
    k_pressure_batch<<<batch_grid(L, L.n_tiles), kTileThreads, (size_t)L.tile_cap * kPressureSlot, L.st>>>(L.chan);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_batch_continuity(const BatchLaunch &L, bool wrap, bool advance_clock, bool fuse_grid)
{
    k_continuity_batch<<<batch_tile_grid(L), kTileThreads, (size_t)L.tile_cap * kContinuitySlot, L.st>>>(
        L.chan, wrap ? 1 : 0, advance_clock ? 1 : 0, fuse_grid ? 1 : 0);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
int csr_batch_wall_shear(const BatchLaunch &L)
{
    k_wall_shear_batch<<<batch_grid(L, shear_ctas(L.n_tiles, L.ncx, L.tile_w)), kTileThreads, (size_t)L.tile_cap * kShearSlot,
                         L.st>>>(L.chan);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

int csr_refresh_mirrors(const PhysParams &P, const StateView &S, cudaStream_t st)
{
    k_refresh_mirrors<<<grid_for(S.ell_stride), kBlock, 0, st>>>(P, S);
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
