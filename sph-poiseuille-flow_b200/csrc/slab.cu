// slab.cu -- x-slab sharding of one long periodic channel over several GPUs.
//
// The reference has no domain decomposition; this is new design, checked against the
// single-domain run.  One context per rank owns the fluid particles whose wrapped x lies in the cell
// columns [col_lo, col_hi) and keeps H halo columns of copies on either side.  Interaction radius = one
// cell, and one reference step consumes four columns of halo validity (list -> density -> KGC ->
// forces -> continuity), so with H = 4 a rank computes a whole step -- redundantly on its halo copies --
// WITHOUT any mid-step exchange.  Once per step, before the list rebuild:
//     pack    owned rows are classified by their new column: stay / migrate left / migrate right, and
//             stayers within H columns of an edge are also copied for that neighbour (deterministic
//             block-scan ranks, no atomics); records = x y vx vy rho drho_dt mass Vol B(4) id.  A row that
//             migrates is marked dead IN PLACE (the list build skips it, so no kernel visits it again)
//     (the caller moves the two buffers to the neighbours and the scalars to everybody: one NCCL all-to-all)
//     unpack  incoming rows are APPENDED: owned += immigrants; halo = received copies + the rows that just
//             left (their new owner sits next door); copies from across the periodic seam are shifted by
//             -+DL exactly like the reference's ghost entries (nb:284,291); wall rows (static) follow
//     compact (now and then) squeezes the dead rows out of the owned region
// plus max|v| (MAX: time step, m:519-527) and the wall-shear partials (SUM) reduced over the ranks.
// The row counts {nf, owned, nt} live in DEVICE memory and every kernel reads them there
// (PhysParams::dev_counts); launch sizes are the capacity.  A step therefore has no host synchronisation
// and can be captured -- kernels and collective -- as one CUDA graph (slab.py).
// Positions stay in GLOBAL coordinates (the cell column is global, counted from col0), so the pair
// arithmetic away from the seam is bit-identical to the single-domain run.
#include "context.cuh"

namespace {

constexpr int kRec = 16;        // doubles per record
constexpr int kHdr = 16;        // doubles of buffer header: [0] migrants, [1] halo copies
constexpr int kSlabBlock = 256;
enum { cStay = 0, cMigL, cHaloL, cMigR, cHaloR, cN };

// ---- peer-to-peer mailboxes ------------------------------------------------------------------------------------
// Every rank owns a MAILBOX in its device memory, two parities of
//     [records from the left neighbour | records from the right neighbour | 4 doubles of scalars per rank]
// A step's sequence number lives in device memory and is bumped by the first kernel of the exchange, so a replayed
// CUDA graph needs no host-side argument; parity = seq & 1.  A sender writes the records straight into the
// receiver's mailbox (peer pointer: NVLink stores), fences system-wide and then writes `seq` into header slot 2; the
// receiver's kernels spin on that slot (with a time-out that raises kErrSlabCap instead of hanging the GPU).  Two
// parities suffice: a rank can be at most one exchange ahead of a neighbour, because it needs that neighbour's
// records (or scalars) of exchange s before it can send exchange s + 1.
struct MailLayout {
    int world, rec_cap;
    __host__ __device__ size_t buf() const { return (size_t)kHdr + (size_t)rec_cap * kRec; }
    __host__ __device__ size_t parity_stride() const { return 2 * buf() + 4 * (size_t)world; }
    __host__ __device__ size_t total() const { return 2 * parity_stride(); }
    __host__ __device__ size_t from_left(int p) const { return (size_t)p * parity_stride(); }
    __host__ __device__ size_t from_right(int p) const { return (size_t)p * parity_stride() + buf(); }
    __host__ __device__ size_t scal(int p, int r) const { return (size_t)p * parity_stride() + 2 * buf() + 4 * (size_t)r; }
};
struct P2PState {
    long long seq;
    unsigned ticket;
    unsigned pad;
    double red[3];
};
struct P2PSend {
    P2PState *state;        // NULL: no peer-to-peer push (the caller moves the send buffers itself)
    double *left_mail;      // mailbox of the left / right neighbour (peer pointers)
    double *right_mail;
    MailLayout L;
};

// loads that bypass L1 (a peer may have written the line while this kernel is resident)
__device__ __forceinline__ double4 ldcg4(const double4 *p)
{
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(p)), b = __ldcg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

constexpr long long kSpinCycles = 4000000000LL;   // ~2 s at 2 GHz: a peer that never arrives is an error, not a hang

// load with acquire semantics at system scope: what follows it (in this thread, and in the threads it then meets at a
// CTA barrier) sees everything the writer released before storing the value -- without the full membar.sys that
// showed up as a quarter of the receive kernel's stall samples (r2 call 22)
__device__ __forceinline__ double ld_acquire_sys(const double *p)
{
#ifdef SPHB_CPUSIM
    return *reinterpret_cast<const volatile double *>(p);
#else
    double v;
    asm volatile("ld.acquire.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
#endif
}

// waits until *slot == seq (a double written by a peer behind a system-wide fence); false on time-out
__device__ __forceinline__ bool mail_wait(const double *slot, double seq)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(slot) != seq) {
        if (clock64() - t0 > kSpinCycles) return false;
        __nanosleep(64);
    }
    return true;
}

struct PeerList {
    double *mail[SlabState::kMaxWorld];
};

struct SlabGeom {
    int col_lo, width, halo, ncx_g;
    double cell, inv_cell, DL;
    const int *cur;   // device {nf, nfo, nt} of the current rows (counts + 12)
};

// class bits of an owned fluid row.  Owned rows are kept wrapped by the continuity kernel, so the wrap
// (nb:270) is a rare branch, and the column is taken with a multiply: a row within rounding of a column
// boundary may be classed one column off the list build's exact binning (nb:253), which only moves it
// between "owned near the edge" and "first halo column" -- both are valid places for it.
__device__ __forceinline__ unsigned slab_class(double x, const SlabGeom &G)
{
    if (!(x >= 0.0 && x < G.DL)) x -= floor(x / G.DL) * G.DL;
    int c = (int)(x * G.inv_cell);
    if (c >= G.ncx_g) c = G.ncx_g - 1;
    if (c < 0) c = 0;
    int d = c - G.col_lo;
    if (d < 0) d += G.ncx_g;
    unsigned m = 0;
    if (d < G.width) {
        m |= 1u << cStay;
        if (d < G.halo) m |= 1u << cHaloL;
        if (d >= G.width - G.halo) m |= 1u << cHaloR;
        // Towards the periodic seam the copies must cover `halo` CELLS of x, not `halo` columns: the last column
        // of the channel is a sliver unless DL is a whole number of cells (nb:254), and a step consumes four full
        // cells of validity.  (Found by the slab fuzz on the simulator: with a 0.08-cell sliver the rows within a
        // cell of x = 0 were off by 1e-6 in drho/dt.)
        else if (G.col_lo + G.width == G.ncx_g && x >= G.DL - (double)G.halo * G.cell) m |= 1u << cHaloR;
    } else if (d < G.width + (G.ncx_g - G.width + 1) / 2) {
        m |= 1u << cMigR;
    } else {
        m |= 1u << cMigL;
    }
    return m;
}

__global__ void __launch_bounds__(kSlabBlock) k_slab_count(const double2 *__restrict__ pos,
                                                          const int *__restrict__ pid, SlabGeom G,
                                                          int *__restrict__ block_tot)
{
    __shared__ int tot[cN];
    if (threadIdx.x < cN) tot[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned m = (i < G.cur[1] && pid[i] >= 0) ? slab_class(pos[i].x, G) : 0u;   // pid < 0: a row that has left
#pragma unroll
    for (int k = 0; k < cN; ++k) {
        const unsigned b = __ballot_sync(0xffffffffu, (m >> k) & 1u);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(&tot[k], __popc(b));
    }
    __syncthreads();
    if (threadIdx.x < cN) block_tot[threadIdx.x * gridDim.x + blockIdx.x] = tot[threadIdx.x];
}

// Peer-to-peer exchange, opening move (one CTA of >= world threads): bumps the sequence number and writes this rank's
// scalars (max|v|, wall-shear partials) into the mailbox of EVERY rank (thread r -> rank r), the sequence number last.
__device__ __forceinline__ void p2p_open_cta(const DevScalars *sc, P2PState *st, const PeerList &peers, const MailLayout &L, int me)
{
    __shared__ long long s_seq;
    if (threadIdx.x == 0) {
        st->seq += 1;
        s_seq = st->seq;
    }
    __syncthreads();
    const int r = (int)threadIdx.x;
    if (r >= L.world) return;
    double *slot = peers.mail[r] + L.scal((int)(s_seq & 1LL), me);
    slot[0] = sc->vmax;
    slot[1] = sc->tau_bottom;
    slot[2] = sc->tau_top;
    __threadfence_system();
    *reinterpret_cast<volatile double *>(slot + 3) = (double)s_seq;
}
// ... and its counterpart (one CTA of >= world threads): waits for the scalars of every rank in the own mailbox,
// reduces them in rank order (MAX / SUM: deterministic), imports max|v| (the time step of the next step, m:519-527)
__device__ __forceinline__ void p2p_scalars_cta(DevScalars *sc, P2PState *st, const double *mail, const MailLayout &L, int *flags)
{
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    const int r = (int)threadIdx.x;
    const int par = (int)(st->seq & 1LL);
    if (r < L.world && !mail_wait(mail + L.scal(par, r) + 3, (double)st->seq)) bad = 1;
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (bad) {
        atomicOr(flags, kErrSlabCap);
        return;
    }
    double vmax = 0.0, tb = 0.0, tt = 0.0;
    for (int q = 0; q < L.world; ++q) {
        const double *t = mail + L.scal(par, q);
        vmax = fmax(vmax, __ldcg(t));
        tb += __ldcg(t + 1);
        tt += __ldcg(t + 2);
    }
    sc->vmax = vmax;
    st->red[0] = vmax;
    st->red[1] = tb;
    st->red[2] = tt;
}

// exclusive scan of the block totals of each class (one CTA per class; blocks <= a few thousand), totals -> counts.
// With peer-to-peer mailboxes one more CTA (blockIdx.x == cN) opens the exchange meanwhile (p2p_open_cta).
__global__ void __launch_bounds__(1024) k_slab_offsets(int *__restrict__ block_tot, int blocks, int *__restrict__ counts,
                                                      const DevScalars *sc, P2PState *p2p, PeerList peers, MailLayout ML, int me)
{
    if (blockIdx.x == cN) {
        if (p2p != nullptr) p2p_open_cta(sc, p2p, peers, ML, me);
        return;
    }
    __shared__ int warp_sum[32];
    __shared__ int chunk_total, carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    {
        const int k = blockIdx.x;
        int *v = block_tot + (size_t)k * blocks;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (int base = 0; base < blocks; base += 1024) {
            const int idx = base + (int)threadIdx.x;
            const int x = idx < blocks ? v[idx] : 0;
            int inc = x;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) warp_sum[wid] = inc;
            __syncthreads();
            if (wid == 0) {
                const int s = warp_sum[lane];
                int si = s;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, si, d);
                    if (lane >= d) si += t;
                }
                warp_sum[lane] = si - s;
                if (lane == 31) chunk_total = si;
            }
            __syncthreads();
            if (idx < blocks) v[idx] = inc - x + warp_sum[wid] + carry;
            __syncthreads();
            if (threadIdx.x == 0) carry += chunk_total;
            __syncthreads();
        }
        if (threadIdx.x == 0) counts[k] = carry;
        __syncthreads();
    }
}

struct SlabArrays {
    double2 *pos, *vel;     // current rows (appended to in place by the unpack)
    double *rho, *drho, *mass, *Vol;
    Mat2 *B;
    double2 *n_pos, *n_vel;
    double *n_rho, *n_drho, *n_mass, *n_Vol;
    Mat2 *n_B;
    int *pid;         // global particle identity, carried through every migration; -1 = the row has left
    int *n_pid;
};

__device__ __forceinline__ void write_record(double *rec, double2 x, double2 v, double rho, double drho, double mass,
                                             double Vol, const Mat2 &B, int pid)
{
    double4 *r = reinterpret_cast<double4 *>(rec);
    r[0] = make_double4(x.x, x.y, v.x, v.y);
    r[1] = make_double4(rho, drho, mass, Vol);
    r[2] = make_double4(B.b11, B.b12, B.b21, B.b22);
    r[3] = make_double4((double)pid, 0.0, 0.0, 0.0);
}

// arrival of a rank's records at its two neighbours: every CTA's stores are fenced system-wide before its ticket; the
// last CTA publishes the sequence number in both mailboxes (all threads of the CTA call this)
__device__ __forceinline__ void p2p_arrive(P2PState *st, double *peer_left, double *peer_right, double seq)
{
    // one system-wide fence per CTA, by the thread that takes the ticket, behind the CTA barrier (fences are cumulative:
    // it covers the stores of every thread that reached the barrier) -- not one per thread
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(&st->ticket, 1u) == gridDim.x - 1) {
            __threadfence_system();
            *reinterpret_cast<volatile double *>(peer_left + 2) = seq;
            *reinterpret_cast<volatile double *>(peer_right + 2) = seq;
            st->ticket = 0u;
        }
    }
}

// SEND: migrants and halo copies -> the two send buffers [migrants | copies], migrants marked dead in place;
// !SEND (now and then, sphb_ctx_slab_compact): the rows that stay -> compacted alternate arrays.
template <bool SEND>
__global__ void __launch_bounds__(kSlabBlock) k_slab_scatter(SlabArrays A, SlabGeom G,
                                                            const int *__restrict__ block_off,
                                                            const int *__restrict__ counts, double *send_left,
                                                            double *send_right, int rec_cap, int *flags, P2PSend X)
{
    // peer-to-peer push: every record also goes straight into the neighbour's mailbox (what this rank sends to its
    // LEFT arrives "from the right" over there), the arrival slot is written by the last CTA to finish
    double *peer_left = nullptr, *peer_right = nullptr;
    double seq = 0.0;
    if (SEND && X.state != nullptr) {
        seq = (double)X.state->seq;
        const int par = (int)(X.state->seq & 1LL);
        peer_left = X.left_mail + X.L.from_right(par);
        peer_right = X.right_mail + X.L.from_left(par);
    }
    __shared__ int warp_off[cN][kSlabBlock / 32];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned m = (i < G.cur[1] && A.pid[i] >= 0) ? slab_class(A.pos[i].x, G) : 0u;
    int rank[cN];
#pragma unroll
    for (int k = 0; k < cN; ++k) {
        const unsigned b = __ballot_sync(0xffffffffu, (m >> k) & 1u);
        rank[k] = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) warp_off[k][wid] = __popc(b);
    }
    __syncthreads();
    if (threadIdx.x < cN) {
        int s = 0;
        for (int w = 0; w < kSlabBlock / 32; ++w) {
            const int t = warp_off[threadIdx.x][w];
            warp_off[threadIdx.x][w] = s;
            s += t;
        }
    }
    __syncthreads();
    if (SEND && i == 0) {
        const int nl = counts[cMigL] + counts[cHaloL], nr = counts[cMigR] + counts[cHaloR];
        send_left[0] = (double)counts[cMigL];
        send_left[1] = (double)counts[cHaloL];
        send_right[0] = (double)counts[cMigR];
        send_right[1] = (double)counts[cHaloR];
        if (peer_left) {
            peer_left[0] = (double)counts[cMigL];
            peer_left[1] = (double)counts[cHaloL];
            peer_right[0] = (double)counts[cMigR];
            peer_right[1] = (double)counts[cHaloR];
        }
        if (nl > rec_cap || nr > rec_cap) atomicOr(flags, kErrSlabCap);
    }
    m &= SEND ? ~(1u << cStay) : (1u << cStay);
    if (m) {
    const double2 x = A.pos[i], v = A.vel[i];
    const double rho = A.rho[i], drho = A.drho[i], mass = A.mass[i], Vol = A.Vol[i];
    const Mat2 B = A.B[i];
    const int pid = A.pid[i];
#pragma unroll
    for (int k = 0; k < cN; ++k) {
        if (!((m >> k) & 1u)) continue;
        const int d = block_off[(size_t)k * gridDim.x + blockIdx.x] + warp_off[k][wid] + rank[k];
        if (k == cStay) {
            A.n_pos[d] = x;
            A.n_vel[d] = v;
            A.n_rho[d] = rho;
            A.n_drho[d] = drho;
            A.n_mass[d] = mass;
            A.n_Vol[d] = Vol;
            A.n_B[d] = B;
            A.n_pid[d] = pid;
        } else {
            const bool left = k == cMigL || k == cHaloL;
            const int base = (k == cHaloL) ? counts[cMigL] : (k == cHaloR ? counts[cMigR] : 0);
            const int slot = base + d;
            if (slot < rec_cap) {
                write_record((left ? send_left : send_right) + kHdr + (size_t)slot * kRec, x, v, rho, drho, mass, Vol, B, pid);
                if (peer_left) write_record((left ? peer_left : peer_right) + kHdr + (size_t)slot * kRec, x, v, rho, drho, mass, Vol, B, pid);
            }
        }
    }
    // a row that migrates stays where it is, marked dead: the neighbour build skips it, so no kernel of the
    // step ever visits it; sphb_ctx_slab_compact squeezes the dead rows out now and then
    if (SEND && (m & ((1u << cMigL) | (1u << cMigR)))) A.pid[i] = -1;
    }
    if (SEND && peer_left) p2p_arrive(X.state, peer_left, peer_right, seq);
}

// ---- pack over the EDGE CANDIDATES only ---------------------------------------------------------------------------
// A row moves far less than a cell column per step (dt <= 0.25 h / (c_f + max|v|), m:519-527, and a column is 2h), so
// the only rows that can have anything to send -- migrants and the copies of the `halo` columns next to an edge -- are
// the ones that sat within halo + 2 columns of an edge when the last list was built.  That build left every row in
// (cell row, cell column) order (the entry array), so the candidates of one cell row and one side are ONE contiguous
// entry range: a warp per (cell row, side) instead of a thread per owned row -- ~2 % of the rows, three small launches.
constexpr int kEdgeBand = 2;     // columns beyond the halo band on the inner side (+1 outwards: the multiply of
                                 // slab_class can class a row on a column boundary one column off the exact binning)
constexpr int kEdgeWarps = 4;    // warps (tasks) per CTA
struct EdgeView {
    const int *ent_key;      // entry -> (row << 3 | kind) of the last build
    const int *cell_start;   // exclusive scan of the cell counts, cell = cell row * ncx + local column
    const GridDesc *g;       // ncy, n_entries of that build
    int ncx, cell_cap;
    int a[2], b[2];          // local columns [a, b) of the left / right candidate band
};
__device__ __forceinline__ bool edge_range(const EdgeView &E, int task, int &e0, int &e1)
{
    // (cell rows beyond the grid of the last build are empty: the scan covers every planned cell, so their ranges
    // collapse -- no need to wait for the grid descriptor first)
    const int row = task >> 1, side = task & 1;
    const int i0 = row * E.ncx + E.a[side], i1 = row * E.ncx + E.b[side];
    if (i1 >= E.cell_cap) {
        const int n_entries = E.g->n_entries;
        e0 = i0 < E.cell_cap ? E.cell_start[i0] : n_entries;
        e1 = n_entries;
    } else {
        e0 = E.cell_start[i0];
        e1 = E.cell_start[i1];
    }
    return e1 > e0;
}
// class bits (without cStay) of the candidate at entry e, 0 for entries that are not live owned rows
__device__ __forceinline__ unsigned edge_class(const EdgeView &E, const SlabArrays &A, const SlabGeom &G, int e, int &row)
{
    const int key = E.ent_key[e];
    row = key >> 3;
    if ((key & 3) != 0 || row >= G.cur[1] || A.pid[row] < 0) return 0u;
    return slab_class(A.pos[row].x, G) & ~(1u << cStay);
}
__global__ void __launch_bounds__(kEdgeWarps * 32) k_slab_edge_count(EdgeView E, SlabArrays A, SlabGeom G, int n_tasks,
                                                                    int *__restrict__ task_tot)
{
    const int lane = threadIdx.x & 31, task = blockIdx.x * kEdgeWarps + (threadIdx.x >> 5);
    if (task >= n_tasks) return;
    int tot[cN] = {0, 0, 0, 0, 0};
    int e0 = 0, e1 = 0;
    if (edge_range(E, task, e0, e1)) {
        for (int base = e0; base < e1; base += 32) {
            int row;
            const unsigned m = base + lane < e1 ? edge_class(E, A, G, base + lane, row) : 0u;
#pragma unroll
            for (int k = 1; k < cN; ++k) tot[k] += __popc(__ballot_sync(0xffffffffu, (m >> k) & 1u));
        }
    }
    if (lane < cN) task_tot[lane * n_tasks + task] = tot[lane];
}

// ... and the scatter over the same candidates (task_off: the exclusive scan of the counts per class, k_slab_offsets)
__global__ void __launch_bounds__(kEdgeWarps * 32) k_slab_edge_scatter(EdgeView E, SlabArrays A, SlabGeom G, int n_tasks,
                                                                      const int *__restrict__ task_off,
                                                                      const int *__restrict__ counts, double *send_left,
                                                                      double *send_right, int rec_cap, int *flags, P2PSend X)
{
    double *peer_left = nullptr, *peer_right = nullptr;
    double seq = 0.0;
    if (X.state != nullptr) {
        seq = (double)X.state->seq;
        const int par = (int)(X.state->seq & 1LL);
        peer_left = X.left_mail + X.L.from_right(par);
        peer_right = X.right_mail + X.L.from_left(par);
    }
    const int lane = threadIdx.x & 31, task = blockIdx.x * kEdgeWarps + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int nl = counts[cMigL] + counts[cHaloL], nr = counts[cMigR] + counts[cHaloR];
        send_left[0] = (double)counts[cMigL];
        send_left[1] = (double)counts[cHaloL];
        send_right[0] = (double)counts[cMigR];
        send_right[1] = (double)counts[cHaloR];
        if (peer_left) {
            peer_left[0] = (double)counts[cMigL];
            peer_left[1] = (double)counts[cHaloL];
            peer_right[0] = (double)counts[cMigR];
            peer_right[1] = (double)counts[cHaloR];
        }
        if (nl > rec_cap || nr > rec_cap) atomicOr(flags, kErrSlabCap);
    }
    int e0 = 0, e1 = 0;
    if (task < n_tasks && edge_range(E, task, e0, e1)) {
        int off[cN];
#pragma unroll
        for (int k = 1; k < cN; ++k) off[k] = task_off[k * n_tasks + task];
        for (int base = e0; base < e1; base += 32) {
            int i = 0;
            const unsigned m = base + lane < e1 ? edge_class(E, A, G, base + lane, i) : 0u;
            double2 x = make_double2(0.0, 0.0), v = x;
            double rho = 0.0, drho = 0.0, mass = 0.0, Vol = 0.0;
            Mat2 B{};
            int pid = -1;
            if (m) {
                x = A.pos[i];
                v = A.vel[i];
                rho = A.rho[i];
                drho = A.drho[i];
                mass = A.mass[i];
                Vol = A.Vol[i];
                B = A.B[i];
                pid = A.pid[i];
            }
#pragma unroll
            for (int k = 1; k < cN; ++k) {
                const unsigned b = __ballot_sync(0xffffffffu, (m >> k) & 1u);
                if ((m >> k) & 1u) {
                    const bool left = k == cMigL || k == cHaloL;
                    const int first = (k == cHaloL) ? counts[cMigL] : (k == cHaloR ? counts[cMigR] : 0);
                    const int slot = first + off[k] + __popc(b & ((1u << lane) - 1u));
                    if (slot < rec_cap) {
                        write_record((left ? send_left : send_right) + kHdr + (size_t)slot * kRec, x, v, rho, drho, mass, Vol, B, pid);
                        if (peer_left)
                            write_record((left ? peer_left : peer_right) + kHdr + (size_t)slot * kRec, x, v, rho, drho, mass, Vol, B, pid);
                    }
                }
                off[k] += __popc(b);
            }
            // a row that migrates stays where it is, marked dead (k_slab_scatter)
            if (m & ((1u << cMigL) | (1u << cMigR))) A.pid[i] = -1;
        }
    }
    if (peer_left) p2p_arrive(X.state, peer_left, peer_right, seq);
}

// The three launches above as ONE (steady state, at most one CTA per SM so that all of them are resident): count | grid
// barrier | every warp sums the totals of the tasks before its own | scatter.  With peer-to-peer mailboxes one more CTA
// opens the exchange before the barrier (sequence number, scalars to every rank) and the last CTA to finish publishes
// the arrival in both neighbours' mailboxes.  counts[15]: barrier arrivals, counts[7]: finish ticket (both reset by the
// last CTA).
__global__ void __launch_bounds__(kEdgeWarps * 32) k_slab_edge_send(EdgeView E, SlabArrays A, SlabGeom G, int n_tasks,
                                                                   int *task_tot, int *counts, double *send_left,
                                                                   double *send_right, int rec_cap, int *flags, P2PSend X,
                                                                   const DevScalars *sc, PeerList peers, int me)
{
    const int lane = threadIdx.x & 31, task = blockIdx.x * kEdgeWarps + (threadIdx.x >> 5);
    const bool open_cta = X.state != nullptr && blockIdx.x == gridDim.x - 1;
    int e0 = 0, e1 = 0;
    bool have = false;
    unsigned m0 = 0u, m1 = 0u;   // class bits and rows of the first two passes, kept across the barrier for the scatter
    int i0 = 0, i1 = 0;
    if (open_cta) {
        if (threadIdx.x == 0) X.state->seq += 1;   // (the scalars travel after the barrier, beside the scatter)
    } else if (task < n_tasks) {
        // ---- count
        int tot[cN] = {0, 0, 0, 0, 0};
        have = edge_range(E, task, e0, e1);
        if (have) {
            int pass = 0;
            for (int base = e0; base < e1; base += 32, ++pass) {
                int row = 0;
                const unsigned m = base + lane < e1 ? edge_class(E, A, G, base + lane, row) : 0u;
                if (pass == 0) {
                    m0 = m;
                    i0 = row;
                } else if (pass == 1) {
                    m1 = m;
                    i1 = row;
                }
#pragma unroll
                for (int k = 1; k < cN; ++k) tot[k] += __popc(__ballot_sync(0xffffffffu, (m >> k) & 1u));
            }
        }
        if (lane == 0) reinterpret_cast<int4 *>(task_tot)[task] = make_int4(tot[1], tot[2], tot[3], tot[4]);
    }
    // ---- grid barrier (every CTA is resident: the host launches at most one per SM)
    __shared__ int s_ok;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile int *arrive = counts + 15;
        atomicAdd(counts + 15, 1);
        const long long t0 = clock64();
        s_ok = 1;
        while (*arrive < (int)gridDim.x) {
            if (clock64() - t0 > kSpinCycles) {
                s_ok = 0;
                break;
            }
            __nanosleep(32);
        }
        if (!s_ok) atomicOr(flags, kErrSlabCap);
        __threadfence();
    }
    __syncthreads();
    if (!s_ok) return;
    // ---- offsets: the totals of the tasks before this one, per class, and the class totals
    double *peer_left = nullptr, *peer_right = nullptr;
    double seq = 0.0;
    if (X.state != nullptr) {
        const long long sq = *reinterpret_cast<volatile long long *>(&X.state->seq);
        seq = (double)sq;
        const int par = (int)(sq & 1LL);
        peer_left = X.left_mail + X.L.from_right(par);
        peer_right = X.right_mail + X.L.from_left(par);
    }
    if (open_cta) {
        // this rank's scalars (max|v|, wall-shear partials) into the mailbox of EVERY rank (thread r -> rank r), the
        // sequence number last (p2p_open_cta without the bump)
        const int r = (int)threadIdx.x;
        if (r < X.L.world) {
            double *slot = peers.mail[r] + X.L.scal((int)((long long)seq & 1LL), me);
            slot[0] = sc->vmax;
            slot[1] = sc->tau_bottom;
            slot[2] = sc->tau_top;
            __threadfence_system();
            *reinterpret_cast<volatile double *>(slot + 3) = seq;
        }
    }
    if (!open_cta && task < n_tasks) {
        // (one 16-byte record of four class totals per task: independent loads, then one warp reduction)
        int off[cN] = {0, 0, 0, 0, 0}, all[cN] = {0, 0, 0, 0, 0};
        {
            int4 before = make_int4(0, 0, 0, 0), total = before;
#pragma unroll 4
            for (int t = lane; t < n_tasks; t += 32) {
                const int4 v = __ldcg(reinterpret_cast<const int4 *>(task_tot) + t);
                total.x += v.x;
                total.y += v.y;
                total.z += v.z;
                total.w += v.w;
                if (t < task) {
                    before.x += v.x;
                    before.y += v.y;
                    before.z += v.z;
                    before.w += v.w;
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                before.x += __shfl_xor_sync(0xffffffffu, before.x, d);
                before.y += __shfl_xor_sync(0xffffffffu, before.y, d);
                before.z += __shfl_xor_sync(0xffffffffu, before.z, d);
                before.w += __shfl_xor_sync(0xffffffffu, before.w, d);
                total.x += __shfl_xor_sync(0xffffffffu, total.x, d);
                total.y += __shfl_xor_sync(0xffffffffu, total.y, d);
                total.z += __shfl_xor_sync(0xffffffffu, total.z, d);
                total.w += __shfl_xor_sync(0xffffffffu, total.w, d);
            }
            off[1] = before.x;
            off[2] = before.y;
            off[3] = before.z;
            off[4] = before.w;
            all[1] = total.x;
            all[2] = total.y;
            all[3] = total.z;
            all[4] = total.w;
        }
        if (task == 0 && lane == 0) {
            const int nl = all[cMigL] + all[cHaloL], nr = all[cMigR] + all[cHaloR];
            counts[cStay] = 0;
#pragma unroll
            for (int k = 1; k < cN; ++k) counts[k] = all[k];
            send_left[0] = (double)all[cMigL];
            send_left[1] = (double)all[cHaloL];
            send_right[0] = (double)all[cMigR];
            send_right[1] = (double)all[cHaloR];
            if (peer_left) {
                peer_left[0] = (double)all[cMigL];
                peer_left[1] = (double)all[cHaloL];
                peer_right[0] = (double)all[cMigR];
                peer_right[1] = (double)all[cHaloR];
            }
            if (nl > rec_cap || nr > rec_cap) atomicOr(flags, kErrSlabCap);
        }
        // ---- scatter
        if (have) {
            int pass = 0;
            for (int base = e0; base < e1; base += 32, ++pass) {
                int i = pass == 0 ? i0 : i1;
                const unsigned m = pass == 0 ? m0 : (pass == 1 ? m1 : (base + lane < e1 ? edge_class(E, A, G, base + lane, i) : 0u));
                double2 x = make_double2(0.0, 0.0), v = x;
                double rho = 0.0, drho = 0.0, mass = 0.0, Vol = 0.0;
                Mat2 B{};
                int pid = -1;
                if (m) {
                    x = A.pos[i];
                    v = A.vel[i];
                    rho = A.rho[i];
                    drho = A.drho[i];
                    mass = A.mass[i];
                    Vol = A.Vol[i];
                    B = A.B[i];
                    pid = A.pid[i];
                }
#pragma unroll
                for (int k = 1; k < cN; ++k) {
                    const unsigned b = __ballot_sync(0xffffffffu, (m >> k) & 1u);
                    if ((m >> k) & 1u) {
                        const bool left = k == cMigL || k == cHaloL;
                        const int first = (k == cHaloL) ? all[cMigL] : (k == cHaloR ? all[cMigR] : 0);
                        const int slot = first + off[k] + __popc(b & ((1u << lane) - 1u));
                        if (slot < rec_cap) {
                            write_record((left ? send_left : send_right) + kHdr + (size_t)slot * kRec, x, v, rho, drho, mass, Vol, B, pid);
                            if (peer_left)
                                write_record((left ? peer_left : peer_right) + kHdr + (size_t)slot * kRec, x, v, rho, drho, mass, Vol, B, pid);
                        }
                    }
                    off[k] += __popc(b);
                }
                if (m & ((1u << cMigL) | (1u << cMigR))) A.pid[i] = -1;   // migrants stay in place, marked dead
            }
        }
    }
    // ---- finish: the last CTA publishes the arrival and re-arms the two counters.  One fence per CTA, by the thread
    // that takes the ticket, behind the CTA barrier (fences are cumulative: it covers the records of every thread of
    // the CTA); a fence per thread was 28 % of this kernel's stall samples (r2 call 22)
    __syncthreads();
    if (threadIdx.x == 0) {
        if (peer_left) __threadfence_system();
        else __threadfence();
        if (atomicAdd(counts + 7, 1) == (int)gridDim.x - 1) {
            if (peer_left) {
                __threadfence_system();
                *reinterpret_cast<volatile double *>(peer_left + 2) = seq;
                *reinterpret_cast<volatile double *>(peer_right + 2) = seq;
            }
            counts[7] = 0;
            counts[15] = 0;
        }
    }
}

// Compaction in CELL order (the slab form of the reference's periodic cell sort, m:529-568): the rows that
// stay are visited through the cell-sorted entry array of the last list build, so the compacted rows come
// out sorted by (cell row, cell column) and the staging loads of the gather kernels stay coalesced.
__global__ void __launch_bounds__(kSlabBlock) k_slab_sort_flags(const int *__restrict__ ent_key,
                                                               const GridDesc *__restrict__ g, int ent_cap,
                                                               SlabArrays A, SlabGeom G, int *__restrict__ flag)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ent_cap) return;
    int f = 0;
    if (e < g->n_entries) {
        const int key = ent_key[e];
        const int i = key >> 3;
        if ((key & 3) == 0 && i < G.cur[1] && A.pid[i] >= 0) f = (slab_class(A.pos[i].x, G) >> cStay) & 1u;
    }
    flag[e] = f;
}
__global__ void __launch_bounds__(kSlabBlock) k_slab_sort_move(const int *__restrict__ ent_key,
                                                              const int *__restrict__ flag,
                                                              const int *__restrict__ ptr, int ent_cap, SlabArrays A)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ent_cap || !flag[e]) return;
    const int i = ent_key[e] >> 3, d = ptr[e];
    A.n_pos[d] = A.pos[i];
    A.n_vel[d] = A.vel[i];
    A.n_rho[d] = A.rho[i];
    A.n_drho[d] = A.drho[i];
    A.n_mass[d] = A.mass[i];
    A.n_Vol[d] = A.Vol[i];
    A.n_B[d] = A.B[i];
    A.n_pid[d] = A.pid[i];
}

struct UnpackSrc {
    const double *buf[4];   // received from left, received from right, sent to left, sent to right
};

// new owned rows = [stayers | migrants from the left | migrants from the right];
// new halo rows  = [copies from the left | copies from the right | own rows that left to the left | ... right]
// The incoming rows are APPENDED to the owned region [0, base): stayers do not move (base = the high-water
// mark of the owned rows, or the number of stayers right after a compaction).
//
// ONE launch for the whole receive side: the appended fluid rows (with their EOS pressure, m:172), the static wall
// rows behind them, the y range of both for the list build that follows (GridFuse: the step's continuity pass has
// reduced the rows it visited and zeroed the cell counters) -- and, by the last CTA to finish, what used to be the
// commit kernel: with peer-to-peer mailboxes the scalars of every rank (max|v| -> the next time step), the new row
// counts {nf, nfo, nt} every kernel of the next step reads, the grid of the coming build.
struct WallStore {
    int n_wall;
    const double2 *w_pos;
    const double *w_mass;
    const double2 *w_wall_vel;
    double rho0, p0;
    double2 *wall_vel;   // per-row wall velocity and pressure of the context (not part of a record)
    double *p;
};
__global__ void __launch_bounds__(kSlabBlock) k_slab_unpack(UnpackSrc U, SlabArrays A, SlabGeom G, int rec_cap,
                                                           int nf_cap, int base_is_stay_count, int *counts,
                                                           int *flags, P2PState *p2p, const double *mail, MailLayout ML,
                                                           WallStore W, DevScalars *sc, GridFuse F)
{
    __shared__ int ok;
    if (p2p != nullptr) {
        // peer-to-peer: the received records are in this rank's own mailbox; wait for both neighbours' arrival slots
        const int par = (int)(p2p->seq & 1LL);
        U.buf[0] = mail + ML.from_left(par);
        U.buf[1] = mail + ML.from_right(par);
        if (threadIdx.x == 0) {
            const double seq = (double)p2p->seq;
            ok = mail_wait(U.buf[0] + 2, seq) && mail_wait(U.buf[1] + 2, seq) ? 1 : 0;
            if (!ok) atomicOr(flags, kErrSlabCap);   // a neighbour never arrived: reported, the step is poisoned
        }
        __syncthreads();
        if (!ok) return;
    }
    // segment sizes from the headers
    const int n_stay = base_is_stay_count ? counts[cStay] : counts[13];
    int seg_n[6], seg_src[6], seg_first[6];
    const int nmL = min((int)__ldcg(U.buf[0]), rec_cap), nhL = min((int)__ldcg(U.buf[0] + 1), rec_cap - nmL);
    const int nmR = min((int)__ldcg(U.buf[1]), rec_cap), nhR = min((int)__ldcg(U.buf[1] + 1), rec_cap - nmR);
    const int smL = min((int)U.buf[2][0], rec_cap), smR = min((int)U.buf[3][0], rec_cap);
    seg_n[0] = nmL; seg_src[0] = 0; seg_first[0] = 0;      // migrants from the left   -> owned
    seg_n[1] = nmR; seg_src[1] = 1; seg_first[1] = 0;      // migrants from the right  -> owned
    seg_n[2] = nhL; seg_src[2] = 0; seg_first[2] = nmL;    // copies from the left     -> halo
    seg_n[3] = nhR; seg_src[3] = 1; seg_first[3] = nmR;    // copies from the right    -> halo
    seg_n[4] = smL; seg_src[4] = 2; seg_first[4] = 0;      // rows that left leftwards -> halo
    seg_n[5] = smR; seg_src[5] = 3; seg_first[5] = 0;      // rows that left rightwards-> halo
    int total = 0;
    for (int k = 0; k < 6; ++k) total += seg_n[k];
    const int n_owned = n_stay + nmL + nmR;
    const int n_fluid = n_stay + total;
    const int nf = min(n_fluid, nf_cap);
    const int stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    long long kmin = LLONG_MAX, kmax = LLONG_MIN;   // GridFuse: y range of the appended rows and the walls
    const double x_lo = (double)G.col_lo * G.cell, x_hi = (double)(G.col_lo + G.width) * G.cell;
    const double margin = (double)(G.halo + 2) * G.cell;
    for (int t = tid; t < total; t += stride) {
        int k = 0, off = t;
        while (off >= seg_n[k]) {
            off -= seg_n[k];
            ++k;
        }
        const int d = n_stay + t;
        if (d >= nf_cap) continue;
        const double4 *r = reinterpret_cast<const double4 *>(U.buf[seg_src[k]] + kHdr + (size_t)(seg_first[k] + off) * kRec);
        const double4 a = ldcg4(r), b = ldcg4(r + 1), c = ldcg4(r + 2);   // (L2: a peer may have written them)
        A.pid[d] = (int)ldcg4(r + 3).x;
        double x = a.x;
        if (k >= 2) {   // a halo copy sits next to the slab: across the periodic seam it is shifted by -+DL
            if (x > x_hi + margin) x -= G.DL;
            else if (x < x_lo - margin) x += G.DL;
        }
        A.pos[d] = make_double2(x, a.y);
        A.vel[d] = make_double2(a.z, a.w);
        A.rho[d] = b.x;
        A.drho[d] = b.y;
        A.mass[d] = b.z;
        A.Vol[d] = b.w;
        Mat2 B;
        B.b11 = c.x;
        B.b12 = c.y;
        B.b21 = c.z;
        B.b22 = c.w;
        A.B[d] = B;
        if (!base_is_stay_count) {   // (after a compaction the loop below visits every fluid row)
            W.p[d] = W.p0 * (b.x / W.rho0 - 1.0);   // EOS of the appended fluid rows (m:172); fluid rows carry no wall velocity
            W.wall_vel[d] = make_double2(0.0, 0.0);
        }
        if (F.g) {
            const long long key = ordered_key(a.y);
            kmin = key < kmin ? key : kmin;
            kmax = key > kmax ? key : kmax;
        }
    }
    if (base_is_stay_count) {
        // a compaction moved the stayers: pressure and wall velocity are per-row arrays that it does not permute.  The
        // appended rows are written by other threads of this launch, so their density comes from the record again.
        for (int t = tid; t < nf; t += stride) {
            double rho_t;
            if (t < n_stay) {
                rho_t = A.rho[t];
            } else {
                int k = 0, off = t - n_stay;
                while (off >= seg_n[k]) {
                    off -= seg_n[k];
                    ++k;
                }
                rho_t = ldcg4(reinterpret_cast<const double4 *>(U.buf[seg_src[k]] + kHdr + (size_t)(seg_first[k] + off) * kRec) + 1).x;
            }
            W.p[t] = W.p0 * (rho_t / W.rho0 - 1.0);
            W.wall_vel[t] = make_double2(0.0, 0.0);
        }
    }
    // wall rows (static) after the fluid rows
    for (int w = tid; w < W.n_wall; w += stride) {
        const int t = nf + w;
        const double2 xw = W.w_pos[w];
        if (F.g) {
            const long long key = ordered_key(xw.y);
            kmin = key < kmin ? key : kmin;
            kmax = key > kmax ? key : kmax;
        }
        A.pos[t] = xw;
        A.vel[t] = make_double2(0.0, 0.0);
        A.rho[t] = W.rho0;
        A.drho[t] = 0.0;
        A.mass[t] = W.w_mass[w];
        A.Vol[t] = W.w_mass[w] / W.rho0;
        Mat2 B;
        B.b11 = 1.0;
        B.b12 = 0.0;
        B.b21 = 0.0;
        B.b22 = 1.0;
        A.B[t] = B;
        W.wall_vel[t] = W.w_wall_vel[w];
        W.p[t] = 0.0;
    }
    if (F.g) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const long long a = __shfl_xor_sync(0xffffffffu, kmin, d);
            const long long b = __shfl_xor_sync(0xffffffffu, kmax, d);
            kmin = a < kmin ? a : kmin;
            kmax = b > kmax ? b : kmax;
        }
        if ((threadIdx.x & 31) == 0) {
            if (kmin != LLONG_MAX) atomicMin(&F.g->ymin_key, kmin);
            if (kmax != LLONG_MIN) atomicMax(&F.g->ymax_key, kmax);
        }
    }
    // ---- commit, by the last CTA to finish (counts[11] = the ticket) ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) ok = atomicAdd(reinterpret_cast<unsigned int *>(counts + 11), 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!ok) return;
    __threadfence();
    if (p2p != nullptr) p2p_scalars_cta(sc, p2p, mail, ML, flags);
    if (threadIdx.x != 0) return;
    counts[8] = n_owned;
    counts[9] = n_fluid - n_owned;
    counts[10] = n_stay;   // first appended row
    if (n_fluid > nf_cap) atomicOr(flags, kErrSlabCap);
    // the new row counts become current: {nf, nfo, nt} read by every kernel of the next step
    counts[12] = nf;
    counts[13] = min(n_owned, nf_cap);
    counts[14] = nf + W.n_wall;
    counts[11] = 0;
    if (F.g != nullptr) grid_derive(F.g, F.cell, F.ncx, F.cell_cap, F.flags, F.totals);
}
__global__ void k_slab_set_counts(int *counts, int nf, int nfo, int nt)
{
    counts[12] = nf;
    counts[13] = nfo;
    counts[14] = nt;
}
__global__ void k_slab_iota(int *p, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}
__global__ void k_slab_scalars_out(const DevScalars *sc, double *out)
{
    out[0] = sc->vmax;
    out[1] = sc->tau_bottom;
    out[2] = sc->tau_top;
}
__global__ void k_slab_vmax_in(DevScalars *sc, const double *in) { sc->vmax = in[0]; }

// the scalars of every rank (triple r at g + off[r], or at g + 3 r when off is NULL) -> max|v| (MAX) and the
// wall-shear sums (SUM, in rank order: deterministic)
__global__ void k_slab_scalars_reduce(DevScalars *sc, const double *g, const long long *off, int world, double *out3)
{
    double vmax = 0.0, tb = 0.0, tt = 0.0;
    for (int r = 0; r < world; ++r) {
        const double *t = g + (off ? off[r] : 3LL * r);
        vmax = fmax(vmax, t[0]);
        tb += t[1];
        tt += t[2];
    }
    sc->vmax = vmax;
    if (out3) {
        out3[0] = vmax;
        out3[1] = tb;
        out3[2] = tt;
    }
}

SlabGeom geom_of(const sphb_ctx *c)
{
    SlabGeom G;
    G.col_lo = c->slab.col_lo;
    G.width = c->slab.col_hi - c->slab.col_lo;
    G.halo = c->slab.halo;
    G.ncx_g = c->slab.ncx_global;
    G.cell = 2.0 * c->prm.h;
    G.inv_cell = 1.0 / G.cell;
    G.DL = c->prm.DL;
    G.cur = c->slab.counts.as<int>() + 12;
    return G;
}

SlabArrays arrays_of(sphb_ctx *c)
{
    SlabArrays A;
    A.pos = c->pos.as<double2>();
    A.vel = c->vel.as<double2>();
    A.rho = c->rho.as<double>();
    A.drho = c->drho_dt.as<double>();
    A.mass = c->mass.as<double>();
    A.Vol = c->Vol.as<double>();
    A.B = c->B.as<Mat2>();
    A.n_pos = c->a_pos.as<double2>();
    A.n_vel = c->a_vel.as<double2>();
    A.n_rho = c->a_rho.as<double>();
    A.n_drho = c->a_drho_dt.as<double>();
    A.n_mass = c->a_mass.as<double>();
    A.n_Vol = c->a_Vol.as<double>();
    A.n_B = c->a_B.as<Mat2>();
    A.pid = c->slab.pid.as<int>();
    A.n_pid = c->slab.a_pid.as<int>();
    return A;
}

// CTAs of k_slab_edge_send that are certainly resident together (its grid barrier needs that): one per SM
unsigned edge_resident_ctas()
{
#ifdef SPHB_CPUSIM
    return 0u;   // (the functional simulator of tests/cpusim runs the CTAs of a launch one after the other: three launches)
#endif
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0u;
    }
    return sms > 0 ? (unsigned)sms : 0u;
}

// The candidate bands of the edge pack (k_slab_edge_*) in the local columns of the last list build, or false when the
// full pack has to run: no list for the current rows, or a slab too narrow for two separate bands.
bool edge_view_of(const sphb_ctx *c, EdgeView &E, int &n_tasks)
{
    const SlabState &s = c->slab;
    const NeighborWork &w = c->nw;
    const int width = s.col_hi - s.col_lo, band = s.halo + kEdgeBand;
    if (!c->have_list || !w.slab || w.ncx_local < 1 || w.cell_cap < 1 || width < 2 * band + 2) return false;
    const int first = s.col_lo - w.col0;   // local column of the first owned column
    E.ent_key = w.ent_key.as<int>();
    E.cell_start = w.cell_start.as<int>();
    E.g = w.grid.as<GridDesc>();
    E.ncx = w.ncx_local;
    E.cell_cap = w.cell_cap;
    E.a[0] = std::max(first - 1, 0);
    E.b[0] = std::min(first + band, w.ncx_local);
    E.a[1] = std::max(first + width - band, 0);
    E.b[1] = std::min(first + width + 1, w.ncx_local);
    n_tasks = 2 * (w.cell_cap / w.ncx_local);
    return n_tasks > 0 && (size_t)cN * (size_t)n_tasks * sizeof(int) <= s.block_tot.bytes;
}

// Host-side values used for launch sizes and by-value kernel parameters.  In steady state they stay at the
// CAPACITY (the rows of a slab change every step and the host never waits for the new counts: the kernels
// read {nf, nfo, nt} from device memory, PhysParams::dev_counts).
void apply_counts(sphb_ctx *c, int n_owned, int n_halo)
{
    const int nf = n_owned + n_halo, nt = nf + c->slab.n_wall;
    c->prm.n_fluid = nf;
    c->prm.n_total = nt;
    c->P.nf = nf;
    c->P.nfo = n_owned;
    c->P.nt = nt;
    c->nw.n_fluid = nf;
    c->nw.n_total = nt;
}

// the current row counts, read back from the device (the one place the host waits for them)
int refresh_counts(sphb_ctx *c)
{
    SlabState &s = c->slab;
    SPHB_CUDA(cudaMemcpyAsync(s.h_counts, s.counts.p, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    s.n_owned = s.h_counts[13];
    s.n_halo = s.h_counts[12] - s.h_counts[13];
    return 0;
}

}  // namespace

extern "C" {

#define SLAB_ENTER()                                                                     \
    clear_error();                                                                       \
    if (!c) return set_error("SPHB:arg", "null context");                                \
    SPHB_TRY(ctx_bind_device(c));                                                        \
    if (!c->slab.on) return set_error("SPHB:state", "the context is not a slab: call sphb_ctx_slab_init")

// The context must have been created with n_fluid = fluid-row CAPACITY (owned + halo copies, with
// head-room) and n_total = that capacity + n_wall.  Host arrays are column-major like sphb_ctx_upload's.
SPHB_API int sphb_ctx_slab_init(sphb_ctx *c, int col_lo, int col_hi, int halo_cols, int n_owned, const double *pos,
                                const double *vel, const double *rho, const double *drho_dt, const double *mass,
                                const int *ids, int n_wall, const double *wall_pos, const double *wall_mass,
                                const double *wall_wall_vel)
{
    clear_error();
    if (!c) return set_error("SPHB:arg", "null context");
    SPHB_TRY(ctx_bind_device(c));
    const bool again = c->slab.on;   // re-initialisation keeps the capacity of the first one
    const int nf_cap = again ? c->slab.nf_cap : c->prm.n_fluid;
    const int ncx_g = std::max(1, (int)std::ceil(c->prm.DL / (2.0 * c->prm.h)));   // nb:254
    if (col_lo < 0 || col_hi <= col_lo || col_hi > ncx_g || halo_cols < 1 || n_owned < 0 || n_owned > nf_cap ||
        n_wall < 0 || (again ? n_wall != c->slab.n_wall : nf_cap + n_wall != c->prm.n_total))
        return set_error("SPHB:arg", "bad slab description");
    if (col_hi - col_lo < halo_cols || ncx_g < (col_hi - col_lo) + 2 * halo_cols + 2)
        return set_error("SPHB:arg", "slab of %d columns with %d halo columns does not fit a channel of %d columns",
                         col_hi - col_lo, halo_cols, ncx_g);
    SlabState &s = c->slab;
    s.on = true;
    s.col_lo = col_lo;
    s.col_hi = col_hi;
    s.halo = halo_cols;
    s.ncx_global = ncx_g;
    s.n_wall = n_wall;
    s.nf_cap = nf_cap;
    ctx_drop_graphs(c);
    // static wall store
    const size_t nw = (size_t)n_wall;
    SPHB_TRY(s.w_pos.reserve((nw + 1) * sizeof(double2)));
    SPHB_TRY(s.w_mass.reserve((nw + 1) * sizeof(double)));
    SPHB_TRY(s.w_wall_vel.reserve((nw + 1) * sizeof(double2)));
    SPHB_TRY(s.counts.reserve(16 * sizeof(int)));
    SPHB_CUDA(cudaMemsetAsync(s.counts.p, 0, 16 * sizeof(int), c->st));   // ([11]: the ticket of the unpack's last CTA)
    SPHB_TRY(s.block_tot.reserve((size_t)cN * (grid_for((size_t)nf_cap, kSlabBlock) + 1) * sizeof(int)));
    if (!s.h_counts) SPHB_CUDA(cudaMallocHost((void **)&s.h_counts, 16 * sizeof(int)));
    SPHB_TRY(s.pid.reserve(((size_t)nf_cap + 1) * sizeof(int)));
    SPHB_TRY(s.a_pid.reserve(((size_t)nf_cap + 1) * sizeof(int)));
    s.squeezed = false;
    if (ids && n_owned)
        SPHB_CUDA(cudaMemcpyAsync(s.pid.p, ids, (size_t)n_owned * sizeof(int), cudaMemcpyHostToDevice, c->st));
    else if (n_owned)
        k_slab_iota<<<grid_for((size_t)n_owned), kBlock, 0, c->st>>>(s.pid.as<int>(), n_owned);
    double *stage = c->stage.as<double>();   // 4 * n_total doubles
    size_t used = 0;   // every array has its own staging region (6 * n_total doubles in all)
    auto up2 = [&](const double *src, size_t n, double2 *dst) -> int {
        if (!n) return 0;
        double *st2 = stage + used;
        used += 2 * n;
        SPHB_CUDA(cudaMemcpyAsync(st2, src, 2 * n * sizeof(double), cudaMemcpyHostToDevice, c->st));
        launch_pack2(st2, dst, (int)n, c->st);
        return 0;
    };
    SPHB_TRY(up2(wall_pos, nw, s.w_pos.as<double2>()));
    SPHB_TRY(up2(wall_wall_vel, nw, s.w_wall_vel.as<double2>()));
    if (nw) SPHB_CUDA(cudaMemcpyAsync(s.w_mass.p, wall_mass, nw * sizeof(double), cudaMemcpyHostToDevice, c->st));
    // owned fluid rows; no halo copies yet (the first exchange brings them)
    const size_t no = (size_t)n_owned;
    SPHB_TRY(up2(pos, no, c->pos.as<double2>()));
    SPHB_TRY(up2(vel, no, c->vel.as<double2>()));
    if (no) {
        SPHB_CUDA(cudaMemcpyAsync(c->rho.p, rho, no * sizeof(double), cudaMemcpyHostToDevice, c->st));
        SPHB_CUDA(cudaMemcpyAsync(c->drho_dt.p, drho_dt, no * sizeof(double), cudaMemcpyHostToDevice, c->st));
        SPHB_CUDA(cudaMemcpyAsync(c->mass.p, mass, no * sizeof(double), cudaMemcpyHostToDevice, c->st));
    }
    apply_counts(c, n_owned, 0);
    s.n_owned = n_owned;
    s.n_halo = 0;
    k_slab_set_counts<<<1, 1, 0, c->st>>>(s.counts.as<int>(), n_owned, n_owned, n_owned + n_wall);
    c->P.dev_counts = nullptr;   // the set-up kernels below run with the exact host values
    // Vol = m / rho (m:120), B = I already
    SPHB_TRY(ctx_volumes_from_mass(c));
    SPHB_TRY(csr_vmax(c->P, c->vel.as<double2>(), c->sc(), c->st));
    // the cell grid of the slab: global columns [col_lo - H - 1, col_hi + H + 2): empty margins on both sides
    c->nw.slab = true;
    c->nw.col0 = col_lo - halo_cols - 1;
    c->nw.ncx_local = (col_hi - col_lo) + 2 * halo_cols + 3;
    double y_lo = 0.0, y_hi = c->prm.DH;
    for (size_t i = 0; i < nw; ++i) {
        y_lo = wall_pos[nw + i] < y_lo ? wall_pos[nw + i] : y_lo;
        y_hi = wall_pos[nw + i] > y_hi ? wall_pos[nw + i] : y_hi;
    }
    for (size_t i = 0; i < no; ++i) {
        y_lo = pos[no + i] < y_lo ? pos[no + i] : y_lo;
        y_hi = pos[no + i] > y_hi ? pos[no + i] : y_hi;
    }
    // plan for the capacity (buffers) and for the actual row density (tile width), then run with the
    // current row counts
    {
        const double rows = std::ceil((y_hi - y_lo + 1e-12) / (2.0 * c->prm.h)) + 1.0;
        const double width = (double)(col_hi - col_lo);
        c->nw.per_cell_hint = ((double)n_owned / width + (double)n_wall / (width + 2.0 * halo_cols)) / rows;
    }
    if (!again || y_lo < c->y_lo || y_hi > c->y_hi) {   // a re-initialisation keeps the planned buffers
        apply_counts(c, nf_cap, 0);
        SPHB_TRY(ctx_plan_lists(c, y_lo, y_hi));
    }
    // (the per-task totals of the edge pack: two tasks per cell row of the planned grid)
    SPHB_TRY(s.block_tot.reserve((size_t)cN * std::max((size_t)grid_for((size_t)nf_cap, kSlabBlock) + 1,
                                                        2 * (size_t)(c->nw.cell_cap / std::max(1, c->nw.ncx_local)) + 2) * sizeof(int)));
    // from here on: launch sizes = capacity, actual counts on the device
    apply_counts(c, nf_cap, 0);
    c->P.nfo = nf_cap;
    c->P.dev_counts = s.counts.as<int>() + 12;
    c->nw.dev_counts = c->P.dev_counts;
    c->have_list = false;
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

SPHB_API int sphb_ctx_slab_pack(sphb_ctx *c, double *send_left_dev, double *send_right_dev, int rec_cap)
{
    SLAB_ENTER();
    if (!send_left_dev || !send_right_dev || rec_cap < 1) return set_error("SPHB:arg", "bad send buffers");
    SlabState &s = c->slab;
    const SlabGeom G = geom_of(c);
    const unsigned blocks = std::max(1u, grid_for((size_t)s.nf_cap, kSlabBlock));
    int *bt = s.block_tot.as<int>(), *cnt = s.counts.as<int>();
    ProfScope ps(&c->kprof, kPfSlab, c->st);
    EdgeView E;
    int n_tasks = 0;
    if (edge_view_of(c, E, n_tasks)) {   // steady state: only the rows near the edges (see k_slab_edge_count)
        const unsigned eb = (unsigned)((n_tasks + kEdgeWarps - 1) / kEdgeWarps);
        if (eb + 1 <= edge_resident_ctas()) {   // one launch (k_slab_edge_send)
            k_slab_edge_send<<<eb, kEdgeWarps * 32, 0, c->st>>>(E, arrays_of(c), G, n_tasks, bt, cnt, send_left_dev,
                                                               send_right_dev, rec_cap, c->nw.flags.as<int>(), P2PSend{},
                                                               c->sc(), PeerList{}, 0);
            c->launches -= 2;
        } else {
            k_slab_edge_count<<<eb, kEdgeWarps * 32, 0, c->st>>>(E, arrays_of(c), G, n_tasks, bt);
            k_slab_offsets<<<cN, 1024, 0, c->st>>>(bt, n_tasks, cnt, nullptr, nullptr, PeerList{}, MailLayout{}, 0);
            k_slab_edge_scatter<<<eb, kEdgeWarps * 32, 0, c->st>>>(E, arrays_of(c), G, n_tasks, bt, cnt, send_left_dev,
                                                                  send_right_dev, rec_cap, c->nw.flags.as<int>(), P2PSend{});
        }
    } else {
        k_slab_count<<<blocks, kSlabBlock, 0, c->st>>>(c->pos.as<double2>(), s.pid.as<int>(), G, bt);
        k_slab_offsets<<<cN, 1024, 0, c->st>>>(bt, (int)blocks, cnt, nullptr, nullptr, PeerList{}, MailLayout{}, 0);
        k_slab_scatter<true><<<blocks, kSlabBlock, 0, c->st>>>(arrays_of(c), G, bt, cnt, send_left_dev, send_right_dev,
                                                               rec_cap, c->nw.flags.as<int>(), P2PSend{});
    }
    c->launches += 3;
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

// Squeeze the dead rows out of the owned region (now and then; between sphb_ctx_slab_pack and
// sphb_ctx_slab_unpack): the rows that stay are compacted -- in cell order, m:529-568 -- into the alternate
// arrays, which become the current ones.
SPHB_API int sphb_ctx_slab_compact(sphb_ctx *c)
{
    SLAB_ENTER();
    SlabState &s = c->slab;
    const SlabGeom G = geom_of(c);
    if (c->have_list) {
        // in cell order, through the entry array of the last build (every live owned row has a real entry)
        const int ent_cap = (int)c->nw.ell_stride;
        SPHB_TRY(s.sort_flag.reserve(((size_t)ent_cap + 1) * sizeof(int)));
        SPHB_TRY(s.sort_ptr.reserve(((size_t)ent_cap + 1) * sizeof(int)));
        ProfScope ps(&c->kprof, kPfSlab, c->st);
        const SlabArrays A = arrays_of(c);
        k_slab_sort_flags<<<grid_for((size_t)ent_cap, kSlabBlock), kSlabBlock, 0, c->st>>>(
            c->nw.ent_key.as<int>(), c->nw.grid.as<GridDesc>(), ent_cap, A, G, s.sort_flag.as<int>());
        // (the total = the rows that stay: what the next unpack appends after -- the edge pack does not count them)
        SPHB_TRY(exclusive_scan_i32(s.sort_flag.as<int>(), s.sort_ptr.as<int>(), (size_t)ent_cap, c->nw.scan_tmp,
                                    s.counts.as<int>() + cStay, c->st, &c->launches));
        k_slab_sort_move<<<grid_for((size_t)ent_cap, kSlabBlock), kSlabBlock, 0, c->st>>>(
            c->nw.ent_key.as<int>(), s.sort_flag.as<int>(), s.sort_ptr.as<int>(), ent_cap, A);
    } else {
        const unsigned blocks = std::max(1u, grid_for((size_t)s.nf_cap, kSlabBlock));
        ProfScope ps(&c->kprof, kPfSlab, c->st);
        k_slab_scatter<false><<<blocks, kSlabBlock, 0, c->st>>>(arrays_of(c), G, s.block_tot.as<int>(),
                                                                s.counts.as<int>(), nullptr, nullptr, 0,
                                                                c->nw.flags.as<int>(), P2PSend{});
    }
    c->launches += 1;
    SPHB_CUDA(cudaGetLastError());
    std::swap(c->pos, c->a_pos);
    std::swap(c->vel, c->a_vel);
    std::swap(c->rho, c->a_rho);
    std::swap(c->drho_dt, c->a_drho_dt);
    std::swap(c->mass, c->a_mass);
    std::swap(c->Vol, c->a_Vol);
    std::swap(c->B, c->a_B);
    std::swap(s.pid, s.a_pid);
    s.squeezed = true;   // the next unpack appends after the stayers counted by the pack
    return 0;
}

SPHB_API int sphb_ctx_slab_unpack(sphb_ctx *c, const double *recv_left_dev, const double *recv_right_dev,
                                  const double *sent_left_dev, const double *sent_right_dev, int rec_cap)
{
    SLAB_ENTER();
    SlabState &s = c->slab;
    const SlabGeom G = geom_of(c);
    UnpackSrc U;
    U.buf[0] = recv_left_dev;
    U.buf[1] = recv_right_dev;
    U.buf[2] = sent_left_dev;
    U.buf[3] = sent_right_dev;
    int *cnt = s.counts.as<int>();
    const SlabArrays A = arrays_of(c);
    {
        ProfScope ps(&c->kprof, kPfSlab, c->st);
        MailLayout ML;
        ML.world = s.p2p_world;
        ML.rec_cap = s.p2p_rec_cap;
        const bool p2p = recv_left_dev == nullptr && recv_right_dev == nullptr;   // sphb_ctx_slab_p2p_recv
        // the grid set-up of the list build that follows: started by the step's continuity pass (y range of the rows it
        // visited, zeroed cell counters), continued by this launch with the appended rows and the walls, derived by
        // its last CTA
        const GridFuse F = c->nw.grid_ready ? neighbor_grid_fuse(c->nw) : GridFuse();
        WallStore W;
        W.n_wall = s.n_wall;
        W.w_pos = s.w_pos.as<double2>();
        W.w_mass = s.w_mass.as<double>();
        W.w_wall_vel = s.w_wall_vel.as<double2>();
        W.rho0 = c->prm.rho0;
        W.p0 = c->prm.p0;
        W.wall_vel = c->wall_vel.as<double2>();
        W.p = c->p.as<double>();
        // appended rows start at the (possibly squeezed) end of the owned region; the walls follow the halo
        k_slab_unpack<<<148, kSlabBlock, 0, c->st>>>(U, A, G, rec_cap, s.nf_cap, s.squeezed ? 1 : 0, cnt,
                                                     c->nw.flags.as<int>(), p2p ? s.p2p_state.as<P2PState>() : nullptr,
                                                     s.mail, ML, W, c->sc(), F);
    }
    c->launches += 1;
    SPHB_CUDA(cudaGetLastError());
    s.squeezed = false;
    c->have_list = false;
    return 0;   // asynchronous: a capacity overflow raises the device flag (sphb_status.error_flags)
}

// ---- peer-to-peer exchange (mailboxes) ------------------------------------------------------------------------------
// Replaces pack -> NCCL all-to-all -> scalar reduction -> unpack by kernels of this library only: the pack kernel stores
// the records straight into the neighbours' mailboxes over NVLink, the unpack kernel waits for its own mailbox to be
// filled.  No collective, no host synchronisation: the whole sharded step is one stream of kernels (one CUDA graph).
SPHB_API int sphb_ctx_slab_p2p_create(sphb_ctx *c, int rank, int world, int rec_cap, void **mailbox_dev, size_t *bytes,
                                      unsigned char ipc_handle[64])
{
    SLAB_ENTER();
    if (rank < 0 || world < 1 || rank >= world || world > SlabState::kMaxWorld || rec_cap < 1)
        return set_error("SPHB:arg", "bad peer-to-peer description (at most %d ranks)", SlabState::kMaxWorld);
    SlabState &s = c->slab;
    if (s.mail) return set_error("SPHB:state", "the slab already has a mailbox");
    MailLayout L;
    L.world = world;
    L.rec_cap = rec_cap;
    s.mail_doubles = L.total();
    SPHB_CUDA(cudaMalloc((void **)&s.mail, s.mail_doubles * sizeof(double)));
    SPHB_CUDA(cudaMemsetAsync(s.mail, 0, s.mail_doubles * sizeof(double), c->st));
    SPHB_TRY(s.p2p_state.reserve(sizeof(P2PState)));
    SPHB_CUDA(cudaMemsetAsync(s.p2p_state.p, 0, sizeof(P2PState), c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    s.p2p_rank = rank;
    s.p2p_world = world;
    s.p2p_rec_cap = rec_cap;
    for (int r = 0; r < SlabState::kMaxWorld; ++r) s.peer[r] = nullptr;
    s.peer[rank] = s.mail;
    if (ipc_handle) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
        cudaIpcMemHandle_t h;
        if (cudaIpcGetMemHandle(&h, s.mail) == cudaSuccess) {
            std::memcpy(ipc_handle, &h, 64);
        } else {   // (no IPC on this platform: mailboxes of one process still work through plain pointers)
            (void)cudaGetLastError();
            std::memset(ipc_handle, 0, 64);
        }
    }
    if (mailbox_dev) *mailbox_dev = s.mail;
    if (bytes) *bytes = s.mail_doubles * sizeof(double);
    return 0;
}

// The mailbox of rank `rank`: either its CUDA IPC handle (another process) or its device pointer (same process; peer
// access between the two devices is enabled here when they differ).
SPHB_API int sphb_ctx_slab_p2p_connect(sphb_ctx *c, int rank, const unsigned char ipc_handle[64], void *same_process_ptr,
                                       int peer_device)
{
    SLAB_ENTER();
    SlabState &s = c->slab;
    if (!s.mail) return set_error("SPHB:state", "create the own mailbox first (sphb_ctx_slab_p2p_create)");
    if (rank < 0 || rank >= s.p2p_world) return set_error("SPHB:arg", "bad rank");
    if (rank == s.p2p_rank) return 0;
    if (same_process_ptr) {
        if (peer_device >= 0 && peer_device != c->device) {
            int can = 0;
            SPHB_CUDA(cudaDeviceCanAccessPeer(&can, c->device, peer_device));
            if (!can) return set_error("SPHB:cuda", "device %d cannot access device %d", c->device, peer_device);
            const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return set_error("SPHB:cuda", "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
            (void)cudaGetLastError();
        }
        s.peer[rank] = static_cast<double *>(same_process_ptr);
        return 0;
    }
    if (!ipc_handle) return set_error("SPHB:arg", "neither a handle nor a pointer");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, 64);
    void *p = nullptr;
    SPHB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s.peer[rank] = static_cast<double *>(p);
    s.peer_opened[rank] = true;
    return 0;
}

// First half of the exchange: sequence number + scalars to every rank, then the pack with the records stored into the
// neighbours' mailboxes.  The local send buffers are still filled (the unpack reads the rows that left from them).
SPHB_API int sphb_ctx_slab_p2p_send(sphb_ctx *c, double *send_left_dev, double *send_right_dev)
{
    SLAB_ENTER();
    SlabState &s = c->slab;
    if (!s.mail) return set_error("SPHB:state", "no mailbox");
    if (!send_left_dev || !send_right_dev) return set_error("SPHB:arg", "bad send buffers");
    const int left = (s.p2p_rank + s.p2p_world - 1) % s.p2p_world, right = (s.p2p_rank + 1) % s.p2p_world;
    PeerList PL;
    for (int r = 0; r < SlabState::kMaxWorld; ++r) PL.mail[r] = r < s.p2p_world ? s.peer[r] : nullptr;
    for (int r = 0; r < s.p2p_world; ++r)
        if (!s.peer[r]) return set_error("SPHB:state", "the mailbox of rank %d is not connected", r);
    MailLayout L;
    L.world = s.p2p_world;
    L.rec_cap = s.p2p_rec_cap;
    P2PSend X;
    X.state = s.p2p_state.as<P2PState>();
    X.left_mail = s.peer[left];
    X.right_mail = s.peer[right];
    X.L = L;
    const SlabGeom G = geom_of(c);
    const unsigned blocks = std::max(1u, grid_for((size_t)s.nf_cap, kSlabBlock));
    int *bt = s.block_tot.as<int>(), *cnt = s.counts.as<int>();
    ProfScope ps(&c->kprof, kPfSlab, c->st);
    // (k_slab_offsets has one more CTA than classes: it bumps the sequence number and sends the scalars to every rank)
    EdgeView E;
    int n_tasks = 0;
    if (edge_view_of(c, E, n_tasks)) {   // steady state: only the rows near the edges (see k_slab_edge_count)
        const unsigned eb = (unsigned)((n_tasks + kEdgeWarps - 1) / kEdgeWarps);
        if (eb + 1 <= edge_resident_ctas()) {   // one launch (k_slab_edge_send; its last CTA opens the exchange)
            k_slab_edge_send<<<eb + 1, kEdgeWarps * 32, 0, c->st>>>(E, arrays_of(c), G, n_tasks, bt, cnt, send_left_dev,
                                                                   send_right_dev, s.p2p_rec_cap, c->nw.flags.as<int>(), X,
                                                                   c->sc(), PL, s.p2p_rank);
            c->launches -= 2;
        } else {
            k_slab_edge_count<<<eb, kEdgeWarps * 32, 0, c->st>>>(E, arrays_of(c), G, n_tasks, bt);
            k_slab_offsets<<<cN + 1, 1024, 0, c->st>>>(bt, n_tasks, cnt, c->sc(), X.state, PL, L, s.p2p_rank);
            k_slab_edge_scatter<<<eb, kEdgeWarps * 32, 0, c->st>>>(E, arrays_of(c), G, n_tasks, bt, cnt, send_left_dev,
                                                                  send_right_dev, s.p2p_rec_cap, c->nw.flags.as<int>(), X);
        }
    } else {
        k_slab_count<<<blocks, kSlabBlock, 0, c->st>>>(c->pos.as<double2>(), s.pid.as<int>(), G, bt);
        k_slab_offsets<<<cN + 1, 1024, 0, c->st>>>(bt, (int)blocks, cnt, c->sc(), X.state, PL, L, s.p2p_rank);
        k_slab_scatter<true><<<blocks, kSlabBlock, 0, c->st>>>(arrays_of(c), G, bt, cnt, send_left_dev, send_right_dev,
                                                               s.p2p_rec_cap, c->nw.flags.as<int>(), X);
    }
    c->launches += 3;
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

// Second half: wait for every rank's scalars (max|v| -> the next time step; wall-shear sums), then for the two
// neighbours' records, and append them (sphb_ctx_slab_unpack's work).  sphb_ctx_slab_compact may run in between.
SPHB_API int sphb_ctx_slab_p2p_recv(sphb_ctx *c, const double *sent_left_dev, const double *sent_right_dev)
{
    SLAB_ENTER();
    SlabState &s = c->slab;
    if (!s.mail) return set_error("SPHB:state", "no mailbox");
    // (the scalars of every rank are awaited and reduced by the commit kernel of the unpack)
    return sphb_ctx_slab_unpack(c, nullptr, nullptr, sent_left_dev, sent_right_dev, s.p2p_rec_cap);
}

// max|v| and the wall-shear sums of the WHOLE channel as of the last peer-to-peer exchange (3 doubles, host)
SPHB_API int sphb_ctx_slab_p2p_monitors(sphb_ctx *c, double *out3)
{
    SLAB_ENTER();
    SlabState &s = c->slab;
    if (!s.mail || !out3) return set_error("SPHB:state", "no mailbox");
    P2PState h;
    SPHB_CUDA(cudaMemcpyAsync(&h, s.p2p_state.p, sizeof(P2PState), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    out3[0] = h.red[0];
    out3[1] = h.red[1];
    out3[2] = h.red[2];
    return 0;
}

// The state a caller on the host needs to carry a slab from one step to the next: pos, vel, rho, drho_dt of the rows
// [0, n_rows) of the owned region (rows that have left included: they are marked dead on the device and stay so).
// to_host != 0: device -> host, else host -> device.  Column-major [n_rows x 2] / [n_rows] arrays, ideally pinned;
// asynchronous on the context's stream (sphb_ctx_sync waits).  The end-to-end leg of bench.py for sharded runs.
SPHB_API int sphb_ctx_slab_rows_io(sphb_ctx *c, int to_host, int n_rows, double *pos, double *vel, double *rho,
                                   double *drho_dt)
{
    SLAB_ENTER();
    if (n_rows < 0 || n_rows > c->slab.nf_cap || !pos || !vel || !rho || !drho_dt) return set_error("SPHB:arg", "bad rows");
    if (n_rows == 0) return 0;
    const size_t n = (size_t)n_rows;
    double *stage = c->stage.as<double>();   // 6 * n_total doubles: two 2-column regions used here
    if (to_host) {
        launch_unpack2(c->pos.as<double2>(), stage, n_rows, c->st);
        launch_unpack2(c->vel.as<double2>(), stage + 2 * n, n_rows, c->st);
        SPHB_CUDA(cudaMemcpyAsync(pos, stage, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        SPHB_CUDA(cudaMemcpyAsync(vel, stage + 2 * n, 2 * n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        SPHB_CUDA(cudaMemcpyAsync(rho, c->rho.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        SPHB_CUDA(cudaMemcpyAsync(drho_dt, c->drho_dt.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    } else {
        SPHB_CUDA(cudaMemcpyAsync(stage, pos, 2 * n * sizeof(double), cudaMemcpyHostToDevice, c->st));
        SPHB_CUDA(cudaMemcpyAsync(stage + 2 * n, vel, 2 * n * sizeof(double), cudaMemcpyHostToDevice, c->st));
        launch_pack2(stage, c->pos.as<double2>(), n_rows, c->st);
        launch_pack2(stage + 2 * n, c->vel.as<double2>(), n_rows, c->st);
        SPHB_CUDA(cudaMemcpyAsync(c->rho.p, rho, n * sizeof(double), cudaMemcpyHostToDevice, c->st));
        SPHB_CUDA(cudaMemcpyAsync(c->drho_dt.p, drho_dt, n * sizeof(double), cudaMemcpyHostToDevice, c->st));
        c->mirrors_stale = true;
    }
    c->launches += 2;
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

// first half of a reference step on the slab: density + KGC, forces, continuity (m:254-270)
SPHB_API int sphb_ctx_slab_step_begin(sphb_ctx *c, double remain)
{
    SLAB_ENTER();
    if (!c->have_list) return set_error("SPHB:state", "no neighbour list: call sphb_ctx_slab_step_end first");
    return ctx_step_physics(c, remain);
}

// second half, after the exchange: list rebuild and wall-shear partials (m:280-283)
SPHB_API int sphb_ctx_slab_step_end(sphb_ctx *c)
{
    SLAB_ENTER();
    c->nw.alive = c->slab.pid.as<int>();   // rows that migrated away are skipped by the build
    SPHB_TRY(ctx_do_neighbor(c));
    SPHB_TRY(ctx_do_shear(c));
    return 0;
}

// [max|v| over owned rows, tau_bottom partial, tau_top partial] -> device buffer of the caller (3 doubles)
SPHB_API int sphb_ctx_slab_scalars_export(sphb_ctx *c, double *out_dev)
{
    SLAB_ENTER();
    k_slab_scalars_out<<<1, 1, 0, c->st>>>(c->sc(), out_dev);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}
// the all-reduced max|v| back into the context (device buffer, 1 double)
SPHB_API int sphb_ctx_slab_vmax_import(sphb_ctx *c, const double *in_dev)
{
    SLAB_ENTER();
    k_slab_vmax_in<<<1, 1, 0, c->st>>>(c->sc(), in_dev);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

// global identity of every fluid row (owned rows first), n_owned + n_halo ints
SPHB_API int sphb_ctx_slab_ids(sphb_ctx *c, int *ids)
{
    SLAB_ENTER();
    SPHB_TRY(refresh_counts(c));
    const size_t n = (size_t)(c->slab.n_owned + c->slab.n_halo);
    if (n) SPHB_CUDA(cudaMemcpyAsync(ids, c->slab.pid.p, n * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    SPHB_CUDA(cudaStreamSynchronize(c->st));
    return 0;
}

// the all-gathered scalars of every rank ([world][3], device) -> max|v| into the context, reduced triple out
SPHB_API int sphb_ctx_slab_scalars_reduce(sphb_ctx *c, const double *gathered_dev, int world, double *out_dev3)
{
    SLAB_ENTER();
    k_slab_scalars_reduce<<<1, 1, 0, c->st>>>(c->sc(), gathered_dev, nullptr, world, out_dev3);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

// same, the triple of rank r at gathered_dev + offsets_dev[r] (doubles): for callers that move the scalars
// inside the same all-to-all as the record buffers
SPHB_API int sphb_ctx_slab_scalars_reduce_at(sphb_ctx *c, const double *gathered_dev, const long long *offsets_dev,
                                             int world, double *out_dev3)
{
    SLAB_ENTER();
    k_slab_scalars_reduce<<<1, 1, 0, c->st>>>(c->sc(), gathered_dev, offsets_dev, world, out_dev3);
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

SPHB_API int sphb_ctx_slab_counts(sphb_ctx *c, int *n_owned, int *n_halo, int *n_wall)
{
    SLAB_ENTER();
    SPHB_TRY(refresh_counts(c));
    if (c->slab.n_owned + c->slab.n_halo > c->slab.nf_cap)
        return set_error("SPHB:capacity", "slab holds %d owned + %d halo fluid rows, capacity %d", c->slab.n_owned,
                         c->slab.n_halo, c->slab.nf_cap);
    if (n_owned) *n_owned = c->slab.n_owned;
    if (n_halo) *n_halo = c->slab.n_halo;
    if (n_wall) *n_wall = c->slab.n_wall;
    return 0;
}

}  // extern "C"
