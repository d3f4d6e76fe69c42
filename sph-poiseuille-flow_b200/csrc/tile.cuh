// tile.cuh -- cell tiles: the unit of work of the neighbour walk and of every gather kernel.
//
// The cell-sorted ENTRY array of the neighbour build (real + periodic ghost entries, cells in row-major
// order, nb:46-103) is cut into tiles of kTileRows cell rows x `tile_w` cell columns.  One CTA owns one
// tile: it stages the entries of the tile's (kTileRows + 2) x (tile_w + 2) cell neighbourhood -- per cell
// row one contiguous entry range plus, for the first / last tile of a row, the periodic wrap column --
// into shared memory with coalesced loads; its threads (one per own entry) then work out of shared
// memory only.
//
// A neighbour is named by its SLOT in the staged array, so a list entry is 16 bits:
//     slot (13 bits) | minimum-image shift + 3 (3 bits)
// and a walker's row is stored eight entries to a 128-bit word, column-major over walkers:
//     ell_q[(k / 8) * stride + entry],   entry k in bits 16 * (7 - k % 8) of the 128-bit word.
// Rows are padded to a multiple of eight with the tile's DUMMY slot (index n_staged): every gather
// kernel stages a far-away, zero-volume particle there, so the padded entries contribute exact zeros
// and the unrolled row loops need no per-entry bound check.
// A ghost entry carries its already shifted position, so interior tiles never shift.
//
// The descriptors are written once per list build (k_tile_setup) and read by every kernel of the step.
#pragma once
#include "common.cuh"

namespace sphb {

constexpr int kTileThreads = 128;   // threads per CTA (one per own entry, strided when a tile holds more)
#ifndef SPHB_TILE_ROWS
#define SPHB_TILE_ROWS 2
#endif
constexpr int kTileRows = SPHB_TILE_ROWS;   // own cell rows of a tile
constexpr int kStageRows = kTileRows + 2;
constexpr int kMaxTileW = 32;       // widest tile in cells
constexpr int kSlotBits = 13;
constexpr int kMaxSlots = (1 << kSlotBits) - 2;   // staged entries per tile at most (+ the dummy slot)
constexpr unsigned kSlotMask = (1u << kSlotBits) - 1u;
constexpr int kShiftBias = 3;       // stored shift = m + 3, m in [-3, 4]
constexpr double kDummyCoord = 1.0e100;   // position of the dummy slot: finite r, q >> 2, W = dW = 0

enum : int { kTileValid = 1, kTileSeam = 2, kTileHasWall = 4 };   // kTileHasWall is set by the walk (it sees the keys)

struct TileSeg {
    int b, e;   // entry range [b, e)
    int off;    // slot of entry b
};

struct __align__(16) TileDesc {
    TileSeg seg[kStageRows * 3];   // [staged row r <-> cell row cy0-1+r][0 main columns a..b-1, 1 wrap column ncx-1, 2 wrap column 0]
    int n_staged;                  // also the dummy slot
    int own_b[kTileRows];          // entries of the tile's own cells, per own row (columns cx0..cx1-1)
    int own_n[kTileRows];
    int own_off[kTileRows];        // slot of entry own_b[k]
    int cy0, cx0, cx1;             // first own cell row; own columns [cx0, cx1)
    int a, b;                      // main range = columns [a, b)
    int flags;
    int pad[(4 - (kStageRows * 9 + 1 + 3 * kTileRows + 6) % 4) % 4 + 4];
};
// cell_start of the main range's cells a .. b (b - a + 1 values) per staged row: only the neighbour walk
// needs it, and loads it from the cell starts itself
typedef int TileCellStarts[kStageRows][kMaxTileW + 3];
static_assert(sizeof(TileDesc) % 16 == 0, "TileDesc is copied as int4");

__host__ __device__ inline int tiles_per_row(int ncx, int w) { return (ncx + w - 1) / w; }
__host__ __device__ inline int tile_rows(int ncy) { return (ncy + kTileRows - 1) / kTileRows; }

// Tile rows whose walkers can have wall neighbours -- the candidates of the wall-shear monitor: [0, nb) and
// [t0, n_trows), i.e. cy0 - 1 <= cyb (bottom wall, y <= 0) or cy0 + kTileRows >= cyt (top wall, y >= DH).
__device__ __forceinline__ void wall_tile_rows(const GridDesc *g, double DH, int n_trows, int &nb, int &t0)
{
    const int cyb = (int)floor((0.0 - g->y_min) / g->cell), cyt = (int)floor((DH - g->y_min) / g->cell);
    nb = cyb + 1 < 0 ? 0 : min(n_trows, (cyb + 1) / kTileRows + 1);
    t0 = max(nb, cyt - kTileRows <= 0 ? 0 : (cyt - 1) / kTileRows);
}

__device__ __forceinline__ unsigned ell_pack(int slot, int m) { return (unsigned)slot | ((unsigned)(m + kShiftBias) << kSlotBits); }
__device__ __forceinline__ int ell_slot(unsigned c) { return (int)(c & kSlotMask); }
__device__ __forceinline__ int ell_shift(unsigned c) { return (int)((c >> kSlotBits) & 7u) - kShiftBias; }

// k-th (k = 0..7) 16-bit entry of a row word
__device__ __forceinline__ unsigned ell_extract(const uint4 &q, int u)
{
    const unsigned w = u < 2 ? q.w : (u < 4 ? q.z : (u < 6 ? q.y : q.x));
    return (u & 1) ? (w & 0xffffu) : (w >> 16);
}

// Appends one 16-bit entry to the row word under construction (a 128-bit shift register).
__device__ __forceinline__ void ell_push(uint4 &q, unsigned v)
{
    q.w = __funnelshift_l(q.z, q.w, 16);
    q.z = __funnelshift_l(q.y, q.z, 16);
    q.y = __funnelshift_l(q.x, q.y, 16);
    q.x = (q.x << 16) | v;
}

// CTA-wide copy of the tile's descriptor into shared memory (callers sync afterwards).
__device__ __forceinline__ void tile_load(TileDesc &T, const TileDesc *__restrict__ tiles, unsigned tile)
{
    const int4 *src = reinterpret_cast<const int4 *>(tiles + tile);
    int4 *dst = reinterpret_cast<int4 *>(&T);
    for (int k = threadIdx.x; k < (int)(sizeof(TileDesc) / 16); k += blockDim.x) dst[k] = src[k];
}

// slot range [sb, se) of cell (staged row r, column cx) of the staged neighbourhood
__device__ __forceinline__ void tile_cell_slots(const TileDesc &T, const TileCellStarts &cs, int r, int cx, int ncx,
                                                int &sb, int &se)
{
    if (cx >= T.a && cx < T.b) {
        const int c0 = cs[r][cx - T.a];
        sb = T.seg[3 * r].off + (c0 - T.seg[3 * r].b);
        se = sb + (cs[r][cx - T.a + 1] - c0);
    } else {
        const TileSeg &s = T.seg[3 * r + (cx == ncx - 1 ? 1 : 2)];
        sb = s.off;
        se = s.off + (s.e - s.b);
    }
}

// entry behind a slot (export paths only)
__device__ __forceinline__ int tile_slot_entry(const TileDesc &T, int slot)
{
    for (int k = 0; k < kStageRows * 3; ++k) {
        const int n = T.seg[k].e - T.seg[k].b;
        if (slot >= T.seg[k].off && slot < T.seg[k].off + n) return T.seg[k].b + (slot - T.seg[k].off);
    }
    return -1;
}

// for (e, slot) over every staged entry, strided over the CTA's threads
#define SPHB_TILE_FOR_STAGED(T, e, slot)                                                                  \
    for (int sg__ = 0; sg__ < kStageRows * 3; ++sg__)                                                     \
        for (int e = (T).seg[sg__].b + (int)threadIdx.x, slot = (T).seg[sg__].off + (int)threadIdx.x;     \
             e < (T).seg[sg__].e; e += (int)blockDim.x, slot += (int)blockDim.x)

#if SPHB_FLAT_STAGING
// ---- flat staging (opt-in: build with SPHB_FLAT_STAGING=1; profiles/r1_v8_source_hotspots.md, item 1) --------
// The segment-by-segment loop above makes a CTA wait for one global-memory round trip per non-empty segment
// and dependent load level (ent_key -> S.<field>[j]) before its row loop can start.  Here the staged slots
// [0, n_staged) form ONE index range, kStageUnroll slots per thread and pass, and a pass is split into
//   load(e)  -> a record of registers: every load of it is unconditional (clamped slot), so the loads of all
//               the thread's slots are in flight together, level by level;
//   store(slot, record) -> shared memory, predicated on slot < n_staged.
#ifndef SPHB_STAGE_UNROLL
#define SPHB_STAGE_UNROLL 3       // 3 x 128 slots cover a lattice tile (4 x 11 cells x 6.8 entries ~ 300) in one pass
#endif
constexpr int kStageUnroll = SPHB_STAGE_UNROLL;

// entry behind a slot: the segment offsets are non-decreasing (an empty segment shares the offset of its
// successor), so the segment of `slot` is the LAST one whose offset is <= slot
__host__ __device__ __forceinline__ int tile_entry_of_slot(const TileDesc &T, int slot)
{
    int k = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 1; j < kStageRows * 3; ++j) k += (T.seg[j].off <= slot) ? 1 : 0;
    return T.seg[k].b + (slot - T.seg[k].off);
}

template <class Load, class Store>
__device__ __forceinline__ void tile_stage_flat(const TileDesc &T, Load &&load, Store &&store)
{
    const int n = T.n_staged;
    if (n <= 0) return;
    for (int base = 0; base < n; base += kStageUnroll * (int)blockDim.x) {
        int slot[kStageUnroll];
        decltype(load(0)) rec[kStageUnroll];
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u) {
            slot[u] = base + u * (int)blockDim.x + (int)threadIdx.x;
            rec[u] = load(tile_entry_of_slot(T, min(slot[u], n - 1)));
        }
#pragma unroll
        for (int u = 0; u < kStageUnroll; ++u)
            if (slot[u] < n) store(slot[u], rec[u]);
    }
}
#endif

// Own entries of a tile as one virtual index range: v in [0, own total) -> own row, entry, slot.
__device__ __forceinline__ int tile_own_total(const TileDesc &T)
{
    int n = 0;
#pragma unroll
    for (int k = 0; k < kTileRows; ++k) n += T.own_n[k];
    return n;
}
__device__ __forceinline__ void tile_own_locate(const TileDesc &T, int v, int &orow, int &ew, int &si)
{
    orow = 0;
#pragma unroll
    for (int k = 0; k + 1 < kTileRows; ++k) {
        if (orow == k && v >= T.own_n[k]) {
            v -= T.own_n[k];
            orow = k + 1;
        }
    }
    ew = T.own_b[orow] + v;
    si = T.own_off[orow] + v;
}

// for (ew, si, orow) over the tile's own entries (ew = entry, si = its slot, orow = own row 0..kTileRows-1):
// one per thread, all own rows in ONE pass (a tile holds <= blockDim own entries by construction of its
// width; a denser tile takes more passes)
#define SPHB_TILE_FOR_OWN(T, ew, si, orow)                                                                \
    for (int v__ = (int)threadIdx.x, nown__ = tile_own_total(T), orow = 0, ew = 0, si = 0;                \
         v__ < nown__ && (tile_own_locate(T, v__, orow, ew, si), true); v__ += (int)blockDim.x)

}  // namespace sphb
