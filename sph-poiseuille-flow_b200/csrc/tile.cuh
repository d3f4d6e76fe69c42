// tile.cuh -- shared-memory staging of neighbour-cell tiles.
//
// The cell-sorted ENTRY array of the neighbour build (real + periodic ghost entries, cells in
// row-major order, nb:46-103) is cut into tiles of kTileCells consecutive cells of one cell row.  One
// CTA owns one tile: it stages the entries of the tile's 3 x (kTileCells + 2) cell neighbourhood --
// at most nine contiguous entry ranges, the periodic wrap columns included -- into shared memory with
// coalesced loads, and its threads then walk their particles' neighbours out of shared memory.  A
// neighbour is named by its ENTRY index, so a ghost entry carries its already shifted position and
// the lookup below is three range compares.  Whatever does not fit (capacity) or lies outside the
// staged ranges (a seam pair met through a ghost entry) falls back to global loads.
#pragma once
#include "common.cuh"

namespace sphb {

constexpr int kTileCells = 16;     // cells of one cell row per CTA
constexpr int kTileThreads = 128;  // threads per CTA

struct TileSeg {
    int b, e;   // entry range [b, e) (already truncated to the staged part)
    int off;    // shared-memory slot of entry b
};

struct TileDesc {
    TileSeg seg[9];    // [row 0..2 = cy-1, cy, cy+1][0 main, 1 wrap column ncx-1, 2 wrap column 0]
    int full_e[9];     // untruncated end of each range (entries past seg.e are read from global)
    int cstart[3][kTileCells + 3];   // cell_start of the main range's cells a .. b (b - a + 1 values)
    int own_b, own_e;  // entries of the tile's own cells (row cy, columns cx0..cx1-1)
    int cy, cx0, cx1;
    int a, b;          // main range = columns [a, b)
    int valid;
};

// ELL entry: accepted ENTRY index (28 bits) | minimum-image shift + 1 (2 bits) | row offset + 1 (2 bits)
constexpr unsigned kEllIdMask = 0x0fffffffu;
__device__ __forceinline__ unsigned ell_pack(int entry, int m, int oy)
{
    return (unsigned)entry | ((unsigned)(m + 1) << 28) | ((unsigned)(oy + 1) << 30);
}
__device__ __forceinline__ int ell_entry(unsigned c) { return (int)(c & kEllIdMask); }
__device__ __forceinline__ int ell_shift(unsigned c) { return (int)((c >> 28) & 3u) - 1; }
__device__ __forceinline__ int ell_row(unsigned c) { return (int)(c >> 30); }

__host__ __device__ inline int tiles_per_row(int ncx) { return (ncx + kTileCells - 1) / kTileCells; }

// Fills the descriptor (the cell starts of the three rows are fetched by 3 x (kTileCells + 3) threads
// in parallel, then thread 0 lays the ranges out); everybody syncs.  `cap` = shared-memory slots.
__device__ __forceinline__ void tile_setup(TileDesc &T, int tile, int ncx, int ncy, const int *__restrict__ cell_start,
                                           int cap)
{
    const int tpr = tiles_per_row(ncx);
    const int cy = tile / tpr, cx0 = (tile % tpr) * kTileCells;
    const int cx1 = min(cx0 + kTileCells, ncx);
    const int a = max(cx0 - 1, 0), b = min(cx1 + 1, ncx);
    const bool valid = cy < ncy;
    {
        const int r = threadIdx.x / 32, k = threadIdx.x % 32;   // one warp per row
        const int row = cy + r - 1;
        if (r < 3 && valid && row >= 0 && row < ncy) {
            if (k <= b - a) T.cstart[r][k] = cell_start[row * ncx + a + k];
            if (k == 24) T.full_e[3 * r + 1] = cell_start[row * ncx + ncx - 1];   // wrap column ncx-1: begin
            if (k == 25) T.full_e[3 * r + 2] = cell_start[row * ncx + 1];         // wrap column 0: end
            if (k == 26) T.seg[3 * r + 1].b = cell_start[row * ncx + ncx];        // (scratch) end of the row
            if (k == 27) T.seg[3 * r + 2].b = cell_start[row * ncx];              // (scratch) start of the row
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        T.cy = cy;
        T.cx0 = cx0;
        T.cx1 = cx1;
        T.a = a;
        T.b = b;
        T.valid = valid;
        int used = 0;
        for (int r = 0; r < 3; ++r) {
            const int row = cy + r - 1;
            TileSeg *s = &T.seg[3 * r];
            int rb[3] = {0, 0, 0}, re[3] = {0, 0, 0};
            if (valid && row >= 0 && row < ncy) {
                rb[0] = T.cstart[r][0];
                re[0] = T.cstart[r][b - a];
                if (cx0 == 0 && b <= ncx - 1) {   // left halo column wraps to ncx-1
                    rb[1] = T.full_e[3 * r + 1];
                    re[1] = s[1].b;
                }
                if (cx1 == ncx && a >= 1) {       // right halo column wraps to 0
                    rb[2] = s[2].b;
                    re[2] = T.full_e[3 * r + 2];
                }
            }
            for (int k = 0; k < 3; ++k) {
                int len = re[k] - rb[k];
                if (len > cap - used) len = cap - used;
                if (len < 0) len = 0;
                s[k].b = rb[k];
                s[k].e = rb[k] + len;
                s[k].off = used;
                T.full_e[3 * r + k] = re[k];
                used += len;
            }
        }
        T.own_b = valid ? T.cstart[1][cx0 - a] : 0;
        T.own_e = valid ? T.cstart[1][cx1 - a] : 0;
    }
    __syncthreads();
}

// entry range [cb, ce) of cell (row r = 0..2, column cx) out of the staged cell starts
__device__ __forceinline__ void tile_cell_range(const TileDesc &T, int r, int cx, int ncx, int &cb, int &ce)
{
    if (cx >= T.a && cx < T.b) {
        cb = T.cstart[r][cx - T.a];
        ce = T.cstart[r][cx - T.a + 1];
    } else if (cx == ncx - 1) {
        cb = T.seg[3 * r + 1].b;
        ce = T.full_e[3 * r + 1];
    } else {
        cb = T.seg[3 * r + 2].b;
        ce = T.full_e[3 * r + 2];
    }
}

// shared-memory slot of entry e met in row `row` (0..2), or -1 when it is not staged
__device__ __forceinline__ int tile_lookup(const TileDesc &T, int e, int row)
{
    const TileSeg *s = &T.seg[3 * row];
    if (e >= s[0].b && e < s[0].e) return s[0].off + (e - s[0].b);
    if (e >= s[1].b && e < s[1].e) return s[1].off + (e - s[1].b);
    if (e >= s[2].b && e < s[2].e) return s[2].off + (e - s[2].b);
    return -1;
}

// run of entries [b, e_end) of ONE cell met in row `row`: slot of b (or -1) and how many are staged
__device__ __forceinline__ int tile_lookup_run(const TileDesc &T, int b, int e_end, int row, int &n_staged)
{
    const TileSeg *s = &T.seg[3 * row];
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (b >= s[k].b && b < s[k].e) {
            n_staged = min(e_end, s[k].e) - b;
            return s[k].off + (b - s[k].b);
        }
    n_staged = 0;
    return -1;
}

// for (e, slot) over every staged entry, strided over the CTA's threads
#define SPHB_TILE_FOR_STAGED(T, e, slot)                                               \
    for (int sg__ = 0; sg__ < 9; ++sg__)                                               \
        for (int e = (T).seg[sg__].b + (int)threadIdx.x, slot = (T).seg[sg__].off + (int)threadIdx.x; \
             e < (T).seg[sg__].e; e += (int)blockDim.x, slot += (int)blockDim.x)

}  // namespace sphb
