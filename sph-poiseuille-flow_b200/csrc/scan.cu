// scan.cu -- hand-written exclusive prefix sum over int32 (no CUB/Thrust on the hot path).
//
// out[i] = sum_{k<i} in[k] for i = 0..n (out has n+1 elements, out[n] = grand total).
// Three launches: (A) per-tile scan, 4096 elements per 1024-thread block, tile totals to tmp;
// (B) one block scans the tile totals with a running carry; (C) tiles add their offset.
// Used for cell starts (N1), half-pair and CSR row pointers (N6) and the cell sort of particles (S1).
#include "common.cuh"

namespace sphb {

namespace {

constexpr int kScanThreads = 1024;
constexpr int kItems = 4;
constexpr int kTile = kScanThreads * kItems;

__device__ __forceinline__ int warp_inclusive_scan(int v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) >= d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across a 1024-thread block; returns the exclusive prefix,
// *block_total receives the sum over the block
__device__ __forceinline__ int block_exclusive_scan(int v, int *block_total)
{
    __shared__ int warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int s = warp_sums[lane];
        int si = warp_inclusive_scan(s);
        warp_sums[lane] = si - s;   // exclusive warp offsets
        if (lane == 31) *block_total = si;
    }
    __syncthreads();
    int res = inc - v + warp_sums[wid];
    return res;
}

__global__ void __launch_bounds__(kScanThreads) scan_tiles(const int *__restrict__ in, int *__restrict__ out,
                                                          int *__restrict__ tile_totals, size_t n)
{
    __shared__ int total_sh;
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
    int v[kItems];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    int pre = block_exclusive_scan(sum, &total_sh);
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        if (base + k < n) out[base + k] = pre;
        pre += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) tile_totals[blockIdx.x] = total_sh;
}

__global__ void __launch_bounds__(kScanThreads) scan_totals(int *__restrict__ tile_totals, size_t n_tiles,
                                                           int *__restrict__ out_end, int *__restrict__ total_out)
{
    __shared__ int total_sh;
    __shared__ int carry_sh;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (size_t start = 0; start < n_tiles; start += kScanThreads) {
        size_t i = start + threadIdx.x;
        int v = (i < n_tiles) ? tile_totals[i] : 0;
        int pre = block_exclusive_scan(v, &total_sh);
        int carry = carry_sh;
        if (i < n_tiles) tile_totals[i] = pre + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = carry + total_sh;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *out_end = carry_sh;
        if (total_out) *total_out = carry_sh;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_add_offsets(int *__restrict__ out,
                                                                const int *__restrict__ tile_totals, size_t n)
{
    const int off = tile_totals[blockIdx.x];
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
        if (base + k < n) out[base + k] += off;
}

}  // namespace

int exclusive_scan_i32(const int *in, int *out, size_t n, DevBuf &tmp, int *total_out_dev, cudaStream_t st,
                       long long *launches)
{
    if (n == 0) {
        SPHB_CUDA(cudaMemsetAsync(out, 0, sizeof(int), st));
        if (total_out_dev) SPHB_CUDA(cudaMemsetAsync(total_out_dev, 0, sizeof(int), st));
        return 0;
    }
    const size_t n_tiles = (n + kTile - 1) / kTile;
    SPHB_TRY(tmp.reserve(n_tiles * sizeof(int)));
    int *totals = tmp.as<int>();
    scan_tiles<<<(unsigned)n_tiles, kScanThreads, 0, st>>>(in, out, totals, n);
    scan_totals<<<1, kScanThreads, 0, st>>>(totals, n_tiles, out + n, total_out_dev);
    if (n_tiles > 1) scan_add_offsets<<<(unsigned)n_tiles, kScanThreads, 0, st>>>(out, totals, n);
    if (launches) *launches += (n_tiles > 1) ? 3 : 2;
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sphb
