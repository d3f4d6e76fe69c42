// scan.cu -- hand-written exclusive prefix sum over int32 (no CUB/Thrust on the hot path).
//
// out[i] = sum_{k<i} in[k] for i = 0..n (out has n+1 elements, out[n] = grand total).
// Two launches: (A) per-tile scan, 4096 elements per 1024-thread block, tile totals to tmp;
// (B) every tile sums the totals before it and adds that offset (one launch when a single tile suffices).
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

__device__ __forceinline__ void scan_tiles_body(const int *__restrict__ in, int *__restrict__ out,
                                                int *__restrict__ tile_totals, size_t n, int *__restrict__ total_out)
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
    if (threadIdx.x == 0) {
        tile_totals[blockIdx.x] = total_sh;
        if (gridDim.x == 1) {   // a single tile is the whole scan
            out[n] = total_sh;
            if (total_out) *total_out = total_sh;
        }
    }
}
__global__ void __launch_bounds__(kScanThreads) scan_tiles(const int *__restrict__ in, int *__restrict__ out,
                                                          int *__restrict__ tile_totals, size_t n,
                                                          int *__restrict__ total_out)
{
    pdl_wait();
    pdl_trigger_top();
    scan_tiles_body(in, out, tile_totals, n, total_out);
    pdl_trigger_end();
}
__global__ void __launch_bounds__(kScanThreads) scan_tiles_batch(const ScanChan *__restrict__ chan, size_t n)
{
    const ScanChan c = chan[blockIdx.y];
    scan_tiles_body(c.in, c.out, c.tile_totals, n, c.total_out);
}

// Every tile sums the totals of the tiles before it (a few hundred values at most) and adds the offset
// to its elements; the last tile also writes the grand total.  No separate "scan of the totals" launch.
__device__ __forceinline__ void scan_finish_body(int *__restrict__ out, const int *__restrict__ tile_totals, size_t n,
                                                 int *__restrict__ total_out)
{
    __shared__ int warp_part[32];
    __shared__ int off_sh;
    int part = 0;
    for (unsigned i = threadIdx.x; i < blockIdx.x; i += kScanThreads) part += tile_totals[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        int s = warp_part[threadIdx.x];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (threadIdx.x == 0) off_sh = s;
    }
    __syncthreads();
    const int off = off_sh;
    const size_t base = (size_t)blockIdx.x * kTile + (size_t)threadIdx.x * kItems;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
        if (base + k < n) out[base + k] += off;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        const int total = off + tile_totals[blockIdx.x];
        out[n] = total;
        if (total_out) *total_out = total;
    }
}
__global__ void __launch_bounds__(kScanThreads) scan_finish(int *__restrict__ out, const int *__restrict__ tile_totals,
                                                           size_t n, int *__restrict__ total_out)
{
    pdl_wait();
    pdl_trigger_top();
    scan_finish_body(out, tile_totals, n, total_out);
    pdl_trigger_end();
}
__global__ void __launch_bounds__(kScanThreads) scan_finish_batch(const ScanChan *__restrict__ chan, size_t n)
{
    const ScanChan c = chan[blockIdx.y];
    scan_finish_body(c.out, c.tile_totals, n, c.total_out);
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
    SPHB_LAUNCH_PDL(scan_tiles, (unsigned)n_tiles, kScanThreads, 0, st, in, out, totals, n, total_out_dev);
    if (n_tiles > 1) SPHB_LAUNCH_PDL(scan_finish, (unsigned)n_tiles, kScanThreads, 0, st, out, totals, n, total_out_dev);
    if (launches) *launches += (n_tiles > 1) ? 2 : 1;
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

size_t scan_workspace_ints(size_t n) { return n == 0 ? 1 : (n + kTile - 1) / kTile; }

int exclusive_scan_i32_batch(const ScanChan *chan_dev, int n_chan, size_t n, cudaStream_t st, long long *launches)
{
    if (n == 0 || n_chan <= 0) return set_error("SPHB:arg", "empty batched scan");
    const size_t n_tiles = (n + kTile - 1) / kTile;
    const dim3 grid((unsigned)n_tiles, (unsigned)n_chan);
    scan_tiles_batch<<<grid, kScanThreads, 0, st>>>(chan_dev, n);
    if (n_tiles > 1) scan_finish_batch<<<grid, kScanThreads, 0, st>>>(chan_dev, n);
    if (launches) *launches += (n_tiles > 1) ? 2 : 1;
    SPHB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace sphb
