// sort.cu -- the periodic cell sort of the fluid rows (row S1; reference: sort_particles_by_cell /
// sort_subset_indices, SPH_Poiseuille.m:529-568) as a device counting sort.
//
// key = (cy, cx) with cx = floor(mod(x, DL) / 2h), cy = floor((y - min fluid y) / 2h); MATLAB's
// sortrows is stable, so equal keys keep their original order; wall rows stay where they are.
// Counting sort: histogram (atomics) -> exclusive scan (scan.cu) -> scatter into atomic slots -> each
// key's range sorted by original index (ranges hold ~7 particles), which makes the result identical to
// the stable sort regardless of atomic ordering.  One gather kernel then permutes every state array.
// Compiled with -fmad=false like neighbor.cu (key arithmetic must match the reference's).
#include "sort.cuh"

#include <climits>
#include <cmath>

namespace sphb {

namespace {

__device__ __forceinline__ long long ordered_key(double d)
{
    long long b = __double_as_longlong(d);
    return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double from_ordered_key(long long k)
{
    return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}

__global__ void k_y0_reset(DevScalars *sc) { sc->y0_key = LLONG_MAX; }

__global__ void __launch_bounds__(256) k_y0(const double2 *__restrict__ pos, int nf, DevScalars *sc)
{
    long long kmin = LLONG_MAX;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) {
        const long long k = ordered_key(pos[i].y);
        kmin = k < kmin ? k : kmin;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, kmin, d);
        kmin = a < kmin ? a : kmin;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(&sc->y0_key, kmin);
}

__device__ __forceinline__ int sort_key(double2 p, double DL, double cell, double y0, int ncx, int key_cap,
                                        int *flags)
{
    double m = fmod(p.x, DL);
    if (m < 0.0) m += DL;
    const int cx = (int)floor(m / cell);
    const int cy = (int)floor((p.y - y0) / cell);
    long long key = (long long)cy * ncx + cx;
    if (key < 0 || key >= key_cap || cx >= ncx) {
        atomicOr(flags, kErrCellCap);
        key = key < 0 ? 0 : key_cap - 1;
    }
    return (int)key;
}

__global__ void __launch_bounds__(kBlock) k_sort_count(const double2 *__restrict__ pos, int nf, double DL,
                                                      double cell, int ncx, int key_cap,
                                                      const DevScalars *__restrict__ sc, int *__restrict__ keys,
                                                      int *__restrict__ cnt, int *flags)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const double y0 = from_ordered_key(sc->y0_key);
    const int key = sort_key(pos[i], DL, cell, y0, ncx, key_cap, flags);
    keys[i] = key;
    atomicAdd(&cnt[key], 1);
}

__global__ void __launch_bounds__(kBlock) k_sort_scatter(int nf, const int *__restrict__ keys,
                                                        const int *__restrict__ start, int *__restrict__ fill,
                                                        int *__restrict__ perm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    const int key = keys[i];
    perm[start[key] + atomicAdd(&fill[key], 1)] = i;
}

__global__ void __launch_bounds__(kBlock) k_sort_canon(int key_cap, const int *__restrict__ start,
                                                      int *__restrict__ perm)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= key_cap) return;
    const int s = start[c], e = start[c + 1];
    for (int a = s + 1; a < e; ++a) {
        const int v = perm[a];
        int b = a - 1;
        while (b >= s && perm[b] > v) {
            perm[b + 1] = perm[b];
            --b;
        }
        perm[b + 1] = v;
    }
}

__global__ void __launch_bounds__(kBlock) k_permute(int nf, int nt, const int *__restrict__ perm, PermuteArrays A)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    const int s = i < nf ? perm[i] : i;
#pragma unroll
    for (int a = 0; a < PermuteArrays::kN2; ++a) A.out2[a][i] = A.in2[a][s];
#pragma unroll
    for (int a = 0; a < PermuteArrays::kN1; ++a) A.out1[a][i] = A.in1[a][s];
    A.outB[i] = A.inB[s];
}

}  // namespace

void SortWork::release_all()
{
    DevBuf *all[] = {&keys, &cnt, &start, &fill, &perm, &scan_tmp};
    for (DevBuf *b : all) b->release();
}

int sort_plan(SortWork &w, int nf, double DL, double h, double y_span_hint)
{
    const double cell = 2.0 * h;
    w.ncx = (int)std::floor(DL / cell) + 1;
    const double ncy = std::floor(y_span_hint / cell) + 1.0 + 8.0;
    if ((double)w.ncx * ncy > 1.0e9) return set_error("SPHB:capacity", "sort key space too large");
    w.key_cap = (int)((double)w.ncx * ncy);
    SPHB_TRY(w.keys.reserve((size_t)nf * sizeof(int)));
    SPHB_TRY(w.perm.reserve((size_t)nf * sizeof(int)));
    SPHB_TRY(w.cnt.reserve(((size_t)w.key_cap + 1) * sizeof(int)));
    SPHB_TRY(w.start.reserve(((size_t)w.key_cap + 1) * sizeof(int)));
    SPHB_TRY(w.fill.reserve(((size_t)w.key_cap + 1) * sizeof(int)));
    SPHB_TRY(w.scan_tmp.reserve(((size_t)w.key_cap / 4096 + 2) * sizeof(int)));
    return 0;
}

int sort_permutation(SortWork &w, const double2 *pos, int nf, double DL, double h, DevScalars *sc, int *flags,
                     cudaStream_t st, long long *launches)
{
    const double cell = 2.0 * h;
    int blocks = (nf + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_y0_reset<<<1, 1, 0, st>>>(sc);
    k_y0<<<blocks, 256, 0, st>>>(pos, nf, sc);
    SPHB_CUDA(cudaMemsetAsync(w.cnt.p, 0, ((size_t)w.key_cap + 1) * sizeof(int), st));
    SPHB_CUDA(cudaMemsetAsync(w.fill.p, 0, ((size_t)w.key_cap + 1) * sizeof(int), st));
    k_sort_count<<<grid_for(nf), kBlock, 0, st>>>(pos, nf, DL, cell, w.ncx, w.key_cap, sc, w.keys.as<int>(),
                                                  w.cnt.as<int>(), flags);
    long long nl = 3;
    SPHB_TRY(exclusive_scan_i32(w.cnt.as<int>(), w.start.as<int>(), (size_t)w.key_cap, w.scan_tmp, nullptr, st, &nl));
    k_sort_scatter<<<grid_for(nf), kBlock, 0, st>>>(nf, w.keys.as<int>(), w.start.as<int>(), w.fill.as<int>(),
                                                    w.perm.as<int>());
    k_sort_canon<<<grid_for((size_t)w.key_cap), kBlock, 0, st>>>(w.key_cap, w.start.as<int>(), w.perm.as<int>());
    nl += 2;
    SPHB_CUDA(cudaGetLastError());
    if (launches) *launches += nl;
    return 0;
}

int sort_apply(const SortWork &w, int nf, int nt, const PermuteArrays &A, cudaStream_t st, long long *launches)
{
    k_permute<<<grid_for(nt), kBlock, 0, st>>>(nf, nt, w.perm.as<int>(), A);
    SPHB_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    return 0;
}

}  // namespace sphb
