// cuda_runtime.h of the CPU FUNCTIONAL SIMULATOR (tests/cpusim) -- TEST INFRASTRUCTURE, never shipped, never loaded
// by the product: sph-poiseuille-flow_b200/__init__.py only ever loads libsph_b200.so (sm_100a) and fails without
// a CUDA device.  tests/cpusim/build.py compiles the product's .cu sources with g++ against this header into
// tests/cpusim/_build/libsph_b200_cpusim.so so that the HOST LOGIC and the KERNEL BODIES of code written without a
// GPU at hand (batched launches, flat staging, ...) can be exercised in this container before GPU minutes are
// spent on them.  It says nothing about performance and is no parity evidence: parity is claimed only from
// `pytest -m gpu` on a B200.
//
// Execution model: a launch runs its CTAs one after the other on the calling thread; the threads of a CTA are
// ucontext fibers switched only at __syncthreads / warp-synchronous intrinsics (which therefore behave as
// barriers over the live threads of the CTA / warp), so atomics are plain read-modify-writes and __shared__ data
// is ordinary static storage.  CUDA graphs record closures and replay them.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#define SPHB_CPUSIM 1
#ifndef __CUDACC__
#define __CUDACC__ 1
#endif

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static

// ---- vector types ----------------------------------------------------------------------------------------
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) double4 { double x, y, z, w; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(8))) int2 { int x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double4 make_double4(double x, double y, double z, double w) { return double4{x, y, z, w}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
constexpr int warpSize = 32;

// ---- runtime API (cpusim_rt.cpp) ---------------------------------------------------------------------------
enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
                      cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1,
                             cudaStreamCaptureModeRelaxed = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
constexpr unsigned cudaStreamNonBlocking = 1;

namespace cpusim {
struct Graph {
    std::vector<std::function<void()>> nodes;
};
struct Stream {
    bool capturing = false;
    Graph *graph = nullptr;
};
struct Event {
    double t_ms = 0.0;
};
// runs `body` for every thread of every CTA of the grid (CTAs in order, threads as fibers)
void run_grid(const std::function<void()> &body, dim3 grid, dim3 block, size_t dyn_smem);
void enqueue(Stream *st, std::function<void()> op);   // runs now, or becomes a node of the graph being captured
unsigned char *dyn_smem();
void cta_barrier();
int cta_barrier_or(int pred);
void warp_barrier();
// warp exchange: every live lane deposits 8 bytes, then reads the deposit of `src` (own value when src is not live)
unsigned long long warp_exchange(unsigned long long mine, int src_lane);
unsigned warp_ballot(int pred);
}  // namespace cpusim

typedef cpusim::Stream *cudaStream_t;
typedef cpusim::Event *cudaEvent_t;
typedef cpusim::Graph *cudaGraph_t;
typedef cpusim::Graph *cudaGraphExec_t;

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaSetDevice(int dev);
cudaError_t cudaGetDevice(int *dev);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocHost(void **p, size_t bytes);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st = nullptr);
cudaError_t cudaMemset(void *dst, int value, size_t bytes);
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t st = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaStreamIsCapturing(cudaStream_t st, cudaStreamCaptureStatus *status);
cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode mode);
cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t *graph);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *exec, cudaGraph_t graph, unsigned long long flags);
cudaError_t cudaGraphLaunch(cudaGraphExec_t exec, cudaStream_t st);
cudaError_t cudaGraphDestroy(cudaGraph_t graph);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t exec);
cudaError_t cudaEventCreate(cudaEvent_t *e);
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }   // (work runs at enqueue time)
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }   // launches run synchronously: always complete
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
template <class K>
inline cudaError_t cudaFuncSetAttribute(K *, cudaFuncAttribute, int)
{
    return cudaSuccess;
}
// peer-to-peer plumbing of slab.cu: one process, one "device": handles are never opened
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorInvalidValue; }
inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorInvalidValue; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
enum { cudaErrorPeerAccessAlreadyEnabled = 704 };
// persistent launches size their grid from these: the simulator pretends to be a 3-SM device with 2 CTAs per SM, so
// that every CTA takes several tiles (the pipelined path) on the small test channels
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int)
{
    *v = 3;
    return cudaSuccess;
}
template <class K>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K *, int, size_t)
{
    *n = 2;
    return cudaSuccess;
}

// kernel<<<grid, block, smem, stream>>>(args...) is rewritten to this by tests/cpusim/build.py
namespace cpusim {
inline dim3 to_dim3(dim3 d) { return d; }
inline dim3 to_dim3(unsigned long long n) { return dim3((unsigned)n); }
void set_launch_error();
template <class K, class G, class B, class... A>
inline void launch(K *kernel, G grid, B block, size_t smem, cudaStream_t st, A... args)
{
    const dim3 g = to_dim3(grid), b = to_dim3(block);
    // what the hardware refuses: more than 227 KB of shared memory per CTA (dynamic + a few KB of static), empty grids
    if (smem + 3072 > (size_t)227 * 1024 || g.x == 0 || g.y == 0 || b.x == 0 || b.x * b.y * b.z > 1024) {
        set_launch_error();
        return;
    }
    enqueue(st, [=]() { run_grid([=]() { kernel(args...); }, g, b, smem); });
}
}  // namespace cpusim

// ---- device intrinsics ---------------------------------------------------------------------------------------
static inline void __syncthreads() { cpusim::cta_barrier(); }
static inline int __syncthreads_or(int pred) { return cpusim::cta_barrier_or(pred); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cpusim::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) {}
// advances on every call, so that a device-side spin with a time-out (slab.cu mail_wait) ends instead of hanging
static inline long long clock64()
{
    static long long t = 0;
    return t += 100000000LL;
}
template <class T>
static inline T __ldcg(const T *p) { return *p; }
static inline void __threadfence_block() {}

namespace cpusim {
template <class T>
inline T shfl_from(T v, int src)
{
    static_assert(sizeof(T) <= 8, "warp exchange moves 8 bytes");
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    bits = warp_exchange(bits, src);
    T out;
    std::memcpy(&out, &bits, sizeof(T));
    return out;
}
inline int lane_id() { return (int)(threadIdx.x & 31u); }
}  // namespace cpusim
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
    const int lane = cpusim::lane_id();
    return cpusim::shfl_from(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32)
{
    (void)width;
    return cpusim::shfl_from(v, cpusim::lane_id() ^ mask);
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32)
{
    (void)width;
    const int lane = cpusim::lane_id();
    return cpusim::shfl_from(v, lane >= (int)delta ? lane - (int)delta : lane);
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32)
{
    (void)width;
    const int lane = cpusim::lane_id();
    return cpusim::shfl_from(v, lane + (int)delta < 32 ? lane + (int)delta : lane);
}
static inline unsigned __ballot_sync(unsigned, int pred) { return cpusim::warp_ballot(pred); }
static inline int __all_sync(unsigned, int pred) { return cpusim::warp_ballot(!pred) == 0u; }
static inline int __any_sync(unsigned, int pred) { return cpusim::warp_ballot(pred) != 0u; }

static inline long long __double_as_longlong(double d)
{
    long long b;
    std::memcpy(&b, &d, 8);
    return b;
}
static inline double __longlong_as_double(long long b)
{
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift)
{
    const unsigned long long v = ((unsigned long long)hi << 32) | lo;
    return (unsigned)((v << (shift & 31u)) >> 32);
}
template <class T>
static inline T __ldg(const T *p)
{
    return *p;
}
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

// atomics: one OS thread, fibers switch only at barriers (the value parameter does not take part in deduction)
namespace cpusim {
template <class T>
struct same {
    typedef T type;
};
}  // namespace cpusim
#define CPUSIM_V(T) typename cpusim::same<T>::type
template <class T>
static inline T atomicAdd(T *p, CPUSIM_V(T) v)
{
    const T old = *p;
    *p = old + v;
    return old;
}
template <class T>
static inline T atomicMax(T *p, CPUSIM_V(T) v)
{
    const T old = *p;
    if (v > old) *p = v;
    return old;
}
template <class T>
static inline T atomicMin(T *p, CPUSIM_V(T) v)
{
    const T old = *p;
    if (v < old) *p = v;
    return old;
}
template <class T>
static inline T atomicOr(T *p, CPUSIM_V(T) v)
{
    const T old = *p;
    *p = old | v;
    return old;
}
template <class T>
static inline T atomicCAS(T *p, CPUSIM_V(T) cmp, CPUSIM_V(T) v)
{
    const T old = *p;
    if (old == cmp) *p = v;
    return old;
}
template <class T>
static inline T atomicExch(T *p, CPUSIM_V(T) v)
{
    const T old = *p;
    *p = v;
    return old;
}

// CUDA's global integer min / max
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }

// the SFU seeds of physmath.cuh (rcp.approx / rsqrt.approx: ~22 good bits), for the rewritten asm statements
namespace cpusim {
inline double approx22(double v)
{
    unsigned long long b;
    std::memcpy(&b, &v, 8);
    b &= ~((1ull << 30) - 1ull);   // keep 22 mantissa bits
    std::memcpy(&v, &b, 8);
    return v;
}
inline double rcp_approx(double x) { return approx22(1.0 / x); }
inline double rsqrt_approx(double x) { return approx22(1.0 / std::sqrt(x)); }
}  // namespace cpusim
