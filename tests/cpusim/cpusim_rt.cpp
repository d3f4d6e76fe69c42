// cpusim_rt.cpp -- runtime of the CPU functional simulator (see include/cuda_runtime.h: TEST INFRASTRUCTURE only).
//
// A launch runs its CTAs one after the other on the calling thread.  The threads of a CTA are ucontext fibers;
// a fiber runs until it finishes or reaches a barrier (CTA-wide: __syncthreads; warp-wide: every *_sync
// intrinsic), where it yields to the scheduler.  Threads that have returned no longer count in any barrier, as on
// the hardware.  A scheduler round in which no fiber makes progress is a deadlock and aborts with a message.
#include <cuda_runtime.h>

#include <chrono>
#include <csetjmp>
#include <mutex>
#include <ucontext.h>

// Fibers are created with makecontext; afterwards they are switched with _setjmp / _longjmp, which (unlike
// swapcontext) do not save and restore the signal mask with a system call on every switch.  AddressSanitizer
// tracks stacks through swapcontext only, so sanitized builds keep it.
#if defined(__SANITIZE_ADDRESS__)
#define CPUSIM_FAST_SWITCH 0
#else
#define CPUSIM_FAST_SWITCH 1
#endif

#include <sys/mman.h>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace cpusim {
namespace {

constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 512 * 1024;
constexpr size_t kDynSmemBytes = 232 * 1024;

struct Fiber {
    ucontext_t ctx;
    jmp_buf jb;
    bool started = false;
    bool done = true;
    bool waiting_cta = false;
    bool waiting_warp = false;
    unsigned long long cta_gen = 0, warp_gen = 0;
};

struct WarpState {
    int live = 0, arrived = 0;
    unsigned long long gen = 0;
    unsigned long long slot[32];
    int pred[32];
    bool lane_live[32];
};

struct Engine {
    ucontext_t sched;
    jmp_buf sched_jb;
    Fiber fib[kMaxThreads];
    unsigned char *stacks = nullptr;
    alignas(16) unsigned char smem[kDynSmemBytes];
    const std::function<void()> *body = nullptr;
    int n_threads = 0, cur = -1;
    int live = 0, arrived = 0, or_acc = 0, or_result = 0;
    unsigned long long gen = 0;
    WarpState warp[kMaxThreads / 32];
    bool progress = false;
    bool in_kernel = false;
} E;

void fail(const char *msg)
{
    std::fprintf(stderr, "cpusim: %s (block %u,%u thread %u)\n", msg, blockIdx.x, blockIdx.y, threadIdx.x);
    std::abort();
}

void yield_to_scheduler()
{
#if CPUSIM_FAST_SWITCH
    if (!_setjmp(E.fib[E.cur].jb)) _longjmp(E.sched_jb, 1);
#else
    swapcontext(&E.fib[E.cur].ctx, &E.sched);
#endif
}

// scheduler -> fiber t
void resume_fiber(int t)
{
    Fiber &f = E.fib[t];
#if CPUSIM_FAST_SWITCH
    if (!_setjmp(E.sched_jb)) {
        if (f.started) _longjmp(f.jb, 1);
        f.started = true;
        setcontext(&f.ctx);   // first entry: onto the fiber's own stack
    }
#else
    swapcontext(&E.sched, &f.ctx);
#endif
}

void release_cta_barrier()
{
    E.arrived = 0;
    E.or_result = E.or_acc;
    E.or_acc = 0;
    E.gen += 1;
    E.progress = true;
}

void release_warp_barrier(WarpState &w)
{
    w.arrived = 0;
    w.gen += 1;
    E.progress = true;
}

void fiber_exit_bookkeeping(int t)
{
    E.fib[t].done = true;
    E.live -= 1;
    WarpState &w = E.warp[t >> 5];
    w.live -= 1;
    w.lane_live[t & 31] = false;
    E.progress = true;
    // threads that have returned no longer take part in barriers
    if (E.live > 0 && E.arrived == E.live) release_cta_barrier();
    if (w.live > 0 && w.arrived == w.live) release_warp_barrier(w);
}

void trampoline()
{
    (*E.body)();
    fiber_exit_bookkeeping(E.cur);
#if CPUSIM_FAST_SWITCH
    _longjmp(E.sched_jb, 1);
#else
    swapcontext(&E.fib[E.cur].ctx, &E.sched);
#endif
    fail("resumed a finished fiber");
}

void run_cta(const std::function<void()> &body, int n_threads)
{
    if (!E.stacks) {
        E.stacks = (unsigned char *)mmap(nullptr, kStackBytes * kMaxThreads, PROT_READ | PROT_WRITE,
                                         MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (E.stacks == MAP_FAILED) fail("cannot map fiber stacks");
    }
    E.body = &body;
    E.n_threads = n_threads;
    E.live = n_threads;
    E.arrived = 0;
    E.or_acc = E.or_result = 0;
    for (int w = 0; w < (n_threads + 31) / 32; ++w) {
        WarpState &ws = E.warp[w];
        ws.live = std::min(32, n_threads - 32 * w);
        ws.arrived = 0;
        for (int l = 0; l < 32; ++l) ws.lane_live[l] = l < ws.live;
    }
    for (int t = 0; t < n_threads; ++t) {
        Fiber &f = E.fib[t];
        f.done = false;
        f.started = false;
        f.waiting_cta = f.waiting_warp = false;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = E.stacks + kStackBytes * (size_t)t;
        f.ctx.uc_stack.ss_size = kStackBytes;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, trampoline, 0);
    }
    // Scheduling order of the fibers: CPUSIM_SCHEDULE = forward (default) | reverse | random.  Correct CUDA code
    // gives the same results under every order; code that leans on an implicit order between barriers does not.
    static const int mode = []() {
        const char *m = std::getenv("CPUSIM_SCHEDULE");
        return !m ? 0 : (m[0] == 'r' && m[1] == 'e') ? 1 : (m[0] == 'r' && m[1] == 'a') ? 2 : 0;
    }();
    static int order[kMaxThreads];
    static unsigned long long rng = 0x2545F4914F6CDD1Dull;
    for (int t = 0; t < n_threads; ++t) order[t] = mode == 1 ? n_threads - 1 - t : t;
    while (E.live > 0) {
        E.progress = false;
        if (mode == 2)
            for (int t = n_threads - 1; t > 0; --t) {   // a fresh shuffle every round
                rng ^= rng << 13;
                rng ^= rng >> 7;
                rng ^= rng << 17;
                std::swap(order[t], order[rng % (unsigned)(t + 1)]);
            }
        for (int k = 0; k < n_threads; ++k) {
            const int t = order[k];
            Fiber &f = E.fib[t];
            if (f.done) continue;
            if (f.waiting_cta && f.cta_gen == E.gen) continue;
            if (f.waiting_warp && f.warp_gen == E.warp[t >> 5].gen) continue;
            f.waiting_cta = f.waiting_warp = false;
            E.cur = t;
            threadIdx.x = (unsigned)t % blockDim.x;
            threadIdx.y = ((unsigned)t / blockDim.x) % blockDim.y;
            threadIdx.z = (unsigned)t / (blockDim.x * blockDim.y);
            E.progress = true;
            resume_fiber(t);
        }
        if (!E.progress) fail("deadlock: every live thread of the CTA waits at a barrier that cannot complete");
    }
    E.cur = -1;
}

}  // namespace

unsigned char *dyn_smem() { return E.smem; }

void cta_barrier()
{
    if (!E.in_kernel) return;
    Fiber &f = E.fib[E.cur];
    E.arrived += 1;
    if (E.arrived == E.live) {
        release_cta_barrier();
        return;
    }
    f.waiting_cta = true;
    f.cta_gen = E.gen;
    yield_to_scheduler();
}

int cta_barrier_or(int pred)
{
    E.or_acc |= pred ? 1 : 0;
    cta_barrier();
    return E.or_result;   // valid until the next CTA barrier is released
}

void warp_barrier()
{
    if (!E.in_kernel) return;
    Fiber &f = E.fib[E.cur];
    WarpState &w = E.warp[E.cur >> 5];
    w.arrived += 1;
    if (w.arrived == w.live) {
        release_warp_barrier(w);
        return;
    }
    f.waiting_warp = true;
    f.warp_gen = w.gen;
    yield_to_scheduler();
}

unsigned long long warp_exchange(unsigned long long mine, int src_lane)
{
    const int t = E.cur;
    WarpState &w = E.warp[t >> 5];
    w.slot[t & 31] = mine;
    warp_barrier();
    unsigned long long got = mine;
    if (src_lane >= 0 && src_lane < 32 && w.lane_live[src_lane]) got = w.slot[src_lane];
    warp_barrier();   // nobody overwrites a slot before everybody has read
    return got;
}

unsigned warp_ballot(int pred)
{
    const int t = E.cur;
    WarpState &w = E.warp[t >> 5];
    w.pred[t & 31] = pred ? 1 : 0;
    warp_barrier();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l)
        if (w.lane_live[l] && w.pred[l]) m |= 1u << l;
    warp_barrier();
    return m;
}

void run_grid(const std::function<void()> &body, dim3 grid, dim3 block, size_t dyn_smem_bytes)
{
    const unsigned long long n = (unsigned long long)block.x * block.y * block.z;
    if (n == 0 || n > kMaxThreads) fail("block size out of range");
    if (dyn_smem_bytes > kDynSmemBytes) fail("dynamic shared memory request exceeds 227 KB");
    if (E.in_kernel) fail("nested launch");
    E.in_kernel = true;
    gridDim = grid;
    blockDim = block;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                blockIdx.x = x;
                blockIdx.y = y;
                blockIdx.z = z;
                run_cta(body, (int)n);
            }
    E.in_kernel = false;
}

// one engine: host threads of the product (e.g. sphb_ensemble_run_dual's launch threads) take turns
std::recursive_mutex &engine_mutex()
{
    static std::recursive_mutex m;
    return m;
}

void enqueue(Stream *st, std::function<void()> op)
{
    if (st && st->capturing) {
        st->graph->nodes.push_back(std::move(op));
        return;
    }
    std::lock_guard<std::recursive_mutex> lock(engine_mutex());
    op();
}

}  // namespace cpusim

// ---- runtime API -------------------------------------------------------------------------------------------------
using cpusim::enqueue;
namespace cpusim {
std::recursive_mutex &engine_mutex();
}

const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cpusim error"; }
namespace cpusim {
static thread_local cudaError_t g_launch_error = cudaSuccess;
void set_launch_error() { g_launch_error = cudaErrorInvalidValue; }
}  // namespace cpusim
cudaError_t cudaGetLastError()
{
    const cudaError_t e = cpusim::g_launch_error;
    cpusim::g_launch_error = cudaSuccess;
    return e;
}
cudaError_t cudaGetDeviceCount(int *n)
{
    *n = 1;
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *dev)
{
    *dev = 0;
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t bytes)
{
    // poisoned, so that reads of memory the product never initialised show up as garbage, as on the device
    *p = std::aligned_alloc(256, (bytes + 255) & ~(size_t)255);
    if (!*p) return cudaErrorMemoryAllocation;
    std::memset(*p, 0xCD, bytes);
    return cudaSuccess;
}
cudaError_t cudaFree(void *p)
{
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t bytes) { return cudaMalloc(p, bytes); }
cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind)
{
    std::memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t st)
{
    enqueue(st, [=]() { std::memmove(dst, src, bytes); });
    return cudaSuccess;
}
cudaError_t cudaMemset(void *dst, int value, size_t bytes)
{
    std::memset(dst, value, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t st)
{
    enqueue(st, [=]() { std::memset(dst, value, bytes); });
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned)
{
    *st = new cpusim::Stream();
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t st)
{
    delete st;
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t st) { return st && st->capturing ? cudaErrorUnknown : cudaSuccess; }
cudaError_t cudaStreamIsCapturing(cudaStream_t st, cudaStreamCaptureStatus *status)
{
    *status = st && st->capturing ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone;
    return cudaSuccess;
}
cudaError_t cudaStreamBeginCapture(cudaStream_t st, cudaStreamCaptureMode)
{
    if (!st || st->capturing) return cudaErrorInvalidValue;
    st->capturing = true;
    st->graph = new cpusim::Graph();
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t st, cudaGraph_t *graph)
{
    if (!st || !st->capturing) return cudaErrorInvalidValue;
    st->capturing = false;
    *graph = st->graph;
    st->graph = nullptr;
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *exec, cudaGraph_t graph, unsigned long long)
{
    *exec = new cpusim::Graph(*graph);
    return cudaSuccess;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t exec, cudaStream_t st)
{
    if (st && st->capturing) return cudaErrorInvalidValue;
    std::lock_guard<std::recursive_mutex> lock(cpusim::engine_mutex());
    for (auto &node : exec->nodes) node();
    return cudaSuccess;
}
cudaError_t cudaGraphDestroy(cudaGraph_t graph)
{
    delete graph;
    return cudaSuccess;
}
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t exec)
{
    delete exec;
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e)
{
    *e = new cpusim::Event();
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e)
{
    delete e;
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st)
{
    if (st && st->capturing) return cudaSuccess;
    e->t_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = (float)(b->t_ms - a->t_ms);
    return cudaSuccess;
}
