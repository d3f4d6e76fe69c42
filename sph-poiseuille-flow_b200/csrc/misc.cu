// misc.cu -- error plumbing, device buffers, layout conversion kernels.
#include "common.cuh"

#include <cstdarg>

namespace sphb {

namespace {
thread_local char g_err_id[128] = "";
thread_local char g_err_msg[512] = "";
}  // namespace

int set_error(const char *id, const char *fmt, ...)
{
    va_list ap;
    snprintf(g_err_id, sizeof(g_err_id), "%s", id ? id : "SPHB:error");
    va_start(ap, fmt);
    vsnprintf(g_err_msg, sizeof(g_err_msg), fmt ? fmt : "", ap);
    va_end(ap);
    return 1;
}

void clear_error()
{
    g_err_id[0] = '\0';
    g_err_msg[0] = '\0';
}

const char *last_error_id() { return g_err_id; }
const char *last_error_msg() { return g_err_msg; }

int ensure_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return set_error("SPHB:cuda", "no usable CUDA device (%s); libsph_b200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    return 0;
}

const char *prof_name(int id)
{
    static const char *names[kPfCount] = {"nbr_grid", "nbr_bin_count", "nbr_scan", "nbr_scatter", "nbr_cell_canon",
                                          "nbr_walk", "density", "kgc",
                                          "visc_transport", "time_step", "pressure", "continuity", "clock",
                                          "wall_shear", "cell_sort", "force_fused", "slab_exchange"};
    return (id >= 0 && id < kPfCount) ? names[id] : "?";
}

size_t Prof::grab()
{
    if (used == pool.size()) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        pool.push_back(e);
    }
    return used++;
}
void Prof::begin(int id, cudaStream_t st)
{
    Rec r;
    r.id = id;
    r.e0 = grab();
    r.e1 = grab();
    cudaEventRecord(pool[r.e0], st);
    recs.push_back(r);
}
void Prof::end(cudaStream_t st)
{
    if (!recs.empty()) cudaEventRecord(pool[recs.back().e1], st);
}
int Prof::collect(cudaStream_t st)
{
    SPHB_CUDA(cudaStreamSynchronize(st));
    for (const Rec &r : recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, pool[r.e0], pool[r.e1]) == cudaSuccess) {
            ms[r.id] += (double)t;
            n[r.id] += 1;
        } else {
            (void)cudaGetLastError();
        }
    }
    recs.clear();
    used = 0;
    return 0;
}
void Prof::reset()
{
    for (int k = 0; k < kPfCount; ++k) {
        ms[k] = 0.0;
        n[k] = 0;
    }
    recs.clear();
    used = 0;
}
void Prof::release()
{
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
    pool.clear();
    recs.clear();
    used = 0;
}

int DevBuf::reserve(size_t need)
{
    if (need <= bytes && p) return 0;
    if (need == 0) need = 16;
    if (p) {
        SPHB_CUDA(cudaFree(p));
        p = nullptr;
        bytes = 0;
    }
    SPHB_CUDA(cudaMalloc(&p, need));
    bytes = need;
    return 0;
}

void DevBuf::release()
{
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
}

namespace {

__global__ void k_pack2(const double *__restrict__ cm, double2 *__restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double2(cm[i], cm[i + n]);
}
__global__ void k_unpack2(const double2 *__restrict__ in, double *__restrict__ cm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double2 v = in[i];
        cm[i] = v.x;
        cm[i + n] = v.y;
    }
}
__global__ void k_pack4(const double *__restrict__ cm, Mat2 *__restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        Mat2 m;
        m.b11 = cm[i];
        m.b12 = cm[i + n];
        m.b21 = cm[i + 2 * (size_t)n];
        m.b22 = cm[i + 3 * (size_t)n];
        out[i] = m;
    }
}
__global__ void k_unpack4(const Mat2 *__restrict__ in, double *__restrict__ cm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const Mat2 m = in[i];
        cm[i] = m.b11;
        cm[i + n] = m.b12;
        cm[i + 2 * (size_t)n] = m.b21;
        cm[i + 3 * (size_t)n] = m.b22;
    }
}

}  // namespace

void launch_pack2(const double *cm, double2 *out, int n, cudaStream_t st)
{
    if (n > 0) k_pack2<<<grid_for(n), kBlock, 0, st>>>(cm, out, n);
}
void launch_unpack2(const double2 *in, double *cm, int n, cudaStream_t st)
{
    if (n > 0) k_unpack2<<<grid_for(n), kBlock, 0, st>>>(in, cm, n);
}
void launch_pack4(const double *cm, Mat2 *out, int n, cudaStream_t st)
{
    if (n > 0) k_pack4<<<grid_for(n), kBlock, 0, st>>>(cm, out, n);
}
void launch_unpack4(const Mat2 *in, double *cm, int n, cudaStream_t st)
{
    if (n > 0) k_unpack4<<<grid_for(n), kBlock, 0, st>>>(in, cm, n);
}

}  // namespace sphb

extern "C" {
SPHB_API const char *sphb_last_error_id(void) { return sphb::last_error_id(); }
SPHB_API const char *sphb_last_error_msg(void) { return sphb::last_error_msg(); }
SPHB_API int sphb_version(void) { return SPHB_VERSION; }
SPHB_API int sphb_device_check(void)
{
    sphb::clear_error();
    return sphb::ensure_device();
}
}
