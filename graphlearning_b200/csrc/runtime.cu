// runtime.cu - version / error string / device info for the C-ABI (include/glb200.h)
#include <stdarg.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>
#include "common.cuh"

namespace glb {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int sm_count()
{
    static int cached = 0;
    if (!cached) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        if (cached <= 0) cached = 148;
    }
    return cached;
}
static bool pool_ready()
{
    static int state = 0;                                   // 0 unknown, 1 pooled, -1 plain cudaMalloc
    if (state == 0) {
        int dev = 0, ok = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&ok, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess && ok &&
            cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;                // never trim on synchronisation
            state = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess ? 1 : -1;
        } else {
            state = -1;
        }
        cudaGetLastError();
    }
    return state == 1;
}
cudaError_t dev_alloc_bytes(void **p, size_t bytes)
{
    if (!bytes) bytes = 1;
    if (pool_ready()) {
        cudaError_t e = cudaMallocAsync(p, bytes, (cudaStream_t)0);
        if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)0);      // usable on every stream from here on
        return e;
    }
    return cudaMalloc(p, bytes);
}
cudaError_t dev_free(void *p)
{
    if (!p) return cudaSuccess;
    if (pool_ready()) return cudaFreeAsync(p, (cudaStream_t)0);
    return cudaFree(p);
}
double PhaseTimer::now()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
static bool phase_timing_enabled()
{
    static const bool on = getenv("GLB_TIMING") != nullptr;       // read once per process
    return on;
}
PhaseTimer::PhaseTimer(const char *w) : on(phase_timing_enabled()), what(w), t0(now()) {}
void PhaseTimer::lap(const char *phase)
{
    if (!on) return;
    cudaDeviceSynchronize();
    const double t = now();
    fprintf(stderr, "[glb timing] %s: %-28s %8.3f ms\n", what, phase, (t - t0) * 1e3);
    t0 = t;
}
}  // namespace glb

extern "C" GLB_API int glb_version(void) { return 200; }

extern "C" GLB_API int glb_release_workspace(void)
{
    int dev = 0;
    cudaMemPool_t pool;
    GLB_CUDA(cudaDeviceSynchronize());
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    cudaGetLastError();
    return 0;
}
extern "C" GLB_API const char *glb_last_error(void) { return glb::g_err; }

extern "C" GLB_API int glb_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *hbm_bytes)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        glb::set_error("glb_device_info: no CUDA device visible");
        return GLB_E_NOGPU;
    }
    int dev = 0;
    GLB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    GLB_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (hbm_bytes) *hbm_bytes = (int64_t)p.totalGlobalMem;
    return 0;
}

extern "C" GLB_API int glb_padded_ld(int c)
{
    if (c <= 0) return GLB_E_INVALID;
    int ld = 4;
    while (ld < c && ld < 128) ld <<= 1;
    if (ld < c) ld = ((c + 127) / 128) * 128;
    return ld;
}
