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
double PhaseTimer::now()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
PhaseTimer::PhaseTimer(const char *w) : on(getenv("GLB_TIMING") != nullptr), what(w), t0(now()) {}
void PhaseTimer::lap(const char *phase)
{
    if (!on) return;
    cudaDeviceSynchronize();
    const double t = now();
    fprintf(stderr, "[glb timing] %s: %-28s %8.3f ms\n", what, phase, (t - t0) * 1e3);
    t0 = t;
}
}  // namespace glb

extern "C" GLB_API int glb_version(void) { return 100; }
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
