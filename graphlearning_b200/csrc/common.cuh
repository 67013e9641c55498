// common.cuh - error plumbing and small device helpers shared by every translation unit of libglb200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/glb200.h"

namespace glb {

void set_error(const char *fmt, ...);

#define GLB_CHECK_ARG(cond, msg)                                          \
    do {                                                                  \
        if (!(cond)) {                                                    \
            glb::set_error("%s: %s", __func__, msg);                      \
            return GLB_E_INVALID;                                         \
        }                                                                 \
    } while (0)

#define GLB_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            glb::set_error("%s: %s failed: %s", __func__, #expr, cudaGetErrorString(e__));    \
            return (int)e__;                                                                  \
        }                                                                                     \
    } while (0)

#define GLB_LAUNCH_CHECK()                                                                    \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) {                                                             \
            glb::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__)); \
            return (int)e__;                                                                  \
        }                                                                                     \
    } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

int sm_count();   // cached, current device

// Device memory of the library's own scratch and plans comes from the CUDA stream-ordered pool (cudaMallocAsync) with the
// release threshold lifted: a freed block stays mapped, so the next plan / search does not pay the driver's map and unmap
// again (53 of the 96 ms of a first fit and 68 of the 83 ms of a kNN search went there, profiles/r2_first_fit_timing.txt).
// Blocks are usable on every stream once dev_alloc has returned; dev_free may be called when the work that uses the block
// has been synchronised.  glb_release_workspace() hands the cached blocks back to the driver.
cudaError_t dev_alloc_bytes(void **p, size_t bytes);
cudaError_t dev_free(void *p);
template <typename T>
static inline cudaError_t dev_alloc(T **p, size_t bytes) { return dev_alloc_bytes(reinterpret_cast<void **>(p), bytes); }

// GLB_TIMING=1: wall-clock phases of the host entry points on stderr (setup-cost experiments)
struct PhaseTimer {
    bool on;
    const char *what;
    double t0;
    static double now();
    explicit PhaseTimer(const char *w);
    void lap(const char *phase);
};

}  // namespace glb
