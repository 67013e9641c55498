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
