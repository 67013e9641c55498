// pingpong_microbench.cu - latency of "store a 16-byte flagged chunk on one SM, see it from another SM".
// CTA 0 and CTA 1 (one warp each, different SMs) bounce a counter through two chunks in global memory.
// Reported: nanoseconds per one-way hop (store -> remote poll sees it) for each store/load flavour.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pingpong tools/pingpong_microbench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

enum St { ST_RELAXED = 0, ST_VOLATILE, ST_FENCE, ST_RELEASE, ST_ATOM, ST_WT, ST_PLAIN };
enum Ld { LD_RELAXED = 0, LD_VOLATILE, LD_ACQUIRE, LD_CG };

template <int ST>
__device__ __forceinline__ void store_chunk(unsigned *p, unsigned v)
{
    if (ST == ST_RELAXED) asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    if (ST == ST_VOLATILE) asm volatile("st.volatile.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    if (ST == ST_FENCE) {
        asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    if (ST == ST_RELEASE) asm volatile("st.release.gpu.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    if (ST == ST_ATOM) {
        asm volatile("{\n.reg .b128 a, b;\nmov.b128 a, {%1,%1,%1,%1};\natom.relaxed.gpu.global.exch.b128 b, [%0], a;\n}\n" ::"l"(p), "r"(v) : "memory");
    }
    if (ST == ST_WT) asm volatile("st.global.wt.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
    if (ST == ST_PLAIN) asm volatile("st.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
template <int LD>
__device__ __forceinline__ unsigned load_chunk(const unsigned *p)
{
    unsigned a, b, c, d;
    if (LD == LD_RELAXED) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    if (LD == LD_VOLATILE) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    if (LD == LD_ACQUIRE) asm volatile("ld.acquire.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    if (LD == LD_CG) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    return d;
}

// nwarps warps per CTA take part (each its own pair of chunks): shows whether more traffic changes the latency
template <int ST, int LD>
__global__ void pingpong(unsigned *buf, int rounds, long long *cycles, int stride_words)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned *mine = buf + (size_t)(warp * 2 + blockIdx.x) * stride_words;          // I write here
    const unsigned *theirs = buf + (size_t)(warp * 2 + (blockIdx.x ^ 1)) * stride_words;   // I poll here
    const long long t0 = clock64();
    if (lane == 0) {
        for (int r = 1; r <= rounds; ++r) {
            if (blockIdx.x == 0) {
                store_chunk<ST>(mine, (unsigned)r);
                while (load_chunk<LD>(theirs) != (unsigned)r) { }
            } else {
                while (load_chunk<LD>(theirs) != (unsigned)r) { }
                store_chunk<ST>(mine, (unsigned)r);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// full-warp version on a full grid: CTA 0 and CTA `partner` bounce 8 rows x 4 chunks (the dataflow kernel's access shape:
// 32 lanes, 8 different 64-byte rows per instruction); every other CTA exits at once.
template <int ST, int LD>
__global__ void pingpong_warp(unsigned *buf, int rounds, int partner, int row_stride_words, unsigned *smid)
{
    if ((blockIdx.x != 0 && blockIdx.x != partner) || threadIdx.x >= 32) return;
    const int side = blockIdx.x == 0 ? 0 : 1;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, li = lane & 3;
    if (threadIdx.x == 0) { unsigned id; asm volatile("mov.u32 %0, %%smid;" : "=r"(id)); smid[side] = id; }
    unsigned *mine = buf + (size_t)(side * 8 + g) * row_stride_words + li * 4;
    const unsigned *theirs = buf + (size_t)((side ^ 1) * 8 + g) * row_stride_words + li * 4;
    for (int r = 1; r <= rounds; ++r) {
        if (side == 0) {
            store_chunk<ST>(mine, (unsigned)r);
            while (load_chunk<LD>(theirs) != (unsigned)r) { }
        } else {
            while (load_chunk<LD>(theirs) != (unsigned)r) { }
            store_chunk<ST>(mine, (unsigned)r);
        }
        __syncwarp();
    }
}

template <int ST, int LD>
static void run_warp(const char *name, int partner)
{
    unsigned *buf, *smid;
    const int row_stride_words = 16 * 37;       // rows scattered over different lines / slices
    CK(cudaMalloc(&buf, 16 * row_stride_words * 4));
    CK(cudaMalloc(&smid, 8));
    const int rounds = 2000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    void *args[] = {&buf, (void *)&rounds, (void *)&partner, (void *)&row_stride_words, &smid};
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(buf, 0, 16 * row_stride_words * 4));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((const void *)pingpong_warp<ST, LD>, dim3(148), dim3(1024), args, 0, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned h[2];
    CK(cudaMemcpy(h, smid, 8, cudaMemcpyDeviceToHost));
    printf("warp 8x64B  %-28s CTA 0 (sm %3u) <-> CTA %3d (sm %3u)  %.1f ns per one-way hop\n", name, h[0], partner, h[1], ms * 1e6 / rounds / 2);
    fflush(stdout);
    CK(cudaFree(buf)); CK(cudaFree(smid));
}

template <int ST, int LD>
static void run(const char *name, int nwarps, int smpair)
{
    unsigned *buf;
    long long *cycles;
    const int stride_words = 64;                // 256 bytes between chunks
    CK(cudaMalloc(&buf, 64 * 2 * stride_words * 4));
    CK(cudaMemset(buf, 0, 64 * 2 * stride_words * 4));
    CK(cudaMalloc(&cycles, 16));
    const int rounds = 2000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    void *args[] = {&buf, (void *)&rounds, &cycles, (void *)&stride_words};
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(buf, 0, 64 * 2 * stride_words * 4));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((const void *)pingpong<ST, LD>, dim3(2), dim3(32 * nwarps), args, 0, 0));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-28s warps=%2d  %.1f ns per one-way hop\n", name, nwarps, ms * 1e6 / rounds / 2);
    CK(cudaFree(buf)); CK(cudaFree(cycles));
    (void)smpair;
}

int main()
{
    CK(cudaSetDevice(0));
    for (int partner : {1, 2, 3, 8, 37, 74, 100, 147}) {
        run_warp<ST_RELAXED, LD_RELAXED>("st.relaxed / ld.relaxed", partner);
    }
    run_warp<ST_VOLATILE, LD_VOLATILE>("st.volatile / ld.volatile", 74);
    run_warp<ST_FENCE, LD_RELAXED>("st.relaxed+fence / ld.relaxed", 74);
    run_warp<ST_ATOM, LD_RELAXED>("atom.exch.b128 / ld.relaxed", 74);
    run_warp<ST_PLAIN, LD_RELAXED>("st plain / ld.relaxed", 74);
    for (int nw : {1, 8, 32}) {
        run<ST_RELAXED, LD_RELAXED>("st.relaxed / ld.relaxed", nw, 0);
        run<ST_VOLATILE, LD_VOLATILE>("st.volatile / ld.volatile", nw, 0);
        run<ST_FENCE, LD_RELAXED>("st.relaxed+fence / ld.relaxed", nw, 0);
        run<ST_RELEASE, LD_ACQUIRE>("st.release / ld.acquire", nw, 0);
        run<ST_ATOM, LD_RELAXED>("atom.exch.b128 / ld.relaxed", nw, 0);
        run<ST_WT, LD_CG>("st.wt / ld.cg", nw, 0);
        run<ST_PLAIN, LD_CG>("st plain / ld.cg", nw, 0);
        run<ST_PLAIN, LD_RELAXED>("st plain / ld.relaxed", nw, 0);
    }
    return 0;
}
