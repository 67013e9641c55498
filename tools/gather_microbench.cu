// gather_microbench.cu - how many random 64-byte rows of an L2-resident table can one SM fetch per cycle?
//
// The Poisson iterate u <- Db + P u on a kNN graph is ~1M random row gathers of the n x 16 fp32 label matrix per
// iteration (one per nonzero).  This tool measures the per-SM ceiling of each path that can bring such a row
// on chip, to decide how poisson.cu should be built:
//   ldg4      4 lanes x LDG.128 per row, 8 rows per warp instruction (what poisson_persistent_kernel does)
//   ldg4s     same with ld.relaxed.gpu (L1 bypass, what a flag-in-data kernel needs)
//   ldg8      8 lanes x LDG.64 per row
//   ldg2      2 lanes x 256-bit loads per row
//   ldg1row   only one row per warp instruction (lanes 0-3 active): cost of one L1 wavefront per instruction
//   bulk      cp.async.bulk (TMA unit, 1-D) 64 B per lane into a per-warp shared-memory ring, then LDS
//   gather4   cp.async.bulk.tensor tile::gather4 (4 rows per instruction) into the ring, then LDS
//   mix       half of the warps run ldg4s, the other half bulk
//   lds       rows already resident in shared memory (upper bound of a staged design)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gather_microbench tools/gather_microbench.cu
// run:   ./gather_microbench <mode> [threads] [stages]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int NIDX = 8192;          // gathers per pass per CTA (6800 in the 70k graph)
constexpr int ROWF = 16;            // floats per row

__device__ __forceinline__ unsigned hash32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float4 ld_strong4(const float *p)
{
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

enum Mode { LDG4 = 0, LDG4S, LDG8, LDG2, LDG1ROW, BULK, GATHER4, MIX, LDS };

struct Params {
    const float *table;
    int nrows;
    int reps;
    float *out;
    long long *cycles;
    int stages;
};

// ---------------------------------------------------------------------------------------------------------
template <int LANES, bool STRONG, bool ONEROW>
__device__ __forceinline__ float ldg_pass(const float *table, const int *s_idx, int tid, int nthreads)
{
    constexpr int U = 4;
    const int li = tid % LANES;
    const int g = tid / LANES;
    const int ng = nthreads / LANES;
    float acc = 0.f;
    if (ONEROW) {
        // one row per warp instruction: lanes 0..3 of each warp active
        const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5;
        if (lane < 4) {
            for (int b = warp; b + nw * (U - 1) < NIDX; b += nw * U) {
                float4 x[U];
#pragma unroll
                for (int u = 0; u < U; ++u) x[u] = __ldg(reinterpret_cast<const float4 *>(table + (size_t)s_idx[b + u * nw] * ROWF) + lane);
#pragma unroll
                for (int u = 0; u < U; ++u) acc += x[u].x + x[u].y + x[u].z + x[u].w;
            }
        }
        return acc;
    }
    for (int b = g; b + ng * (U - 1) < NIDX; b += ng * U) {
        if (LANES == 4) {
            float4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float *p = table + (size_t)s_idx[b + u * ng] * ROWF + li * 4;
                x[u] = STRONG ? ld_strong4(p) : __ldg(reinterpret_cast<const float4 *>(p));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += x[u].x + x[u].y + x[u].z + x[u].w;
        } else if (LANES == 8) {
            float2 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = __ldg(reinterpret_cast<const float2 *>(table + (size_t)s_idx[b + u * ng] * ROWF) + li);
#pragma unroll
            for (int u = 0; u < U; ++u) acc += x[u].x + x[u].y;
        } else {   // LANES == 2: 256-bit loads
            float v[U][8];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float *p = table + (size_t)s_idx[b + u * ng] * ROWF + li * 8;
                asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]), "=f"(v[u][4]), "=f"(v[u][5]), "=f"(v[u][6]), "=f"(v[u][7])
                             : "l"(p) : "memory");
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc += v[u][q];
        }
    }
    return acc;
}

constexpr int SLOT = 80;            // bytes between row slots of a ring stage (64 B row + 16 B pad: conflict-free LDS.128)

// ring fetch through the TMA unit; warp-private ring of `stages` stages x 32 rows
template <bool G4>
__device__ __forceinline__ float tma_pass(const Params &p, const CUtensorMap *map, const int *s_idx, unsigned char *ring, uint64_t *bars,
                                          int warp_slot, int nwarps_tma, int lane, unsigned &phase_bits, int stages)
{
    // rows of this warp: s_idx[warp_slot*32 + lane + k * nwarps_tma*32]
    float acc = 0.f;
    const int stride = nwarps_tma * 32;
    const int nbatch = NIDX / stride;
    unsigned char *wring = ring + (size_t)warp_slot * stages * 32 * SLOT;
    uint64_t *wbar = bars + warp_slot * stages;
    auto issue = [&](int k) {
        const int s = k % stages;
        if (lane == 0) mbar_expect_tx(&wbar[s], 32 * 64);
        __syncwarp();
        const int base = warp_slot * 32 + k * stride;
        if (G4) {
            if (lane < 8) {
                const int4 r = *reinterpret_cast<const int4 *>(&s_idx[base + lane * 4]);
                tma_gather4(wring + (size_t)s * 32 * SLOT + lane * 256, map, 0, r.x, r.y, r.z, r.w, &wbar[s]);
            }
        } else {
            const int r = s_idx[base + lane];
            bulk_g2s(wring + (size_t)s * 32 * SLOT + lane * SLOT, p.table + (size_t)r * ROWF, 64, &wbar[s]);
        }
    };
    const int pre = stages - 1 < nbatch ? stages - 1 : nbatch;
    for (int k = 0; k < pre; ++k) issue(k);
    for (int k = 0; k < nbatch; ++k) {
        if (k + pre < nbatch) issue(k + pre);
        const int s = k % stages;
        mbar_wait(&wbar[s], (phase_bits >> s) & 1u);
        phase_bits ^= 1u << s;
        const unsigned char *row = wring + (size_t)s * 32 * SLOT + (G4 ? lane * 64 : lane * SLOT);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 x = *reinterpret_cast<const float4 *>(row + q * 16);
            acc += x.x + x.y + x.z + x.w;
        }
        __syncwarp();       // every lane has read the stage before it is refilled
    }
    return acc;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench_kernel(Params p, const __grid_constant__ CUtensorMap map)
{
    extern __shared__ __align__(128) unsigned char smem[];
    int *s_idx = reinterpret_cast<int *>(smem);                               // NIDX ints = 32 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + NIDX * 4);           // up to 32*8 barriers = 2 KB
    unsigned char *ring = smem + NIDX * 4 + 2048;
    const int tid = threadIdx.x, nthreads = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    for (int i = tid; i < NIDX; i += nthreads) s_idx[i] = (int)(hash32(i * 2654435761u + blockIdx.x * 97u + 12345u) % (unsigned)p.nrows);
    if (MODE == BULK || MODE == GATHER4 || MODE == MIX) {
        if (tid < nwarps * p.stages) mbar_init(&bars[tid], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (MODE == LDS) {
        // table slice resident in shared memory: 2048 rows x 64 B = 128 KB
        float4 *s_tab = reinterpret_cast<float4 *>(ring);
        for (int i = tid; i < 2048 * 4; i += nthreads) s_tab[i] = reinterpret_cast<const float4 *>(p.table)[i];
    }
    __syncthreads();
    unsigned phase_bits = 0;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep) {
        if (MODE == LDG4) acc += ldg_pass<4, false, false>(p.table, s_idx, tid, nthreads);
        if (MODE == LDG4S) acc += ldg_pass<4, true, false>(p.table, s_idx, tid, nthreads);
        if (MODE == LDG8) acc += ldg_pass<8, false, false>(p.table, s_idx, tid, nthreads);
        if (MODE == LDG2) acc += ldg_pass<2, false, false>(p.table, s_idx, tid, nthreads);
        if (MODE == LDG1ROW) acc += ldg_pass<4, false, true>(p.table, s_idx, tid, nthreads);
        if (MODE == BULK) acc += tma_pass<false>(p, &map, s_idx, ring, bars, warp, nwarps, lane, phase_bits, p.stages);
        if (MODE == GATHER4) acc += tma_pass<true>(p, &map, s_idx, ring, bars, warp, nwarps, lane, phase_bits, p.stages);
        if (MODE == MIX) {
            // first half of the index list by LDG warps, second half by TMA warps (each half NIDX/2 rows)
            const int half = nwarps / 2;
            if (warp < half) acc += ldg_pass<4, true, false>(p.table, s_idx, tid, nthreads / 2) * 0.5f;   // rows [0, NIDX) strided: see note in main
            else acc += tma_pass<false>(p, &map, s_idx, ring, bars, warp - half, half, lane, phase_bits, p.stages);
        }
        if (MODE == LDS) {
            const float4 *s_tab = reinterpret_cast<const float4 *>(ring);
            const int li = tid & 3, g = tid >> 2, ng = nthreads >> 2;
            for (int b = g; b < NIDX; b += ng) {
                const float4 x = s_tab[(s_idx[b] & 2047) * 4 + li];
                acc += x.x + x.y + x.z + x.w;
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) p.cycles[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) p.out[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    const char *names[] = {"ldg4", "ldg4s", "ldg8", "ldg2", "ldg1row", "bulk", "gather4", "mix", "lds"};
    if (argc < 2) { printf("usage: %s <mode> [threads] [stages]\n", argv[0]); return 1; }
    int mode = -1;
    for (int i = 0; i < 9; ++i) if (!strcmp(argv[1], names[i])) mode = i;
    if (mode < 0) { printf("unknown mode\n"); return 1; }
    int threads = argc > 2 ? atoi(argv[2]) : 1024;
    int stages = argc > 3 ? atoi(argv[3]) : 4;
    const int nrows = 70000, reps = 200;
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *table, *out;
    long long *cycles;
    CK(cudaMalloc(&table, (size_t)nrows * ROWF * 4));
    CK(cudaMemset(table, 0, (size_t)nrows * ROWF * 4));
    CK(cudaMalloc(&out, 16));
    CK(cudaMalloc(&cycles, sizeof(long long) * sms));

    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (mode == GATHER4) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 3; }
        cuuint64_t gdim[2] = {ROWF, (cuuint64_t)nrows};
        cuuint64_t gstr[1] = {ROWF * 4};
        cuuint32_t box[2] = {ROWF, 1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, table, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 3; }
    }
    const int nwarps = threads / 32;
    size_t smem = NIDX * 4 + 2048;
    if (mode == BULK || mode == GATHER4 || mode == MIX) smem += (size_t)nwarps * stages * 32 * SLOT;
    if (mode == LDS) smem += 2048 * 64;
    Params p{table, nrows, reps, out, cycles, stages};
    void (*kern)(Params, const CUtensorMap) = nullptr;
    switch (mode) {
        case LDG4: kern = bench_kernel<LDG4>; break;
        case LDG4S: kern = bench_kernel<LDG4S>; break;
        case LDG8: kern = bench_kernel<LDG8>; break;
        case LDG2: kern = bench_kernel<LDG2>; break;
        case LDG1ROW: kern = bench_kernel<LDG1ROW>; break;
        case BULK: kern = bench_kernel<BULK>; break;
        case GATHER4: kern = bench_kernel<GATHER4>; break;
        case MIX: kern = bench_kernel<MIX>; break;
        case LDS: kern = bench_kernel<LDS>; break;
    }
    if (smem > 227 * 1024) { printf("smem %zu too large\n", smem); return 1; }
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<<<sms, threads, smem>>>(p, map);          // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    kern<<<sms, threads, smem>>>(p, map);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long *h = (long long *)malloc(sizeof(long long) * sms);
    CK(cudaMemcpy(h, cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    long long mx = 0; double avg = 0;
    for (int i = 0; i < sms; ++i) { if (h[i] > mx) mx = h[i]; avg += (double)h[i] / sms; }
    // rows actually fetched per pass per CTA
    double rows = NIDX;
    if (mode == LDG1ROW) rows = (double)(NIDX / (nwarps * 4)) * (nwarps * 4);
    if (mode == MIX) rows = NIDX * 1.5;       // LDG half walks the whole list with half the groups (NIDX rows), TMA half NIDX/2... see below
    if (mode == MIX) {
        // ldg_pass over nthreads/2 threads covers all NIDX rows; tma_pass with `half` warps covers all NIDX rows too
        rows = 2.0 * NIDX;
    }
    const double cyc_per_row = avg / reps / rows;
    printf("%-8s threads=%d stages=%d  %.3f ms  avg %.0f cyc/pass (max %.0f)  %.3f cyc/row/SM  -> 6800 rows = %.2f us @1.94GHz  | chip %.2f Grows/s %.2f TB/s\n",
           names[mode], threads, stages, ms, avg / reps, (double)mx / reps, cyc_per_row, 6800 * cyc_per_row / 1940.0,
           rows * reps * sms / (ms * 1e-3) / 1e9, rows * reps * sms * 64.0 / (ms * 1e-3) / 1e12);
    return 0;
}
