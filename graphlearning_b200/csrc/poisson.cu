// poisson.cu - the Poisson-learning iterate u <- Db + P u on sm_100a.
//
// Replaces the hot loop of ssl.poisson._fit, gradient-descent branch
// (reference graphlearning/ssl.py:667-669; third-party arithmetic: scipy _sparsetools csr_matvecs).
//
// Two kernels:
//   poisson_step_kernel        one iteration per launch.  CSR (col,val) streamed coalesced, rows of the
//                              row-major n x ldu label matrix gathered with 128-bit loads, D^-1 folded into
//                              the values, "+ Db" fused into the store.  For graphs too big for the
//                              persistent kernel; genuinely HBM/L2-gather bound.
//   poisson_persistent_kernel  T iterations in ONE cooperative launch, one CTA per SM.  Each CTA stages the
//                              CSR slab and Db slab of its row block in shared memory once, so per iteration
//                              only the u gathers (L2) and the u store touch global memory; iterations are
//                              separated by a hand-rolled grid barrier.  For the 70k-node north-star
//                              graph the whole working set is L2 resident and a launch per iteration
//                              (~2-3 us) would cost as much as the iteration itself.
//
// Lane mapping (both kernels): LANES = ldu/4 lanes own one matrix row; lane li holds the float4 of output
// columns [4*li, 4*li+4) (for ldu > 128 it loops over column tiles).  One warp-wide LDG.128 therefore
// gathers 32/LANES complete rows of u, each row = one 64-byte (ldu=16) piece of a single 128-byte line,
// which is what the L1TEX tag stage likes (one tag look-up per gathered row).  Each lane walks the
// nonzeros of its row UNROLL at a time: all UNROLL gathers are issued before the first FMA so that
// every thread keeps UNROLL L2 requests in flight; there is no cross-lane reduction at all.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace glb {

constexpr int kUnroll = 8;      // nonzeros in flight per lane in the one-launch-per-iteration kernel

__device__ __forceinline__ void fma4(float4 &acc, float a, const float4 &x)
{
    acc.x = fmaf(a, x.x, acc.x);
    acc.y = fmaf(a, x.y, acc.y);
    acc.z = fmaf(a, x.z, acc.z);
    acc.w = fmaf(a, x.w, acc.w);
}

template <bool NC>
__device__ __forceinline__ float4 load_u4(const float *p)
{
    if (NC) return __ldg(reinterpret_cast<const float4 *>(p));
    return *reinterpret_cast<const float4 *>(p);   // coherent at L1 after the grid barrier's fence
}

// Nonzero j as (byte offset of row col[j] inside u, value).  Offsets are 32-bit: n * ldu * 4 < 2^32 is checked
// on the host.  The persistent kernel keeps the pairs in shared memory, precomputed once per launch.
struct CsrGlobal {
    const int *__restrict__ col;
    const float *__restrict__ val;
    unsigned row_bytes;
    __device__ __forceinline__ void get(int j, unsigned &off, float &a) const
    {
        off = (unsigned)__ldg(col + j) * row_bytes;
        a = __ldg(val + j);
    }
};
struct CsrShared {
    const int2 *cv;
    __device__ __forceinline__ void get(int j, unsigned &off, float &a) const
    {
        const int2 e = cv[j];
        off = (unsigned)e.x;
        a = __int_as_float(e.y);
    }
};

// sum_j val[j] * u[col[j], 4 columns] over the nonzeros [beg, end) of one row; ubase = u + column offset.
// Full batches of UNROLL run without predicates (all gathers issued before the first FMA); the tail is predicated.
template <bool NC, int UNROLL, typename Csr>
__device__ __forceinline__ float4 row_times_u(const Csr &csr, int beg, int end, const char *__restrict__ ubase)
{
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = beg;
    for (; j + UNROLL <= end; j += UNROLL) {
        unsigned off[UNROLL];
        float a[UNROLL];
        float4 x[UNROLL];
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) csr.get(j + i, off[i], a[i]);
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) x[i] = load_u4<NC>(reinterpret_cast<const float *>(ubase + off[i]));
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) fma4(acc, a[i], x[i]);
    }
    if (j < end) {
        unsigned off[UNROLL];
        float a[UNROLL];
        float4 x[UNROLL];
#pragma unroll
        for (int i = 0; i < UNROLL - 1; ++i) {
            off[i] = 0; a[i] = 0.f;
            if (j + i < end) csr.get(j + i, off[i], a[i]);
        }
#pragma unroll
        for (int i = 0; i < UNROLL - 1; ++i) {
            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j + i < end) x[i] = load_u4<NC>(reinterpret_cast<const float *>(ubase + off[i]));
        }
#pragma unroll
        for (int i = 0; i < UNROLL - 1; ++i) fma4(acc, a[i], x[i]);      // a = 0, x = 0 past the end of the row
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// K1: one iteration per launch, CSR read from global memory
// ------------------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(256)
poisson_step_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const float *__restrict__ val,
                    const float *__restrict__ Db, const float *__restrict__ u_in, float *__restrict__ u_out,
                    int n, int ldu)
{
    const int li = threadIdx.x % LANES;
    const long long rid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const long long nrid = ((long long)gridDim.x * blockDim.x) / LANES;
    const CsrGlobal csr{col, val, (unsigned)ldu * 4u};
    for (long long row = rid; row < n; row += nrid) {
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        for (int coff = li * 4; coff < ldu; coff += LANES * 4) {
            float4 acc = row_times_u<true, kUnroll>(csr, beg, end, reinterpret_cast<const char *>(u_in + coff));
            const size_t o = (size_t)row * ldu + coff;
            const float4 b = __ldg(reinterpret_cast<const float4 *>(Db + o));
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
            *reinterpret_cast<float4 *>(u_out + o) = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2: persistent, T iterations per launch
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_add(unsigned *p, unsigned v)
{
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(unsigned *p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Grid-wide barrier between iterations (all CTAs are co-resident: cooperative launch, one per SM).
// bar.sync orders the CTA's u stores before thread 0's release fence; the arrival is a relaxed red/st; the
// waiters poll with relaxed loads (no L1 invalidation per poll) and issue ONE acquire fence when the epoch
// is complete, which also drops this SM's stale L1 lines before the next iteration's gathers.
//   FLAGS = false: one monotone counter, spin until it reaches epoch * gridDim.x
//   FLAGS = true : one flag word per CTA (128 bytes apart); thread i spins on the flag of CTA i
constexpr int kFlagStride = 32;          // unsigned words = 128 bytes
template <bool FLAGS>
__device__ __forceinline__ void grid_barrier(unsigned *sync_words, unsigned epoch)
{
    __syncthreads();
    if (FLAGS) {
        if (threadIdx.x == 0) {
            fence_acq_rel_gpu();
            st_relaxed(sync_words + (size_t)blockIdx.x * kFlagStride, epoch);
        }
        if (threadIdx.x < gridDim.x) {
            while (ld_relaxed(sync_words + (size_t)threadIdx.x * kFlagStride) < epoch) { }
            fence_acq_rel_gpu();
        }
    } else {
        if (threadIdx.x == 0) {
            fence_acq_rel_gpu();
            red_relaxed_add(sync_words, 1u);
            while (ld_relaxed(sync_words) < epoch * gridDim.x) { }
            fence_acq_rel_gpu();
        }
    }
    __syncthreads();
}

constexpr int kLongRow = 32;     // rows with more nonzeros are split over the lane groups of a whole warp

template <int LANES, int THREADS, int UNROLL, bool FLAGS>
__global__ void __launch_bounds__(THREADS, 1)
poisson_persistent_kernel(const int *__restrict__ rowptr, const int *__restrict__ col,
                          const float *__restrict__ val, const float *__restrict__ Db, float *u0, float *u1, int n,
                          int ldu, int T, const int *__restrict__ cta_rows, int max_rows, int slab_cap,
                          unsigned *sync_words)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: (col,val) [slab_cap] int2 | rp [max_rows+1] i32 | long rows [max_rows] i32 | n_long i32 |
    //         has-source flag [max_rows] u8
    int2 *s_cv = reinterpret_cast<int2 *>(smem_raw);
    int *s_rp = reinterpret_cast<int *>(s_cv + slab_cap);
    int *s_long = s_rp + (max_rows + 1);
    int *s_nlong = s_long + max_rows;
    unsigned char *s_src = reinterpret_cast<unsigned char *>(s_nlong + 1);

    const int r0 = cta_rows[blockIdx.x];
    const int r1 = cta_rows[blockIdx.x + 1];
    const int nrows = r1 - r0;
    const int nz0 = rowptr[r0];
    const int nnz_slab = rowptr[r1] - nz0;
    if (threadIdx.x == 0) *s_nlong = 0;
    for (int i = threadIdx.x; i < nnz_slab; i += THREADS)
        s_cv[i] = make_int2((int)((unsigned)col[nz0 + i] * (unsigned)ldu * 4u), __float_as_int(val[nz0 + i]));
    for (int i = threadIdx.x; i <= nrows; i += THREADS) s_rp[i] = rowptr[r0 + i] - nz0;
    __syncthreads();
    // The Poisson source Db is zero except on the labelled rows: remember which rows of this block have
    // one instead of streaming n x ldu zeros every iteration.  Rows too long for one lane group are listed.
    for (int lr = threadIdx.x; lr < nrows; lr += THREADS) {
        bool nz = false;
        const float4 *b = reinterpret_cast<const float4 *>(Db + (size_t)(r0 + lr) * ldu);
        for (int q = 0; q < ldu / 4; ++q) {
            const float4 v = __ldg(b + q);
            nz |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);      // NaN != 0 is true
        }
        s_src[lr] = nz ? 1 : 0;
        if (s_rp[lr + 1] - s_rp[lr] > kLongRow) s_long[atomicAdd(s_nlong, 1)] = lr;
    }
    __syncthreads();
    const int n_long = *s_nlong;

    const int li = threadIdx.x % LANES;
    const int rid = threadIdx.x / LANES;
    constexpr int RPP = THREADS / LANES;            // rows per pass of the CTA
    constexpr int NG = 32 / LANES;                  // lane groups per warp
    const int lane = threadIdx.x & 31;
    const int grp = lane / LANES;
    const int warp = threadIdx.x >> 5;
    const CsrShared csr{s_cv};

    for (int t = 0; t < T; ++t) {
        const float *u_in = (t & 1) ? u1 : u0;
        float *u_out = (t & 1) ? u0 : u1;
        // phase A: one lane group per (short) row
        for (int lr = rid; lr < nrows; lr += RPP) {
            const int beg = s_rp[lr], end = s_rp[lr + 1];
            if (end - beg > kLongRow) continue;
            for (int coff = li * 4; coff < ldu; coff += LANES * 4) {
                float4 acc = row_times_u<false, UNROLL>(csr, beg, end, reinterpret_cast<const char *>(u_in + coff));
                const size_t o = (size_t)(r0 + lr) * ldu + coff;
                if (s_src[lr]) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(Db + o));
                    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
                }
                *reinterpret_cast<float4 *>(u_out + o) = acc;
            }
        }
        // phase B: one warp per long row, nonzeros split over its NG lane groups, fixed-order shuffle reduction
        for (int q = warp; q < n_long; q += THREADS / 32) {
            const int lr = s_long[q];
            const int beg = s_rp[lr], end = s_rp[lr + 1];
            const int per = (end - beg + NG - 1) / NG;
            const int gb = min(end, beg + grp * per), ge = min(end, gb + per);
            for (int coff = li * 4; coff < ldu; coff += LANES * 4) {
                float4 acc = row_times_u<false, UNROLL>(csr, gb, ge, reinterpret_cast<const char *>(u_in + coff));
#pragma unroll
                for (int off = LANES; off < 32; off <<= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
                    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
                }
                if (grp == 0) {
                    const size_t o = (size_t)(r0 + lr) * ldu + coff;
                    if (s_src[lr]) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(Db + o));
                        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
                    }
                    *reinterpret_cast<float4 *>(u_out + o) = acc;
                }
            }
        }
        if (t + 1 < T) grid_barrier<FLAGS>(sync_words, (unsigned)(t + 1));
    }
}

// ------------------------------------------------------------------------------------------------
// mixing vector v <- RW v (fp64) and max|v - vinf|     (ssl.py:667,669)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_max_to_global(double m, unsigned long long *out)
{
    // non-negative doubles (and NaN, which sorts above +inf) compare like their bit patterns
    for (int off = 16; off > 0; off >>= 1) {
        double o = __shfl_xor_sync(0xffffffffu, m, off);
        m = (__double_as_longlong(o) > __double_as_longlong(m)) ? o : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

__global__ void __launch_bounds__(256)
mixing_step_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                   const double *__restrict__ vinf, const double *__restrict__ v_in, double *__restrict__ v_out,
                   int n, unsigned long long *err_out)
{
    constexpr int G = 8;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngroups = ((long long)gridDim.x * blockDim.x) / G;
    double worst = 0.0;
    for (long long rb = 0; rb < n; rb += ngroups) {          // uniform trip count: shuffles stay converged
        const long long row = rb + gid;
        double s = 0.0;
        if (row < n) {
            const int beg = rowptr[row], end = rowptr[row + 1];
            for (int j = beg + gl; j < end; j += G) s += val[j] * v_in[col[j]];
        }
        for (int off = G / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off, G);
        if (row < n && gl == 0) {
            v_out[row] = s;
            const double d = fabs(s - vinf[row]);
            // fabs(NaN) is NaN: keep it (np.max propagates NaN, ssl.py:667)
            worst = (d > worst || d != d) ? d : worst;
        }
    }
    block_max_to_global(worst, err_out);
}

__global__ void __launch_bounds__(256)
maxdiff_kernel(const double *__restrict__ v, const double *__restrict__ vinf, int n, unsigned long long *err_out)
{
    double worst = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d = fabs(v[i] - vinf[i]);
        worst = (d > worst || d != d) ? d : worst;
    }
    block_max_to_global(worst, err_out);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int LANES>
static int launch_step(const int *rp, const int *col, const float *val, const float *Db, const float *u_in,
                       float *u_out, int64_t n, int ldu, cudaStream_t st)
{
    const int threads = 256;
    const int64_t rows_per_block = threads / LANES;
    int64_t blocks = (n + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)sm_count() * 8 * 8;           // 8 resident CTAs/SM x 8 waves, then grid-stride
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    poisson_step_kernel<LANES><<<(unsigned)blocks, threads, 0, st>>>(rp, col, val, Db, u_in, u_out, (int)n, ldu);
    return 0;
}

static int dispatch_step(const int *rp, const int *col, const float *val, const float *Db, const float *u_in,
                         float *u_out, int64_t n, int ldu, cudaStream_t st)
{
    switch (ldu) {
        case 4: return launch_step<1>(rp, col, val, Db, u_in, u_out, n, ldu, st);
        case 8: return launch_step<2>(rp, col, val, Db, u_in, u_out, n, ldu, st);
        case 16: return launch_step<4>(rp, col, val, Db, u_in, u_out, n, ldu, st);
        case 32: return launch_step<8>(rp, col, val, Db, u_in, u_out, n, ldu, st);
        case 64: return launch_step<16>(rp, col, val, Db, u_in, u_out, n, ldu, st);
        default: return launch_step<32>(rp, col, val, Db, u_in, u_out, n, ldu, st);
    }
}

static bool valid_ld(int ldu)
{
    if (ldu < 4) return false;
    if (ldu <= 128) return (ldu & (ldu - 1)) == 0;
    return ldu % 128 == 0;
}

}  // namespace glb

using namespace glb;

extern "C" GLB_API int glb_poisson_step(const int32_t *d_rowptr, const int32_t *d_col, const float *d_val, const float *d_Db,
                                const float *d_u_in, float *d_u_out, int64_t n, int ldu, void *stream)
{
    GLB_CHECK_ARG(d_rowptr && d_col && d_val && d_Db && d_u_in && d_u_out, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31), "n out of range");
    GLB_CHECK_ARG(valid_ld(ldu), "ldu must be a power of two in [4,128] or a multiple of 128");
    GLB_CHECK_ARG((double)n * ldu * 4.0 < 4294967296.0, "label matrix larger than 4 GiB: 32-bit row offsets overflow");
    GLB_CHECK_ARG(d_u_in != d_u_out, "u_in and u_out must differ");
    dispatch_step(d_rowptr, d_col, d_val, d_Db, d_u_in, d_u_out, n, ldu, (cudaStream_t)stream);
    GLB_LAUNCH_CHECK();
    return 0;
}

struct glb_poisson_plan {
    int64_t n, nnz;
    int ldu;
    int persistent;
    int grid, max_rows, slab_cap;
    int threads;
    const void *fn;
    size_t smem_bytes;
    unsigned *d_counter;      // barrier words: kFlagStride * grid unsigned
    int *d_cta_rows;          // grid + 1 row boundaries of the work-balanced partition
};

// Launch geometry of the persistent kernel.  Default: 1024 threads, 4 gathers in flight per lane, counter
// barrier (fastest in the r1b sweep, profiles/).  GLB_POISSON_VARIANT="threads,unroll,flags" (e.g. "512,16,0") selects another instantiation for
// experiments; unknown combinations fall back to the default.
struct PersistVariant { int threads, unroll, flags; };

static PersistVariant persist_variant()
{
    PersistVariant v{1024, 4, 0};
    const char *e = getenv("GLB_POISSON_VARIANT");
    if (e) {
        int t = 0, u = 0, f = 0;
        if (sscanf(e, "%d,%d,%d", &t, &u, &f) == 3) { v.threads = t; v.unroll = u; v.flags = f; }
    }
    return v;
}

template <int LANES>
static const void *persistent_fn(const PersistVariant &v, int *threads)
{
#define GLB_PV(T_, U_, F_)                                                      \
    if (v.threads == T_ && v.unroll == U_ && v.flags == F_) {                   \
        *threads = T_;                                                          \
        return (const void *)poisson_persistent_kernel<LANES, T_, U_, (F_) != 0>; \
    }
    GLB_PV(1024, 8, 1) GLB_PV(1024, 8, 0) GLB_PV(1024, 4, 1) GLB_PV(1024, 4, 0)
    GLB_PV(512, 8, 1) GLB_PV(512, 16, 1) GLB_PV(512, 16, 0) GLB_PV(768, 8, 1)
#undef GLB_PV
    *threads = 1024;
    return (const void *)poisson_persistent_kernel<LANES, 1024, 4, false>;
}

static const void *pick_persistent(int ldu, int *threads)
{
    const PersistVariant v = persist_variant();
    switch (ldu) {
        case 4: return persistent_fn<1>(v, threads);
        case 8: return persistent_fn<2>(v, threads);
        case 16: return persistent_fn<4>(v, threads);
        case 32: return persistent_fn<8>(v, threads);
        case 64: return persistent_fn<16>(v, threads);
        default: return persistent_fn<32>(v, threads);
    }
}

extern "C" GLB_API int glb_poisson_plan_create(glb_poisson_plan **plan, const int32_t *d_rowptr, int64_t n, int64_t nnz,
                                       int ldu, void *stream)
{
    GLB_CHECK_ARG(plan && d_rowptr, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "size out of range");
    GLB_CHECK_ARG(valid_ld(ldu), "bad ldu");
    GLB_CHECK_ARG((double)n * ldu * 4.0 < 4294967296.0, "label matrix larger than 4 GiB: 32-bit row offsets overflow");
    cudaStream_t st = (cudaStream_t)stream;
    glb_poisson_plan *p = new glb_poisson_plan();
    p->n = n; p->nnz = nnz; p->ldu = ldu; p->persistent = 0; p->d_counter = nullptr; p->d_cta_rows = nullptr;
    p->grid = 0; p->max_rows = 0; p->slab_cap = 0; p->smem_bytes = 0; p->threads = 0; p->fn = nullptr;
    *plan = p;

    int dev = 0, coop = 0, max_smem = 0;
    GLB_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int sms = sm_count();
    // Upper bound for a graph that could fit the shared-memory slabs at all (8 bytes per nonzero per CTA).
    if (!coop || (double)nnz * 8.0 / sms > (double)max_smem) return 0;
    int grid = (int)((n + 127) / 128);          // tiny graphs: at least ~128 rows per CTA
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;

    // work-balanced contiguous row partition from the host copy of rowptr: cost(row) = nnz(row) + 4
    std::vector<int> h_rp((size_t)n + 1);
    GLB_CUDA(cudaMemcpyAsync(h_rp.data(), d_rowptr, sizeof(int) * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    std::vector<int> bounds((size_t)grid + 1, 0);
    const double total = (double)h_rp[n] + 4.0 * (double)n;
    {
        int b = 1;
        for (int64_t i = 0; i < n && b < grid; ++i) {
            const double pref = (double)h_rp[i + 1] + 4.0 * (double)(i + 1);
            while (b < grid && pref >= total * b / grid) bounds[b++] = (int)(i + 1);
        }
        for (; b <= grid; ++b) bounds[b] = (int)n;
        bounds[grid] = (int)n;
    }
    int cap = 0, max_rows = 0;
    for (int b = 0; b < grid; ++b) {
        cap = std::max(cap, h_rp[bounds[b + 1]] - h_rp[bounds[b]]);
        max_rows = std::max(max_rows, bounds[b + 1] - bounds[b]);
    }
    cap = (cap + 3) & ~3;
    const size_t smem = (size_t)cap * 8 + (size_t)(2 * max_rows + 2) * 4 + (size_t)((max_rows + 15) & ~15) + 16;
    if (smem > (size_t)max_smem) return 0;
    int threads = 0;
    const void *fn = pick_persistent(ldu, &threads);
    GLB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1 || grid > per_sm * sms || grid > threads) return 0;
    GLB_CUDA(cudaMalloc(&p->d_counter, sizeof(unsigned) * kFlagStride * grid));
    GLB_CUDA(cudaMalloc(&p->d_cta_rows, sizeof(int) * (grid + 1)));
    GLB_CUDA(cudaMemcpyAsync(p->d_cta_rows, bounds.data(), sizeof(int) * (grid + 1), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaStreamSynchronize(st));        // bounds is a local
    p->persistent = 1;
    p->grid = grid; p->max_rows = max_rows; p->slab_cap = cap; p->smem_bytes = smem;
    p->threads = threads; p->fn = fn;
    return 0;
}

extern "C" GLB_API int glb_poisson_plan_destroy(glb_poisson_plan *plan)
{
    if (!plan) return 0;
    if (plan->d_counter) cudaFree(plan->d_counter);
    if (plan->d_cta_rows) cudaFree(plan->d_cta_rows);
    delete plan;
    return 0;
}

extern "C" GLB_API int glb_poisson_plan_is_persistent(const glb_poisson_plan *plan) { return plan ? plan->persistent : 0; }

extern "C" GLB_API int glb_poisson_iterate(glb_poisson_plan *plan, const int32_t *d_rowptr, const int32_t *d_col,
                                   const float *d_val, const float *d_Db, float *d_u0, float *d_u1, int T,
                                   int *result_in_u1, int *launches, void *stream)
{
    GLB_CHECK_ARG(plan && d_rowptr && d_col && d_val && d_Db && d_u0 && d_u1, "null pointer");
    GLB_CHECK_ARG(T >= 0, "T must be >= 0");
    GLB_CHECK_ARG(d_u0 != d_u1, "u0 and u1 must differ");
    cudaStream_t st = (cudaStream_t)stream;
    if (result_in_u1) *result_in_u1 = T & 1;
    if (T == 0) return 0;
    if (plan->persistent) {
        GLB_CUDA(cudaMemsetAsync(plan->d_counter, 0, sizeof(unsigned) * kFlagStride * plan->grid, st));
        int n = (int)plan->n, ldu = plan->ldu, max_rows = plan->max_rows, cap = plan->slab_cap;
        void *args[] = {(void *)&d_rowptr, (void *)&d_col, (void *)&d_val, (void *)&d_Db, (void *)&d_u0, (void *)&d_u1,
                        (void *)&n, (void *)&ldu, (void *)&T, (void *)&plan->d_cta_rows, (void *)&max_rows, (void *)&cap,
                        (void *)&plan->d_counter};
        GLB_CUDA(cudaLaunchCooperativeKernel(plan->fn, dim3(plan->grid), dim3(plan->threads), args,
                                             plan->smem_bytes, st));
        if (launches) *launches += 1;
    } else {
        for (int t = 0; t < T; ++t) {
            const float *in = (t & 1) ? d_u1 : d_u0;
            float *out = (t & 1) ? d_u0 : d_u1;
            dispatch_step(d_rowptr, d_col, d_val, d_Db, in, out, plan->n, plan->ldu, st);
        }
        GLB_LAUNCH_CHECK();
        if (launches) *launches += T;
    }
    return 0;
}

extern "C" GLB_API int glb_poisson_mixing_T(const int32_t *d_rw_rowptr, const int32_t *d_rw_col, const double *d_rw_val,
                                    const double *d_vinf, double *d_v, double *d_tmp, int64_t n, int min_iter,
                                    int max_iter, int *T_out, int *launches, void *stream)
{
    GLB_CHECK_ARG(d_rw_rowptr && d_rw_col && d_rw_val && d_vinf && d_v && d_tmp && T_out, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31), "n out of range");
    cudaStream_t st = (cudaStream_t)stream;
    const int BATCH = 64;
    unsigned long long *d_err = nullptr;
    GLB_CUDA(cudaMalloc(&d_err, sizeof(unsigned long long) * (BATCH + 1)));
    unsigned long long h_err[BATCH + 1];
    const double thr = 1.0 / (double)n;
    const int threads = 256;
    int blocks = ceil_div(n * 8, threads);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    int T = 0, rc = 0;
    bool done = false;
    // err_0
    cudaMemsetAsync(d_err, 0, sizeof(unsigned long long) * (BATCH + 1), st);
    maxdiff_kernel<<<blocks, threads, 0, st>>>(d_v, d_vinf, (int)n, d_err);
    if (launches) *launches += 1;
    cudaMemcpyAsync(h_err, d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    double err;
    memcpy(&err, &h_err[0], sizeof(double));
    double *cur = d_v, *nxt = d_tmp;
    while (!done) {
        // condition of ssl.py:667 evaluated BEFORE each step, with err = max|v_T - vinf|
        if (!((T < min_iter || err > thr) && T < max_iter)) break;
        int steps = max_iter - T;
        if (steps > BATCH) steps = BATCH;
        cudaMemsetAsync(d_err, 0, sizeof(unsigned long long) * (BATCH + 1), st);
        for (int s = 0; s < steps; ++s) {
            mixing_step_kernel<<<blocks, threads, 0, st>>>(d_rw_rowptr, d_rw_col, d_rw_val, d_vinf, cur, nxt, (int)n,
                                                           d_err + s);
            double *t = cur; cur = nxt; nxt = t;
        }
        if (launches) *launches += steps;
        cudaError_t e = cudaMemcpyAsync(h_err, d_err, sizeof(unsigned long long) * steps, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = (int)e; set_error("glb_poisson_mixing_T: %s", cudaGetErrorString(e)); break; }
        for (int s = 0; s < steps; ++s) {
            T += 1;
            memcpy(&err, &h_err[s], sizeof(double));
            if (!((T < min_iter || err > thr) && T < max_iter)) { done = true; break; }
        }
    }
    cudaFree(d_err);
    if (rc == 0) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { rc = (int)e; set_error("glb_poisson_mixing_T: %s", cudaGetErrorString(e)); }
    }
    *T_out = T;
    return rc;
}
