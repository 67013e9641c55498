// cg.cu - multi-right-hand-side conjugate gradient on sm_100a.
//
// Replaces utils.conjgrad (reference graphlearning/utils.py:483-532), the solver behind ssl.laplace._fit
// (graphlearning/ssl.py:1249) and the default solver of ssl.poisson._fit (graphlearning/ssl.py:624-629):
//
//     x = 0 (or x0); r = b - A x; p = r; rsold = sum(r*r, axis=0)
//     while err > tol and i < max_iter:
//         Ap = A p; alpha = rsold / sum(p*Ap, axis=0); x += alpha p; r -= alpha Ap
//         rsnew = sum(r*r, axis=0); err = sqrt(sum(rsnew)); p = r + (rsnew/rsold) p; rsold = rsnew
//
// per-column alpha/beta, ONE stopping norm over all columns (utils.py:528), at least one iteration.
//
// Everything is fp64 like the reference (A values, x, r, p, Ap: row-major n x ldu, ldu = glb_padded_ld(c)); dot
// products are accumulated in a fixed order (deterministic run to run).  fp32 storage with fp64 reductions was
// tried first and dropped: on ill-conditioned graphs (two-moons: 1475 iterations of the
// singular normalised Laplacian in the reference) the fp32 recurrence residual never reaches the reference's
// tolerance, so neither the iteration count nor the scores can be matched.  A row of c = 10 doubles padded to 16
// is exactly one 128-byte line, i.e. one L1 wavefront per gathered row - the same count as the fp32 layout.
//
// Three kernels per iteration, no host round trip inside a batch of iterations:
//   cg_spmm_dot      Ap = A p (CSR gather, one lane group per row as in poisson.cu) fused with sum(p*Ap);
//                    the last CTA to finish folds the per-CTA partials and publishes alpha
//   cg_update        x += alpha p, r -= alpha Ap fused with sum(r*r); the last CTA publishes beta, err,
//                    records err in the history and raises `done` when err <= tol
//   cg_direction     p = r + beta p
// Every kernel returns at once when `done` is set, so the host enqueues iterations in batches and reads the
// state back once per batch.  All three are HBM/L2-streaming or gather kernels; algorithmic bytes per
// iteration (doubled for fp64 values): nnz*12 + (n+1)*4 + 11*n*c*8.
#include <math.h>
#include <vector>
#include "common.cuh"

namespace glb {

struct __align__(32) real4 { double x, y, z, w; };
__device__ __forceinline__ real4 ld4(const double *p)            // 32-byte aligned
{
    const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
    real4 v; v.x = a.x; v.y = a.y; v.z = b.x; v.w = b.y;
    return v;
}
__device__ __forceinline__ real4 ld4_rw(const double *p)         // data written earlier in the same kernel sequence by this thread
{
    const double2 a = reinterpret_cast<const double2 *>(p)[0], b = reinterpret_cast<const double2 *>(p)[1];
    real4 v; v.x = a.x; v.y = a.y; v.z = b.x; v.w = b.y;
    return v;
}
__device__ __forceinline__ void st4(double *p, const real4 &v)
{
    reinterpret_cast<double2 *>(p)[0] = make_double2(v.x, v.y);
    reinterpret_cast<double2 *>(p)[1] = make_double2(v.z, v.w);
}

constexpr int kCgMaxLd = 128;
constexpr int kCgThreads = 256;
constexpr int kCgHist = 64;                 // iterations per batch = length of the err history window
constexpr int kCgLongRow = 64;              // rows with more nonzeros are spread over a whole warp ...
constexpr int kCgHubRow = 768;              // ... and beyond this over a whole CTA

struct CgState {
    double rsold[kCgMaxLd];
    double alpha[kCgMaxLd];
    double beta[kCgMaxLd];
    double err;
    double err_hist[kCgHist];
    long long iters;
    int done;
    unsigned ticket[2];                     // last-CTA tickets of cg_spmm_dot / cg_update
};

// ---- block-level deterministic reduction of per-thread column partials -----------------------------------
// thread layout: li = threadIdx.x % LANES owns columns [4 li, 4 li + 4); d[0..3] are its fp64 partials.
// Result: partial[blockIdx.x * ldu + k] for k < ldu.
template <int LANES>
__device__ __forceinline__ void block_reduce_columns(double d[4], double *sh, double *partial_out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, li = threadIdx.x % LANES;
#pragma unroll
    for (int off = LANES; off < 32; off <<= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) d[q] += __shfl_xor_sync(0xffffffffu, d[q], off);
    }
    constexpr int LDU = LANES * 4;
    if (lane < LANES) {
#pragma unroll
        for (int q = 0; q < 4; ++q) sh[warp * LDU + li * 4 + q] = d[q];
    }
    __syncthreads();
    if (threadIdx.x < LDU) {
        double s = 0.0;
        for (int w = 0; w < kCgThreads / 32; ++w) s += sh[w * LDU + threadIdx.x];
        partial_out[(size_t)blockIdx.x * LDU + threadIdx.x] = s;
    }
}

// The last CTA to arrive folds the per-CTA partials in block order.  Returns true in that CTA with the totals in
// sh_tot[0..LDU) (valid after the __syncthreads inside).
template <int LDU>
__device__ __forceinline__ bool last_block_totals(const double *partial, unsigned *ticket, double *sh, double *sh_tot)
{
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    constexpr int NSEG = kCgThreads / LDU;                  // LDU <= 128 -> at least 2 segments
    const int k = threadIdx.x % LDU, seg = threadIdx.x / LDU;
    double s = 0.0;
    for (unsigned b = seg; b < gridDim.x; b += NSEG) s += __ldcg(partial + (size_t)b * LDU + k);
    sh[seg * LDU + k] = s;
    __syncthreads();
    if (threadIdx.x < LDU) {
        double tot = 0.0;
        for (int g = 0; g < NSEG; ++g) tot += sh[g * LDU + threadIdx.x];
        sh_tot[threadIdx.x] = tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) *ticket = 0u;                     // ready for the next launch
    return true;
}

// ---- K1: Ap = A p, pAp = sum(p * Ap) -> alpha ---------------------------------------------------------------
// MODE 0: CG iteration (dot with p, publishes alpha).  MODE 1: plain product out = A p (initial residual).
template <int LANES, int MODE>
__global__ void __launch_bounds__(kCgThreads)
cg_spmm_dot(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
            const double *__restrict__ p, double *__restrict__ Ap, int n, int c, double *partial, CgState *st)
{
    constexpr int LDU = LANES * 4;
    __shared__ double sh[(kCgThreads / 32) * LDU > kCgThreads ? (kCgThreads / 32) * LDU : kCgThreads];
    __shared__ double sh_tot[LDU];
    if (MODE == 0 && st->done) return;
    const int li = threadIdx.x % LANES;
    const int rid = (blockIdx.x * kCgThreads + threadIdx.x) / LANES;
    const int nrid = (gridDim.x * kCgThreads) / LANES;
    double d[4] = {0.0, 0.0, 0.0, 0.0};
    for (int row = rid; row < n; row += nrid) {
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        if (end - beg > kCgLongRow) continue;                         // hub rows: second loop, a whole warp per row
        real4 acc; acc.x = acc.y = acc.z = acc.w = 0.0;
        int j = beg;
        for (; j + 4 <= end; j += 4) {
            int cj[4]; double a[4]; real4 x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { cj[i] = __ldg(col + j + i); a[i] = __ldg(val + j + i); }
#pragma unroll
            for (int i = 0; i < 4; ++i) x[i] = ld4(p + (size_t)cj[i] * LDU + li * 4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc.x = fma(a[i], x[i].x, acc.x); acc.y = fma(a[i], x[i].y, acc.y);
                acc.z = fma(a[i], x[i].z, acc.z); acc.w = fma(a[i], x[i].w, acc.w);
            }
        }
        for (; j < end; ++j) {
            const double a = __ldg(val + j);
            const real4 x = ld4(p + (size_t)__ldg(col + j) * LDU + li * 4);
            acc.x = fma(a, x.x, acc.x); acc.y = fma(a, x.y, acc.y); acc.z = fma(a, x.z, acc.z); acc.w = fma(a, x.w, acc.w);
        }
        st4(Ap + (size_t)row * LDU + li * 4, acc);
        if (MODE == 0) {
            const real4 pr = ld4(p + (size_t)row * LDU + li * 4);
            d[0] += pr.x * acc.x; d[1] += pr.y * acc.y; d[2] += pr.z * acc.z; d[3] += pr.w * acc.w;
        }
    }
    // Rows with more than kCgLongRow nonzeros (the hubs of high-dimensional kNN graphs: thousands of neighbours at
    // d = 512): one lane group walking such a row alone would hold the whole launch.  Every warp looks through its
    // own contiguous share of the rows (a fixed partition: deterministic) and spreads each long row it finds over
    // its 32 / LANES lane groups; the partial sums are folded with shuffles in a fixed order.
    {
        constexpr int NG = 32 / LANES;
        const int lane = threadIdx.x & 31, g = lane / LANES;
        const int gw = (blockIdx.x * kCgThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kCgThreads) >> 5;
        const int chunk = (n + nwarps - 1) / nwarps;
        const int r0 = min(n, gw * chunk), r1 = min(n, r0 + chunk);
        for (int base = r0; base < r1; base += 32) {
            const int mine = base + lane;
            const int mylen = mine < r1 ? __ldg(rowptr + mine + 1) - __ldg(rowptr + mine) : 0;
            const bool is_long = mylen > kCgLongRow && mylen <= kCgHubRow;
            unsigned todo = __ballot_sync(0xffffffffu, is_long);
            while (todo) {
                const int row = base + __ffs(todo) - 1;
                todo &= todo - 1;
                const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
                real4 acc; acc.x = acc.y = acc.z = acc.w = 0.0;
                int j = beg + g;
                for (; j + 3 * NG < end; j += 4 * NG) {
                    int cj[4]; double a[4]; real4 x[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { cj[i] = __ldg(col + j + i * NG); a[i] = __ldg(val + j + i * NG); }
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = ld4(p + (size_t)cj[i] * LDU + li * 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc.x = fma(a[i], x[i].x, acc.x); acc.y = fma(a[i], x[i].y, acc.y);
                        acc.z = fma(a[i], x[i].z, acc.z); acc.w = fma(a[i], x[i].w, acc.w);
                    }
                }
                for (; j < end; j += NG) {
                    const double a = __ldg(val + j);
                    const real4 x = ld4(p + (size_t)__ldg(col + j) * LDU + li * 4);
                    acc.x = fma(a, x.x, acc.x); acc.y = fma(a, x.y, acc.y); acc.z = fma(a, x.z, acc.z); acc.w = fma(a, x.w, acc.w);
                }
#pragma unroll
                for (int off = LANES; off < 32; off <<= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
                }
                if (g == 0) {
                    st4(Ap + (size_t)row * LDU + li * 4, acc);
                    if (MODE == 0) {
                        const real4 pr = ld4(p + (size_t)row * LDU + li * 4);
                        d[0] += pr.x * acc.x; d[1] += pr.y * acc.y; d[2] += pr.z * acc.z; d[3] += pr.w * acc.w;
                    }
                }
            }
        }
    }
    // Rows with more than kCgHubRow nonzeros (5 000 neighbours happen at d = 512): the whole CTA takes one such row - 64
    // lane groups x 4 gathers in flight instead of a warp's 8 x 4 - and folds the partial sums through shared memory
    // in a fixed order.  Every CTA looks through its own contiguous share of the rows.
    {
        constexpr int NGB = kCgThreads / LANES;                        // lane groups of the CTA
        const int gb = threadIdx.x / LANES, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int chunk = (n + (int)gridDim.x - 1) / (int)gridDim.x;
        const int r0 = min(n, (int)blockIdx.x * chunk), r1 = min(n, r0 + chunk);
        for (int base = r0; base < r1; base += kCgThreads) {
            const int mine = base + threadIdx.x;
            const bool hub = mine < r1 && __ldg(rowptr + mine + 1) - __ldg(rowptr + mine) > kCgHubRow;
            if (!__syncthreads_or(hub)) continue;
            for (int t = 0; t < kCgThreads && base + t < r1; ++t) {
                const int row = base + t;
                const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
                if (end - beg <= kCgHubRow) continue;                  // block-uniform
                real4 acc; acc.x = acc.y = acc.z = acc.w = 0.0;
                int j = beg + gb;
                for (; j + 3 * NGB < end; j += 4 * NGB) {
                    int cj[4]; double a[4]; real4 x[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { cj[i] = __ldg(col + j + i * NGB); a[i] = __ldg(val + j + i * NGB); }
#pragma unroll
                    for (int i = 0; i < 4; ++i) x[i] = ld4(p + (size_t)cj[i] * LDU + li * 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc.x = fma(a[i], x[i].x, acc.x); acc.y = fma(a[i], x[i].y, acc.y);
                        acc.z = fma(a[i], x[i].z, acc.z); acc.w = fma(a[i], x[i].w, acc.w);
                    }
                }
                for (; j < end; j += NGB) {
                    const double a = __ldg(val + j);
                    const real4 x = ld4(p + (size_t)__ldg(col + j) * LDU + li * 4);
                    acc.x = fma(a, x.x, acc.x); acc.y = fma(a, x.y, acc.y); acc.z = fma(a, x.z, acc.z); acc.w = fma(a, x.w, acc.w);
                }
#pragma unroll
                for (int off = LANES; off < 32; off <<= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
                }
                __syncthreads();                                       // sh is free (previous hub row / nothing yet)
                if (lane < LANES) { sh[warp * LDU + li * 4 + 0] = acc.x; sh[warp * LDU + li * 4 + 1] = acc.y; sh[warp * LDU + li * 4 + 2] = acc.z; sh[warp * LDU + li * 4 + 3] = acc.w; }
                __syncthreads();
                if (threadIdx.x < LANES) {
                    real4 tot; tot.x = tot.y = tot.z = tot.w = 0.0;
                    for (int w = 0; w < kCgThreads / 32; ++w) {
                        tot.x += sh[w * LDU + li * 4 + 0]; tot.y += sh[w * LDU + li * 4 + 1];
                        tot.z += sh[w * LDU + li * 4 + 2]; tot.w += sh[w * LDU + li * 4 + 3];
                    }
                    st4(Ap + (size_t)row * LDU + li * 4, tot);
                    if (MODE == 0) {
                        const real4 pr = ld4(p + (size_t)row * LDU + li * 4);
                        d[0] += pr.x * tot.x; d[1] += pr.y * tot.y; d[2] += pr.z * tot.z; d[3] += pr.w * tot.w;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (MODE != 0) return;
    __syncthreads();
    block_reduce_columns<LANES>(d, sh, partial);
    if (last_block_totals<LDU>(partial, &st->ticket[0], sh, sh_tot)) {
        if (threadIdx.x < LDU) st->alpha[threadIdx.x] = threadIdx.x < c ? st->rsold[threadIdx.x] / sh_tot[threadIdx.x] : 0.0;
    }
}

// ---- K2: x += alpha p; r -= alpha Ap; rsnew = sum(r*r) -> beta, err, done ----------------------------------
// INIT: r = b - Ap (Ap = A x0, or r = b when Ap is NULL), p = r, rsold = sum(r*r); no x update.
template <int LANES, bool INIT>
__global__ void __launch_bounds__(kCgThreads)
cg_update(double *__restrict__ x, double *__restrict__ r, double *__restrict__ p, const double *__restrict__ Ap,
          const double *__restrict__ b, int n, int c, double tol, double *partial, CgState *st)
{
    constexpr int LDU = LANES * 4;
    __shared__ double sh[(kCgThreads / 32) * LDU > kCgThreads ? (kCgThreads / 32) * LDU : kCgThreads];
    __shared__ double sh_tot[LDU];
    if (!INIT && st->done) return;
    const int li = threadIdx.x % LANES;
    double al[4] = {0.0, 0.0, 0.0, 0.0};
    if (!INIT) {
#pragma unroll
        for (int q = 0; q < 4; ++q) al[q] = st->alpha[li * 4 + q];
    }
    double d[4] = {0.0, 0.0, 0.0, 0.0};
    const long long total = (long long)n * LANES;                     // groups of 4 doubles
    const long long stride = (long long)gridDim.x * kCgThreads;       // multiple of LANES: li is loop invariant
    for (long long i = (long long)blockIdx.x * kCgThreads + threadIdx.x; i < total; i += stride) {
        real4 rv;
        if (INIT) {
            rv = ld4(b + i * 4);
            if (Ap) {
                const real4 av = ld4(Ap + i * 4);
                rv.x -= av.x; rv.y -= av.y; rv.z -= av.z; rv.w -= av.w;
            }
            st4(p + i * 4, rv);
        } else {
            const real4 pv = ld4_rw(p + i * 4);
            const real4 av = ld4(Ap + i * 4);
            real4 xv = ld4_rw(x + i * 4);
            rv = ld4_rw(r + i * 4);
            // x += alpha*p ; r -= alpha*Ap  (utils.py:525-526; numpy rounds the product, then the sum)
            xv.x += al[0] * pv.x; xv.y += al[1] * pv.y; xv.z += al[2] * pv.z; xv.w += al[3] * pv.w;
            rv.x -= al[0] * av.x; rv.y -= al[1] * av.y; rv.z -= al[2] * av.z; rv.w -= al[3] * av.w;
            st4(x + i * 4, xv);
        }
        st4(r + i * 4, rv);
        d[0] += rv.x * rv.x; d[1] += rv.y * rv.y; d[2] += rv.z * rv.z; d[3] += rv.w * rv.w;
    }
    block_reduce_columns<LANES>(d, sh, partial);
    if (last_block_totals<LDU>(partial, &st->ticket[1], sh, sh_tot)) {
        if (INIT) {
            if (threadIdx.x < LDU) st->rsold[threadIdx.x] = sh_tot[threadIdx.x];
        } else {
            if (threadIdx.x < LDU) {
                const double rsnew = sh_tot[threadIdx.x], rsold = st->rsold[threadIdx.x];
                st->beta[threadIdx.x] = threadIdx.x < c ? rsnew / rsold : 0.0;
                st->rsold[threadIdx.x] = rsnew;
            }
            if (threadIdx.x == 0) {
                double s = 0.0;
                for (int k = 0; k < c; ++k) s += sh_tot[k];                // np.sum(rsnew), utils.py:528
                const double err = sqrt(s);
                const long long it = st->iters + 1;
                st->iters = it;
                st->err = err;
                st->err_hist[(it - 1) % kCgHist] = err;
                if (!(err > tol)) st->done = 1;                            // loop test `err > tol`; NaN stops too
            }
        }
    }
}

// ---- K3: p = r + beta p -------------------------------------------------------------------------------------
template <int LANES>
__global__ void __launch_bounds__(kCgThreads)
cg_direction(const double *__restrict__ r, double *__restrict__ p, int n, const CgState *st)
{
    if (st->done) return;          // x is final; the reference's last update of p is never used
    const int li = threadIdx.x % LANES;
    double be[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) be[q] = st->beta[li * 4 + q];
    const long long total = (long long)n * LANES;
    const long long stride = (long long)gridDim.x * kCgThreads;
    for (long long i = (long long)blockIdx.x * kCgThreads + threadIdx.x; i < total; i += stride) {
        const real4 rv = ld4(r + i * 4);
        real4 pv = ld4_rw(p + i * 4);
        pv.x = rv.x + be[0] * pv.x; pv.y = rv.y + be[1] * pv.y; pv.z = rv.z + be[2] * pv.z; pv.w = rv.w + be[3] * pv.w;
        st4(p + i * 4, pv);
    }
}

__global__ void cg_state_init(CgState *st)
{
    for (int k = threadIdx.x; k < kCgMaxLd; k += blockDim.x) { st->rsold[k] = 0.0; st->alpha[k] = 0.0; st->beta[k] = 0.0; }
    if (threadIdx.x == 0) { st->err = 1.0; st->iters = 0; st->done = 0; st->ticket[0] = 0; st->ticket[1] = 0; }
}

// n x c fp64 <-> n x ldu fp32 (plain layout, zero padded)
__global__ void __launch_bounds__(256)
cg_pack_kernel(const double *__restrict__ src, long long n, int c, double *__restrict__ dst, int ldu)
{
    const long long total = n * ldu;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ldu;
        const int k = (int)(i - r * ldu);
        dst[i] = k < c ? src[r * c + k] : 0.0;
    }
}
__global__ void __launch_bounds__(256)
cg_unpack_kernel(const double *__restrict__ src, long long n, int c, int ldu, double *__restrict__ dst)
{
    const long long total = n * c;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c;
        dst[i] = src[r * ldu + (i - r * c)];
    }
}
static int cg_grid() { return sm_count() * 4; }      // 4 x 256 threads per SM: gather kernels, latency bound

static size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }

template <int LANES>
static int cg_run(const int *rp, const int *col, const double *val, int64_t n, const double *b, const double *x0, int c,
                  double tol, int64_t max_iter, double *x, void *work, int64_t *iters_out, double *err_out, int *launches,
                  cudaStream_t st)
{
    constexpr int LDU = LANES * 4;
    const int grid = cg_grid();
    unsigned char *w = (unsigned char *)(((uintptr_t)work + 255) & ~(uintptr_t)255);
    const size_t vec = a256((size_t)n * LDU * sizeof(double));
    double *r = (double *)w;  w += vec;
    double *p = (double *)w;  w += vec;
    double *Ap = (double *)w; w += vec;
    double *partial = (double *)w; w += a256((size_t)grid * LDU * sizeof(double));
    CgState *state = (CgState *)w;
    int nl = 0;
    cg_state_init<<<1, 128, 0, st>>>(state); ++nl;
    if (x0) {
        GLB_CUDA(cudaMemcpyAsync(x, x0, (size_t)n * LDU * sizeof(double), cudaMemcpyDeviceToDevice, st));
        cg_spmm_dot<LANES, 1><<<grid, kCgThreads, 0, st>>>(rp, col, val, x, Ap, (int)n, c, partial, state); ++nl;
        cg_update<LANES, true><<<grid, kCgThreads, 0, st>>>(x, r, p, Ap, b, (int)n, c, tol, partial, state); ++nl;
    } else {
        GLB_CUDA(cudaMemsetAsync(x, 0, (size_t)n * LDU * sizeof(double), st));
        cg_update<LANES, true><<<grid, kCgThreads, 0, st>>>(x, r, p, nullptr, b, (int)n, c, tol, partial, state); ++nl;
    }
    GLB_LAUNCH_CHECK();
    CgState h;
    int64_t enq = 0;                       // iterations enqueued so far
    h.done = 0; h.iters = 0; h.err = 1.0;
    // the reference's loop test is evaluated before every iteration with err initialised to 1 (utils.py:519-521)
    if (!(1.0 > tol) || max_iter <= 0) { if (iters_out) *iters_out = 0; if (err_out) *err_out = 1.0; if (launches) *launches += nl; return 0; }
    int64_t ramp = 16;                     // batches of 16, 32, 64, 64, ...: a solve that ends early leaves few idle launches behind
    while (!h.done && enq < max_iter) {
        int64_t batch = max_iter - enq;
        if (batch > ramp) batch = ramp;
        if (ramp < kCgHist) ramp *= 2;
        for (int64_t i = 0; i < batch; ++i) {
            cg_spmm_dot<LANES, 0><<<grid, kCgThreads, 0, st>>>(rp, col, val, p, Ap, (int)n, c, partial, state);
            cg_update<LANES, false><<<grid, kCgThreads, 0, st>>>(x, r, p, Ap, b, (int)n, c, tol, partial, state);
            cg_direction<LANES><<<grid, kCgThreads, 0, st>>>(r, p, (int)n, state);
        }
        nl += 3 * (int)batch;
        enq += batch;
        GLB_LAUNCH_CHECK();
        GLB_CUDA(cudaMemcpyAsync(&h, state, sizeof(CgState), cudaMemcpyDeviceToHost, st));
        GLB_CUDA(cudaStreamSynchronize(st));
    }
    if (iters_out) *iters_out = h.iters;
    if (err_out) *err_out = h.err;
    if (launches) *launches += nl;
    return 0;
}

}  // namespace glb

using namespace glb;

extern "C" GLB_API int64_t glb_cg_work_bytes(int64_t n, int c)
{
    if (n <= 0 || c <= 0) return GLB_E_INVALID;
    const int ldu = glb_padded_ld(c);
    if (ldu > kCgMaxLd) return GLB_E_UNSUPPORTED;
    return (int64_t)(3 * a256((size_t)n * ldu * sizeof(double)) + a256((size_t)cg_grid() * ldu * sizeof(double)) + a256(sizeof(CgState)) + 512);
}

extern "C" GLB_API int glb_cg_solve(const int32_t *d_rowptr, const int32_t *d_col, const double *d_val, int64_t n, int64_t nnz,
                                    const double *d_b, const double *d_x0, int c, double tol, int64_t max_iter, double *d_x,
                                    void *d_work, int64_t work_bytes, int64_t *iters, double *err, int *launches, void *stream)
{
    GLB_CHECK_ARG(d_rowptr && (nnz == 0 || (d_col && d_val)) && d_b && d_x && d_work, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "size out of range");
    GLB_CHECK_ARG(c > 0, "c must be positive");
    const int ldu = glb_padded_ld(c);
    if (ldu > kCgMaxLd) { set_error("glb_cg_solve: more than %d right-hand sides", kCgMaxLd); return GLB_E_UNSUPPORTED; }
    GLB_CHECK_ARG(work_bytes >= glb_cg_work_bytes(n, c), "workspace too small");
    GLB_CHECK_ARG((double)n * ldu < 2147483648.0 * 4.0, "label matrix too large");
    cudaStream_t st = (cudaStream_t)stream;
    switch (ldu) {
        case 4: return cg_run<1>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
        case 8: return cg_run<2>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
        case 16: return cg_run<4>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
        case 32: return cg_run<8>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
        case 64: return cg_run<16>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
        default: return cg_run<32>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
    }
}

// Host-buffer entry point: utils.conjgrad(A, b, x0, max_iter, tol) with the reference's own types
// (scipy CSR int32/float64, numpy float64 n x c).  Synchronous.
extern "C" GLB_API int glb_cg_host(const int32_t *h_rowptr, const int32_t *h_col, const double *h_val, int64_t n, int64_t nnz,
                                   const double *h_b, const double *h_x0, int c, double tol, int64_t max_iter, double *h_x,
                                   int64_t *iters, double *err, int *launches)
{
    GLB_CHECK_ARG(h_rowptr && (nnz == 0 || (h_col && h_val)) && h_b && h_x, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "size out of range");
    GLB_CHECK_ARG(c > 0, "c must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("glb_cg_host: no CUDA device visible");
        return GLB_E_NOGPU;
    }
    const int64_t wb = glb_cg_work_bytes(n, c);
    if (wb < 0) { set_error("glb_cg_host: unsupported shape (c = %d)", c); return (int)wb; }
    const int ldu = glb_padded_ld(c);
    cudaStream_t st = 0;
    struct Arena { std::vector<void *> v; ~Arena() { for (void *p : v) dev_free(p); } } A;
    auto alloc = [&](void **p, size_t bytes) { cudaError_t e = dev_alloc(p, bytes ? bytes : 1); if (e == cudaSuccess) A.v.push_back(*p); return e; };
    int *rp, *col; double *val, *b, *x, *x0 = nullptr, *stage; void *work;
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    GLB_CUDA(alloc((void **)&rp, (n + 1) * sizeof(int)));   GLB_CUDA(alloc((void **)&col, nz * sizeof(int)));
    GLB_CUDA(alloc((void **)&val, nz * sizeof(double)));
    GLB_CUDA(alloc((void **)&b, (size_t)n * ldu * sizeof(double))); GLB_CUDA(alloc((void **)&x, (size_t)n * ldu * sizeof(double)));
    GLB_CUDA(alloc((void **)&stage, (size_t)n * c * sizeof(double))); GLB_CUDA(alloc(&work, (size_t)wb));
    GLB_CUDA(cudaMemcpyAsync(rp, h_rowptr, (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(col, h_col, nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(val, h_val, nnz * sizeof(double), cudaMemcpyHostToDevice, st));
    const int blocks = sm_count() * 8;
    int nl = 0;
    GLB_CUDA(cudaMemcpyAsync(stage, h_b, (size_t)n * c * sizeof(double), cudaMemcpyHostToDevice, st));
    cg_pack_kernel<<<blocks, 256, 0, st>>>(stage, n, c, b, ldu); ++nl;
    if (h_x0) {
        GLB_CUDA(alloc((void **)&x0, (size_t)n * ldu * sizeof(double)));
        GLB_CUDA(cudaMemcpyAsync(stage, h_x0, (size_t)n * c * sizeof(double), cudaMemcpyHostToDevice, st));
        cg_pack_kernel<<<blocks, 256, 0, st>>>(stage, n, c, x0, ldu); ++nl;
    }
    GLB_LAUNCH_CHECK();
    int rc = glb_cg_solve(rp, col, val, n, nnz, b, x0, c, tol, max_iter, x, work, wb, iters, err, &nl, st);
    if (rc) return rc;
    cg_unpack_kernel<<<blocks, 256, 0, st>>>(x, n, c, ldu, stage); ++nl;
    GLB_LAUNCH_CHECK();
    GLB_CUDA(cudaMemcpyAsync(h_x, stage, (size_t)n * c * sizeof(double), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    if (launches) *launches = nl;
    return 0;
}
