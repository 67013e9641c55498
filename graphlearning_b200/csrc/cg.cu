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
// is exactly one 128-byte line; 8 lanes x 16 bytes fetch it with one L1 wavefront - the same count as the fp32 layout.
//
// Three kernels per iteration, no host round trip inside a batch of iterations:
//   cg_spmm_dot      Ap = A p (CSR gather, one lane group per row, 8 gathers per lane in flight; rows of more than 64
//                    nonzeros as a static list of warp-wide work items) fused with sum(p*Ap); the last CTA to
//                    finish folds the per-CTA partials and publishes alpha
//   cg_update        x += alpha p, r -= alpha Ap fused with sum(r*r); the last CTA publishes beta, err,
//                    records err in the history and raises `done` when err <= tol
//   cg_direction     p = r + beta p
// Every kernel returns at once when `done` is set, so the host enqueues iterations in batches and reads the
// state back once per batch.  All three are HBM/L2-streaming or gather kernels; algorithmic bytes per
// iteration (doubled for fp64 values): nnz*12 + (n+1)*4 + 11*n*c*8.
#include <math.h>
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace glb {

// ---- layout ----------------------------------------------------------------------------------------------------------
// A row of the n x LDU matrices is LDU = LPR x CPL doubles: LPR lanes per row, CPL columns (2, or 4 at LDU = 128) per lane,
// so that one warp-wide 16-byte load fetches a whole 128-byte row of c = 10 classes with 8 lanes - one L1 wavefront per
// gathered row (the 4-lane x 32-byte layout of round 1 needed two).  Lanes whose columns are all padding (li * CPL >= c:
// 3 of 8 at c = 10) neither load nor store: the padding beyond them is never read.
template <int CPL>
__device__ __forceinline__ void ldv(const double *p, double (&v)[CPL])       // read-only data of this launch
{
#pragma unroll
    for (int q = 0; q < CPL / 2; ++q) { const double2 t = __ldg(reinterpret_cast<const double2 *>(p) + q); v[2 * q] = t.x; v[2 * q + 1] = t.y; }
}
template <int CPL>
__device__ __forceinline__ void ldv_rw(const double *p, double (&v)[CPL])    // data this thread rewrites
{
#pragma unroll
    for (int q = 0; q < CPL / 2; ++q) { const double2 t = reinterpret_cast<const double2 *>(p)[q]; v[2 * q] = t.x; v[2 * q + 1] = t.y; }
}
template <int CPL>
__device__ __forceinline__ void stv(double *p, const double (&v)[CPL])
{
#pragma unroll
    for (int q = 0; q < CPL / 2; ++q) reinterpret_cast<double2 *>(p)[q] = make_double2(v[2 * q], v[2 * q + 1]);
}

constexpr int kCgMaxLd = 128;
constexpr int kCgThreads = 256;
constexpr int kCgHist = 64;                 // iterations per batch = length of the err history window
constexpr int kCgLongRow = 64;              // rows with more nonzeros are cut into work items of a whole warp ...
constexpr int kCgItem = 256;                // ... of at most this many nonzeros each

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

// Work lists of the rows with more than kCgLongRow nonzeros (the hubs of high-dimensional kNN graphs: 2 400 neighbours at
// d = 512), built on the host once per solve from the row pointers.  Static lists dealt round-robin keep the launch
// balanced AND the summation order fixed: rows of up to kCgItem nonzeros go to warps (item i to warp i % warps), longer
// ones to CTAs (row m to CTA m % grid, its rounds of 32 entries dealt over the 8 warps, partial sums folded through
// shared memory in warp order).  Inside a round the entries are dealt to the warp's lane groups in blocks of LPR.
struct CgLong {
    const int4 *items;                      // {row, first entry, last entry + 1, 0}: rows of 65 .. kCgItem nonzeros
    const int4 *hubs;                       // the same for longer rows
    int n_items, n_hubs;
};

// ---- block-level deterministic reduction of per-thread column partials -----------------------------------
// thread layout: li = threadIdx.x % LPR owns columns [CPL li, CPL li + CPL); d[] are its fp64 partials.
// Result: partial_out[blockIdx.x * LDU + k] for k < LDU.
template <int LPR, int CPL>
__device__ __forceinline__ void block_reduce_columns(double (&d)[CPL], double *sh, double *partial_out)
{
    constexpr int LDU = LPR * CPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, li = threadIdx.x % LPR;
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
        for (int q = 0; q < CPL; ++q) d[q] += __shfl_xor_sync(0xffffffffu, d[q], off);
    }
    if (lane < LPR) {
#pragma unroll
        for (int q = 0; q < CPL; ++q) sh[warp * LDU + li * CPL + q] = d[q];
    }
    __syncthreads();
    if (threadIdx.x < LDU) {
        double s = 0.0;
        for (int w = 0; w < kCgThreads / 32; ++w) s += sh[w * LDU + threadIdx.x];
        partial_out[(size_t)blockIdx.x * LDU + threadIdx.x] = s;
    }
}

// Is this the last CTA of the launch to get here?  Everything the other CTAs wrote before is visible to it.
__device__ __forceinline__ bool last_block(unsigned *ticket)
{
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) *ticket = 0u;                          // ready for the next launch
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

// The last CTA folds the per-CTA partials in block order: totals in sh_tot[0..LDU) (valid after the __syncthreads inside).
template <int LDU>
__device__ __forceinline__ void fold_partials(const double *partial, double *sh, double *sh_tot)
{
    constexpr int NSEG = kCgThreads / LDU;                  // LDU <= 128 -> at least 2 segments
    const int k = threadIdx.x % LDU, seg = threadIdx.x / LDU;
    double s = 0.0;
#pragma unroll 8
    for (unsigned b = seg; b < gridDim.x; b += NSEG) s += __ldcg(partial + (size_t)b * LDU + k);   // loads independent, adds in block order
    sh[seg * LDU + k] = s;
    __syncthreads();
    if (threadIdx.x < LDU) {
        double tot = 0.0;
        for (int g = 0; g < NSEG; ++g) tot += sh[g * LDU + threadIdx.x];
        sh_tot[threadIdx.x] = tot;
    }
    __syncthreads();
}

// ---- K1: Ap = A p, pAp = sum(p * Ap) -> alpha ---------------------------------------------------------------
// MODE 0: CG iteration (dot with p, publishes alpha).  MODE 1: plain product out = A p (initial residual).
// Rows of up to kCgLongRow nonzeros: one lane group per row, RPW = 32 / LPR rows per warp.  The entries of a row are
// fetched LPR (at least 8) at a time - lane li of the group loads entry li, coalesced - and handed round with shuffles;
// every lane then has 8 label-row gathers in flight.  acc = fma(a_j, x_j, acc) in stored order.
template <int LPR, int CPL, int MODE>
__global__ void __launch_bounds__(kCgThreads, CPL == 2 ? 3 : 1)
cg_spmm_dot(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
            const double *__restrict__ p, double *__restrict__ Ap, int n, int c, double *partial, CgState *st, const CgLong L)
{
    constexpr int LDU = LPR * CPL, RPW = 32 / LPR;
    constexpr int EPL = LPR >= 8 ? 1 : 8 / LPR;                  // entries a lane fetches per round
    constexpr int E = LPR * EPL;                                 // entries of a row per round: 8, 16 or 32
    constexpr int SB = LPR < 8 ? LPR : 8;                        // gathers per lane and sub-batch of an item round
    __shared__ double sh[(kCgThreads / 32) * LDU > kCgThreads ? (kCgThreads / 32) * LDU : kCgThreads];
    __shared__ double sh_tot[LDU];
    if (MODE == 0 && st->done) return;
    const int lane = threadIdx.x & 31, li = lane % LPR, g = lane / LPR;
    const int wid = (blockIdx.x * kCgThreads + threadIdx.x) >> 5, nw = (gridDim.x * kCgThreads) >> 5;
    const bool active = li * CPL < c;
    const double *pl = p + li * CPL;
    double d[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) d[k] = 0.0;
    int nbeg = 0, nend = 0;                                      // row pointers of the NEXT visit, fetched one visit ahead
    if (wid * RPW + g < n) { nbeg = __ldg(rowptr + wid * RPW + g); nend = __ldg(rowptr + wid * RPW + g + 1); }
    for (int base = wid * RPW; base < n; base += nw * RPW) {
        const int row = base + g;
        int beg = nbeg, len = nend - nbeg;
        bool is_long = false;
        if (len > kCgLongRow) { is_long = true; len = 0; }      // a work item of the list below
        {
            const long long nrow = (long long)row + (long long)nw * RPW;
            nbeg = nend = 0;
            if (nrow < n) { nbeg = __ldg(rowptr + nrow); nend = __ldg(rowptr + nrow + 1); }
        }
        const int maxlen = __reduce_max_sync(0xffffffffu, len);
        double acc[CPL];
#pragma unroll
        for (int k = 0; k < CPL; ++k) acc[k] = 0.0;
        // entries of the first round; the next round's entries are fetched while this round's gathers are in flight
        int cj[EPL];
        double aj[EPL];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
            const int t = e * LPR + li;
            const bool ok = t < len;
            cj[e] = ok ? __ldg(col + beg + t) : 0;
            aj[e] = ok ? __ldg(val + beg + t) : 0.0;
        }
        for (int j0 = 0; j0 < maxlen; j0 += E) {
            int cn[EPL];
            double an[EPL];
#pragma unroll
            for (int e = 0; e < EPL; ++e) {
                const int t = j0 + E + e * LPR + li;
                const bool ok = t < len;
                cn[e] = ok ? __ldg(col + beg + t) : 0;
                an[e] = ok ? __ldg(val + beg + t) : 0.0;
            }
#pragma unroll
            for (int sb = 0; sb < E; sb += 8) {
                if (j0 + sb < maxlen) {                              // warp-uniform
                    int cc[8];
                    double aa[8], x[8][CPL];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int src = g * LPR + (sb + q) % LPR;
                        cc[q] = __shfl_sync(0xffffffffu, cj[(sb + q) / LPR], src);
                        aa[q] = __shfl_sync(0xffffffffu, aj[(sb + q) / LPR], src);
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (active && j0 + sb + q < len) ldv<CPL>(pl + (size_t)cc[q] * LDU, x[q]);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (active && j0 + sb + q < len) {
#pragma unroll
                            for (int k = 0; k < CPL; ++k) acc[k] = fma(aa[q], x[q][k], acc[k]);
                        }
                }
            }
#pragma unroll
            for (int e = 0; e < EPL; ++e) { cj[e] = cn[e]; aj[e] = an[e]; }
        }
        if (row < n && active && !is_long) {
            stv<CPL>(Ap + (size_t)row * LDU + li * CPL, acc);
            if (MODE == 0) {
                double pr[CPL];
                ldv<CPL>(pl + (size_t)row * LDU, pr);
#pragma unroll
                for (int k = 0; k < CPL; ++k) d[k] += pr[k] * acc[k];
            }
        }
    }
    // One warp-wide pass over entries [e0, e1) in rounds of 32 (one coalesced load per lane, the next round's entries in
    // flight beside the gathers), rounds r0, r0 + rstep, ...; group g gathers entries [g LPR, g LPR + LPR) of a round.
    // Returns with the groups folded: every lane group holds the sum.
    auto warp_pass = [&](int e0, int e1, int r0, int rstep, double (&acc)[CPL]) {
#pragma unroll
        for (int k = 0; k < CPL; ++k) acc[k] = 0.0;
        int j0 = e0 + r0 * 32;
        int cj = j0 + lane < e1 ? __ldg(col + j0 + lane) : 0;
        double aj = j0 + lane < e1 ? __ldg(val + j0 + lane) : 0.0;
        for (; j0 < e1; j0 += rstep * 32) {
            const int jn = j0 + rstep * 32;
            const bool okn = jn + lane < e1;
            const int cn = okn ? __ldg(col + jn + lane) : 0;
            const double an = okn ? __ldg(val + jn + lane) : 0.0;
#pragma unroll
            for (int sb = 0; sb < LPR; sb += SB) {
                int cc[SB];
                double aa[SB], x[SB][CPL];
#pragma unroll
                for (int q = 0; q < SB; ++q) {
                    cc[q] = __shfl_sync(0xffffffffu, cj, g * LPR + sb + q);
                    aa[q] = __shfl_sync(0xffffffffu, aj, g * LPR + sb + q);
                }
#pragma unroll
                for (int q = 0; q < SB; ++q)
                    if (active && j0 + g * LPR + sb + q < e1) ldv<CPL>(pl + (size_t)cc[q] * LDU, x[q]);
#pragma unroll
                for (int q = 0; q < SB; ++q)
                    if (active && j0 + g * LPR + sb + q < e1) {
#pragma unroll
                        for (int k = 0; k < CPL; ++k) acc[k] = fma(aa[q], x[q][k], acc[k]);
                    }
            }
            cj = cn; aj = an;
        }
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], off);
        }
    };
    for (int it = wid; it < L.n_items; it += nw) {               // item i belongs to warp i % nw
        const int4 item = __ldg(L.items + it);
        double acc[CPL];
        warp_pass(item.y, item.z, 0, 1, acc);
        if (g == 0 && active) {
            stv<CPL>(Ap + (size_t)item.x * LDU + li * CPL, acc);
            if (MODE == 0) {
                double pr[CPL];
                ldv<CPL>(pl + (size_t)item.x * LDU, pr);
#pragma unroll
                for (int k = 0; k < CPL; ++k) d[k] += pr[k] * acc[k];
            }
        }
    }
    for (int m = blockIdx.x; m < L.n_hubs; m += gridDim.x) {     // hub m belongs to CTA m % grid
        const int4 hub = __ldg(L.hubs + m);
        const int warp = threadIdx.x >> 5;
        double acc[CPL];
        warp_pass(hub.y, hub.z, warp, kCgThreads / 32, acc);
        __syncthreads();                                         // sh is free (previous hub)
        if (g == 0) {
#pragma unroll
            for (int k = 0; k < CPL; ++k) sh[warp * LDU + li * CPL + k] = active ? acc[k] : 0.0;
        }
        __syncthreads();
        if (warp == 0 && g == 0 && active) {
            double tot[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                tot[k] = 0.0;
                for (int w = 0; w < kCgThreads / 32; ++w) tot[k] += sh[w * LDU + li * CPL + k];
            }
            stv<CPL>(Ap + (size_t)hub.x * LDU + li * CPL, tot);
            if (MODE == 0) {
                double pr[CPL];
                ldv<CPL>(pl + (size_t)hub.x * LDU, pr);
#pragma unroll
                for (int k = 0; k < CPL; ++k) d[k] += pr[k] * tot[k];
            }
        }
    }
    if (MODE != 0) return;
    __syncthreads();
    block_reduce_columns<LPR, CPL>(d, sh, partial);
    if (!last_block(&st->ticket[0])) return;
    fold_partials<LDU>(partial, sh, sh_tot);
    if (threadIdx.x < LDU) st->alpha[threadIdx.x] = threadIdx.x < c ? st->rsold[threadIdx.x] / sh_tot[threadIdx.x] : 0.0;
}

// ---- K2: x += alpha p; r -= alpha Ap; rsnew = sum(r*r) -> beta, err, done ----------------------------------
// INIT: r = b - Ap (Ap = A x0, or r = b when Ap is NULL), p = r, rsold = sum(r*r); no x update.
template <int LPR, int CPL, bool INIT>
__global__ void __launch_bounds__(kCgThreads)
cg_update(double *__restrict__ x, double *__restrict__ r, double *__restrict__ p, const double *__restrict__ Ap,
          const double *__restrict__ b, int n, int c, double tol, double *partial, CgState *st)
{
    constexpr int LDU = LPR * CPL;
    __shared__ double sh[(kCgThreads / 32) * LDU > kCgThreads ? (kCgThreads / 32) * LDU : kCgThreads];
    __shared__ double sh_tot[LDU];
    if (!INIT && st->done) return;
    const int li = threadIdx.x % LPR;
    const bool active = li * CPL < c;
    double al[CPL], d[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) { al[q] = INIT ? 0.0 : st->alpha[li * CPL + q]; d[q] = 0.0; }
    const long long total = (long long)n * LPR;                       // groups of CPL doubles
    const long long stride = (long long)gridDim.x * kCgThreads;       // multiple of LPR: li is loop invariant
    if (active)
#pragma unroll 2
        for (long long i = (long long)blockIdx.x * kCgThreads + threadIdx.x; i < total; i += stride) {
            double rv[CPL];
            if (INIT) {
                ldv<CPL>(b + i * CPL, rv);
                if (Ap) {
                    double av[CPL];
                    ldv<CPL>(Ap + i * CPL, av);
#pragma unroll
                    for (int q = 0; q < CPL; ++q) rv[q] -= av[q];
                }
                stv<CPL>(p + i * CPL, rv);
            } else {
                double pv[CPL], av[CPL], xv[CPL];
                ldv_rw<CPL>(p + i * CPL, pv);
                ldv<CPL>(Ap + i * CPL, av);
                ldv_rw<CPL>(x + i * CPL, xv);
                ldv_rw<CPL>(r + i * CPL, rv);
                // x += alpha*p ; r -= alpha*Ap  (utils.py:525-526; numpy rounds the product, then the sum)
#pragma unroll
                for (int q = 0; q < CPL; ++q) { xv[q] += al[q] * pv[q]; rv[q] -= al[q] * av[q]; }
                stv<CPL>(x + i * CPL, xv);
            }
            stv<CPL>(r + i * CPL, rv);
#pragma unroll
            for (int q = 0; q < CPL; ++q) d[q] += rv[q] * rv[q];
        }
    block_reduce_columns<LPR, CPL>(d, sh, partial);
    if (!last_block(&st->ticket[1])) return;
    fold_partials<LDU>(partial, sh, sh_tot);
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

// ---- K3: p = r + beta p -------------------------------------------------------------------------------------
template <int LPR, int CPL>
__global__ void __launch_bounds__(kCgThreads)
cg_direction(const double *__restrict__ r, double *__restrict__ p, int n, int c, const CgState *st)
{
    if (st->done) return;          // x is final; the reference's last update of p is never used
    const int li = threadIdx.x % LPR;
    if (li * CPL >= c) return;     // padding
    double be[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) be[q] = st->beta[li * CPL + q];
    const long long total = (long long)n * LPR;
    const long long stride = (long long)gridDim.x * kCgThreads;
#pragma unroll 4
    for (long long i = (long long)blockIdx.x * kCgThreads + threadIdx.x; i < total; i += stride) {
        double rv[CPL], pv[CPL];
        ldv<CPL>(r + i * CPL, rv);
        ldv_rw<CPL>(p + i * CPL, pv);
#pragma unroll
        for (int q = 0; q < CPL; ++q) pv[q] = rv[q] + be[q] * pv[q];
        stv<CPL>(p + i * CPL, pv);
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

// Work lists of the long rows (CgLong), from the row pointers on the host: one download of n + 1 integers per solve.
struct CgLongPlan {
    int4 *d_items = nullptr, *d_hubs = nullptr;
    CgLong L{nullptr, nullptr, 0, 0};
    ~CgLongPlan() { dev_free(d_items); dev_free(d_hubs); }
};

static int cg_plan_long_rows(const int *d_rowptr, int64_t n, CgLongPlan &P, cudaStream_t st)
{
    std::vector<int> rp((size_t)n + 1);
    GLB_CUDA(cudaMemcpyAsync(rp.data(), d_rowptr, sizeof(int) * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    std::vector<int4> items, hubs;
    for (int64_t r = 0; r < n; ++r) {
        const int beg = rp[(size_t)r], len = rp[(size_t)r + 1] - beg;
        if (len > kCgItem) hubs.push_back(make_int4((int)r, beg, beg + len, 0));
        else if (len > kCgLongRow) items.push_back(make_int4((int)r, beg, beg + len, 0));
    }
    // longest hubs first: CTA m % grid gets hub m, so the big ones are spread before the small ones fill in
    std::stable_sort(hubs.begin(), hubs.end(), [](const int4 &a, const int4 &b) { return a.z - a.y > b.z - b.y; });
    if (!items.empty()) {
        GLB_CUDA(dev_alloc(&P.d_items, sizeof(int4) * items.size()));
        GLB_CUDA(cudaMemcpyAsync(P.d_items, items.data(), sizeof(int4) * items.size(), cudaMemcpyHostToDevice, st));
    }
    if (!hubs.empty()) {
        GLB_CUDA(dev_alloc(&P.d_hubs, sizeof(int4) * hubs.size()));
        GLB_CUDA(cudaMemcpyAsync(P.d_hubs, hubs.data(), sizeof(int4) * hubs.size(), cudaMemcpyHostToDevice, st));
    }
    GLB_CUDA(cudaStreamSynchronize(st));                     // the staging vectors are locals
    P.L = CgLong{P.d_items, P.d_hubs, (int)items.size(), (int)hubs.size()};
    return 0;
}

template <int LPR, int CPL>
static int cg_run(const int *rp, const int *col, const double *val, int64_t n, const double *b, const double *x0, int c,
                  double tol, int64_t max_iter, double *x, void *work, int64_t *iters_out, double *err_out, int *launches,
                  cudaStream_t st)
{
    constexpr int LDU = LPR * CPL;
    const int grid = cg_grid();
    unsigned char *w = (unsigned char *)(((uintptr_t)work + 255) & ~(uintptr_t)255);
    const size_t vec = a256((size_t)n * LDU * sizeof(double));
    double *r = (double *)w;  w += vec;
    double *p = (double *)w;  w += vec;
    double *Ap = (double *)w; w += vec;
    double *partial = (double *)w; w += a256((size_t)grid * LDU * sizeof(double));
    CgState *state = (CgState *)w;
    CgLongPlan plan;
    int rc = cg_plan_long_rows(rp, n, plan, st);
    if (rc) return rc;
    const CgLong L = plan.L;
    // K1 holds 8 gathers per lane in registers: fewer resident CTAs than the streaming kernels, and one full wave only
    int per_sm = 0;
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_spmm_dot<LPR, CPL, 0>, kCgThreads, 0));
    const int grid1 = std::max(1, std::min(grid, std::max(1, per_sm) * sm_count()));
    int nl = 0;
    cg_state_init<<<1, 128, 0, st>>>(state); ++nl;
    if (x0) {
        GLB_CUDA(cudaMemcpyAsync(x, x0, (size_t)n * LDU * sizeof(double), cudaMemcpyDeviceToDevice, st));
        cg_spmm_dot<LPR, CPL, 1><<<grid1, kCgThreads, 0, st>>>(rp, col, val, x, Ap, (int)n, c, partial, state, L); ++nl;
        cg_update<LPR, CPL, true><<<grid, kCgThreads, 0, st>>>(x, r, p, Ap, b, (int)n, c, tol, partial, state); ++nl;
    } else {
        GLB_CUDA(cudaMemsetAsync(x, 0, (size_t)n * LDU * sizeof(double), st));
        cg_update<LPR, CPL, true><<<grid, kCgThreads, 0, st>>>(x, r, p, nullptr, b, (int)n, c, tol, partial, state); ++nl;
    }
    GLB_LAUNCH_CHECK();
    CgState h;
    int64_t enq = 0;                       // iterations enqueued so far
    h.done = 0; h.iters = 0; h.err = 1.0;
    // the reference's loop test is evaluated before every iteration with err initialised to 1 (utils.py:519-521)
    if (!(1.0 > tol) || max_iter <= 0) {
        GLB_CUDA(cudaStreamSynchronize(st));                 // the work list is freed on return
        if (iters_out) *iters_out = 0; if (err_out) *err_out = 1.0; if (launches) *launches += nl; return 0;
    }
    int64_t ramp = 16;                     // batches of 16, 32, 64, 64, ...: a solve that ends early leaves few idle launches behind
    while (!h.done && enq < max_iter) {
        int64_t batch = max_iter - enq;
        if (batch > ramp) batch = ramp;
        if (ramp < kCgHist) ramp *= 2;
        for (int64_t i = 0; i < batch; ++i) {
            cg_spmm_dot<LPR, CPL, 0><<<grid1, kCgThreads, 0, st>>>(rp, col, val, p, Ap, (int)n, c, partial, state, L);
            cg_update<LPR, CPL, false><<<grid, kCgThreads, 0, st>>>(x, r, p, Ap, b, (int)n, c, tol, partial, state);
            cg_direction<LPR, CPL><<<grid, kCgThreads, 0, st>>>(r, p, (int)n, c, state);
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
        case 4: return cg_run<2, 2>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
        case 8: return cg_run<4, 2>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
        case 16: return cg_run<8, 2>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
        case 32: return cg_run<16, 2>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
        case 64: return cg_run<32, 2>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
        default: return cg_run<32, 4>(d_rowptr, d_col, d_val, n, d_b, d_x0, c, tol, max_iter, d_x, d_work, iters, err, launches, st);
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
