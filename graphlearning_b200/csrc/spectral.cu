// spectral.cu - fp64 block kernels behind graph.eigen_decomp / utils.randomized_svd on sm_100a.
//
// The reference computes the leading singular pairs of A = D^-1/2 W D^-1/2 (or 2 max(deg) I - L) with ARPACK svds
// (graphlearning/graph.py:728-765) or with a randomized SVD whose cost is the repeated product Y <- A (A^T Y) on a
// block of c = 2k columns (graphlearning/utils.py:611-621).  Both reduce to three block operations on tall-skinny
// row-major n x ld fp64 matrices with c <= 256 columns, all HBM-streaming:
//
//   glb_spmm_f64        Z = alpha * A X  +  beta * Y1 * diag(bcol)  +  gamma * Y2     one warp per matrix row, lanes
//                       own column pairs (16-byte loads, a neighbour row is read as one coalesced segment), the
//                       three-term recurrence of the Chebyshev filter is fused into the epilogue
//   glb_gram_f64        G = X^T Y  (c1 x c2)   shared-memory tiles, per-CTA partial sums reduced in a fixed order
//   glb_right_mul_f64   Y = X S    (n x c1 times c1 x c2)   S and a row tile of X staged in shared memory
//
// Sums inside a row / a CTA run in a fixed order, so results are reproducible run to run.
#include <algorithm>
#include "common.cuh"

namespace glb {
namespace {

// ---------------------------------------------------------------------------------------------------------
// SpMM: one warp per row, lane l owns columns {2l, 2l+1} + 64 p, p < CP; sums run in stored order
// ---------------------------------------------------------------------------------------------------------
template <int CP, int UN>
__global__ void __launch_bounds__(256)
spmm_f64_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val, long long n,
                const double *__restrict__ X, int ldx, double *Z, int ldz, int c, double alpha,
                const double *Y1, int ldy1, double beta, const double *__restrict__ bcol,
                const double *Y2, int ldy2, double gamma)      // Z may alias Y1 / Y2 (in-place recurrences): no __restrict__ on those
{
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    double2 acc[CP];
#pragma unroll
    for (int p = 0; p < CP; ++p) acc[p] = make_double2(0.0, 0.0);
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    for (int k = beg; k < end; k += UN) {                       // UN neighbour rows in flight (UN * CP 16-byte loads per lane)
        double a[UN];
        double2 v[UN][CP];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const bool on = k + u < end;
            a[u] = on ? __ldg(val + k + u) : 0.0;
            const double *x = X + (size_t)(on ? __ldg(col + k + u) : 0) * ldx;
#pragma unroll
            for (int p = 0; p < CP; ++p) {
                const int cc = 2 * lane + 64 * p;
                v[u][p] = make_double2(0.0, 0.0);
                if (on && cc < c) v[u][p] = *reinterpret_cast<const double2 *>(x + cc);
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u)
#pragma unroll
            for (int p = 0; p < CP; ++p) {
                acc[p].x = fma(a[u], v[u][p].x, acc[p].x);
                acc[p].y = fma(a[u], v[u][p].y, acc[p].y);
            }
    }
#pragma unroll
    for (int p = 0; p < CP; ++p) {
        const int cc = 2 * lane + 64 * p;
        if (cc >= c) continue;
        double2 r = make_double2(alpha * acc[p].x, alpha * acc[p].y);
        if (Y1) {
            const double2 y = *reinterpret_cast<const double2 *>(Y1 + (size_t)row * ldy1 + cc);
            const double b0 = bcol ? beta * bcol[cc] : beta, b1 = bcol ? (cc + 1 < c ? beta * bcol[cc + 1] : 0.0) : beta;
            r.x = fma(b0, y.x, r.x); r.y = fma(b1, y.y, r.y);
        }
        if (Y2) {
            const double2 y = *reinterpret_cast<const double2 *>(Y2 + (size_t)row * ldy2 + cc);
            r.x = fma(gamma, y.x, r.x); r.y = fma(gamma, y.y, r.y);
        }
        if (cc + 1 >= c) r.y = 0.0;                             // odd c: keep the padding column zero
        *reinterpret_cast<double2 *>(Z + (size_t)row * ldz + cc) = r;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Gram: G = X^T Y.  CTA b reduces rows [b*chunk, (b+1)*chunk) into part[b] (c1 x c2); 256 threads as 16 x 16,
// thread (ty,tx) owns outputs (ty + 16 a, tx + 16 b), a < TA, b < TB.
// ---------------------------------------------------------------------------------------------------------
constexpr int kGramRows = 16;      // rows per shared-memory tile
template <int TA, int TB>
__global__ void __launch_bounds__(256)
gram_partial_kernel(const double *__restrict__ X, int ldx, int c1, const double *__restrict__ Y, int ldy, int c2, long long n,
                    long long chunk, double *__restrict__ part)
{
    extern __shared__ double sm[];
    double *sx = sm, *sy = sm + kGramRows * (16 * TA);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[TA][TB];
#pragma unroll
    for (int a = 0; a < TA; ++a)
#pragma unroll
        for (int b = 0; b < TB; ++b) acc[a][b] = 0.0;
    const long long r0 = (long long)blockIdx.x * chunk, r1 = min(n, r0 + chunk);
    for (long long rb = r0; rb < r1; rb += kGramRows) {
        const int nr = (int)min((long long)kGramRows, r1 - rb);
        for (int i = threadIdx.x; i < kGramRows * 16 * TA; i += 256) {
            const int r = i / (16 * TA), cc = i % (16 * TA);
            sx[i] = (r < nr && cc < c1) ? X[(size_t)(rb + r) * ldx + cc] : 0.0;
        }
        for (int i = threadIdx.x; i < kGramRows * 16 * TB; i += 256) {
            const int r = i / (16 * TB), cc = i % (16 * TB);
            sy[i] = (r < nr && cc < c2) ? Y[(size_t)(rb + r) * ldy + cc] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < kGramRows; ++r) {
            double xa[TA], yb[TB];
#pragma unroll
            for (int a = 0; a < TA; ++a) xa[a] = sx[r * 16 * TA + ty + 16 * a];
#pragma unroll
            for (int b = 0; b < TB; ++b) yb[b] = sy[r * 16 * TB + tx + 16 * b];
#pragma unroll
            for (int a = 0; a < TA; ++a)
#pragma unroll
                for (int b = 0; b < TB; ++b) acc[a][b] = fma(xa[a], yb[b], acc[a][b]);
        }
        __syncthreads();
    }
    double *out = part + (size_t)blockIdx.x * c1 * c2;
#pragma unroll
    for (int a = 0; a < TA; ++a)
#pragma unroll
        for (int b = 0; b < TB; ++b) {
            const int i = ty + 16 * a, j = tx + 16 * b;
            if (i < c1 && j < c2) out[(size_t)i * c2 + j] = acc[a][b];
        }
}

__global__ void __launch_bounds__(256) gram_reduce_kernel(const double *__restrict__ part, int nparts, int sz, double *__restrict__ G)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sz) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(size_t)p * sz + i];
    G[i] = s;
}

// ---------------------------------------------------------------------------------------------------------
// Right multiply: Y = X S.  CTA = 64 rows, the inner dimension is walked in chunks of 16 staged in shared memory;
// thread t owns rows t/8 and t/8 + 32 and output columns t%8 + 8 j, j < TJ.
// ---------------------------------------------------------------------------------------------------------
constexpr int kRmRows = 64, kRmK = 16;
template <int TJ>
__global__ void __launch_bounds__(256)
right_mul_kernel(const double *__restrict__ X, int ldx, long long n, int c1, const double *__restrict__ S, int lds, int c2,
                 double *__restrict__ Y, int ldy, int wlim)
{
    // S: c1 x c2 panel with row stride lds; columns c2..wlim-1 of Y are written as zeros
    constexpr int W = 8 * TJ;
    __shared__ double ss[kRmK * W];            // chunk of S, zero padded to W columns
    __shared__ double sx[kRmRows * (kRmK + 1)];
    const long long r0 = (long long)blockIdx.x * kRmRows;
    const int r = threadIdx.x >> 3, cg = threadIdx.x & 7;
    double acc0[TJ], acc1[TJ];
#pragma unroll
    for (int j = 0; j < TJ; ++j) acc0[j] = acc1[j] = 0.0;
    for (int k0 = 0; k0 < c1; k0 += kRmK) {
        const int kc = min(kRmK, c1 - k0);
        for (int i = threadIdx.x; i < kRmK * W; i += 256) {
            const int kk = i / W, cc = i % W;
            ss[i] = (kk < kc && cc < c2) ? S[(size_t)(k0 + kk) * lds + cc] : 0.0;
        }
        for (int i = threadIdx.x; i < kRmRows * kRmK; i += 256) {
            const int rr = i / kRmK, kk = i % kRmK;
            sx[rr * (kRmK + 1) + kk] = (r0 + rr < n && kk < kc) ? X[(size_t)(r0 + rr) * ldx + k0 + kk] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < kRmK; ++kk) {
            const double x0 = sx[r * (kRmK + 1) + kk], x1 = sx[(r + 32) * (kRmK + 1) + kk];
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
                const double sv = ss[kk * W + cg + 8 * j];
                acc0[j] = fma(x0, sv, acc0[j]);
                acc1[j] = fma(x1, sv, acc1[j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < TJ; ++j) {
        const int cc = cg + 8 * j;
        if (cc >= wlim) continue;
        if (r0 + r < n) Y[(size_t)(r0 + r) * ldy + cc] = cc < c2 ? acc0[j] : 0.0;           // padding columns are zeroed
        if (r0 + r + 32 < n) Y[(size_t)(r0 + r + 32) * ldy + cc] = cc < c2 ? acc1[j] : 0.0;
    }
}

template <int TA, int TB>
int launch_gram(const double *X, int ldx, int c1, const double *Y, int ldy, int c2, long long n, int nparts, long long chunk,
                double *part, cudaStream_t st)
{
    const size_t smem = (size_t)kGramRows * 16 * (TA + TB) * sizeof(double);
    gram_partial_kernel<TA, TB><<<nparts, 256, smem, st>>>(X, ldx, c1, Y, ldy, c2, n, chunk, part);
    return 0;
}

template <int TJ>
int launch_right_mul(const double *X, int ldx, long long n, int c1, const double *S, int lds, int c2, double *Y, int ldy, int wlim,
                     cudaStream_t st)
{
    right_mul_kernel<TJ><<<ceil_div(n, kRmRows), 256, 0, st>>>(X, ldx, n, c1, S, lds, c2, Y, ldy, wlim);
    return 0;
}

}  // namespace
}  // namespace glb

using namespace glb;

static const int kMaxBlockCols = 256;

extern "C" GLB_API int glb_spmm_f64(const int32_t *d_rowptr, const int32_t *d_col, const double *d_val, int64_t n,
                                    const double *d_X, int ldx, double *d_Z, int ldz, int c, double alpha,
                                    const double *d_Y1, int ldy1, double beta, const double *d_bcol, const double *d_Y2,
                                    int ldy2, double gamma, void *stream)
{
    GLB_CHECK_ARG(d_rowptr && d_X && d_Z, "null pointer");
    GLB_CHECK_ARG(n > 0 && c > 0 && c <= kMaxBlockCols, "block width must be 1..256");
    const int need = (c + 1) & ~1;
    GLB_CHECK_ARG(ldx >= need && ldz >= need && !(ldx & 1) && !(ldz & 1), "leading dimensions must be even and >= c rounded up to even");
    GLB_CHECK_ARG((!d_Y1 || (ldy1 >= need && !(ldy1 & 1))) && (!d_Y2 || (ldy2 >= need && !(ldy2 & 1))), "bad leading dimension of Y1/Y2");
    GLB_CHECK_ARG(d_Z != d_X, "Z must not alias X");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = ceil_div(n * 32, 256);
    const int cp = (c + 63) / 64;
#define GLB_SPMM(CP, UN) spmm_f64_kernel<CP, UN><<<grid, 256, 0, st>>>(d_rowptr, d_col, d_val, n, d_X, ldx, d_Z, ldz, c, alpha, d_Y1, ldy1, beta, d_bcol, d_Y2, ldy2, gamma)
    if (cp == 1) GLB_SPMM(1, 8); else if (cp == 2) GLB_SPMM(2, 4); else if (cp == 3) GLB_SPMM(3, 2); else GLB_SPMM(4, 2);
#undef GLB_SPMM
    GLB_LAUNCH_CHECK();
    return 0;
}

extern "C" GLB_API int64_t glb_gram_work_bytes(int c1, int c2)
{
    return ((int64_t)sm_count() * 2 * c1 * c2 + 128 * 128) * (int64_t)sizeof(double);
}

extern "C" GLB_API int glb_gram_f64(const double *d_X, int ldx, int c1, const double *d_Y, int ldy, int c2, int64_t n,
                                    double *d_G, void *d_work, int64_t work_bytes, void *stream)
{
    GLB_CHECK_ARG(d_X && d_Y && d_G && d_work, "null pointer");
    GLB_CHECK_ARG(n > 0 && c1 > 0 && c2 > 0 && c1 <= kMaxBlockCols && c2 <= kMaxBlockCols && ldx >= c1 && ldy >= c2, "bad shape");
    GLB_CHECK_ARG(work_bytes >= glb_gram_work_bytes(c1, c2), "workspace too small (glb_gram_work_bytes)");
    cudaStream_t st = (cudaStream_t)stream;
    int nparts = sm_count() * 2;
    long long chunk = (n + nparts - 1) / nparts;
    chunk = (chunk + kGramRows - 1) / kGramRows * kGramRows;
    nparts = (int)((n + chunk - 1) / chunk);
    double *part = (double *)d_work;
    const int ta = (c1 + 15) / 16, tb = (c2 + 15) / 16;
    // tile shapes: round both up to a power of two (1, 2, 4, 8, 16); 16 x 16 accumulators would spill, so widths above
    // 128 run as several column panels of Y
    auto up = [](int t) { int r = 1; while (r < t) r <<= 1; return r; };
    const int TA = up(ta);
    GLB_CHECK_ARG(TA <= 16, "c1 too wide");
    // process Y in panels of <= 128 columns and X in panels of <= 128 rows of G
    for (int i0 = 0; i0 < c1; i0 += 128) {
        const int ci = std::min(128, c1 - i0);
        for (int j0 = 0; j0 < c2; j0 += 128) {
            const int cj = std::min(128, c2 - j0);
            const int A = up((ci + 15) / 16), B = up((cj + 15) / 16);
            double *pp = part;     // partials of this panel: nparts x ci x cj (fits: ci*cj <= c1*c2)
#define GLB_GRAM(TA_, TB_) launch_gram<TA_, TB_>(d_X + i0, ldx, ci, d_Y + j0, ldy, cj, n, nparts, chunk, pp, st)
            if (A == 1 && B == 1) GLB_GRAM(1, 1); else if (A == 1 && B == 2) GLB_GRAM(1, 2); else if (A == 1 && B == 4) GLB_GRAM(1, 4); else if (A == 1 && B == 8) GLB_GRAM(1, 8);
            else if (A == 2 && B == 1) GLB_GRAM(2, 1); else if (A == 2 && B == 2) GLB_GRAM(2, 2); else if (A == 2 && B == 4) GLB_GRAM(2, 4); else if (A == 2 && B == 8) GLB_GRAM(2, 8);
            else if (A == 4 && B == 1) GLB_GRAM(4, 1); else if (A == 4 && B == 2) GLB_GRAM(4, 2); else if (A == 4 && B == 4) GLB_GRAM(4, 4); else if (A == 4 && B == 8) GLB_GRAM(4, 8);
            else if (A == 8 && B == 1) GLB_GRAM(8, 1); else if (A == 8 && B == 2) GLB_GRAM(8, 2); else if (A == 8 && B == 4) GLB_GRAM(8, 4); else GLB_GRAM(8, 8);
#undef GLB_GRAM
            GLB_LAUNCH_CHECK();
            if (ci == c1 && cj == c2) {
                gram_reduce_kernel<<<ceil_div(c1 * c2, 256), 256, 0, st>>>(pp, nparts, c1 * c2, d_G);
            } else {
                // panel: reduce into a dense ci x cj scratch at the end of the workspace, then scatter rows into G
                double *scratch = part + (size_t)nparts * ci * cj;
                gram_reduce_kernel<<<ceil_div(ci * cj, 256), 256, 0, st>>>(pp, nparts, ci * cj, scratch);
                GLB_CUDA(cudaMemcpy2DAsync(d_G + (size_t)i0 * c2 + j0, (size_t)c2 * sizeof(double), scratch, (size_t)cj * sizeof(double),
                                           (size_t)cj * sizeof(double), ci, cudaMemcpyDeviceToDevice, st));
            }
            GLB_LAUNCH_CHECK();
        }
    }
    return 0;
}

extern "C" GLB_API int glb_right_mul_f64(const double *d_X, int ldx, int64_t n, int c1, const double *d_S, int c2, double *d_Y,
                                         int ldy, void *stream)
{
    GLB_CHECK_ARG(d_X && d_S && d_Y, "null pointer");
    GLB_CHECK_ARG(n > 0 && c1 > 0 && c2 > 0 && c1 <= kMaxBlockCols && c2 <= kMaxBlockCols && ldx >= c1 && ldy >= c2, "bad shape");
    GLB_CHECK_ARG(d_Y != d_X, "Y must not alias X");
    cudaStream_t st = (cudaStream_t)stream;
    // panels of <= 128 output columns (16 accumulators per row and thread); the last panel also zeroes the padding
    int rc = 0;
    for (int j0 = 0; j0 < c2 && !rc; j0 += 128) {
        const int cj = std::min(128, c2 - j0);
        const bool last = j0 + cj == c2;
        const int ldp = last ? std::min(ldy - j0, 128) : cj;             // columns this panel may write
        const int tj = (std::max(cj, ldp) + 7) / 8;
        // a panel sees S and Y shifted by j0 columns; its "c2" is cj, its padding limit ldp
        if (tj <= 2) rc = launch_right_mul<2>(d_X, ldx, n, c1, d_S + j0, c2, cj, d_Y + j0, ldy, ldp, st);
        else if (tj <= 4) rc = launch_right_mul<4>(d_X, ldx, n, c1, d_S + j0, c2, cj, d_Y + j0, ldy, ldp, st);
        else if (tj <= 8) rc = launch_right_mul<8>(d_X, ldx, n, c1, d_S + j0, c2, cj, d_Y + j0, ldy, ldp, st);
        else rc = launch_right_mul<16>(d_X, ldx, n, c1, d_S + j0, c2, cj, d_Y + j0, ldy, ldp, st);
    }
    if (rc) return rc;
    GLB_LAUNCH_CHECK();
    return 0;
}
