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
// SpMM: one warp per row, lane l owns columns {2l, 2l+1} + 64 p, p < CP
// ---------------------------------------------------------------------------------------------------------
template <int CP>
__global__ void __launch_bounds__(256)
spmm_f64_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val, long long n,
                const double *__restrict__ X, int ldx, double *__restrict__ Z, int ldz, int c, double alpha,
                const double *__restrict__ Y1, int ldy1, double beta, const double *__restrict__ bcol,
                const double *__restrict__ Y2, int ldy2, double gamma)
{
    const int lane = threadIdx.x & 31;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    double2 acc[CP];
#pragma unroll
    for (int p = 0; p < CP; ++p) acc[p] = make_double2(0.0, 0.0);
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    int k = beg;
    for (; k + 1 < end; k += 2) {                               // two neighbour rows in flight
        const int j0 = __ldg(col + k), j1 = __ldg(col + k + 1);
        const double a0 = __ldg(val + k), a1 = __ldg(val + k + 1);
        const double *x0 = X + (size_t)j0 * ldx, *x1 = X + (size_t)j1 * ldx;
        double2 v0[CP], v1[CP];
#pragma unroll
        for (int p = 0; p < CP; ++p) {
            const int cc = 2 * lane + 64 * p;
            v0[p] = v1[p] = make_double2(0.0, 0.0);
            if (cc < c) { v0[p] = *reinterpret_cast<const double2 *>(x0 + cc); v1[p] = *reinterpret_cast<const double2 *>(x1 + cc); }
        }
#pragma unroll
        for (int p = 0; p < CP; ++p) {
            acc[p].x = fma(a0, v0[p].x, acc[p].x); acc[p].y = fma(a0, v0[p].y, acc[p].y);
            acc[p].x = fma(a1, v1[p].x, acc[p].x); acc[p].y = fma(a1, v1[p].y, acc[p].y);
        }
    }
    if (k < end) {
        const int j0 = __ldg(col + k);
        const double a0 = __ldg(val + k);
        const double *x0 = X + (size_t)j0 * ldx;
#pragma unroll
        for (int p = 0; p < CP; ++p) {
            const int cc = 2 * lane + 64 * p;
            if (cc < c) {
                const double2 v = *reinterpret_cast<const double2 *>(x0 + cc);
                acc[p].x = fma(a0, v.x, acc[p].x); acc[p].y = fma(a0, v.y, acc[p].y);
            }
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
// Right multiply: Y = X S.  CTA = 32 rows; thread t: row t/8, output columns t%8 + 8 j.
// ---------------------------------------------------------------------------------------------------------
template <int TJ>
__global__ void __launch_bounds__(256)
right_mul_kernel(const double *__restrict__ X, int ldx, long long n, int c1, const double *__restrict__ S, int c2,
                 double *__restrict__ Y, int ldy)
{
    extern __shared__ double sm[];
    double *ss = sm;                       // c1 x (8*TJ), zero padded
    double *sx = sm + (size_t)c1 * 8 * TJ; // 32 x c1
    const int W = 8 * TJ;
    for (int i = threadIdx.x; i < c1 * W; i += 256) {
        const int r = i / W, cc = i % W;
        ss[i] = cc < c2 ? S[(size_t)r * c2 + cc] : 0.0;
    }
    const long long r0 = (long long)blockIdx.x * 32;
    for (int i = threadIdx.x; i < 32 * c1; i += 256) {
        const int r = i / c1, cc = i % c1;
        sx[i] = (r0 + r < n) ? X[(size_t)(r0 + r) * ldx + cc] : 0.0;
    }
    __syncthreads();
    const int r = threadIdx.x >> 3, cg = threadIdx.x & 7;
    double acc[TJ];
#pragma unroll
    for (int j = 0; j < TJ; ++j) acc[j] = 0.0;
    for (int kk = 0; kk < c1; ++kk) {
        const double x = sx[r * c1 + kk];
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[j] = fma(x, ss[kk * W + cg + 8 * j], acc[j]);
    }
    if (r0 + r < n) {
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
            const int cc = cg + 8 * j;
            if (cc < ldy) Y[(size_t)(r0 + r) * ldy + cc] = cc < c2 ? acc[j] : 0.0;      // padding columns are zeroed
        }
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
int launch_right_mul(const double *X, int ldx, long long n, int c1, const double *S, int c2, double *Y, int ldy, cudaStream_t st)
{
    const size_t smem = ((size_t)c1 * 8 * TJ + 32 * (size_t)c1) * sizeof(double);
    GLB_CUDA(cudaFuncSetAttribute(right_mul_kernel<TJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    right_mul_kernel<TJ><<<ceil_div(n, 32), 256, smem, st>>>(X, ldx, n, c1, S, c2, Y, ldy);
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
#define GLB_SPMM(CP) spmm_f64_kernel<CP><<<grid, 256, 0, st>>>(d_rowptr, d_col, d_val, n, d_X, ldx, d_Z, ldz, c, alpha, d_Y1, ldy1, beta, d_bcol, d_Y2, ldy2, gamma)
    if (cp == 1) GLB_SPMM(1); else if (cp == 2) GLB_SPMM(2); else if (cp == 3) GLB_SPMM(3); else GLB_SPMM(4);
#undef GLB_SPMM
    GLB_LAUNCH_CHECK();
    return 0;
}

extern "C" GLB_API int64_t glb_gram_work_bytes(int c1, int c2)
{
    return (int64_t)sm_count() * 2 * c1 * c2 * (int64_t)sizeof(double);
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
    const int tj = (std::max(c2, std::min(ldy, kMaxBlockCols)) + 7) / 8;      // cover the padding columns of Y too
    int rc;
    if (tj <= 4) rc = launch_right_mul<4>(d_X, ldx, n, c1, d_S, c2, d_Y, ldy, st);
    else if (tj <= 8) rc = launch_right_mul<8>(d_X, ldx, n, c1, d_S, c2, d_Y, ldy, st);
    else if (tj <= 16) rc = launch_right_mul<16>(d_X, ldx, n, c1, d_S, c2, d_Y, ldy, st);
    else rc = launch_right_mul<32>(d_X, ldx, n, c1, d_S, c2, d_Y, ldy, st);
    if (rc) return rc;
    GLB_LAUNCH_CHECK();
    return 0;
}
