// mbo.cu - device pieces of the label post-processing that follows the iterate: volume-constrained label projection
// (ssl.volume_label_projection, reference graphlearning/ssl.py:172-209, the projection step of PoissonMBO :826-829 and of
// every model with class_priors :476-477), one-hot encoding and the max-norm reductions of the fixed-point loops
// (graph.page_rank, graphlearning/graph.py:1406-1410).
//
// volume_label_projection is a loop of up to 10^4 rounds of  predict -> class sizes -> weight update  over the n x k score
// matrix; on the host each round is an argmax pass in numpy (1.4 ms at n = 70 000).  Here the whole loop runs in ONE launch:
// a single CTA of 1024 threads walks the rows (the scores stay in L2, 5.6 MB; no grid-wide barrier is needed), counts the
// classes in shared memory and thread 0 updates the k weights between two block barriers - exactly the reference's fp64
// operations in the reference's order (no fused multiply-add), so the weights, the labels and the number of rounds are
// those of the reference.
#include <float.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include "common.cuh"

namespace glb {

constexpr int kProjThreads = 1024;
constexpr int kProjMaxK = 64;

// labels = argmax_j (scores[i][j] * w[j]) with scores = (prob - min) / max(prob - min)   (ssl.py:257-264; argmin when the
// model is a dissimilarity).  first maximum wins, NaN never wins unless it comes first (numpy argmax semantics: a NaN is
// treated as the maximum) - handled like numpy: a NaN entry wins as soon as it is met.
__device__ __forceinline__ int row_label(const double *__restrict__ row, int k, double pmin, double inv_range_den, const double *w, bool similarity)
{
    int best = 0;
    double bv = __dmul_rn(__ddiv_rn(__dsub_rn(row[0], pmin), inv_range_den), w[0]);
    for (int j = 1; j < k; ++j) {
        const double v = __dmul_rn(__ddiv_rn(__dsub_rn(row[j], pmin), inv_range_den), w[j]);
        if (bv != bv) break;                                 // numpy: the first NaN is the arg-extremum
        if (similarity ? (v > bv || v != v) : (v < bv || v != v)) { bv = v; best = j; }
    }
    return best;
}

__global__ void __launch_bounds__(kProjThreads, 1)
volume_projection_kernel(const double *__restrict__ prob, long long n, int k, int ld, const double *__restrict__ priors, int similarity,
                         int max_rounds, double tol, double *__restrict__ weights, long long *__restrict__ labels, double *__restrict__ err_out,
                         int *__restrict__ rounds_out)
{
    __shared__ double s_w[kProjMaxK];
    __shared__ unsigned s_cnt[kProjMaxK];
    __shared__ double s_red[kProjThreads / 32];
    __shared__ double s_min, s_den, s_err;
    const int tid = threadIdx.x;
    // global min and max of prob (ssl.py:257-258): scores = prob - min; scores = scores / max(scores)
    double lo = INFINITY, hi = -INFINITY;
    bool has_nan = false;
    for (long long i = tid; i < n * k; i += kProjThreads) {
        const double v = prob[(i / k) * ld + (i % k)];
        has_nan |= v != v;
        lo = fmin(lo, v); hi = fmax(hi, v);
    }
    for (int off = 16; off > 0; off >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off)); }
    if ((tid & 31) == 0) s_red[tid >> 5] = lo;
    __syncthreads();
    if (tid == 0) { double m = s_red[0]; for (int i = 1; i < kProjThreads / 32; ++i) m = fmin(m, s_red[i]); s_min = m; }
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = hi;
    __syncthreads();
    if (tid == 0) {
        double m = s_red[0];
        for (int i = 1; i < kProjThreads / 32; ++i) m = fmax(m, s_red[i]);
        s_den = __dsub_rn(m, s_min);                         // max(prob - min)
        s_err = 1.0;
    }
    if (tid < k) s_w[tid] = weights[tid];
    __syncthreads();
    (void)has_nan;
    const double pmin = s_min, den = s_den;
    const double dt = similarity ? -0.1 : 0.1;               // ssl.py:191-193
    int round = 0;
    while (round < max_rounds && s_err > tol) {              // ssl.py:198 (evaluated by every thread on the same shared values)
        ++round;
        if (tid < k) s_cnt[tid] = 0u;
        __syncthreads();
        for (long long i = tid; i < n; i += kProjThreads)
            atomicAdd(&s_cnt[row_label(prob + i * ld, k, pmin, den, s_w, similarity != 0)], 1u);
        __syncthreads();
        if (tid == 0) {
            double err = 0.0;
            for (int j = 0; j < k; ++j) {
                const double grad = __dsub_rn(__ddiv_rn((double)s_cnt[j], (double)n), priors[j]);      // np.mean of the one-hot columns
                err = fmax(err, fabs(grad));
                s_w[j] = __dadd_rn(s_w[j], __dmul_rn(dt, grad));
            }
            const double w0 = s_w[0];
            for (int j = 0; j < k; ++j) s_w[j] = __ddiv_rn(s_w[j], w0);
            s_err = err;
        }
        __syncthreads();
    }
    for (long long i = tid; i < n; i += kProjThreads) labels[i] = row_label(prob + i * ld, k, pmin, den, s_w, similarity != 0);
    if (tid < k) weights[tid] = s_w[tid];
    if (tid == 0) { *err_out = s_err; *rounds_out = round; }
}

// dst (n x ld fp64, zero padded) <- one-hot of labels (width k)
__global__ void __launch_bounds__(256) onehot_kernel(const long long *__restrict__ labels, long long n, int k, int ld, double *__restrict__ dst)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * ld; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ld;
        const int j = (int)(i - r * ld);
        dst[i] = (j < k && labels[r] == j) ? 1.0 : 0.0;
    }
}

// out[0] = max_i |x_i - y_i| as the bit pattern of a non-negative double (atomicMax on the bits; NaN sorts above +inf)
__global__ void __launch_bounds__(256) max_abs_diff_kernel(const double *__restrict__ x, const double *__restrict__ y, long long n, int c, int ldx, int ldy,
                                                           unsigned long long *__restrict__ out)
{
    double worst = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * c; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c;
        const int j = (int)(i - r * c);
        const double d = fabs(x[r * ldx + j] - (y ? y[r * ldy + j] : 0.0));
        worst = (d > worst || d != d) ? d : worst;
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, worst, off);
        worst = (__double_as_longlong(o) > __double_as_longlong(worst)) ? o : worst;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(worst));
}

// One step of the centred-kernel fixed point (ssl.centered_kernel._fit, graphlearning/ssl.py:1409-1413):
//     w = (1/alpha) (y - 1 mean_y^T) - u;  w[train] = 0;  err = max |w|;  u = u + w
__global__ void __launch_bounds__(256)
centered_step_kernel(const double *__restrict__ y, int ldy, const double *__restrict__ mean_y, double inv_alpha, double *__restrict__ u, int ldu,
                     const unsigned char *__restrict__ labelled, long long n, int c, unsigned long long *__restrict__ err_bits)
{
    double worst = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * c; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c;
        const int j = (int)(i - r * c);
        const double uo = u[r * ldu + j];
        double w = __dsub_rn(__dmul_rn(inv_alpha, __dsub_rn(y[r * ldy + j], mean_y[j])), uo);
        if (labelled[r]) w = 0.0;
        const double a = fabs(w);
        worst = (a > worst || a != a) ? a : worst;
        u[r * ldu + j] = __dadd_rn(uo, w);
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, worst, off);
        worst = (__double_as_longlong(o) > __double_as_longlong(worst)) ? o : worst;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(err_bits, (unsigned long long)__double_as_longlong(worst));
}

// out = min over the n x c entries, for non-negative matrices (bit patterns order like the values): the `np.min(F) == 0`
// test of the grow loop of clustering.incres (graphlearning/clustering.py:357)
__global__ void __launch_bounds__(256)
min_nonneg_kernel(const double *__restrict__ x, long long n, int c, int ld, unsigned long long *__restrict__ out)
{
    unsigned long long best = ~0ull;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n * c; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c;
        const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(x[r * ld + (i - r * c)]));
        best = b < best ? b : best;
    }
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, off);
        best = o < best ? o : best;
    }
    if ((threadIdx.x & 31) == 0) atomicMin(out, best);
}

// labels[i] = first column holding the row maximum (np.argmax(F, axis=1), clustering.py:361)
__global__ void __launch_bounds__(256)
argmax_rows_kernel(const double *__restrict__ x, long long n, int c, int ld, long long *__restrict__ labels)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        const double *row = x + r * ld;
        int best = 0;
        double bv = row[0];
        bool nan = bv != bv;                          // numpy: the first NaN wins
        for (int j = 1; j < c && !nan; ++j) {
            const double v = row[j];
            if (v != v) { best = j; nan = true; }
            else if (v > bv) { bv = v; best = j; }
        }
        labels[r] = best;
    }
}

}  // namespace glb

using namespace glb;

extern "C" GLB_API int glb_centered_step_f64(const double *d_y, int ldy, const double *d_mean_y, double inv_alpha, double *d_u, int ldu,
                                             const unsigned char *d_labelled, int64_t n, int c, double *h_err, void *stream)
{
    GLB_CHECK_ARG(d_y && d_mean_y && d_u && d_labelled && h_err && n > 0 && c > 0 && ldy >= c && ldu >= c, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    static thread_local unsigned long long *d_out = nullptr;
    if (!d_out) GLB_CUDA(dev_alloc(&d_out, sizeof(unsigned long long)));
    GLB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), st));
    const int blocks = (int)std::min<int64_t>((n * c + 255) / 256, (int64_t)sm_count() * 8);
    centered_step_kernel<<<blocks, 256, 0, st>>>(d_y, ldy, d_mean_y, inv_alpha, d_u, ldu, d_labelled, n, c, d_out);
    unsigned long long bits = 0;
    GLB_CUDA(cudaMemcpyAsync(&bits, d_out, sizeof(bits), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    memcpy(h_err, &bits, sizeof(double));
    return 0;
}

extern "C" GLB_API int glb_min_nonneg_f64(const double *d_x, int64_t n, int c, int ld, double *h_out, void *stream)
{
    GLB_CHECK_ARG(d_x && h_out && n > 0 && c > 0 && ld >= c, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    static thread_local unsigned long long *d_out = nullptr;
    if (!d_out) GLB_CUDA(dev_alloc(&d_out, sizeof(unsigned long long)));
    GLB_CUDA(cudaMemsetAsync(d_out, 0xff, sizeof(unsigned long long), st));
    const int blocks = (int)std::min<int64_t>((n * c + 255) / 256, (int64_t)sm_count() * 8);
    min_nonneg_kernel<<<blocks, 256, 0, st>>>(d_x, n, c, ld, d_out);
    unsigned long long bits = 0;
    GLB_CUDA(cudaMemcpyAsync(&bits, d_out, sizeof(bits), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    memcpy(h_out, &bits, sizeof(double));
    return 0;
}

extern "C" GLB_API int glb_argmax_rows_f64(const double *d_x, int64_t n, int c, int ld, int64_t *d_labels, void *stream)
{
    GLB_CHECK_ARG(d_x && d_labels && n > 0 && c > 0 && ld >= c, "bad argument");
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
    argmax_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_x, n, c, ld, (long long *)d_labels);
    GLB_LAUNCH_CHECK();
    return 0;
}

extern "C" GLB_API int glb_volume_projection(const double *d_prob, int64_t n, int k, int ld, const double *d_priors, int similarity,
                                             int max_rounds, double tol, double *d_weights, int64_t *d_labels, double *d_err, int *d_rounds,
                                             void *stream)
{
    GLB_CHECK_ARG(d_prob && d_priors && d_weights && d_labels && d_err && d_rounds, "null pointer");
    GLB_CHECK_ARG(n > 0 && k > 0 && ld >= k, "bad shape");
    if (k > kProjMaxK) { set_error("glb_volume_projection: k = %d (> %d classes) is not supported", k, kProjMaxK); return GLB_E_UNSUPPORTED; }
    volume_projection_kernel<<<1, kProjThreads, 0, (cudaStream_t)stream>>>(d_prob, n, k, ld, d_priors, similarity, max_rounds, tol, d_weights,
                                                                           (long long *)d_labels, d_err, d_rounds);
    GLB_LAUNCH_CHECK();
    return 0;
}

extern "C" GLB_API int glb_onehot_f64(const int64_t *d_labels, int64_t n, int k, int ld, double *d_dst, void *stream)
{
    GLB_CHECK_ARG(d_labels && d_dst && n > 0 && k > 0 && ld >= k, "bad argument");
    const int blocks = (int)std::min<int64_t>((n * ld + 255) / 256, (int64_t)sm_count() * 16);
    onehot_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const long long *)d_labels, n, k, ld, d_dst);
    GLB_LAUNCH_CHECK();
    return 0;
}

extern "C" GLB_API int glb_max_abs_diff_f64(const double *d_x, const double *d_y, int64_t n, int c, int ldx, int ldy, double *h_out, void *stream)
{
    GLB_CHECK_ARG(d_x && h_out && n > 0 && c > 0 && ldx >= c && (!d_y || ldy >= c), "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    static thread_local unsigned long long *d_out = nullptr;
    if (!d_out) GLB_CUDA(dev_alloc(&d_out, sizeof(unsigned long long)));
    GLB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), st));
    const int blocks = (int)std::min<int64_t>((n * c + 255) / 256, (int64_t)sm_count() * 8);
    max_abs_diff_kernel<<<blocks, 256, 0, st>>>(d_x, d_y, n, c, ldx, ldy, d_out);
    unsigned long long bits = 0;
    GLB_CUDA(cudaMemcpyAsync(&bits, d_out, sizeof(bits), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    memcpy(h_out, &bits, sizeof(double));
    return 0;
}
