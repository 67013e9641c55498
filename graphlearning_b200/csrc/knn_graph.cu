// knn_graph.cu - kNN result (indices + kernel weights) -> symmetric CSR weight matrix, on sm_100a.
//
// Replaces the sparse assembly of weightmatrix.knn (reference graphlearning/weightmatrix.py:166-186):
//     W = coo_matrix((weights, (self_ind, knn_ind))).tocsr()      # duplicates summed, columns ascending
//     W = (W + W.transpose()) / 2                                 # symmetrize (gaussian / user kernels; sparse_max and the
//                                                                 # symgaussian rule for the other kernels, :176-181)
//     W.setdiag(0); W.eliminate_zeros()
// - 0.18 s of scipy COO/CSC conversions at n = 70 000, k = 10, four times the GPU kNN search that feeds it.
// The kernel weights themselves (exp(-4 d^2 / d_k^2) etc.) stay numpy on the host so that they are bit-identical to
// the reference's; the assembly is exact arithmetic on them: an entry of the result is fl(w_ij + w_ji) / 2 (addition
// is commutative, the halving exact), so the CSR produced here equals scipy's bit for bit.
//
// One 64-bit key per directed entry, (row << 32 | col) for W and - when symmetrizing - (col << 32 | row) for W^T;
// CUB radix sort by key; every run of equal keys is one output entry (sum of the run, halved when symmetrizing);
// diagonal and exactly-zero entries are dropped; CUB exclusive scan of the keep flags gives the output positions and a
// binary search over the compacted keys the row pointers.  HBM-streaming, ~2 n k 16 bytes per pass.
#include <cub/cub.cuh>
#include <vector>
#include "common.cuh"

namespace glb {
namespace {

typedef unsigned long long u64;

__global__ void __launch_bounds__(256)
kg_keys_kernel(const long long *__restrict__ ind, const double *__restrict__ w, long long n, int k, int symmetrize,
               u64 *__restrict__ keys, double *__restrict__ vals, int *bad)
{
    const long long total = n * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / k, j = ind[e];
        if (j < 0 || j >= n) { atomicOr(bad, 1); continue; }
        keys[e] = ((u64)i << 32) | ((u64)j << 1);                 // lowest bit: 0 = entry of W, 1 = entry of W^T
        vals[e] = w[e];
        if (symmetrize) {
            keys[total + e] = ((u64)j << 32) | ((u64)i << 1) | 1ull;
            vals[total + e] = w[e];
        }
    }
}

// head of a run of equal (row, column): value of the output entry, keep flag.  Inside a run the entries of W come before
// those of W^T (origin bit), each group in its original order (stable sort): a = W_ij, b = W^T_ij with duplicates summed
// as scipy's COO -> CSR conversion sums them, then the reference's rule (weightmatrix.py:176-183):
//   mode 1  (a + b) / 2                                    gaussian and user kernels
//   mode 2  max(a, b)                                      utils.sparse_max: 'distance', 'uniform', 'singular'
//   mode 3  b > a ? (a + b) - a : a                        'symgaussian': W + W^T.multiply(W^T > W) - W.multiply(W^T > W)
__global__ void __launch_bounds__(256)
kg_runs_kernel(const u64 *__restrict__ keys, const double *__restrict__ vals, long long m, int mode,
               double *__restrict__ run_val, int *__restrict__ keep)
{
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (long long)gridDim.x * blockDim.x) {
        const u64 key = keys[p] >> 1;
        int kp = 0;
        double s = 0.0;
        if (p == 0 || (keys[p - 1] >> 1) != key) {
            double a = 0.0, b = 0.0;
            bool has_a = false, has_b = false;
            for (long long q = p; q < m && (keys[q] >> 1) == key; ++q) {               // at most two entries unless the kNN list repeats a column
                if (keys[q] & 1ull) { b = has_b ? b + vals[q] : vals[q]; has_b = true; }
                else { a = has_a ? a + vals[q] : vals[q]; has_a = true; }
            }
            if (mode == 0) s = a;
            else if (mode == 1) s = (has_a && has_b ? a + b : (has_a ? a : b)) / 2;
            else if (mode == 2) s = b > a ? b : a;
            else s = b > a ? __dsub_rn(__dadd_rn(a, b), a) : a;
            kp = ((unsigned)(key >> 31) != (unsigned)(key & 0x7fffffffull)) && (s != 0.0);   // setdiag(0) + eliminate_zeros
        }
        run_val[p] = s;
        keep[p] = kp;
    }
}

__global__ void __launch_bounds__(256)
kg_scatter_kernel(const u64 *__restrict__ keys, const double *__restrict__ run_val, const int *__restrict__ keep,
                  const int *__restrict__ pos, long long m, u64 *__restrict__ out_keys, int *__restrict__ col, double *__restrict__ val)
{
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < m; p += (long long)gridDim.x * blockDim.x) {
        if (!keep[p]) continue;
        const int o = pos[p];
        out_keys[o] = keys[p];
        col[o] = (int)((unsigned)keys[p] >> 1);
        val[o] = run_val[p];
    }
}

__global__ void __launch_bounds__(256)
kg_rowptr_kernel(const u64 *__restrict__ keys, long long nnz, long long n, int *__restrict__ rowptr)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (long long)gridDim.x * blockDim.x) {
        long long lo = 0, hi = nnz;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if ((long long)(keys[mid] >> 32) < i) lo = mid + 1; else hi = mid;
        }
        rowptr[i] = (int)lo;
    }
}

struct Arena {
    std::vector<void *> ptrs;
    ~Arena() { for (void *p : ptrs) dev_free(p); }
    template <typename T>
    cudaError_t alloc(T **p, size_t count)
    {
        void *q = nullptr;
        cudaError_t e = dev_alloc(&q, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(q);
        *p = (T *)q;
        return e;
    }
};

int blocks_for(long long work)
{
    long long b = (work + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace
}  // namespace glb

using namespace glb;

extern "C" GLB_API int glb_knn_weights_csr_host(const int64_t *h_ind, const double *h_w, int64_t n, int k, int symmetrize,
                                                int32_t *h_rowptr, int32_t *h_col, double *h_val, int64_t cap, int64_t *nnz_out,
                                                int *launches)
{
    GLB_CHECK_ARG(h_ind && h_w && h_rowptr && h_col && h_val && nnz_out, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31) && k > 0 && n * (long long)k * 2 < (1ll << 31), "size out of range");
    GLB_CHECK_ARG(symmetrize >= 0 && symmetrize <= 3, "symmetrize: 0 none, 1 average, 2 max, 3 symgaussian");
    const long long m = n * (long long)k * (symmetrize ? 2 : 1);
    GLB_CHECK_ARG(cap >= m, "output capacity must be at least n*k (2*n*k when symmetrizing)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("glb_knn_weights_csr_host: no CUDA device visible (there is no CPU fallback)");
        return GLB_E_NOGPU;
    }
    cudaStream_t st = 0;
    Arena A;
    long long *ind; double *w, *vals_in, *vals_out, *run_val, *val; u64 *keys_in, *keys_out, *okeys; int *keep, *pos, *col, *rowptr, *bad;
    GLB_CUDA(A.alloc(&ind, (size_t)(n * k)));   GLB_CUDA(A.alloc(&w, (size_t)(n * k)));
    GLB_CUDA(A.alloc(&keys_in, (size_t)m));     GLB_CUDA(A.alloc(&keys_out, (size_t)m));
    GLB_CUDA(A.alloc(&vals_in, (size_t)m));     GLB_CUDA(A.alloc(&vals_out, (size_t)m));
    GLB_CUDA(A.alloc(&run_val, (size_t)m));     GLB_CUDA(A.alloc(&keep, (size_t)m));   GLB_CUDA(A.alloc(&pos, (size_t)m + 1));
    GLB_CUDA(A.alloc(&okeys, (size_t)m));       GLB_CUDA(A.alloc(&col, (size_t)m));    GLB_CUDA(A.alloc(&val, (size_t)m));
    GLB_CUDA(A.alloc(&rowptr, (size_t)n + 1));  GLB_CUDA(A.alloc(&bad, 1));
    GLB_CUDA(cudaMemcpyAsync(ind, h_ind, (size_t)(n * k) * sizeof(long long), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(w, h_w, (size_t)(n * k) * sizeof(double), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
    size_t sort_bytes = 0, scan_bytes = 0;
    int hi_bits = 1;
    while ((1ll << hi_bits) < n) ++hi_bits;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys_in, keys_out, vals_in, vals_out, (int)m, 0, 32 + hi_bits, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, keep, pos, (int)m, st);
    unsigned char *temp;
    GLB_CUDA(A.alloc(&temp, std::max(sort_bytes, scan_bytes)));
    kg_keys_kernel<<<blocks_for(n * k), 256, 0, st>>>(ind, w, n, k, symmetrize, keys_in, vals_in, bad);
    GLB_CUDA(cub::DeviceRadixSort::SortPairs(temp, sort_bytes, keys_in, keys_out, vals_in, vals_out, (int)m, 0, 32 + hi_bits, st));
    kg_runs_kernel<<<blocks_for(m), 256, 0, st>>>(keys_out, vals_out, m, symmetrize, run_val, keep);
    GLB_CUDA(cub::DeviceScan::ExclusiveSum(temp, scan_bytes, keep, pos, (int)m, st));
    kg_scatter_kernel<<<blocks_for(m), 256, 0, st>>>(keys_out, run_val, keep, pos, m, okeys, col, val);
    GLB_LAUNCH_CHECK();
    int last_pos = 0, last_keep = 0, h_bad = 0;
    GLB_CUDA(cudaMemcpyAsync(&last_pos, pos + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaMemcpyAsync(&last_keep, keep + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    if (h_bad) { set_error("glb_knn_weights_csr_host: neighbour index out of range"); return GLB_E_INVALID; }
    const long long nnz = (long long)last_pos + last_keep;
    kg_rowptr_kernel<<<blocks_for(n + 1), 256, 0, st>>>(okeys, nnz, n, rowptr);
    GLB_LAUNCH_CHECK();
    GLB_CUDA(cudaMemcpyAsync(h_rowptr, rowptr, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (nnz) {
        GLB_CUDA(cudaMemcpyAsync(h_col, col, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToHost, st));
        GLB_CUDA(cudaMemcpyAsync(h_val, val, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    GLB_CUDA(cudaStreamSynchronize(st));
    *nnz_out = nnz;
    if (launches) *launches = 6;
    return 0;
}
