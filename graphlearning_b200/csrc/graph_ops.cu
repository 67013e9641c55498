// graph_ops.cu - graph normalisation on the device: degrees, CSR transpose, D^-1 scaling, dtype packing.
//
// Replaces graph.degree_vector / degree_matrix (reference graphlearning/graph.py:108-122, 210-233) and the
// sparse products D*W.transpose(), W.transpose()*D, D*source of ssl.poisson._fit (graphlearning/ssl.py:634-644).
// All of it is HBM-bound streaming work: one coalesced pass per array, fp64 like the reference's setup.
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"

namespace glb {

// deg[i] = sum of the row, accumulated in stored order like scipy's csr_matvec (graph.py:121)
__global__ void __launch_bounds__(256)
degree_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val, int n,
              bool skip_diag, double *__restrict__ deg)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int j = rowptr[i]; j < rowptr[i + 1]; ++j)
            if (!(skip_diag && col[j] == i)) s += val[j];
        deg[i] = s;
    }
}

// keys[j] = (col << 32) | row, payload = j
__global__ void __launch_bounds__(256)
transpose_keys_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, int n,
                      unsigned long long *__restrict__ keys, int *__restrict__ perm)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long row = warp; row < n; row += nwarps)
        for (int j = rowptr[row] + lane; j < rowptr[row + 1]; j += 32) {
            keys[j] = ((unsigned long long)(unsigned)col[j] << 32) | (unsigned long long)(unsigned)row;
            perm[j] = j;
        }
}

__global__ void __launch_bounds__(256)
transpose_fill_kernel(const unsigned long long *__restrict__ keys, const int *__restrict__ perm,
                      const double *__restrict__ val, long long nnz, int *__restrict__ t_col,
                      double *__restrict__ t_val)
{
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += (long long)gridDim.x * blockDim.x) {
        t_col[j] = (int)(keys[j] & 0xffffffffull);
        t_val[j] = val[perm[j]];
    }
}

// t_rowptr[i] = first sorted position whose key's high word is >= i
__global__ void __launch_bounds__(256)
transpose_rowptr_kernel(const unsigned long long *__restrict__ keys, long long nnz, int n, int *__restrict__ t_rowptr)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (long long)gridDim.x * blockDim.x) {
        long long lo = 0, hi = nnz;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if ((long long)(keys[mid] >> 32) < i) lo = mid + 1; else hi = mid;
        }
        t_rowptr[i] = (int)lo;
    }
}

__global__ void __launch_bounds__(256)
poisson_scale_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                     const double *__restrict__ deg, int n, float *__restrict__ P_val, double *__restrict__ RW_val)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long row = warp; row < n; row += nwarps) {
        const double dinv = 1.0 / deg[row];                    // degree_matrix(p=-1): d**-1
        for (int j = rowptr[row] + lane; j < rowptr[row + 1]; j += 32) {
            const int c = col[j];
            const double w = (c == row) ? 0.0 : val[j];        // W - diag(W), ssl.py:615-616
            if (P_val) P_val[j] = (float)(dinv * w);
            if (RW_val) RW_val[j] = w * (1.0 / deg[c]);
        }
    }
}

static int stream_blocks(int64_t work, int threads = 256)
{
    int64_t b = (work + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace glb

using namespace glb;

extern "C" GLB_API int glb_csr_degree(const int32_t *d_rowptr, const int32_t *d_col, const double *d_val, int64_t n,
                              int skip_diagonal, double *d_deg, void *stream)
{
    GLB_CHECK_ARG(d_rowptr && d_val && d_deg, "null pointer");
    GLB_CHECK_ARG(!skip_diagonal || d_col, "col is needed to skip the diagonal");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31), "n out of range");
    degree_kernel<<<stream_blocks(n), 256, 0, (cudaStream_t)stream>>>(d_rowptr, d_col, d_val, (int)n,
                                                                      skip_diagonal != 0, d_deg);
    GLB_LAUNCH_CHECK();
    return 0;
}

static size_t sort_temp_bytes(int64_t nnz)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const int *)nullptr, (int *)nullptr, (int)nnz);
    return bytes;
}

extern "C" GLB_API int64_t glb_csr_transpose_work_bytes(int64_t n, int64_t nnz)
{
    (void)n;
    if (nnz < 0 || nnz >= (1ll << 31)) return GLB_E_INVALID;
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    return (int64_t)(2 * align256(nz * 8) + 2 * align256(nz * 4) + align256(sort_temp_bytes(nnz)) + 256);
}

extern "C" GLB_API int glb_csr_transpose(const int32_t *d_rowptr, const int32_t *d_col, const double *d_val, int64_t n,
                                 int64_t nnz, int32_t *d_t_rowptr, int32_t *d_t_col, double *d_t_val, void *d_work,
                                 int64_t work_bytes, void *stream)
{
    GLB_CHECK_ARG(d_rowptr && d_col && d_val && d_t_rowptr && d_t_col && d_t_val && d_work, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "size out of range");
    GLB_CHECK_ARG(work_bytes >= glb_csr_transpose_work_bytes(n, nnz), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1);
    unsigned char *w = (unsigned char *)(((uintptr_t)d_work + 255) & ~(uintptr_t)255);
    unsigned long long *keys_in = (unsigned long long *)w;  w += align256(nz * 8);
    unsigned long long *keys_out = (unsigned long long *)w; w += align256(nz * 8);
    int *perm_in = (int *)w;  w += align256(nz * 4);
    int *perm_out = (int *)w; w += align256(nz * 4);
    void *temp = w;
    size_t temp_bytes = sort_temp_bytes(nnz);
    if (nnz > 0) {
        transpose_keys_kernel<<<stream_blocks(n * 32), 256, 0, st>>>(d_rowptr, d_col, (int)n, keys_in, perm_in);
        // only the bits that can be set need sorting: 32 low (row) + enough high bits for n columns
        int hi_bits = 1;
        while ((1ll << hi_bits) < n) ++hi_bits;
        GLB_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, perm_in, perm_out, (int)nnz, 0,
                                                 32 + hi_bits, st));
        transpose_fill_kernel<<<stream_blocks(nnz), 256, 0, st>>>(keys_out, perm_out, d_val, nnz, d_t_col, d_t_val);
    }
    transpose_rowptr_kernel<<<stream_blocks(n + 1), 256, 0, st>>>(keys_out, nnz, (int)n, d_t_rowptr);
    GLB_LAUNCH_CHECK();
    return 0;
}

extern "C" GLB_API int glb_poisson_scale(const int32_t *d_t_rowptr, const int32_t *d_t_col, const double *d_t_val,
                                 const double *d_deg, int64_t n, float *d_P_val, double *d_RW_val, void *stream)
{
    GLB_CHECK_ARG(d_t_rowptr && d_t_col && d_t_val && d_deg, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31), "n out of range");
    poisson_scale_kernel<<<stream_blocks(n * 32), 256, 0, (cudaStream_t)stream>>>(d_t_rowptr, d_t_col, d_t_val, d_deg,
                                                                                   (int)n, d_P_val, d_RW_val);
    GLB_LAUNCH_CHECK();
    return 0;
}

