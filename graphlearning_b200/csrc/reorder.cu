// reorder.cu - locality ordering of the graph and the device-side relabelling of CSR / label matrices.
//
// Not in the reference: scipy's csr_matvecs walks rows in storage order on one core and does not care about
// node numbering.  On the GPU every iteration gathers nnz rows of u; with a bandwidth-reducing ordering
// (reverse Cuthill-McKee) the rows gathered by one CTA overlap ~2.5-4x, so most gathers hit that SM's L1
// instead of L2.  The ordering is structural preprocessing (integers only, no arithmetic of the path):
// it runs once per graph on the host from the CSR pattern the caller already holds there; all relabelling of
// matrices and label matrices is done on the device.  Results are returned in the caller's numbering.
#include <algorithm>
#include <numeric>
#include <vector>
#include <cub/device/device_scan.cuh>
#include "common.cuh"

namespace glb {

__global__ void __launch_bounds__(256)
perm_rowlen_kernel(const int *__restrict__ rowptr, const int *__restrict__ perm, int n, int *__restrict__ len,
                   int *__restrict__ iperm)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (long long)gridDim.x * blockDim.x) {
        if (i < n) {
            const int o = perm[i];
            len[i] = rowptr[o + 1] - rowptr[o];
            iperm[o] = (int)i;
        } else {
            len[i] = 0;
        }
    }
}

__global__ void __launch_bounds__(256)
perm_fill_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const float *__restrict__ val,
                 const int *__restrict__ perm, const int *__restrict__ iperm, const int *__restrict__ new_rowptr, int n,
                 int *__restrict__ new_col, float *__restrict__ new_val)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp; i < n; i += nwarps) {
        const int o = perm[i];
        const int src = rowptr[o], len = rowptr[o + 1] - src, dst = new_rowptr[i];
        for (int j = lane; j < len; j += 32) {
            new_col[dst + j] = iperm[col[src + j]];
            new_val[dst + j] = val[src + j];
        }
    }
}

}  // namespace glb

using namespace glb;

// Reverse Cuthill-McKee on a pattern given as CSR (possibly directed); components are started from their lowest-degree
// unvisited node.  order[new] = old.
static void rcm_order(const int *rp, const int *col, int64_t n, std::vector<int> &out)
{
    // start candidates in (degree, index) order: counting sort, degrees are small integers
    std::vector<int> deg((size_t)n), by_deg((size_t)n), order((size_t)n);
    std::vector<char> seen((size_t)n, 0);
    int maxdeg = 0;
    for (int64_t i = 0; i < n; ++i) { deg[i] = rp[i + 1] - rp[i]; maxdeg = std::max(maxdeg, deg[i]); }
    {
        std::vector<int64_t> first((size_t)maxdeg + 2, 0);
        for (int64_t i = 0; i < n; ++i) first[(size_t)deg[i] + 1]++;
        for (int d = 0; d <= maxdeg; ++d) first[(size_t)d + 1] += first[(size_t)d];
        for (int64_t i = 0; i < n; ++i) by_deg[(size_t)first[(size_t)deg[i]]++] = (int)i;
    }
    std::vector<unsigned long long> nb;                           // (degree << 32 | node): one integer compare per pair
    size_t tail = 0;
    for (int64_t s = 0; s < n; ++s) {
        const int start = by_deg[s];
        if (seen[start]) continue;
        seen[start] = 1;
        size_t head = tail;
        order[tail++] = start;
        while (head < tail) {
            const int v = order[head++];
            nb.clear();
            for (int j = rp[v]; j < rp[v + 1]; ++j) {
                const int w = col[j];
                if (w >= 0 && w < n && !seen[w]) { seen[w] = 1; nb.push_back(((unsigned long long)(unsigned)deg[w] << 32) | (unsigned)w); }
            }
            if (nb.size() <= 24) {                                // the usual case: a handful of new neighbours
                for (size_t a = 1; a < nb.size(); ++a) {
                    const unsigned long long key = nb[a];
                    size_t b = a;
                    for (; b > 0 && nb[b - 1] > key; --b) nb[b] = nb[b - 1];
                    nb[b] = key;
                }
            } else {
                std::sort(nb.begin(), nb.end());
            }
            for (unsigned long long key : nb) order[tail++] = (int)(unsigned)(key & 0xffffffffull);
        }
    }
    out.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) out[i] = order[(size_t)(n - 1 - i)];
}

extern "C" GLB_API int glb_locality_order_host(const int32_t *h_rowptr, const int32_t *h_col, int64_t n, int32_t *h_perm)
{
    GLB_CHECK_ARG(h_rowptr && h_perm && (h_col || h_rowptr[n] == 0), "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31), "n out of range");
    std::vector<int> order;
    rcm_order(h_rowptr, h_col, n, order);
    std::copy(order.begin(), order.end(), h_perm);
    return 0;
}

// B = Pi A Pi^T for an fp32 CSR matrix: new row i is old row perm[i], old column c becomes iperm[c].
// d_iperm (n ints) is written.  The order of the entries inside a row is preserved.
extern "C" GLB_API int glb_csr_permute(const int32_t *d_rowptr, const int32_t *d_col, const float *d_val, int64_t n,
                                       int64_t nnz, const int32_t *d_perm, int32_t *d_iperm, int32_t *d_out_rowptr,
                                       int32_t *d_out_col, float *d_out_val, void *stream)
{
    GLB_CHECK_ARG(d_rowptr && d_col && d_val && d_perm && d_iperm && d_out_rowptr && d_out_col && d_out_val, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "size out of range");
    cudaStream_t st = (cudaStream_t)stream;
    int *len = nullptr;
    void *temp = nullptr;
    size_t temp_bytes = 0;
    GLB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, (const int *)nullptr, (int *)nullptr, (int)(n + 1), st));
    GLB_CUDA(dev_alloc(&len, sizeof(int) * (size_t)(n + 1)));
    cudaError_t e = dev_alloc(&temp, temp_bytes ? temp_bytes : 1);
    if (e != cudaSuccess) { dev_free(len); set_error("glb_csr_permute: %s", cudaGetErrorString(e)); return (int)e; }
    const int blocks = std::min<int64_t>((n + 256) / 256, (int64_t)sm_count() * 16);
    perm_rowlen_kernel<<<blocks, 256, 0, st>>>(d_rowptr, d_perm, (int)n, len, d_iperm);
    e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, len, d_out_rowptr, (int)(n + 1), st);
    const int wblocks = std::min<int64_t>((n * 32 + 255) / 256, (int64_t)sm_count() * 16);
    perm_fill_kernel<<<wblocks, 256, 0, st>>>(d_rowptr, d_col, d_val, d_perm, d_iperm, d_out_rowptr, (int)n, d_out_col,
                                              d_out_val);
    cudaError_t e2 = cudaStreamSynchronize(st);        // len/temp are freed below
    dev_free(len);
    dev_free(temp);
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("glb_csr_permute: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
