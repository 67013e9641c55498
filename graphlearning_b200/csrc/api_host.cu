// api_host.cu - host-buffer entry points of the C-ABI: what the reference's Python binds in place of its own
// CPU loops.  Inputs and outputs are HOST arrays in the reference's own dtypes (scipy CSR int32/float64,
// numpy float64); all device memory is allocated, used and freed inside the call.
#include <vector>
#include "common.cuh"

namespace glb {

// reorder the graph when T * nnz exceeds this (the host RCM costs about as much as ~100 iterations)
constexpr long long kReorderMinWork = 200ll * 1000 * 1000;

struct DeviceArena {           // frees everything on scope exit, whatever the return path
    std::vector<void *> ptrs;
    ~DeviceArena() { for (void *p : ptrs) cudaFree(p); }
    template <typename T>
    cudaError_t alloc(T **p, size_t count)
    {
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) ptrs.push_back(q);
        *p = (T *)q;
        return e;
    }
};

// Db[i,:] = (1/deg[i]) * source[i,:]   (ssl.py:636), packed to n x ldu fp32
__global__ void __launch_bounds__(256)
scaled_pack_kernel(const double *__restrict__ src, const double *__restrict__ deg, long long n, int c,
                   float *__restrict__ dst, int ldu, const int *__restrict__ perm)
{
    const long long total = n * ldu;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ldu;
        const int k = (int)(i - r * ldu);
        const long long sr = perm ? perm[r] : r;
        dst[i] = (k < c) ? (float)((1.0 / deg[sr]) * src[sr * c + k]) : 0.f;
    }
}

// deterministic two-pass sum of deg (fixed grid, fixed order)
__global__ void __launch_bounds__(256) partial_sum_kernel(const double *__restrict__ x, long long n, double *__restrict__ part)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += x[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

// vinf = deg / sum(deg)  (ssl.py:642-643);  v = indicator(train)/m  (ssl.py:639-641)
__global__ void __launch_bounds__(256)
mixing_init_kernel(const double *__restrict__ deg, const double *__restrict__ part, int nparts, long long n,
                   double *__restrict__ vinf, double *__restrict__ v)
{
    double tot = 0.0;
    for (int i = 0; i < nparts; ++i) tot += part[i];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        vinf[i] = deg[i] / tot;
        v[i] = 0.0;
    }
}

__global__ void mixing_seed_kernel(const long long *__restrict__ train_ind, long long m, long long n, double *__restrict__ v)
{
    // v[train_ind] = 1 (duplicates collapse, numpy fancy assignment), then v /= sum(v)
    for (long long i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        long long t = train_ind[i];
        if (t < 0) t += n;
        if (t >= 0 && t < n) v[t] = 1.0;
    }
}

__global__ void __launch_bounds__(256) scale_by_sum_kernel(double *__restrict__ v, const double *__restrict__ part, int nparts, long long n)
{
    double tot = 0.0;
    for (int i = 0; i < nparts; ++i) tot += part[i];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) v[i] = v[i] / tot;
}

}  // namespace glb

using namespace glb;

extern "C" GLB_API int glb_poisson_gd_host(const int32_t *h_rowptr, const int32_t *h_col, const double *h_val, int64_t n,
                                   int64_t nnz, const double *h_source, int c, const int64_t *h_train_ind, int64_t m,
                                   int min_iter, int max_iter, double *h_u_out, int *T_done, int *launches)
{
    GLB_CHECK_ARG(h_rowptr && h_col && h_val && h_source && h_u_out, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "size out of range");
    GLB_CHECK_ARG(c > 0, "c must be positive");
    GLB_CHECK_ARG(min_iter >= 0 && max_iter >= 0, "iteration counts must be >= 0");
    GLB_CHECK_ARG(m == 0 || h_train_ind, "train_ind is null");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("glb_poisson_gd_host: no CUDA device visible");
        return GLB_E_NOGPU;
    }
    const int ldu = glb_padded_ld(c);
    cudaStream_t st = 0;
    DeviceArena A;
    int nl = 0;

    int *rp, *col, *t_rp, *t_col;
    double *val, *t_val, *deg, *src, *rw_val, *vinf, *v, *vtmp, *part, *u64;
    float *P_val, *Db, *u0, *u1;
    long long *tind;
    void *work;
    const int64_t work_bytes = glb_csr_transpose_work_bytes(n, nnz);
    const int NPART = 256;
    GLB_CUDA(A.alloc(&rp, n + 1));      GLB_CUDA(A.alloc(&col, nnz));      GLB_CUDA(A.alloc(&val, nnz));
    GLB_CUDA(A.alloc(&t_rp, n + 1));    GLB_CUDA(A.alloc(&t_col, nnz));    GLB_CUDA(A.alloc(&t_val, nnz));
    GLB_CUDA(A.alloc(&deg, n));         GLB_CUDA(A.alloc(&src, n * c));    GLB_CUDA(A.alloc(&rw_val, nnz));
    GLB_CUDA(A.alloc(&vinf, n));        GLB_CUDA(A.alloc(&v, n));          GLB_CUDA(A.alloc(&vtmp, n));
    GLB_CUDA(A.alloc(&part, NPART));    GLB_CUDA(A.alloc(&P_val, nnz));    GLB_CUDA(A.alloc(&Db, n * ldu));
    GLB_CUDA(A.alloc(&u0, n * ldu));    GLB_CUDA(A.alloc(&u1, n * ldu));   GLB_CUDA(A.alloc(&tind, m));
    GLB_CUDA(A.alloc((unsigned char **)&work, (size_t)work_bytes));
    u64 = src;                          // reused for the fp64 result once Db is built

    GLB_CUDA(cudaMemcpyAsync(rp, h_rowptr, (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(col, h_col, nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(val, h_val, nnz * sizeof(double), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(src, h_source, n * c * sizeof(double), cudaMemcpyHostToDevice, st));
    if (m) GLB_CUDA(cudaMemcpyAsync(tind, h_train_ind, m * sizeof(long long), cudaMemcpyHostToDevice, st));

    // W <- W - diag(W) (ssl.py:615-616) is never materialised: the degree kernel skips diagonal entries and
    // glb_poisson_scale writes zeros for them.
    int rc;
    if ((rc = glb_csr_transpose(rp, col, val, n, nnz, t_rp, t_col, t_val, work, work_bytes, st))) return rc;
    nl += 4;
    if ((rc = glb_csr_degree(rp, col, val, n, 1, deg, st))) return rc;
    nl += 1;
    if ((rc = glb_poisson_scale(t_rp, t_col, t_val, deg, n, P_val, rw_val, st))) return rc;
    nl += 1;
    // Iteration count first (the ordering below only pays off for long runs).

    // iteration count by the reference's stopping rule
    int T = max_iter;
    if (min_iter < max_iter) {
        partial_sum_kernel<<<NPART, 256, 0, st>>>(deg, n, part);
        mixing_init_kernel<<<sm_count() * 4, 256, 0, st>>>(deg, part, NPART, n, vinf, v);
        if (m) mixing_seed_kernel<<<ceil_div(m, 256), 256, 0, st>>>(tind, m, n, v);
        partial_sum_kernel<<<NPART, 256, 0, st>>>(v, n, part);
        scale_by_sum_kernel<<<sm_count() * 4, 256, 0, st>>>(v, part, NPART, n);
        nl += 5;
        GLB_LAUNCH_CHECK();
        if ((rc = glb_poisson_mixing_T(t_rp, t_col, rw_val, vinf, v, vtmp, n, min_iter, max_iter, &T, &nl, st))) return rc;
    }

    // Locality ordering for long runs: RCM on the host pattern, relabelling on the device.
    const int *it_rp = t_rp, *it_col = t_col;
    const float *it_val = P_val;
    int *perm = nullptr;
    if ((int64_t)T * nnz >= (int64_t)kReorderMinWork && nnz > 0) {
        std::vector<int> h_perm((size_t)n);
        if ((rc = glb_locality_order_host(h_rowptr, h_col, n, h_perm.data()))) return rc;
        int *iperm, *p_rp, *p_col;
        float *p_val;
        GLB_CUDA(A.alloc(&perm, n));   GLB_CUDA(A.alloc(&iperm, n));
        GLB_CUDA(A.alloc(&p_rp, n + 1)); GLB_CUDA(A.alloc(&p_col, nnz)); GLB_CUDA(A.alloc(&p_val, nnz));
        GLB_CUDA(cudaMemcpyAsync(perm, h_perm.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
        if ((rc = glb_csr_permute(t_rp, t_col, P_val, n, nnz, perm, iperm, p_rp, p_col, p_val, st))) return rc;
        nl += 3;
        it_rp = p_rp; it_col = p_col; it_val = p_val;
    }
    scaled_pack_kernel<<<sm_count() * 8, 256, 0, st>>>(src, deg, n, c, Db, ldu, perm);
    nl += 1;
    GLB_LAUNCH_CHECK();

    glb_poisson_plan *plan = nullptr;
    if ((rc = glb_poisson_plan_create(&plan, it_rp, n, nnz, ldu, st))) return rc;
    GLB_CUDA(cudaMemsetAsync(u0, 0, n * ldu * sizeof(float), st));
    GLB_CUDA(cudaMemsetAsync(u1, 0, n * ldu * sizeof(float), st));
    int in_u1 = 0;
    rc = glb_poisson_iterate(plan, it_rp, it_col, it_val, Db, u0, u1, T, &in_u1, &nl, st);
    glb_poisson_plan_destroy(plan);
    if (rc) return rc;
    if ((rc = glb_unpack_f32_to_f64(in_u1 ? u1 : u0, n, c, ldu, u64, perm, st))) return rc;
    nl += 1;
    GLB_CUDA(cudaMemcpyAsync(h_u_out, u64, n * c * sizeof(double), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    if (T_done) *T_done = T;
    if (launches) *launches = nl;
    return 0;
}
