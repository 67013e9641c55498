// laplace.cu - graph Laplacians and the Dirichlet sub-system of Laplace learning, assembled on the device.
//
// Replaces graph.laplacian (reference graphlearning/graph.py:469-513: D - W, I - D^-1 W, I - D^-1/2 W D^-1/2 as scipy
// sparse products and differences, 74 ms at 70k nodes) and the system assembly of ssl.laplace._fit
// (graphlearning/ssl.py:1222-1246: tau + L, fancy-index slicing of the unlabelled rows and columns, right-hand side
// -L[:, train] F, Jacobi scaling M A M / M b with M = diag(1 / sqrt(diag(A) + 1e-10)); ~0.3 s there).
//
// All three normalisations are one formula, L = Diag(dg) - Diag(left) W Diag(right):
//     combinatorial  dg = d,  left = right = none          (graph.py:497)
//     randomwalk     dg = 1,  left = d^-1, right = none    (graph.py:499-500)
//     normalized     dg = 1,  left = right = d^-1/2        (graph.py:502-503)
// The caller computes d^p on the host exactly as the reference (numpy `d**p`, n values); the O(nnz) work is here.  Every
// entry is rounded as scipy rounds it: (left_i * w_ij) * right_j, then dg_i - that on the diagonal / its negative off
// the diagonal; results that are exactly zero are dropped, as scipy's sparse subtraction does; columns stay sorted.  W has
// to be canonical (sorted columns, no duplicates) - what gl.graph holds.
#include <cub/device/device_scan.cuh>
#include <vector>
#include "common.cuh"

namespace glb {

struct LapSpec {
    const int *rp, *col;
    const double *val, *left, *right, *dg, *tau;
    int n;
    __device__ __forceinline__ double prod(int i, int q) const
    {
        double p = val[q];
        if (left) p = __dmul_rn(left[i], p);
        if (right) p = __dmul_rn(p, right[col[q]]);
        return p;
    }
    // value of the diagonal entry of row i (has_diag: W stores w_ii at q) and of an off-diagonal entry
    __device__ __forceinline__ double diag(int i, bool has_diag, int q) const
    {
        double v = has_diag ? __dsub_rn(dg[i], prod(i, q)) : dg[i];
        if (tau) v = __dadd_rn(tau[i], v);                  // spdiags(tau) + L   (ssl.py:1222)
        return v;
    }
    __device__ __forceinline__ double off(int i, int q) const { return -prod(i, q); }
};

// ---- full Laplacian ------------------------------------------------------------------------------------------------
// pass 0 counts the stored entries of every row, pass 1 writes them (diagonal merged in at its sorted position)
template <bool FILL>
__global__ void __launch_bounds__(256)
laplacian_rows_kernel(LapSpec S, int *__restrict__ counts, const int *__restrict__ out_rp, int *__restrict__ out_col,
                      double *__restrict__ out_val)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += gridDim.x * blockDim.x) {
        const int b = S.rp[i], e = S.rp[i + 1];
        int o = FILL ? out_rp[i] : 0;
        bool diag_done = false;
        auto emit = [&](int c, double v) {
            if (v == 0.0) return;                           // scipy's csr binop stores non-zero results only
            if (FILL) { out_col[o] = c; out_val[o] = v; }
            ++o;
        };
        for (int q = b; q < e; ++q) {
            const int j = S.col[q];
            if (!diag_done && j >= i) {
                emit(i, S.diag(i, j == i, q));
                diag_done = true;
                if (j == i) continue;
            }
            emit(j, S.off(i, q));
        }
        if (!diag_done) emit(i, S.diag(i, false, 0));
        if (!FILL) counts[i] = o;
    }
}

// ---- Dirichlet sub-system of Laplace learning ---------------------------------------------------------------------------
// lab[i] = -1 for an unlabelled node, else the row of F that holds its label; pos[i] = index of node i among the unlabelled.
__global__ void __launch_bounds__(256) sys_init_lab_kernel(int *__restrict__ lab, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) lab[i] = -1;
}
__global__ void __launch_bounds__(256) sys_mark_kernel(const long long *__restrict__ train, int m, int *__restrict__ lab)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < m; t += gridDim.x * blockDim.x) lab[train[t]] = t;
}
__global__ void __launch_bounds__(256) sys_flag_kernel(const int *__restrict__ lab, int n, int *__restrict__ flag)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) flag[i] = (i < n && lab[i] < 0) ? 1 : 0;
}

// M = 1 / sqrt(diag(A) + 1e-10)   (ssl.py:1244-1245), and the stored entries of every row of A = L[idx][:, idx]
__global__ void __launch_bounds__(256)
sys_diag_count_kernel(LapSpec S, const int *__restrict__ lab, const int *__restrict__ pos, double *__restrict__ M,
                      int *__restrict__ counts)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += gridDim.x * blockDim.x) {
        if (lab[i] >= 0) continue;
        const int b = S.rp[i], e = S.rp[i + 1];
        int cnt = 0;
        double d = 0.0;
        bool seen = false;
        for (int q = b; q < e; ++q) {
            const int j = S.col[q];
            if (j == i) { d = S.diag(i, true, q); seen = true; }
            else if (lab[j] < 0 && S.off(i, q) != 0.0) ++cnt;
        }
        if (!seen) d = S.diag(i, false, 0);
        if (d != 0.0) ++cnt;
        M[pos[i]] = 1.0 / sqrt(d + 1e-10);
        counts[pos[i]] = cnt;
    }
}

// rows of M A M (CSR over the unlabelled nodes) and of M b, b = -L[:, train] F restricted to the unlabelled rows
// (ssl.py:1236-1249).  F: m x c row-major; Mb: nu x ldb (zero padded, the layout of glb_cg_solve).
__global__ void __launch_bounds__(256)
sys_fill_kernel(LapSpec S, const int *__restrict__ lab, const int *__restrict__ pos, const double *__restrict__ M,
                const double *__restrict__ F, int c, int ldb, const int *__restrict__ a_rp, int *__restrict__ a_col,
                double *__restrict__ a_val, double *__restrict__ Mb)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n; i += gridDim.x * blockDim.x) {
        if (lab[i] >= 0) continue;
        const int r = pos[i];
        const double mi = M[r];
        int o = a_rp[r];
        double *brow = Mb + (size_t)r * ldb;
        for (int k = 0; k < ldb; ++k) brow[k] = 0.0;
        bool diag_done = false;
        auto emit = [&](int j, double v) {
            if (v == 0.0) return;
            a_col[o] = pos[j];
            a_val[o] = __dmul_rn(__dmul_rn(mi, v), M[pos[j]]);           // (M A) M
            ++o;
        };
        const int b = S.rp[i], e = S.rp[i + 1];
        for (int q = b; q < e; ++q) {
            const int j = S.col[q];
            if (!diag_done && j >= i) {
                emit(i, S.diag(i, j == i, q));
                diag_done = true;
                if (j == i) continue;
            }
            const double v = S.off(i, q);
            if (lab[j] < 0) { emit(j, v); continue; }
            const double *f = F + (size_t)lab[j] * c;                      // labelled neighbour: b_i -= L_ij F_j
            for (int k = 0; k < c; ++k) brow[k] = fma(-v, f[k], brow[k]);
        }
        if (!diag_done) emit(i, S.diag(i, false, 0));
        for (int k = 0; k < c; ++k) brow[k] = __dmul_rn(mi, brow[k]);
    }
}

// u[idx] = M v, u[train] = F   (ssl.py:1250-1255)
__global__ void __launch_bounds__(256)
sys_scatter_kernel(const int *__restrict__ lab, const int *__restrict__ pos, const double *__restrict__ M,
                   const double *__restrict__ x, int ldb, const double *__restrict__ F, int c, int n, double *__restrict__ u)
{
    const long long total = (long long)n * c;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / c), k = (int)(t % c);
        u[t] = lab[i] >= 0 ? F[(size_t)lab[i] * c + k] : __dmul_rn(M[pos[i]], x[(size_t)pos[i] * ldb + k]);
    }
}

struct LapArena {
    std::vector<void *> v;
    ~LapArena() { for (void *p : v) dev_free(p); }
    template <typename T>
    cudaError_t alloc(T **p, size_t count)
    {
        void *q = nullptr;
        cudaError_t e = dev_alloc(&q, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) v.push_back(q);
        *p = (T *)q;
        return e;
    }
};

static int upload_spec(LapArena &A, const int32_t *h_rowptr, const int32_t *h_col, const double *h_val, int64_t n, int64_t nnz,
                       const double *h_left, const double *h_right, const double *h_diag, const double *h_tau, LapSpec &S,
                       cudaStream_t st)
{
    int *rp, *col;
    double *val, *left = nullptr, *right = nullptr, *dg, *tau = nullptr;
    GLB_CUDA(A.alloc(&rp, n + 1)); GLB_CUDA(A.alloc(&col, nnz)); GLB_CUDA(A.alloc(&val, nnz)); GLB_CUDA(A.alloc(&dg, n));
    GLB_CUDA(cudaMemcpyAsync(rp, h_rowptr, (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(col, h_col, nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(val, h_val, nnz * sizeof(double), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(dg, h_diag, n * sizeof(double), cudaMemcpyHostToDevice, st));
    auto vec = [&](const double *h, double **d) -> int {
        if (!h) return 0;
        GLB_CUDA(A.alloc(d, n));
        GLB_CUDA(cudaMemcpyAsync(*d, h, n * sizeof(double), cudaMemcpyHostToDevice, st));
        return 0;
    };
    int rc;
    if ((rc = vec(h_left, &left)) || (rc = vec(h_right, &right)) || (rc = vec(h_tau, &tau))) return rc;
    S = LapSpec{rp, col, val, left, right, dg, tau, (int)n};
    return 0;
}

static int exclusive_scan(LapArena &A, const int *in, int *out, int count, cudaStream_t st)
{
    void *temp = nullptr;
    size_t bytes = 0;
    GLB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, count, st));
    GLB_CUDA(A.alloc((unsigned char **)&temp, bytes));
    GLB_CUDA(cub::DeviceScan::ExclusiveSum(temp, bytes, in, out, count, st));
    return 0;
}

static int check_gpu(const char *who)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("%s: no CUDA device visible", who);
        return GLB_E_NOGPU;
    }
    return 0;
}

}  // namespace glb

using namespace glb;

extern "C" GLB_API int glb_laplacian_csr_host(const int32_t *h_rowptr, const int32_t *h_col, const double *h_val, int64_t n,
                                              int64_t nnz, const double *h_left, const double *h_right, const double *h_diag,
                                              int32_t *h_out_rowptr, int32_t *h_out_col, double *h_out_val)
{
    GLB_CHECK_ARG(h_rowptr && (nnz == 0 || (h_col && h_val)) && h_diag && h_out_rowptr && h_out_col && h_out_val, "null pointer");
    GLB_CHECK_ARG(n > 0 && nnz >= 0 && n + nnz < (1ll << 31), "size out of range");
    int rc;
    if ((rc = check_gpu("glb_laplacian_csr_host"))) return rc;
    cudaStream_t st = 0;
    LapArena A;
    LapSpec S;
    if ((rc = upload_spec(A, h_rowptr, h_col, h_val, n, nnz, h_left, h_right, h_diag, nullptr, S, st))) return rc;
    int *counts, *out_rp, *out_col;
    double *out_val;
    GLB_CUDA(A.alloc(&counts, n + 1)); GLB_CUDA(A.alloc(&out_rp, n + 1));
    GLB_CUDA(A.alloc(&out_col, nnz + n)); GLB_CUDA(A.alloc(&out_val, nnz + n));
    GLB_CUDA(cudaMemsetAsync(counts + n, 0, sizeof(int), st));
    const int blocks = std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
    laplacian_rows_kernel<false><<<blocks, 256, 0, st>>>(S, counts, nullptr, nullptr, nullptr);
    if ((rc = exclusive_scan(A, counts, out_rp, (int)n + 1, st))) return rc;
    laplacian_rows_kernel<true><<<blocks, 256, 0, st>>>(S, nullptr, out_rp, out_col, out_val);
    GLB_LAUNCH_CHECK();
    GLB_CUDA(cudaMemcpyAsync(h_out_rowptr, out_rp, (n + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    const int64_t out_nnz = h_out_rowptr[n];
    GLB_CUDA(cudaMemcpyAsync(h_out_col, out_col, out_nnz * sizeof(int), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaMemcpyAsync(h_out_val, out_val, out_nnz * sizeof(double), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// Device-resident weight matrix and Laplacian scalings of one graph: what every fit on that graph shares (ssl_trials runs
// hundreds of fits on one W; the reference rebuilds L inside each _fit, ssl.py:1208-1222).
struct glb_laplace_graph {
    int64_t n = 0, nnz = 0;
    LapArena A;
    LapSpec S{};
};

extern "C" GLB_API int glb_laplace_graph_create(glb_laplace_graph **out, const int32_t *h_rowptr, const int32_t *h_col,
                                                const double *h_val, int64_t n, int64_t nnz, const double *h_left,
                                                const double *h_right, const double *h_diag, const double *h_tau)
{
    GLB_CHECK_ARG(out && h_rowptr && (nnz == 0 || (h_col && h_val)) && h_diag, "null pointer");
    GLB_CHECK_ARG(n > 0 && nnz >= 0 && n + nnz < (1ll << 31), "size out of range");
    int rc;
    if ((rc = check_gpu("glb_laplace_graph_create"))) return rc;
    glb_laplace_graph *g = new glb_laplace_graph();
    g->n = n; g->nnz = nnz;
    cudaStream_t st = 0;
    if ((rc = upload_spec(g->A, h_rowptr, h_col, h_val, n, nnz, h_left, h_right, h_diag, h_tau, g->S, st))) { delete g; return rc; }
    cudaError_t e = cudaStreamSynchronize(st);                // the host arrays are the caller's
    if (e != cudaSuccess) { delete g; set_error("glb_laplace_graph_create: %s", cudaGetErrorString(e)); return (int)e; }
    *out = g;
    return 0;
}

extern "C" GLB_API int glb_laplace_graph_destroy(glb_laplace_graph *g)
{
    delete g;
    return 0;
}

extern "C" GLB_API int glb_laplace_graph_fit(glb_laplace_graph *g, const int64_t *h_train_ind, int64_t m, const double *h_F, int c,
                                             double tol, double *h_u, int64_t *iters, double *err, int *launches, double *h_ms)
{
    GLB_CHECK_ARG(g && h_train_ind && h_F && h_u, "null pointer");
    const int64_t n = g->n, nnz = g->nnz;
    GLB_CHECK_ARG(m > 0 && m < n && c > 0, "need 0 < m < n labelled nodes and c > 0 classes");
    for (int64_t t = 0; t < m; ++t) GLB_CHECK_ARG(h_train_ind[t] >= 0 && h_train_ind[t] < n, "train_ind out of range");
    int rc;
    const int64_t wb = glb_cg_work_bytes(n, c);
    if (wb < 0) { set_error("glb_laplace_graph_fit: unsupported number of classes (c = %d)", c); return (int)wb; }
    cudaStream_t st = 0;
    PhaseTimer tm("laplace_fit");
    LapArena A;                                              // per-fit scratch (pooled allocations)
    const LapSpec &S = g->S;
    const int ldb = glb_padded_ld(c);
    long long *train;
    int *lab, *flag, *pos, *counts, *a_rp, *a_col;
    double *F, *M, *a_val, *Mb, *x, *u;
    void *work;
    GLB_CUDA(A.alloc(&train, m)); GLB_CUDA(A.alloc(&F, m * c)); GLB_CUDA(A.alloc(&lab, n)); GLB_CUDA(A.alloc(&flag, n + 1));
    GLB_CUDA(A.alloc(&pos, n + 1)); GLB_CUDA(A.alloc(&counts, n + 1)); GLB_CUDA(A.alloc(&a_rp, n + 1));
    GLB_CUDA(A.alloc(&a_col, nnz + n)); GLB_CUDA(A.alloc(&a_val, nnz + n)); GLB_CUDA(A.alloc(&M, n));
    GLB_CUDA(A.alloc(&Mb, (size_t)n * ldb)); GLB_CUDA(A.alloc(&x, (size_t)n * ldb)); GLB_CUDA(A.alloc(&u, (size_t)n * c));
    GLB_CUDA(A.alloc((unsigned char **)&work, (size_t)wb));
    GLB_CUDA(cudaMemcpyAsync(train, h_train_ind, m * sizeof(long long), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(F, h_F, m * c * sizeof(double), cudaMemcpyHostToDevice, st));
    const int blocks = std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
    int nl = 0;
    sys_init_lab_kernel<<<blocks, 256, 0, st>>>(lab, (int)n); ++nl;
    sys_mark_kernel<<<ceil_div(m, 256), 256, 0, st>>>(train, (int)m, lab); ++nl;
    sys_flag_kernel<<<blocks, 256, 0, st>>>(lab, (int)n, flag); ++nl;
    if ((rc = exclusive_scan(A, flag, pos, (int)n + 1, st))) return rc;
    ++nl;
    int nu = 0;
    GLB_CUDA(cudaMemcpyAsync(&nu, pos + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaMemsetAsync(counts, 0, (n + 1) * sizeof(int), st));
    sys_diag_count_kernel<<<blocks, 256, 0, st>>>(S, lab, pos, M, counts); ++nl;
    if ((rc = exclusive_scan(A, counts, a_rp, (int)n + 1, st))) return rc;
    ++nl;
    sys_fill_kernel<<<blocks, 256, 0, st>>>(S, lab, pos, M, F, c, ldb, a_rp, a_col, a_val, Mb); ++nl;
    GLB_LAUNCH_CHECK();
    GLB_CUDA(cudaStreamSynchronize(st));                     // nu
    if (nu <= 0) { set_error("glb_laplace_graph_fit: every node is labelled"); return GLB_E_INVALID; }
    int a_nnz = 0;
    GLB_CUDA(cudaMemcpy(&a_nnz, a_rp + nu, sizeof(int), cudaMemcpyDeviceToHost));
    tm.lap("upload + system assembly");
    int64_t it = 0;
    double e = 0.0;
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    cudaEventRecord(ev0, st);
    rc = glb_cg_solve(a_rp, a_col, a_val, nu, a_nnz, Mb, nullptr, c, tol, 100000, x, work, wb, &it, &e, &nl, st);
    cudaEventRecord(ev1, st);
    if (rc == 0 && cudaEventSynchronize(ev1) == cudaSuccess && h_ms) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        h_ms[0] = (double)ms;                                 // device time of the CG iterations
        h_ms[1] = (double)a_nnz;                              // stored entries of the system matrix
        h_ms[2] = (double)nu;                                 // unknowns
    }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    if (rc) return rc;
    sys_scatter_kernel<<<sm_count() * 8, 256, 0, st>>>(lab, pos, M, x, ldb, F, c, (int)n, u); ++nl;
    GLB_LAUNCH_CHECK();
    GLB_CUDA(cudaMemcpyAsync(h_u, u, (size_t)n * c * sizeof(double), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    tm.lap("cg + scatter + download");
    if (iters) *iters = it;
    if (err) *err = e;
    if (launches) *launches = nl;
    return 0;
}

extern "C" GLB_API int glb_laplace_fit_host(const int32_t *h_rowptr, const int32_t *h_col, const double *h_val, int64_t n,
                                            int64_t nnz, const double *h_left, const double *h_right, const double *h_diag,
                                            const double *h_tau, const int64_t *h_train_ind, int64_t m, const double *h_F, int c,
                                            double tol, double *h_u, int64_t *iters, double *err, int *launches, double *h_ms)
{
    glb_laplace_graph *g = nullptr;
    int rc = glb_laplace_graph_create(&g, h_rowptr, h_col, h_val, n, nnz, h_left, h_right, h_diag, h_tau);
    if (rc) return rc;
    rc = glb_laplace_graph_fit(g, h_train_ind, m, h_F, c, tol, h_u, iters, err, launches, h_ms);
    glb_laplace_graph_destroy(g);
    return rc;
}
