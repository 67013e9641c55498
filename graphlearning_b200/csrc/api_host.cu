// api_host.cu - host-buffer entry points of the C-ABI: what the reference's Python binds in place of its own
// CPU loops.  Inputs and outputs are HOST arrays in the reference's own dtypes (scipy CSR int32/float64,
// numpy float64); all device memory is allocated, used and freed inside the call.
#include <string.h>
#include <algorithm>
#include <thread>
#include <utility>
#include <vector>
#include "common.cuh"

namespace glb {

// auto mode relabels every graph that is big enough for the ordering to pay for its host time: the dataflow kernel's
// first gather attempt goes through L1 and the slab kernel streams HBM, both live on neighbouring rows sharing columns
// (profiles/r2_dataflow_pair_stream_ab.txt: 7.76 us/iteration in the caller's numbering, 6.40 relabelled)
constexpr int64_t kReorderMinNodes = 4096;

struct DeviceArena {           // frees everything on scope exit, whatever the return path
    std::vector<void *> ptrs;
    ~DeviceArena() { for (void *p : ptrs) dev_free(p); }
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

// source[idx[i], :] = rows[i, :] for the m labelled nodes (idx unique and in range: prepared on the host); the rest of
// the source term is zero (ssl.py:619-622)
__global__ void __launch_bounds__(256)
scatter_source_rows_kernel(const long long *__restrict__ idx, const double *__restrict__ rows, long long m, int c, double *__restrict__ src)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m * c; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c;
        src[idx[r] * c + (i - r * c)] = rows[i];
    }
}

}  // namespace glb

using namespace glb;

// Device-resident state of one weight matrix: everything ssl.poisson._fit recomputes on every call
// (graph.graph(W), D, P = D^-1 W^T, RW = W^T D^-1, deg/sum(deg); ssl.py:615-617, 634-644) is built once here.
struct glb_poisson_graph {
    int64_t n = 0, nnz = 0;
    DeviceArena A;                       // owns the per-graph device buffers below
    DeviceArena Aw;                      // owns the per-width buffers (src64, Db, u0, u1, tind): released when the width changes
    int *t_rp = nullptr, *t_col = nullptr;               // pattern of W^T
    float *P_val = nullptr;
    double *rw_val = nullptr, *deg = nullptr, *vinf = nullptr, *v = nullptr, *vtmp = nullptr, *part = nullptr;
    // iterate arrays (relabelled when a locality ordering is in use)
    const int *it_rp = nullptr, *it_col = nullptr;
    const float *it_val = nullptr;
    int *perm = nullptr;
    // per-width work buffers and plan (kept for the last width used)
    int ldu = 0, c_plan = 0;
    int64_t rows = 0;                    // rows of Db / u0 / u1 (glb_poisson_plan_rows)
    int64_t m_cap = 0;
    double *src64 = nullptr;             // n x c staging (source in, result out)
    float *Db = nullptr, *u0 = nullptr, *u1 = nullptr;
    long long *tind = nullptr;
    unsigned char *rows_stage = nullptr; // labelled rows of a sparse source: m x (8 + 8c) bytes, indices first
    int64_t rows_cap = 0;                // bytes
    glb_poisson_plan *plan = nullptr;
    int setup_launches = 0;
    ~glb_poisson_graph() { if (plan) glb_poisson_plan_destroy(plan); }
};

static const int NPART = 256;

extern "C" GLB_API int glb_poisson_graph_create(glb_poisson_graph **out, const int32_t *h_rowptr, const int32_t *h_col,
                                                const double *h_val, int64_t n, int64_t nnz, int reorder)
{
    GLB_CHECK_ARG(out && h_rowptr && (nnz == 0 || (h_col && h_val)), "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "size out of range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("glb_poisson_graph_create: no CUDA device visible");
        return GLB_E_NOGPU;
    }
    cudaStream_t st = 0;
    PhaseTimer tm("graph_create");
    // Locality ordering (reorder < 0 = auto): host work on the pattern the caller holds, started now so that it runs beside
    // the upload and the device-side transposition / scaling
    const bool want = (reorder > 0 || (reorder < 0 && n >= kReorderMinNodes)) && nnz > 0;
    std::vector<int> h_perm;
    int order_rc = 0;
    std::thread order_thread;
    struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{order_thread};
    if (want) {
        h_perm.resize((size_t)n);
        order_thread = std::thread([&]() { order_rc = glb_locality_order_host(h_rowptr, h_col, n, h_perm.data()); });
    }
    glb_poisson_graph *g = new glb_poisson_graph();
    struct Guard { glb_poisson_graph *g; ~Guard() { delete g; } } guard{g};
    g->n = n; g->nnz = nnz;
    int *rp, *col;
    double *val, *t_val;
    void *work;
    DeviceArena tmp;                    // freed when this function returns
    const int64_t work_bytes = glb_csr_transpose_work_bytes(n, nnz);
    GLB_CUDA(tmp.alloc(&rp, n + 1));      GLB_CUDA(tmp.alloc(&col, nnz));      GLB_CUDA(tmp.alloc(&val, nnz));
    GLB_CUDA(tmp.alloc(&t_val, nnz));     GLB_CUDA(tmp.alloc((unsigned char **)&work, (size_t)work_bytes));
    GLB_CUDA(g->A.alloc(&g->t_rp, n + 1)); GLB_CUDA(g->A.alloc(&g->t_col, nnz)); GLB_CUDA(g->A.alloc(&g->P_val, nnz));
    GLB_CUDA(g->A.alloc(&g->rw_val, nnz)); GLB_CUDA(g->A.alloc(&g->deg, n));     GLB_CUDA(g->A.alloc(&g->vinf, n));
    GLB_CUDA(g->A.alloc(&g->v, n));        GLB_CUDA(g->A.alloc(&g->vtmp, n));    GLB_CUDA(g->A.alloc(&g->part, NPART));
    GLB_CUDA(cudaMemcpyAsync(rp, h_rowptr, (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(col, h_col, nnz * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(val, h_val, nnz * sizeof(double), cudaMemcpyHostToDevice, st));

    tm.lap("alloc + upload");
    // W <- W - diag(W) (ssl.py:615-616) is never materialised: the degree kernel skips diagonal entries and
    // glb_poisson_scale writes zeros for them.
    int rc;
    if ((rc = glb_csr_transpose(rp, col, val, n, nnz, g->t_rp, g->t_col, t_val, work, work_bytes, st))) return rc;
    if ((rc = glb_csr_degree(rp, col, val, n, 1, g->deg, st))) return rc;
    if ((rc = glb_poisson_scale(g->t_rp, g->t_col, t_val, g->deg, n, g->P_val, g->rw_val, st))) return rc;
    partial_sum_kernel<<<NPART, 256, 0, st>>>(g->deg, n, g->part);
    mixing_init_kernel<<<sm_count() * 4, 256, 0, st>>>(g->deg, g->part, NPART, n, g->vinf, g->v);
    g->setup_launches = 4 + 1 + 1 + 2;
    GLB_LAUNCH_CHECK();

    tm.lap("transpose/degree/scale");
    g->it_rp = g->t_rp; g->it_col = g->t_col; g->it_val = g->P_val;
    if (want) {
        order_thread.join();
        if (order_rc) return order_rc;
        int *iperm, *p_rp, *p_col;
        float *p_val;
        GLB_CUDA(g->A.alloc(&g->perm, n));  GLB_CUDA(tmp.alloc(&iperm, n));
        GLB_CUDA(g->A.alloc(&p_rp, n + 1)); GLB_CUDA(g->A.alloc(&p_col, nnz)); GLB_CUDA(g->A.alloc(&p_val, nnz));
        GLB_CUDA(cudaMemcpyAsync(g->perm, h_perm.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
        if ((rc = glb_csr_permute(g->t_rp, g->t_col, g->P_val, n, nnz, g->perm, iperm, p_rp, p_col, p_val, st))) return rc;
        g->setup_launches += 3;
        g->it_rp = p_rp; g->it_col = p_col; g->it_val = p_val;
    }
    GLB_CUDA(cudaStreamSynchronize(st));
    tm.lap("locality order");
    guard.g = nullptr;
    *out = g;
    return 0;
}

extern "C" GLB_API int glb_poisson_graph_destroy(glb_poisson_graph *g)
{
    delete g;
    return 0;
}

// One fit.  The source term is either dense (h_source: n x c) or given by its nonzero rows (h_rows: m_rows x c at the nodes
// h_row_ind, everything else zero - what ssl.py:619-622 builds: a few hundred bytes instead of n x c x 8 over PCIe).
static int poisson_graph_fit_impl(glb_poisson_graph *g, const double *h_source, const int64_t *h_row_ind, const double *h_rows,
                                  int64_t m_rows, int c, const int64_t *h_train_ind, int64_t m, int min_iter, int max_iter,
                                  double *h_u_out, int *T_done, int *launches)
{
    GLB_CHECK_ARG(g && (h_source || m_rows == 0 || (h_row_ind && h_rows)) && h_u_out, "null pointer");
    GLB_CHECK_ARG(c > 0, "c must be positive");
    GLB_CHECK_ARG(min_iter >= 0 && max_iter >= 0, "iteration counts must be >= 0");
    GLB_CHECK_ARG(m == 0 || h_train_ind, "train_ind is null");
    const int64_t n = g->n, nnz = g->nnz;
    cudaStream_t st = 0;
    int nl = 0, rc;
    PhaseTimer tm("graph_fit");
    if (c != g->c_plan) {                                // (re)build the plan and the per-width buffers
        if (g->plan) { glb_poisson_plan_destroy(g->plan); g->plan = nullptr; g->c_plan = 0; }
        for (void *q : g->Aw.ptrs) dev_free(q);                    // buffers of the previous width (nothing is in flight: fits are synchronous)
        g->Aw.ptrs.clear();
        g->src64 = nullptr; g->Db = g->u0 = g->u1 = nullptr; g->tind = nullptr; g->m_cap = 0;
        g->rows_stage = nullptr; g->rows_cap = 0;
        if ((rc = glb_poisson_plan_create(&g->plan, g->it_rp, g->it_col, g->it_val, n, nnz, c, GLB_POISSON_KIND_AUTO, st)))
            return rc;
        const int ld = glb_poisson_plan_ld(g->plan);
        const int64_t rows = glb_poisson_plan_rows(g->plan);          // n, or n + 1 with the library's scratch row
        GLB_CUDA(g->Aw.alloc(&g->src64, n * c));  GLB_CUDA(g->Aw.alloc(&g->Db, rows * ld));
        GLB_CUDA(g->Aw.alloc(&g->u0, rows * ld)); GLB_CUDA(g->Aw.alloc(&g->u1, rows * ld));
        GLB_CUDA(cudaMemsetAsync(g->Db, 0, rows * ld * sizeof(float), st));
        GLB_CUDA(cudaMemsetAsync(g->u1, 0, rows * ld * sizeof(float), st));
        g->rows = rows;
        g->ldu = ld; g->c_plan = c;
        tm.lap("plan_create + buffers");
    }
    const int ldu = g->ldu;
    if (m > g->m_cap) { GLB_CUDA(g->Aw.alloc(&g->tind, m)); g->m_cap = m; }
    if (h_source) {
        GLB_CUDA(cudaMemcpyAsync(g->src64, h_source, n * c * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
        // numpy's source[ind] = rows: negative indices count from the end, of a repeated index the LAST row stays
        std::vector<std::pair<long long, long long>> order((size_t)m_rows);
        for (int64_t i = 0; i < m_rows; ++i) {
            long long t = h_row_ind[i];
            if (t < 0) t += n;
            GLB_CHECK_ARG(t >= 0 && t < n, "source row index out of range");
            order[(size_t)i] = {t, (long long)i};
        }
        std::sort(order.begin(), order.end());
        const size_t stride = 8 + 8 * (size_t)c;
        std::vector<unsigned char> stage((size_t)m_rows * stride + 8);
        long long *s_idx = reinterpret_cast<long long *>(stage.data());
        int64_t mu = 0;
        for (int64_t i = 0; i < m_rows; ++i)
            if (i + 1 == m_rows || order[(size_t)i + 1].first != order[(size_t)i].first) s_idx[mu++] = order[(size_t)i].second;
        // s_idx holds positions for now; rows go behind the mu indices
        double *s_rows = reinterpret_cast<double *>(stage.data() + 8 * (size_t)mu);
        for (int64_t q = 0; q < mu; ++q) {
            const long long pos = s_idx[q];
            memcpy(s_rows + (size_t)q * c, h_rows + (size_t)pos * c, sizeof(double) * (size_t)c);
            long long t = h_row_ind[pos];
            s_idx[q] = t < 0 ? t + n : t;
        }
        const int64_t bytes = (int64_t)(mu * stride);
        if (bytes > g->rows_cap) { GLB_CUDA(g->Aw.alloc(&g->rows_stage, (size_t)bytes)); g->rows_cap = bytes; }
        GLB_CUDA(cudaMemsetAsync(g->src64, 0, n * c * sizeof(double), st));
        if (mu) {
            // pageable source: the call returns once the bytes sit in the driver's staging buffer, `stage` may go out of scope
            GLB_CUDA(cudaMemcpyAsync(g->rows_stage, stage.data(), (size_t)bytes, cudaMemcpyHostToDevice, st));
            scatter_source_rows_kernel<<<ceil_div(mu * c, 256), 256, 0, st>>>(reinterpret_cast<const long long *>(g->rows_stage),
                                                                             reinterpret_cast<const double *>(g->rows_stage + 8 * (size_t)mu), mu, c, g->src64);
            nl += 1;
        }
    }
    if ((rc = glb_poisson_pack(g->plan, g->src64, g->deg, g->perm, g->Db, st))) return rc;
    nl += 1;

    // iteration count by the reference's stopping rule (ssl.py:639-644, 667, 669)
    int T = max_iter;
    if (min_iter < max_iter) {
        GLB_CUDA(cudaMemcpyAsync(g->tind, h_train_ind, m * sizeof(long long), cudaMemcpyHostToDevice, st));
        GLB_CUDA(cudaMemsetAsync(g->v, 0, n * sizeof(double), st));
        if (m) mixing_seed_kernel<<<ceil_div(m, 256), 256, 0, st>>>(g->tind, m, n, g->v);
        partial_sum_kernel<<<NPART, 256, 0, st>>>(g->v, n, g->part);
        scale_by_sum_kernel<<<sm_count() * 4, 256, 0, st>>>(g->v, g->part, NPART, n);
        nl += 3;
        GLB_LAUNCH_CHECK();
        if ((rc = glb_poisson_mixing_T(g->t_rp, g->t_col, g->rw_val, g->vinf, g->v, g->vtmp, n, min_iter, max_iter, &T, &nl,
                                       st)))
            return rc;
    }
    GLB_CUDA(cudaMemsetAsync(g->u0, 0, g->rows * ldu * sizeof(float), st));
    int in_u1 = 0;
    if ((rc = glb_poisson_iterate(g->plan, g->Db, g->u0, g->u1, T, &in_u1, &nl, st))) return rc;
    if ((rc = glb_poisson_unpack(g->plan, in_u1 ? g->u1 : g->u0, g->perm, g->src64, st))) return rc;
    nl += 1;
    GLB_CUDA(cudaMemcpyAsync(h_u_out, g->src64, n * c * sizeof(double), cudaMemcpyDeviceToHost, st));
    if ((rc = glb_poisson_plan_check(g->plan, st))) return rc;          // synchronises
    tm.lap("upload + iterate + download");
    if (T_done) *T_done = T;
    if (launches) *launches = nl;
    return 0;
}

extern "C" GLB_API int glb_poisson_graph_fit(glb_poisson_graph *g, const double *h_source, int c, const int64_t *h_train_ind,
                                             int64_t m, int min_iter, int max_iter, double *h_u_out, int *T_done,
                                             int *launches)
{
    GLB_CHECK_ARG(h_source, "null pointer");
    return poisson_graph_fit_impl(g, h_source, nullptr, nullptr, 0, c, h_train_ind, m, min_iter, max_iter, h_u_out, T_done, launches);
}

extern "C" GLB_API int glb_poisson_graph_fit_rows(glb_poisson_graph *g, const int64_t *h_row_ind, const double *h_rows,
                                                  int64_t m_rows, int c, const int64_t *h_train_ind, int64_t m, int min_iter,
                                                  int max_iter, double *h_u_out, int *T_done, int *launches)
{
    GLB_CHECK_ARG(m_rows >= 0, "m_rows must be >= 0");
    return poisson_graph_fit_impl(g, nullptr, h_row_ind, h_rows, m_rows, c, h_train_ind, m, min_iter, max_iter, h_u_out, T_done,
                                  launches);
}

// Page-locked host memory for the callers' result buffers: a device-to-host copy into it runs at PCIe speed and needs no
// staging copy (a 5.6 MB pageable numpy array costs ~1 ms, this ~0.15 ms).
extern "C" GLB_API int glb_host_alloc(int64_t bytes, void **h_ptr)
{
    GLB_CHECK_ARG(bytes > 0 && h_ptr, "bad argument");
    GLB_CUDA(cudaHostAlloc(h_ptr, (size_t)bytes, cudaHostAllocDefault));
    return 0;
}
extern "C" GLB_API int glb_host_free(void *h_ptr)
{
    if (h_ptr) GLB_CUDA(cudaFreeHost(h_ptr));
    return 0;
}

extern "C" GLB_API int glb_poisson_gd_host(const int32_t *h_rowptr, const int32_t *h_col, const double *h_val, int64_t n,
                                           int64_t nnz, const double *h_source, int c, const int64_t *h_train_ind,
                                           int64_t m, int min_iter, int max_iter, double *h_u_out, int *T_done,
                                           int *launches)
{
    GLB_CHECK_ARG(h_rowptr && h_col && h_val && h_source && h_u_out, "null pointer");
    glb_poisson_graph *g = nullptr;
    int rc = glb_poisson_graph_create(&g, h_rowptr, h_col, h_val, n, nnz, -1);
    if (rc) return rc;
    rc = glb_poisson_graph_fit(g, h_source, c, h_train_ind, m, min_iter, max_iter, h_u_out, T_done, launches);
    if (rc == 0 && launches) *launches += g->setup_launches;
    glb_poisson_graph_destroy(g);
    return rc;
}
