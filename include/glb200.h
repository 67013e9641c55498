/* glb200.h - C-ABI of libglb200.so: the B200 (sm_100a) implementation of GraphLearning's
 * data-parallel hot path (kNN graph build -> Laplacian normalisation -> repeated sparse x dense
 * iterate of ssl.poisson / ssl.laplace).
 *
 * The reference (jwcalder/GraphLearning v1.7.5) has no plugin registry; its only FFI is the CPython
 * module c_code/cextensions.cpp:346-373 (raw pointers + sizes, in-place mutation, None return).
 * Every entry point below states which reference code it replaces (file:line under /root/reference).
 *
 * Conventions
 *  - every function returns 0 on success, a negative GLB_E_* for argument errors, a positive
 *    cudaError_t for CUDA failures; glb_last_error() returns the message of the calling thread.
 *  - "d_" pointers are DEVICE pointers, "h_" pointers are HOST pointers.  The caller owns every
 *    buffer; the library owns only opaque plan handles (explicit create/destroy).
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Device entry points are
 *    stream-ordered and do not synchronise unless stated.
 *  - CSR matrices: int32 rowptr[n+1], int32 col[nnz], values fp32 (iterate) or fp64 (setup).
 *  - dense label matrices on the device are row-major n x ld fp32 in the layout of the glb_poisson_plan
 *    they are used with (glb_poisson_plan_ld; see the Poisson section).
 */
#ifndef GLB200_H
#define GLB200_H
#include <stdint.h>

#if defined(__GNUC__)
#define GLB_API __attribute__((visibility("default")))
#else
#define GLB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GLB_E_INVALID   (-1)   /* bad argument (null pointer, negative size, unsupported option) */
#define GLB_E_NOGPU     (-2)   /* no CUDA device / wrong architecture */
#define GLB_E_NOMEM     (-3)
#define GLB_E_UNSUPPORTED (-4)
#define GLB_E_TIMEOUT   (-5)   /* a polling loop of a persistent kernel hit its watchdog; results are invalid */

GLB_API int glb_version(void);
/* The library keeps freed device scratch (plans, search buffers) mapped in the CUDA stream-ordered memory pool so that the
 * next call does not pay the driver's map / unmap again; this returns all of it to the driver.  Synchronises the device. */
GLB_API int glb_release_workspace(void);
GLB_API const char *glb_last_error(void);
/* Number of SMs etc. of the current device; fails with GLB_E_NOGPU when there is none. */
GLB_API int glb_device_info(int *sm_count, int *cc_major, int *cc_minor, int64_t *hbm_bytes);
/* Leading dimension of the plain layout of an n x c label matrix: power of two in {4..128} or a multiple of 128. */
GLB_API int glb_padded_ld(int c);

/* ---------------------------------------------------------------------------------------------
 * Graph normalisation.   graph.degree_vector (graphlearning/graph.py:108-122), degree_matrix(p=-1)
 * (:210-233) and the setup lines of ssl.poisson._fit (graphlearning/ssl.py:615-616, 634-644).
 * ------------------------------------------------------------------------------------------- */
/* deg[i] = sum_j W[i,j] (fp64, stored order).  Replaces W*ones (graph.py:121).  skip_diagonal != 0 leaves
 * W[i,i] out (the degrees of W - diag(W), ssl.py:615-617); d_col may be NULL otherwise. */
GLB_API int glb_csr_degree(const int32_t *d_rowptr, const int32_t *d_col, const double *d_val, int64_t n, int skip_diagonal,
                   double *d_deg, void *stream);

/* Transpose a CSR matrix (fp64 values); output is canonical (columns sorted inside each row).
 * Replaces W.transpose() + the CSC->CSR conversion scipy does inside ssl.py:635,644.
 * d_work: scratch of glb_csr_transpose_work_bytes(nnz) bytes. */
GLB_API int64_t glb_csr_transpose_work_bytes(int64_t n, int64_t nnz);
GLB_API int glb_csr_transpose(const int32_t *d_rowptr, const int32_t *d_col, const double *d_val, int64_t n, int64_t nnz,
                      int32_t *d_t_rowptr, int32_t *d_t_col, double *d_t_val, void *d_work, int64_t work_bytes,
                      void *stream);

/* From Wt = W^T (CSR, fp64; diagonal entries are treated as zero, ssl.py:615-616) and deg = W*1:
 *   P_val[j]  = (float)( Wt_val[j] / deg[row(j)] )      P  = D^-1 W^T   (ssl.py:634-635)
 *   RW_val[j] =          Wt_val[j] / deg[col(j)]        RW = W^T D^-1   (ssl.py:644)
 * Either output may be NULL. */
GLB_API int glb_poisson_scale(const int32_t *d_t_rowptr, const int32_t *d_t_col, const double *d_t_val, const double *d_deg,
                      int64_t n, float *d_P_val, double *d_RW_val, void *stream);

/* Locality ordering (no counterpart in the reference: scipy's single-core SpMM does not care about node
 * numbering).  Reverse Cuthill-McKee on the HOST from the CSR pattern the caller holds there; integers only.
 * h_perm[new] = old.  glb_csr_permute relabels an fp32 CSR matrix on the device: B = Pi A Pi^T. */
GLB_API int glb_locality_order_host(const int32_t *h_rowptr, const int32_t *h_col, int64_t n, int32_t *h_perm);
#define GLB_DATAFLOW_LONG_ROW 32   /* dataflow kernel: rows with more nonzeros are dealt over whole warps */
GLB_API int glb_csr_permute(const int32_t *d_rowptr, const int32_t *d_col, const float *d_val, int64_t n, int64_t nnz,
                            const int32_t *d_perm, int32_t *d_iperm, int32_t *d_out_rowptr, int32_t *d_out_col,
                            float *d_out_val, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Poisson iterate  u <- Db + P u.   Replaces the loop body graphlearning/ssl.py:668 (CPU: scipy
 * csr_matvecs) and :658 (torch.sparse.addmm).
 *
 * A plan binds the CSR matrix P (device arrays owned by the caller, which must outlive the plan) and the
 * number of label columns c to one of three kernels:
 *   GLB_POISSON_KIND_DATAFLOW  all T iterations in ONE persistent launch without any grid barrier: every
 *                              16-byte chunk of a label row carries the iteration that produced it, so the
 *                              gathers themselves synchronise.  Needs a structurally symmetric pattern,
 *                              c <= 96 and a graph whose per-SM slab fits shared memory.
 *   GLB_POISSON_KIND_BARRIER   all T iterations in one cooperative launch with a grid barrier per iteration
 *                              (directed graphs).
 *   GLB_POISSON_KIND_STEP      one launch per iteration (graphs too large for the shared-memory slabs).
 * GLB_POISSON_KIND_AUTO picks the first applicable one in that order; asking for a specific kind that does
 * not apply returns GLB_E_UNSUPPORTED.
 *
 * Device label matrices (Db, u) are row-major glb_poisson_plan_rows(plan) x glb_poisson_plan_ld(plan) fp32 in the
 * plan's layout: plain (columns 0..c-1, zero padded) or, for the dataflow kernel, chunks {x[3q], x[3q+1], x[3q+2],
 * epoch}.  rows = n, or n + 256 for the dataflow kernel: rows n.. are scratch rows owned by the library (the padding
 * entries of its slabs point there); allocate them, never read them.
 * glb_poisson_pack / glb_poisson_unpack convert rows 0..n-1 from / to the reference's n x c float64 arrays.
 * ------------------------------------------------------------------------------------------- */
#define GLB_POISSON_KIND_AUTO     (-1)
#define GLB_POISSON_KIND_STEP     0
#define GLB_POISSON_KIND_BARRIER  1
#define GLB_POISSON_KIND_DATAFLOW 2
typedef struct glb_poisson_plan glb_poisson_plan;
GLB_API int glb_poisson_plan_create(glb_poisson_plan **plan, const int32_t *d_rowptr, const int32_t *d_col,
                                    const float *d_val, int64_t n, int64_t nnz, int c, int kind, void *stream);
GLB_API int glb_poisson_plan_destroy(glb_poisson_plan *plan);
GLB_API int glb_poisson_plan_kind(const glb_poisson_plan *plan);
GLB_API int glb_poisson_plan_ld(const glb_poisson_plan *plan);
GLB_API int64_t glb_poisson_plan_rows(const glb_poisson_plan *plan);
/* Synchronises the stream and reports GLB_E_TIMEOUT if a polling loop of the dataflow kernel hit its ~2 s watchdog since
 * the last check (the kernel then drains instead of hanging the GPU; the results of that launch are invalid). */
GLB_API int glb_poisson_plan_check(glb_poisson_plan *plan, void *stream);
/* nnz / stored entries of the dataflow kernel's sliced-ELL slabs (1.0 = no padding; 0 for other kinds) */
GLB_API double glb_poisson_plan_fill(const glb_poisson_plan *plan);
/* iterations between two re-alignment gates of the dataflow kernel (tuned at plan time; 0 for other kinds) */
GLB_API int glb_poisson_plan_gate(const glb_poisson_plan *plan);
/* Host-only self-check of the dataflow slab builder (no GPU): builds the slabs for `grid` CTAs and walks them as the
 * kernel does, computing y = P x in double precision.  out4 = {max |y - P x| / max |P x|, warp-wide gather
 * instructions per iteration, fill, rows not stored exactly once}. */
GLB_API int glb_dataflow_slabs_check_host(const int32_t *h_rowptr, const int32_t *h_col, const float *h_val, int64_t n, int c,
                                          int grid, double *out4);

/* d_dst (n x ld fp32, plan layout) <- d_src (n x c fp64), each row divided by d_deg[row] when d_deg is not
 * NULL (Db = D^-1 source, ssl.py:636).  d_perm (may be NULL) is a locality ordering: device row r holds the
 * caller's row d_perm[r].  glb_poisson_unpack is the inverse (no scaling). */
GLB_API int glb_poisson_pack(const glb_poisson_plan *plan, const double *d_src, const double *d_deg, const int32_t *d_perm,
                             float *d_dst, void *stream);
GLB_API int glb_poisson_unpack(const glb_poisson_plan *plan, const float *d_src, const int32_t *d_perm, double *d_dst,
                               void *stream);

/* One iteration, one kernel launch: d_u_out = d_Db + P * d_u_in (any plan kind).  u_in != u_out. */
GLB_API int glb_poisson_step(const glb_poisson_plan *plan, const float *d_Db, const float *d_u_in, float *d_u_out,
                             void *stream);

/* T iterations starting from d_u0.  The result is in d_u0 when T is even, d_u1 when T is odd;
 * *result_in_u1 (host int, may be NULL) says which.  launches (host, may be NULL) += kernels launched. */
GLB_API int glb_poisson_iterate(glb_poisson_plan *plan, const float *d_Db, float *d_u0, float *d_u1, int T,
                                int *result_in_u1, int *launches, void *stream);

/* Stopping rule of ssl.py:667,669: v <- RW v (fp64) from v0 = indicator(train)/m until
 * T >= min_iter and max|v - vinf| <= 1/n, or T == max_iter.  Synchronises the stream (T is a host
 * value).  d_v is overwritten; d_tmp is n doubles of scratch. */
GLB_API int glb_poisson_mixing_T(const int32_t *d_rw_rowptr, const int32_t *d_rw_col, const double *d_rw_val,
                         const double *d_vinf, double *d_v, double *d_tmp, int64_t n, int min_iter, int max_iter,
                         int *T_out, int *launches, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Row-slab Poisson iterate: one block of rows of u <- Db + P u per launch (the loop body
 * graphlearning/ssl.py:668 for rows [r0, r1)), for graphs beyond the shared-memory kernels above and for the
 * row-partitioned multi-GPU run (one slab per GPU, one process per GPU).  No counterpart in the reference, which is
 * single process; what it replaces is scipy's csr_matvecs on the row block.
 *
 * Local index space of a slab: own rows 0..m-1, then its n_halo halo rows (label rows owned by other ranks that a
 * column of the slab points to), then one all-zero scratch row.  glb_slab_create takes the slab's CSR in that space
 * (h_col in [0, m + n_halo), fp32 values = D^-1 W^T rounded once, rows summed in stored order), the rows that must
 * run first (h_boundary[r] != 0: rows a peer needs or rows that read halo rows; NULL = none) and, per row, the puts
 * that deliver it (CSR over the m rows: h_send_ptr[m+1], peer rank, destination row in THAT peer's local space).
 * HOST pointers; the device copy (sliced ELL, see csrc/slab.cu) is owned by the handle.
 *
 * Label matrices live in one "region" per rank: 64 uint32 flag words followed by three (m + n_halo + 1) x
 * glb_padded_ld(c) fp32 buffers (version v of u in buffer v % 3).  The owner allocates its region with
 * glb_ipc_alloc(glb_slab_region_bytes) and sends the 64-byte handle to its peers, which map it with glb_ipc_open
 * (CUDA IPC; all GPUs on one NVLink/NVSwitch node).  glb_slab_attach binds the slab to the regions of all ranks
 * (region_base[rank] = own region; neighbour_mask = ranks this one exchanges rows with, its own bit clear;
 * world <= 16).  A single-GPU run is world = 1, neighbour_mask = 0.
 *
 * glb_slab_reset zeroes the three buffers (u_0 = 0, ssl.py:638); synchronise stream and ranks before
 * glb_slab_iterate, which enqueues T launches of the step kernel: boundary rows first - they wait for the
 * neighbours' flags of this version, are written locally AND into every peer that gathers them (stores to mapped
 * peer memory over NVLink, from the kernel that computes them) and are followed by a release of this rank's flag in
 * every neighbour's region - then the interior rows, which overlap the transfer.  d_Db: m x ld, plain layout.
 * *result_buffer = T % 3.  glb_slab_check synchronises and reports GLB_E_TIMEOUT if a wait for a neighbour ran into
 * its ~3 s watchdog.  glb_slab_pack / glb_slab_unpack convert the m own rows from / to m x c float64.
 * ------------------------------------------------------------------------------------------- */
typedef struct glb_slab glb_slab;
GLB_API int glb_slab_create(glb_slab **slab, const int32_t *h_rowptr, const int32_t *h_col, const float *h_val, int64_t m,
                            int64_t n_halo, int c, const uint8_t *h_boundary, const int64_t *h_send_ptr,
                            const int32_t *h_send_peer, const int32_t *h_send_dst, void *stream);
GLB_API int glb_slab_destroy(glb_slab *slab);
/* Host-only self-check of the slab builder (no GPU): builds the entry stream of one slab as glb_slab_create does for a device of
 * `sms` SMs and walks it tile by tile, warp by warp as slab_step_kernel does, computing y = P x in double precision.
 * out8 = {max |y - P x| / max |P x|, structural errors (rows not stored exactly once, padding that is not (scratch row, 0), warp
 * parts beyond their region), rows on the wrong side of the boundary / interior split, fill, slices per tile, bytes of a warp's
 * stream region, longest warp part / mean warp part, tiles}. */
GLB_API int glb_slab_check_host(const int32_t *h_rowptr, const int32_t *h_col, const float *h_val, int64_t m, int64_t n_halo, int c,
                                const uint8_t *h_boundary, int sms, double *out8);
GLB_API int64_t glb_slab_rows(const glb_slab *slab);          /* m + n_halo + 1 */
GLB_API int glb_slab_ld(const glb_slab *slab);
GLB_API double glb_slab_fill(const glb_slab *slab);           /* nnz / stored entries of the sliced ELL */
GLB_API int glb_slab_tile_slices(const glb_slab *slab);       /* slices per tile the library chose for this slab (16 or 32) */
GLB_API int64_t glb_slab_region_bytes(const glb_slab *slab);
GLB_API int glb_slab_attach(glb_slab *slab, int rank, int world, void *const *region_base, const int64_t *region_rows,
                            uint32_t neighbour_mask);
GLB_API int glb_slab_buffer(const glb_slab *slab, int v, float **d_buf);
GLB_API int glb_slab_pack(const glb_slab *slab, const double *d_src, const double *d_deg, float *d_dst, void *stream);
GLB_API int glb_slab_unpack(const glb_slab *slab, int v, double *d_dst, void *stream);
GLB_API int glb_slab_reset(glb_slab *slab, void *stream);
GLB_API int glb_slab_iterate(glb_slab *slab, const float *d_Db, int T, int *result_buffer, int *launches, void *stream);
GLB_API int glb_slab_check(glb_slab *slab, void *stream);
/* Peer-mappable device memory (cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle).  handle64: 64 bytes. */
GLB_API int glb_ipc_alloc(int64_t bytes, void **d_ptr, void *handle64);
GLB_API int glb_ipc_open(const void *handle64, void **d_ptr);
GLB_API int glb_ipc_close(void *d_ptr);
GLB_API int glb_ipc_free(void *d_ptr);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer entry points: the gradient-descent branch of ssl.poisson._fit (graphlearning/ssl.py:615-670)
 * with HOST inputs/outputs in the reference's own dtypes - what the reference's Python binds in place of
 * its `while` loop.
 *
 * glb_poisson_graph_create uploads the scipy CSR weight matrix W (canonical or not, diagonal ignored) and
 * builds, on the device, everything the reference recomputes in every _fit call: degrees, P = D^-1 W^T,
 * RW = W^T D^-1, vinf = deg/sum(deg).  reorder: 0 = keep the node numbering, 1 = relabel with a locality
 * ordering (results are still returned in the caller's numbering), -1 = automatic.
 * glb_poisson_graph_fit runs one fit on that graph: source is the n x c fp64 Poisson source term
 * (ssl.py:619-622), train_ind the m labelled nodes (used by the stopping rule of ssl.py:639-641,667,669
 * when min_iter < max_iter; T = max_iter otherwise), u_out the n x c fp64 scores.  Synchronous.
 * glb_poisson_graph_fit_rows is the same fit with the source term given by its nonzero rows, exactly as
 * ssl.py:619-622 builds it (source = zeros((n, c)); source[row_ind] = rows - numpy semantics: negative indices
 * count from the end, of a repeated index the last row stays): m_rows x c doubles cross PCIe instead of n x c.
 * glb_poisson_gd_host = create + fit + destroy.  T_done / launches may be NULL.
 * glb_host_alloc / glb_host_free: page-locked host memory for result buffers (h_u_out): the download then runs at
 * PCIe speed without a staging copy.  Any host pointer is accepted by the entry points; pinned ones are faster.
 * ------------------------------------------------------------------------------------------- */
typedef struct glb_poisson_graph glb_poisson_graph;
GLB_API int glb_poisson_graph_create(glb_poisson_graph **graph, const int32_t *h_rowptr, const int32_t *h_col,
                                     const double *h_val, int64_t n, int64_t nnz, int reorder);
GLB_API int glb_poisson_graph_destroy(glb_poisson_graph *graph);
GLB_API int glb_poisson_graph_fit(glb_poisson_graph *graph, const double *h_source, int c, const int64_t *h_train_ind,
                                  int64_t m, int min_iter, int max_iter, double *h_u_out, int *T_done, int *launches);
GLB_API int glb_poisson_graph_fit_rows(glb_poisson_graph *graph, const int64_t *h_row_ind, const double *h_rows, int64_t m_rows,
                                       int c, const int64_t *h_train_ind, int64_t m, int min_iter, int max_iter,
                                       double *h_u_out, int *T_done, int *launches);
GLB_API int glb_host_alloc(int64_t bytes, void **h_ptr);
GLB_API int glb_host_free(void *h_ptr);
GLB_API int glb_poisson_gd_host(const int32_t *h_rowptr, const int32_t *h_col, const double *h_val, int64_t n, int64_t nnz,
                                const double *h_source, int c, const int64_t *h_train_ind, int64_t m, int min_iter,
                                int max_iter, double *h_u_out, int *T_done, int *launches);

/* ---------------------------------------------------------------------------------------------
 * Graph Laplacians and the linear system of Laplace learning, assembled on the device (csrc/laplace.cu).
 *
 * glb_laplacian_csr_host replaces graph.laplacian (graphlearning/graph.py:469-513):
 *     L = Diag(h_diag) - Diag(h_left) W Diag(h_right)        h_left / h_right may be NULL (= identity)
 *     combinatorial: diag = d;   randomwalk: diag = 1, left = d^-1;   normalized: diag = 1, left = right = d^-1/2
 * with d = W 1 and d^p computed by the caller exactly as the reference does (numpy, n values).  W: canonical scipy
 * CSR (int32 / float64).  Every entry is rounded as scipy rounds it ((left_i w_ij) right_j, then the difference);
 * exact zeros are dropped, columns are sorted, the diagonal is always present unless it is zero.  Outputs: HOST
 * arrays, h_out_col / h_out_val with room for nnz + n entries; h_out_rowptr[n] is the number stored.
 *
 * glb_laplace_fit_host replaces ssl.laplace._fit for order = 1 (graphlearning/ssl.py:1222-1255): the matrix
 * tau + L above, restricted to the unlabelled nodes and Jacobi scaled (M A M, M b with M = diag(A)^-1/2, b = -L[:, train] F),
 * is assembled in HBM, solved by glb_cg_solve (tol, at most 1e5 iterations, as utils.conjgrad) and scattered back:
 * h_u (n x c) = M v on the unlabelled nodes, F on the labelled ones.  h_tau: n values or NULL; h_train_ind: m distinct
 * nodes in [0, n); h_F: m x c one-hot labels (utils.labels_to_onehot).  iters / err / launches / h_ms may be NULL;
 * h_ms[3] = {device milliseconds of the CG iterations (CUDA events), stored entries of the system matrix, unknowns}.
 * ------------------------------------------------------------------------------------------- */
GLB_API int glb_laplacian_csr_host(const int32_t *h_rowptr, const int32_t *h_col, const double *h_val, int64_t n, int64_t nnz,
                                   const double *h_left, const double *h_right, const double *h_diag, int32_t *h_out_rowptr,
                                   int32_t *h_out_col, double *h_out_val);
GLB_API int glb_laplace_fit_host(const int32_t *h_rowptr, const int32_t *h_col, const double *h_val, int64_t n, int64_t nnz,
                                 const double *h_left, const double *h_right, const double *h_diag, const double *h_tau,
                                 const int64_t *h_train_ind, int64_t m, const double *h_F, int c, double tol, double *h_u,
                                 int64_t *iters, double *err, int *launches, double *h_ms);
/* The same fit with the weight matrix and the scalings resident in HBM across fits (ssl_trials: hundreds of fits on one
 * graph): create uploads W, left/right/diag/tau once, every fit then moves only train_ind, F and the result.
 * glb_laplace_fit_host = create + fit + destroy. */
typedef struct glb_laplace_graph glb_laplace_graph;
GLB_API int glb_laplace_graph_create(glb_laplace_graph **graph, const int32_t *h_rowptr, const int32_t *h_col, const double *h_val,
                                     int64_t n, int64_t nnz, const double *h_left, const double *h_right, const double *h_diag,
                                     const double *h_tau);
GLB_API int glb_laplace_graph_fit(glb_laplace_graph *graph, const int64_t *h_train_ind, int64_t m, const double *h_F, int c,
                                  double tol, double *h_u, int64_t *iters, double *err, int *launches, double *h_ms);
GLB_API int glb_laplace_graph_destroy(glb_laplace_graph *graph);

/* ---------------------------------------------------------------------------------------------
 * Conjugate gradient with c right-hand sides.  Replaces utils.conjgrad (graphlearning/utils.py:483-532):
 * per-column alpha/beta, one stopping norm err = sqrt(sum over all columns of |r|^2), loop test
 * `err > tol and i < max_iter` with err initialised to 1.  Called by ssl.laplace._fit (ssl.py:1249) and the
 * default solver of ssl.poisson._fit (ssl.py:624-629).  fp64 throughout, dot products in a fixed order.
 * c <= 128 (GLB_E_UNSUPPORTED otherwise).
 *
 * glb_cg_solve: device arrays; A = CSR fp64, b / x0 / x row-major n x glb_padded_ld(c) fp64 (plain layout, zero
 * padded), x0 may be NULL.  Synchronises the stream once per 64 iterations (iters / err are host values).
 * glb_cg_host: the same with HOST buffers in the reference's own types (scipy CSR int32/float64, numpy
 * float64 n x c; a 1-D right-hand side is c = 1).
 * ------------------------------------------------------------------------------------------- */
GLB_API int64_t glb_cg_work_bytes(int64_t n, int c);
GLB_API int glb_cg_solve(const int32_t *d_rowptr, const int32_t *d_col, const double *d_val, int64_t n, int64_t nnz,
                         const double *d_b, const double *d_x0, int c, double tol, int64_t max_iter, double *d_x, void *d_work,
                         int64_t work_bytes, int64_t *iters, double *err, int *launches, void *stream);
GLB_API int glb_cg_host(const int32_t *h_rowptr, const int32_t *h_col, const double *h_val, int64_t n, int64_t nnz,
                        const double *h_b, const double *h_x0, int c, double tol, int64_t max_iter, double *h_x,
                        int64_t *iters, double *err, int *launches);

/* ---------------------------------------------------------------------------------------------
 * Exact k-nearest-neighbour search, Euclidean.  Replaces weightmatrix.knnsearch
 * (graphlearning/weightmatrix.py:297-429): the exact branches (cKDTree :349-352, brute force :354-361) and,
 * as a quality superset, the approximate annoy default (:368-409).  For similarity='angular' pass
 * row-normalised features (:344-345).
 *
 * X: n x d float64 row-major.  ind: n x k int64, dist: n x k float64, neighbours INCLUDING self, ascending by
 * fp64 distance (equal distances ordered by index).  The result is that of an exact fp64 ranking: an fp32
 * tiled distance pass picks 32..128 candidates per row, they are re-ranked in fp64 from the original
 * features, and a per-row error-margin certificate proves that no true neighbour was missed; rows without
 * certificate are redone by fp64 brute force (*fallback_rows counts them, may be NULL).  k <= 112.
 * glb_knn_search: device pointers, allocates its scratch internally, synchronises the stream before
 * returning.  glb_knn_search_host: host pointers.
 * ------------------------------------------------------------------------------------------- */
GLB_API int glb_knn_search(const double *d_X, int64_t n, int d, int k, int64_t *d_ind, double *d_dist, int *launches,
                           int *fallback_rows, void *stream);
GLB_API int glb_knn_search_host(const double *h_X, int64_t n, int d, int k, int64_t *h_ind, double *h_dist, int *launches,
                                int *fallback_rows);

/* ---------------------------------------------------------------------------------------------
 * p-Laplace / AMLE neighbour sweeps.  Replace the reference's C extension entry points
 * cextensions.lp_iterate / cextensions.lip_iterate (c_code/cextensions.cpp:19-107), i.e.
 * lp_iterate_main (c_code/lp_iterate.cpp:35-125, Jacobi, upper + lower barrier function),
 * lip_iterate_main (:129-187, Gauss-Seidel) and lip_iterate_weighted_main (:190-259, Gauss-Seidel with a
 * 30-step bisection per row), called from graph.plaplace / graph.amle (graphlearning/graph.py:1261,1276,1330).
 *
 * Same arguments as the C functions they replace (lp_iterate.h:33-35): the graph as row-sorted COO triplets
 * exactly as graph.__ccode_init__ (graph.py:69-84) prepares them - nbr = neighbour (column) index of entry k
 * [the reference's `I`/`II` argument], row = row index, ascending [`J`], w = weight; M entries; ind/val = m
 * Dirichlet nodes and values; uu/ul/u are n doubles, read as the initial iterate and overwritten with the
 * result as the reference leaves it in the caller's arrays (lp: the last ODD sweep, because the reference swaps
 * its buffer pointers, :116-123).  weighted != 0 selects lip_iterate_weighted_main (alpha/beta ignored).
 * The iterates are bit-identical to the reference's at every sweep count (fp64, same operation order), the
 * Gauss-Seidel sweeps included: they run as a dependency dataflow, not as a Jacobi substitute.
 * HOST pointers; synchronous; *sweeps = sweeps executed; sweeps / launches may be NULL.
 * ------------------------------------------------------------------------------------------- */
GLB_API int glb_lp_iterate_host(double *h_uu, double *h_ul, const int32_t *h_nbr, const int32_t *h_row, const double *h_w,
                                const int32_t *h_ind, const double *h_val, double p, int T, double tol, int n, int M, int m,
                                int *sweeps, int *launches);
GLB_API int glb_lip_iterate_host(double *h_u, const int32_t *h_nbr, const int32_t *h_row, const double *h_w,
                                 const int32_t *h_ind, const double *h_val, int T, double tol, int weighted, double alpha,
                                 double beta, int n, int M, int m, int *sweeps, int *launches);
/* The same sweeps for c right-hand sides at once - the one-vs-rest loop of ssl.fit (graphlearning/ssl.py:469-474) over
 * ssl.plaplace / ssl.amle, whose classes share the graph and the Dirichlet rows.  u: n x c row-major (in/out), val: m x c
 * row-major, sweeps: c ints (each class keeps its own stopping sweep).  Column k of the result is bit-identical to
 * glb_lip_iterate_host on column k.  c <= 32. */
GLB_API int glb_lip_iterate_multi_host(double *h_u, const int32_t *h_nbr, const int32_t *h_row, const double *h_w,
                                       const int32_t *h_ind, const double *h_val, int T, double tol, int weighted,
                                       double alpha, double beta, int n, int M, int m, int c, int *sweeps, int *launches);

/* ---------------------------------------------------------------------------------------------
 * fp64 block operations for the spectral path: graph.eigen_decomp (graphlearning/graph.py:623-806, ARPACK svds
 * at :734,756 and utils.randomized_svd at :736,758) and utils.randomized_svd itself
 * (graphlearning/utils.py:576-642, whose cost is the repeated product Y <- A (A^T Y), :618-621).
 * Tall-skinny matrices are row-major n x ld fp64 with c <= 256 used columns; ld even and >= c rounded up to
 * even for glb_spmm_f64 (16-byte accesses), ld >= c for the other two.  Device pointers, stream-ordered.
 *
 * glb_spmm_f64:       Z = alpha * A X + beta * Y1 * diag(bcol) + gamma * Y2.   A = CSR fp64; Y1, bcol, Y2 may be
 *                     NULL (bcol NULL = all ones); Z may alias Y1 / Y2 but not X.  Replaces scipy csr_matvecs /
 *                     csr_matmat at the call sites above, with the Chebyshev / power recurrence fused in.
 * glb_gram_f64:       G = X^T Y, c1 x c2 row-major (device), partial sums reduced in a fixed order.
 *                     d_work: glb_gram_work_bytes(c1, c2) bytes.
 * glb_right_mul_f64:  Y = X S, S = c1 x c2 row-major (device); padding columns of Y (c2..ldy-1) are zeroed.
 * ------------------------------------------------------------------------------------------- */
GLB_API int glb_spmm_f64(const int32_t *d_rowptr, const int32_t *d_col, const double *d_val, int64_t n, const double *d_X,
                         int ldx, double *d_Z, int ldz, int c, double alpha, const double *d_Y1, int ldy1, double beta,
                         const double *d_bcol, const double *d_Y2, int ldy2, double gamma, void *stream);
GLB_API int64_t glb_gram_work_bytes(int c1, int c2);
GLB_API int glb_gram_f64(const double *d_X, int ldx, int c1, const double *d_Y, int ldy, int c2, int64_t n, double *d_G,
                         void *d_work, int64_t work_bytes, void *stream);
GLB_API int glb_right_mul_f64(const double *d_X, int ldx, int64_t n, int c1, const double *d_S, int c2, double *d_Y, int ldy,
                              void *stream);

/* ---------------------------------------------------------------------------------------------
 * Label post-processing that follows the iterate (device pointers, stream-ordered).
 *
 * glb_volume_projection: ssl.volume_label_projection (graphlearning/ssl.py:172-209) - the projection step of PoissonMBO
 *   (:826-829) and of every model fitted with class_priors (:476-477): up to max_rounds (10^4) rounds of
 *   labels = predict() (global min-max scaling, argmax of scores * weights, ssl.py:257-264; argmin when similarity == 0),
 *   grad = class sizes - priors, weights += -+0.1 * grad, weights /= weights[0], until max|grad| <= tol (1e-3) - in ONE
 *   launch, the reference's fp64 operations in the reference's order.  prob: n x ld fp64 (k used columns); weights (k,
 *   in/out), labels (n, int64), err (1 double), rounds (1 int) are device arrays.  k <= 64.
 * glb_onehot_f64:        dst (n x ld fp64) = one-hot of labels with k columns (utils.labels_to_onehot, utils.py:536-572).
 * glb_max_abs_diff_f64:  *h_out = max |x - y| over n x c entries (y may be NULL) - the stopping tests of the fixed-point
 *   loops (graph.page_rank, graphlearning/graph.py:1408).  Synchronises the stream (host result).
 * ------------------------------------------------------------------------------------------- */
GLB_API int glb_volume_projection(const double *d_prob, int64_t n, int k, int ld, const double *d_priors, int similarity,
                                  int max_rounds, double tol, double *d_weights, int64_t *d_labels, double *d_err, int *d_rounds,
                                  void *stream);
GLB_API int glb_onehot_f64(const int64_t *d_labels, int64_t n, int k, int ld, double *d_dst, void *stream);
GLB_API int glb_max_abs_diff_f64(const double *d_x, const double *d_y, int64_t n, int c, int ldx, int ldy, double *h_out, void *stream);
/* glb_centered_step_f64: one step of the fixed point of ssl.centered_kernel._fit (graphlearning/ssl.py:1409-1413) on
 *   device label matrices: w = inv_alpha (y - 1 mean_y^T) - u, w[labelled] = 0, *h_err = max |w|, u += w.  d_mean_y: c
 *   doubles, d_labelled: n bytes (non-zero = labelled node).  Synchronises the stream (host result).
 * glb_min_nonneg_f64:    *h_out = min over the n x c entries of a non-negative matrix (the `np.min(F) == 0` test of the grow
 *   loop of clustering.incres, graphlearning/clustering.py:357).  Synchronises the stream.
 * glb_argmax_rows_f64:   labels[i] = first column of the row maximum (np.argmax(F, axis=1), clustering.py:361). */
GLB_API int glb_centered_step_f64(const double *d_y, int ldy, const double *d_mean_y, double inv_alpha, double *d_u, int ldu,
                                  const unsigned char *d_labelled, int64_t n, int c, double *h_err, void *stream);
GLB_API int glb_min_nonneg_f64(const double *d_x, int64_t n, int c, int ld, double *h_out, void *stream);
GLB_API int glb_argmax_rows_f64(const double *d_x, int64_t n, int c, int ld, int64_t *d_labels, void *stream);

/* ---------------------------------------------------------------------------------------------
 * kNN result -> CSR weight matrix.  Replaces the sparse assembly of weightmatrix.knn
 * (graphlearning/weightmatrix.py:166-186): coo_matrix((weights, (self_ind, knn_ind))).tocsr(), the symmetrisation
 * (:176-183), setdiag(0) and eliminate_zeros() (:185-186).  symmetrize: 0 = none, 1 = (W + W^T) / 2 (gaussian and
 * user kernels), 2 = utils.sparse_max(W, W^T) ('distance', 'uniform', 'singular'; graphlearning/utils.py:263-286),
 * 3 = W + W^T.multiply(W^T > W) - W.multiply(W^T > W) ('symgaussian').
 * ind: n x k int64 neighbour indices (row i = neighbours of i, self included), w: n x k float64 kernel weights
 * (computed by the caller exactly as the reference does, :139-164).  Output: canonical CSR (int32 rowptr[n+1],
 * int32 col, float64 val, columns ascending), bit-identical to scipy's.  cap = capacity of col / val in entries,
 * at least n*k (2*n*k when symmetrize != 0).  HOST pointers, synchronous.
 * ------------------------------------------------------------------------------------------- */
GLB_API int glb_knn_weights_csr_host(const int64_t *h_ind, const double *h_w, int64_t n, int k, int symmetrize,
                                     int32_t *h_rowptr, int32_t *h_col, double *h_val, int64_t cap, int64_t *nnz,
                                     int *launches);

#ifdef __cplusplus
}
#endif
#endif /* GLB200_H */
