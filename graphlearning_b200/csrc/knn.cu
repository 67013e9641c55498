// knn.cu - exact k-nearest-neighbour search (Euclidean) on sm_100a.
//
// Replaces weightmatrix.knnsearch, exact branches (reference graphlearning/weightmatrix.py:349-361: cKDTree
// `query` and the brute-force `norm(Y - Y[i])` + `argsort` loop) and stands in for its approximate `annoy` default
// (:368-409).  Output contract of the reference: (n, k) neighbour indices INCLUDING self, ascending by fp64
// Euclidean distance, and the (n, k) float64 distances.
//
// Exactness.  The reference ranks in fp64.  Ranking n^2 fp32 Gram-form distances mis-orders near ties (SURVEY.md
// 7.3(3)), so the search is two-stage with a checked error margin:
//   1. knn_prepare      centre the features (distances are translation invariant, smaller norms = smaller absolute
//                       error), round to fp32, squared norms, R = max norm.
//   2. knn_dist_kernel  tiled fp32 distance block  D[q][j] = |q|^2 + |x_j|^2 - 2 q.x_j  for a block of QB queries
//                       against all n points (128 x 128 x 16 register-tiled, operands staged through shared memory).
//   3. knn_select       one warp per query: streams its row of D once and keeps the C = 32 * NPL smallest
//                       (key = distance bits << 32 | index) in a register-resident sorted list spread over the
//                       lanes; insertions are shuffles, a ballot against the running threshold filters the stream.
//   4. knn_rerank       one warp per query: recomputes the C candidates as sum((x_i - x_j)^2) in fp64 from the
//                       ORIGINAL fp64 features, sorts them by (distance, index) and writes the first k.
//                       The row is certified when  approx[C-1] - approx[k-1] > 2 E_i  with the rigorous bound
//                       E_i = 2 (d + 8) 2^-24 (|x_i| + R)^2  on the error of any approximate distance of row i:
//                       then every point outside the candidate list is provably farther than k points inside it.
//   5. knn_exact_rows   rows that fail the certificate (near-degenerate data) are redone by brute force in fp64.
// So the returned indices are those of an exact fp64 ranking (ties between exactly equal distances are broken by
// index; the reference's tie order is unspecified, weightmatrix.py:352,359-361).
//
//   2b. knn_dist_tc_kernel  the same block on the tensor cores when d >= 64: bf16 hi/lo split, three tcgen05.mma per
//                       k-step into one fp32 TMEM accumulator (see the kernel).
// Roofline: stage 2 is 2 n^2 d flops (fp32 FMA for d < 64, 3 x bf16 tensor-core MMAs for d >= 64); stages 2+3 move
// 2 * 4 n^2 bytes through HBM/L2 (the distance block), stage 4 gathers C rows of fp64 features per query.
#include <cuda.h>
#include <cuda_bf16.h>
#include <float.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "common.cuh"

namespace glb {

typedef unsigned long long u64;

constexpr int BM = 128, BN = 128, BK = 16;          // distance tile
constexpr int kDistThreads = 256;

// ---------------------------------------------------------------------------------------------------------
// 1. prepare
// ---------------------------------------------------------------------------------------------------------
// column sums (fp64) -> mean; deterministic: one block per column chunk, fixed order
__global__ void __launch_bounds__(256)
knn_colsum_kernel(const double *__restrict__ X, long long n, int d, double *__restrict__ mean)
{
    __shared__ double sh[256];
    const int col = blockIdx.x;
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) s += X[i * d + col];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) mean[col] = sh[0] / (double)n;
}

// Xc (n x dpad fp32, zero padded) = fp32(X - mean); norm2[i] = fp32(|Xc_i|^2 in fp64); rmax2 = max_i |Xc_i|^2
__global__ void __launch_bounds__(256)
knn_center_kernel(const double *__restrict__ X, const double *__restrict__ mean, long long n, int d, int dpad,
                  float *__restrict__ Xc, float *__restrict__ norm2, unsigned long long *rmax2_bits)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    double worst = 0.0;
    for (long long i = warp; i < n; i += nwarps) {
        double s = 0.0;
        for (int t = lane; t < dpad; t += 32) {
            float v = 0.f;
            if (t < d) v = (float)(X[i * d + t] - mean[t]);
            Xc[i * dpad + t] = v;
            s += (double)v * (double)v;
        }
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) norm2[i] = (float)s;
        worst = fmax(worst, s);
    }
    if (lane == 0) atomicMax(rmax2_bits, (unsigned long long)__double_as_longlong(worst));   // s >= 0: bit order = value order
}

// ---------------------------------------------------------------------------------------------------------
// 2. distance block: D[q - q0][j] = max(0, nq + nj - 2 <Xc_q, Xc_j>), q in [q0, q0 + nq), j in [0, n); j >= n -> +inf
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDistThreads, 2)
knn_dist_kernel(const float *__restrict__ Xc, const float *__restrict__ norm2, int n, int dpad, int q0, int nq,
                float *__restrict__ D, long long ldD)
{
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;                  // 16 x 16 threads, 8 x 8 outputs each
    const int row0 = q0 + blockIdx.y * BM;                   // first query of the tile (global index)
    const int col0 = blockIdx.x * BN;                        // first database point of the tile
    // global -> register staging: thread loads 2 float4 of A and 2 of B per K step (128 rows x 16 floats = 512 float4)
    const int lr = tid / 4, lc = (tid % 4) * 4;              // row within tile (0..63, +64), float offset within BK
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    auto load_tile = [&](int k0, float4 ra[2], float4 rb[2]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + h * 64;
            const int qa = row0 + r, pb = col0 + r;
            ra[h] = (qa < q0 + nq && qa < n) ? __ldg(reinterpret_cast<const float4 *>(Xc + (size_t)qa * dpad + k0 + lc)) : make_float4(0.f, 0.f, 0.f, 0.f);
            rb[h] = (pb < n) ? __ldg(reinterpret_cast<const float4 *>(Xc + (size_t)pb * dpad + k0 + lc)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_tile = [&](int buf, const float4 ra[2], const float4 rb[2]) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = lr + h * 64;
            As[buf][lc + 0][r] = ra[h].x; As[buf][lc + 1][r] = ra[h].y; As[buf][lc + 2][r] = ra[h].z; As[buf][lc + 3][r] = ra[h].w;
            Bs[buf][lc + 0][r] = rb[h].x; Bs[buf][lc + 1][r] = rb[h].y; Bs[buf][lc + 2][r] = rb[h].z; Bs[buf][lc + 3][r] = rb[h].w;
        }
    };
    float4 ra[2], rb[2];
    load_tile(0, ra, rb);
    store_tile(0, ra, rb);
    __syncthreads();
    const int nk = dpad / BK;
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tile((kt + 1) * BK, ra, rb);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tile(buf ^ 1, ra, rb);
            __syncthreads();
        }
    }
    // epilogue
    float nb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int p = col0 + tx * 8 + j;
        nb[j] = p < n ? __ldg(norm2 + p) : INFINITY;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int q = row0 + ty * 8 + i;
        if (q >= q0 + nq || q >= n) continue;
        const float na = __ldg(norm2 + q);
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(0.f, fmaf(-2.f, acc[i][j], na + nb[j]));     // +inf stays +inf
        float *dst = D + (size_t)(q - q0) * ldD + col0 + tx * 8;
        *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4 *>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
}


// ---------------------------------------------------------------------------------------------------------
// 2b. distance block on the 5th-generation tensor cores (d >= 64): tcgen05.mma, accumulator in TMEM
// ---------------------------------------------------------------------------------------------------------
// The certificate of stage 4 needs the cross term to ~2^-16 relative, bf16 carries 2^-9.  The centred fp32 features are
// split x = hi + lo (hi = bf16(x), lo = bf16(x - hi), |x - hi - lo| <= 2^-18 |x|) and the cross term is accumulated as
// hi.hi + hi.lo + lo.hi - three bf16 MMAs per k-step into the same fp32 TMEM accumulator (the dropped lo.lo term and the
// two split residuals are <= 3 * 2^-18 |x||y|).
//
// One CTA (128 threads) = one 128 x 128 tile of D.  Operands are K-major bf16 in shared memory in the canonical
// no-swizzle UMMA layout: 8 x 16-byte core matrices, K-adjacent core matrices 128 B apart (LBO), 8-row groups
// (BK/8) * 128 B apart (SBO); filled with 16-byte cp.async (out-of-range rows zero-filled), made visible to the async
// proxy with fence.proxy.async.  Two stages of BK = 64: thread 0 issues the 3 * 4 tcgen05.mma (M = N = 128, K = 16) of a
// stage and commits them to the stage's mbarrier, so the next stage's loads overlap the tensor-core work.  Epilogue:
// warp w reads TMEM lanes 32w..32w+31 (one query row per thread) with tcgen05.ld 32x32b.x32, applies
// max(0, |q|^2 + |x_j|^2 - 2 acc) and writes 128-byte runs of its row of D.
constexpr int TBM = 128, TBN = 128, TBK = 32;
constexpr int kTcThreads = 128;
constexpr int kTcStageBytes = 4 * TBM * TBK * 2;                 // A hi, A lo, B hi, B lo
constexpr int kTcSmemBytes = 2 * kTcStageBytes + 64;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr)
{
    // start address [0,14) | leading byte offset [16,30) = 128 B | stride byte offset [32,46) = (TBK/8)*128 B | version [46,48) = 1
    // | layout type [61,64) = 0 (no swizzle); all offsets in 16-byte units
    return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | ((unsigned long long)(128u >> 4) << 16) |
           ((unsigned long long)((TBK / 8 * 128u) >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity)
{
    // bounded: a tensor-core batch completes within microseconds; ~1 s of polling means a broken descriptor -> trap
    for (unsigned spins = 0; spins < (1u << 26); ++spins) {
        unsigned done;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

__global__ void __launch_bounds__(kTcThreads, 3)
knn_dist_tc_kernel(const __nv_bfloat16 *__restrict__ Xh, const __nv_bfloat16 *__restrict__ Xl, const float *__restrict__ norm2, int n,
                   int dpad, int q0, int nq, float *__restrict__ D, long long ldD)
{
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(tc_smem + 2 * kTcStageBytes);       // [0,1] stage free, [2] done
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(tc_smem + 2 * kTcStageBytes + 32);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = q0 + blockIdx.y * TBM;                  // first query of the tile
    const int col0 = blockIdx.x * TBN;                       // first database point of the tile
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = *tmem_slot;

    // one stage = 4 operand blocks of 128 rows x TBK bf16; 16-byte chunk (r, c) of a block lives at
    // (r/8) * SBO + c * 128 + (r%8) * 16 with SBO = (TBK/8) * 128
    constexpr int CPR = TBK / 8;                                 // 16-byte chunks per row
    auto fill = [&](int stage, int k0) {
        unsigned char *base = tc_smem + stage * kTcStageBytes;
#pragma unroll
        for (int it = 0; it < 4 * TBM * CPR / kTcThreads; ++it) {
            const int ch = it * kTcThreads + tid;
            const int blk = ch / (TBM * CPR), r = (ch / CPR) % TBM, c = ch % CPR;       // consecutive threads: the chunks of a row
            const int grow = (blk < 2 ? row0 : col0) + r;
            const bool ok = blk < 2 ? (grow < q0 + nq && grow < n) : (grow < n);
            const __nv_bfloat16 *src = ((blk & 1) ? Xl : Xh) + (size_t)(ok ? grow : 0) * dpad + k0 + c * 8;
            const unsigned dst = smem_u32(base + blk * (TBM * TBK * 2) + (r >> 3) * (CPR * 128) + c * 128 + (r & 7) * 16);
            const int bytes = ok ? 16 : 0;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = BF16 (bits 7-9, 10-12 = 1), K-major both, N >> 3 at 17, M >> 4 at 24
    const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(TBN >> 3) << 17) | ((unsigned)(TBM >> 4) << 24);
    const int nkb = dpad / TBK;
    fill(0, 0);
    for (int kb = 0; kb < nkb; ++kb) {
        const int stage = kb & 1;
        if (kb + 1 < nkb) {
            if (kb + 1 >= 2) mbar_wait(smem_u32(mbar + ((kb + 1) & 1)), (unsigned)(((kb - 1) >> 1) & 1));   // MMAs of block kb-1 done with that stage
            fill((kb + 1) & 1, (kb + 1) * TBK);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned sb = smem_u32(tc_smem + stage * kTcStageBytes);
            const unsigned a_hi = sb, a_lo = sb + TBM * TBK * 2, b_hi = sb + 2 * TBM * TBK * 2, b_lo = sb + 3 * TBM * TBK * 2;
#pragma unroll
            for (int ks = 0; ks < TBK / 16; ++ks) {            // K = 16 per instruction = two core matrices = 256 B
                const unsigned o = ks * 256;
                umma_bf16(tmem_d, umma_desc(a_hi + o), umma_desc(b_hi + o), idesc, (kb | ks) ? 1u : 0u);
                umma_bf16(tmem_d, umma_desc(a_hi + o), umma_desc(b_lo + o), idesc, 1u);
                umma_bf16(tmem_d, umma_desc(a_lo + o), umma_desc(b_hi + o), idesc, 1u);
            }
            // arrives on the barrier when every MMA issued so far has completed (implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar + stage)) : "memory");
            if (kb + 1 == nkb)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar + 2)) : "memory");
        }
    }
    mbar_wait(smem_u32(mbar + 2), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: thread = query row (TMEM lane), 4 x 32 columns.  The 32 x 32 block a warp holds is transposed through
    // shared memory (the operand stages are free now) so that every store instruction writes one 128-byte run of a row.
    const int q = row0 + warp * 32 + lane;
    const float na = (q < q0 + nq && q < n) ? __ldg(norm2 + q) : 0.f;
    float *stg = reinterpret_cast<float *>(tc_smem) + warp * (32 * 33);
#pragma unroll 1
    for (int cb = 0; cb < TBN / 32; ++cb) {
        unsigned v[32];
        const unsigned taddr = tmem_d + ((unsigned)(warp * 32) << 16) + (unsigned)(cb * 32);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                       "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                       "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int p = col0 + cb * 32 + lane;                 // this lane's column after the transpose
        const float nb = p < n ? __ldg(norm2 + p) : INFINITY;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = fmaf(-2.f, __uint_as_float(v[j]), na);      // row = lane, column j
        __syncwarp();
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            const int qr = row0 + warp * 32 + r;
            if (qr < q0 + nq && qr < n)
                D[(size_t)(qr - q0) * ldD + p] = fmaxf(0.f, stg[r * 33 + lane] + nb);                   // +inf stays +inf
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_d) : "memory");
}

// hi / lo bf16 split of the centred features (n x dpad fp32 -> 2 x n x dpad bf16)
__global__ void __launch_bounds__(256)
knn_split_kernel(const float *__restrict__ Xc, long long total, __nv_bfloat16 *__restrict__ Xh, __nv_bfloat16 *__restrict__ Xl)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float x = Xc[i];
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        Xh[i] = h;
        Xl[i] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
}

// ---------------------------------------------------------------------------------------------------------
// 3. selection: one warp per query row, C = 32 * NPL candidates
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 make_key(float d2, int idx) { return ((u64)__float_as_uint(d2) << 32) | (u64)(unsigned)idx; }

// sorted list of C keys spread over the warp: lane l holds positions [l * NPL, (l + 1) * NPL)
template <int NPL>
struct WarpList {
    u64 e[NPL];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < NPL; ++i) e[i] = ~0ull;
    }
    __device__ __forceinline__ u64 last(int lane) const { (void)lane; return __shfl_sync(0xffffffffu, e[NPL - 1], 31); }
    // insert key (warp-uniform value) keeping the list sorted; the largest element falls off
    __device__ __forceinline__ void insert(u64 key, int lane)
    {
        int below = 0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) below += (e[i] < key) ? 1 : 0;
        const int p = __reduce_add_sync(0xffffffffu, below);            // insertion position
        u64 carry = __shfl_up_sync(0xffffffffu, e[NPL - 1], 1);          // last element of the previous lane
#pragma unroll
        for (int i = NPL - 1; i >= 0; --i) {
            const int g = lane * NPL + i;
            const u64 prev = (i == 0) ? carry : e[i - 1];
            if (g > p) e[i] = prev;
            else if (g == p) e[i] = key;
        }
    }
};

template <int NPL>
__global__ void __launch_bounds__(256)
knn_select_kernel(const float *__restrict__ D, long long ldD, int n_pad, int nq, u64 *__restrict__ cand)
{
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    const float *row = D + (size_t)q * ldD;
    WarpList<NPL> list;
    list.init();
    float tau = INFINITY;                      // current C-th smallest distance; a candidate must be <= tau
    u64 tau_key = ~0ull;
    for (int base = 0; base < n_pad; base += 128) {
        const float4 v4 = __ldg(reinterpret_cast<const float4 *>(row + base) + lane);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int idx = base + lane * 4 + c;
            unsigned m = __ballot_sync(0xffffffffu, v[c] <= tau && v[c] < INFINITY);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float dv = __shfl_sync(0xffffffffu, v[c], src);
                const int di = __shfl_sync(0xffffffffu, idx, src);
                const u64 key = make_key(dv, di);
                if (key < tau_key) {
                    list.insert(key, lane);
                    tau_key = list.last(lane);
                    tau = __uint_as_float((unsigned)(tau_key >> 32));
                    if (tau_key == ~0ull) tau = INFINITY;
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) cand[(size_t)q * (32 * NPL) + lane * NPL + i] = list.e[i];
}

// ---------------------------------------------------------------------------------------------------------
// 4. exact re-rank of the candidates + certificate
// ---------------------------------------------------------------------------------------------------------
struct ExactKey { double d2; int idx; };
__device__ __forceinline__ bool key_less(double da, int ia, double db, int ib) { return da < db || (da == db && ia < ib); }

template <int NPL>
__global__ void __launch_bounds__(256)
knn_rerank_kernel(const double *__restrict__ X, int n, int d, const float *__restrict__ norm2,
                  const unsigned long long *__restrict__ rmax2_bits, const u64 *__restrict__ cand, int q0, int nq, int k,
                  long long *__restrict__ out_ind, double *__restrict__ out_dist, int *__restrict__ fail_rows,
                  int *__restrict__ fail_count, double err_rel, const float *__restrict__ outside_bound)
{
    constexpr int C = 32 * NPL;
    const int lane = threadIdx.x & 31;
    const int wq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wq >= nq) return;
    const int q = q0 + wq;
    const double *xq = X + (size_t)q * d;
    u64 keys[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) keys[i] = cand[(size_t)wq * C + lane * NPL + i];
    // certificate: approximate distances of list positions k-1 and C-1 (list is sorted ascending by approximate key;
    // position p is slot p % NPL of lane p / NPL)
    const int pk = k - 1;
    u64 mine_k = keys[0];
#pragma unroll
    for (int i = 1; i < NPL; ++i) if (i == pk % NPL) mine_k = keys[i];
    const u64 key_k = __shfl_sync(0xffffffffu, mine_k, pk / NPL);
    const u64 key_c = __shfl_sync(0xffffffffu, keys[NPL - 1], 31);
    const float ak = __uint_as_float((unsigned)(key_k >> 32)), ac = __uint_as_float((unsigned)(key_c >> 32));
    const double R = sqrt(__longlong_as_double((long long)*rmax2_bits));
    const double ni = sqrt((double)norm2[q]);
    const double E = 2.0 * err_rel * (ni + R) * (ni + R);      // err_rel: bound on |approx cross term - exact| / (|x||y|)
    // fewer than C points in total (n <= C): the list holds everything, nothing can be missing.
    // outside_bound (fused search): a lower bound on the approximate distance of every point that is NOT in the list.
    const bool certified = outside_bound ? (key_k != ~0ull && (double)outside_bound[wq] - (double)ak > 2.0 * E)
                                         : ((key_c == ~0ull) || ((double)ac - (double)ak > 2.0 * E));
    // Only candidates with approx <= approx[k-1] + 2E can be among the k nearest: the k first of the list have a true
    // distance <= approx[k-1] + E, every other point beyond that threshold a true distance > approx[k-1] + E.  The list is
    // sorted, so they are its first m positions; position p goes to lane p % 32 (all lanes busy, one candidate each per
    // round) and is evaluated as sum((x_i - x_j)^2) in fp64 from the original features, coordinates in order.
    int m = 0;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const bool in = keys[i] != ~0ull && (double)__uint_as_float((unsigned)(keys[i] >> 32)) <= (double)ak + 2.0 * E;
        m += __popc(__ballot_sync(0xffffffffu, in));
    }
    if (key_k == ~0ull) m = C;                                 // fewer than k valid candidates: keep whatever there is
    double d2[NPL];
    int idx[NPL];
#pragma unroll
    for (int r = 0; r < NPL; ++r) {
        d2[r] = INFINITY; idx[r] = 0x7fffffff;
        if (r * 32 >= m) continue;                             // warp-uniform
        const int pos = r * 32 + lane;
        u64 key = ~0ull;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const u64 got = __shfl_sync(0xffffffffu, keys[i], pos / NPL);
            if (i == pos % NPL) key = got;
        }
        if (pos < m && key != ~0ull) {
            idx[r] = (int)(unsigned)(key & 0xffffffffull);
            const double *xc = X + (size_t)idx[r] * d;
            double s = 0.0;
            if ((d & 1) == 0) {                                // 16-byte loads, same order of the sum
                for (int t = 0; t < d; t += 2) {
                    const double2 a = *reinterpret_cast<const double2 *>(xq + t), b = __ldg(reinterpret_cast<const double2 *>(xc + t));
                    const double d0 = a.x - b.x, d1 = a.y - b.y;
                    s = fma(d0, d0, s); s = fma(d1, d1, s);
                }
            } else {
                for (int t = 0; t < d; ++t) { const double df = xq[t] - xc[t]; s = fma(df, df, s); }
            }
            d2[r] = s;
        }
    }
    // rank of every candidate among the C by (exact d2, index): O(C) shuffles per element, C <= 128
    int rank[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) rank[i] = 0;
    for (int src = 0; src < 32; ++src) {
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            if (j * 32 >= m) continue;                         // warp-uniform: nothing was evaluated in that round
            const double od = __shfl_sync(0xffffffffu, d2[j], src);
            const int oi = __shfl_sync(0xffffffffu, idx[j], src);
#pragma unroll
            for (int i = 0; i < NPL; ++i) rank[i] += key_less(od, oi, d2[i], idx[i]) ? 1 : 0;
        }
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        if (rank[i] < k) {
            out_ind[(size_t)q * k + rank[i]] = idx[i];
            out_dist[(size_t)q * k + rank[i]] = sqrt(d2[i]);
        }
    }
    if (!certified && lane == 0) fail_rows[atomicAdd(fail_count, 1)] = q;
}

// ---------------------------------------------------------------------------------------------------------
// 5. brute force in fp64 for the rows without certificate: one CTA per row
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
knn_exact_rows_kernel(const double *__restrict__ X, int n, int d, const int *__restrict__ rows, const int *__restrict__ nrows_ptr,
                      int k, double *__restrict__ scratch, long long *__restrict__ out_ind, double *__restrict__ out_dist)
{
    __shared__ double sh_d[256];
    __shared__ int sh_i[256];
    const int nrows = *nrows_ptr;
    for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
        const int q = rows[r];
        double *dist = scratch + (size_t)blockIdx.x * n;
        const double *xq = X + (size_t)q * d;
        for (int j = threadIdx.x; j < n; j += 256) {
            const double *xc = X + (size_t)j * d;
            double s = 0.0;
            for (int t = 0; t < d; ++t) { const double df = xq[t] - xc[t]; s = fma(df, df, s); }
            dist[j] = s;
        }
        __syncthreads();
        for (int sel = 0; sel < k; ++sel) {                 // k rounds of block-wide argmin by (distance, index)
            double bd = INFINITY; int bi = 0x7fffffff;
            for (int j = threadIdx.x; j < n; j += 256) {
                const double v = dist[j];
                if (key_less(v, j, bd, bi)) { bd = v; bi = j; }
            }
            sh_d[threadIdx.x] = bd; sh_i[threadIdx.x] = bi;
            __syncthreads();
            for (int off = 128; off > 0; off >>= 1) {
                if ((int)threadIdx.x < off && key_less(sh_d[threadIdx.x + off], sh_i[threadIdx.x + off], sh_d[threadIdx.x], sh_i[threadIdx.x])) {
                    sh_d[threadIdx.x] = sh_d[threadIdx.x + off]; sh_i[threadIdx.x] = sh_i[threadIdx.x + off];
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                out_ind[(size_t)q * k + sel] = sh_i[0];
                out_dist[(size_t)q * k + sel] = sqrt(sh_d[0]);
                if (sh_i[0] < n) dist[sh_i[0]] = INFINITY;
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// 2c + 3c. fused distance block + candidate filter on the tensor cores: D never leaves the SM
// ---------------------------------------------------------------------------------------------------------
// One CTA = 128 queries against ALL database points, 128 at a time.  Warp roles (192 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor (SASS UTMALDG) of 128 x 64 bf16 operand tiles (hi and lo halves of the split
//            features) into 128-byte-swizzled shared memory, a ring of stages with full/empty mbarriers.  The query tiles
//            are loaded once and stay resident when they fit (d <= 256), the database tiles stream.
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M = N = 128, K = 16, SASS UTCHMMA) hi.hi + hi.lo + lo.hi per k-step
//            into one of TWO 128-column fp32 accumulators in TMEM, commits the stage (frees it for the producer) and, after
//            the last k-block, the accumulator (hands it to the epilogue) - the tensor pipe starts the next database tile
//            while the epilogue drains the previous one.
//   warps 2-9  epilogue: thread = (query row = TMEM lane, half of the tile's 128 columns); two warps per scheduler so that the
//            dependent ALU chains of one hide behind the other.  tcgen05.ld 32 columns at a time, d = |q|^2 + |x_j|^2 - 2 acc
//            for all 32 (branch-free, a hit mask), then only the hits:
//            MODE_SAMPLE  keep the R smallest distances in registers -> written per (row, half); the R-th smallest of the
//                         union is tau_row
//            MODE_EMIT    append (d, j) to the (row, half) candidate buffer in HBM when d <= tau_row (0.2 % of the points)
// Two passes: MODE_SAMPLE against a fixed pseudo-random sample of s database points gives every row a threshold tau_row
// whose expected rank in the full set is n R / s (137 for n = 70 000, R = 16, s = 8192); MODE_EMIT then writes only the
// points below it.  A row's candidate list = its C smallest emitted points (knn_pick_kernel); every point outside the list
// has an approximate distance >= min(tau_row, C-th smallest emitted), which is what the certificate of stage 4 needs.
// Rows whose buffer overflows (> cap candidates) or that miss the certificate go to the exact fp64 fallback.
constexpr int FBM = 128, FBN = 128, FBK = 64;
constexpr int kFusedThreads = 320;                               // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int kEpiThreads = 256;
constexpr int kFTile = FBM * FBK * 2;                            // 16 KB: one 128 x 64 bf16 operand tile, 128-byte swizzle
constexpr int kSampleR = 16;                                     // rank of the threshold inside the sample
constexpr int MODE_SAMPLE = 0, MODE_EMIT = 1;

__device__ __forceinline__ void mbar_init_n(unsigned mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned mbar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned mbar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned mbar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}
__device__ __forceinline__ unsigned long long umma_desc_sw128(unsigned saddr)
{
    // K-major, 128-byte swizzle: start address [0,14) | LBO [16,30) = 1 (unused) | SBO [32,46) = 1024 B (8 rows x 128 B)
    // | version [46,48) = 1 | layout type [61,64) = 2 (SWIZZLE_128B); offsets in 16-byte units
    return (unsigned long long)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((unsigned long long)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_commit(unsigned mbar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

struct FusedArgs {
    const float *norm_q;             // |q|^2 of the queries (index = global query row)
    const float *norm_d;             // |x_j|^2 of the database (or sample) points
    int q0, nq;                      // query rows [q0, q0 + nq)
    int nd;                          // database points
    int kb_per_tile;                 // dpad / 64
    int stages;                      // ring depth
    int a_resident;                  // query tiles loaded once
    // MODE_SAMPLE: the kSampleR smallest sample distances of every (row, column half), ascending
    float *best_out;                 // [((row - q0) * 2 + half) * kSampleR]
    // MODE_EMIT
    const float *best_in;            // the same array: tau_row = kSampleR-th smallest of the two lists
    float *tau_out;                  // [row - q0] (written by half 0; knn_pick_kernel needs it)
    u64 *cand_buf;                   // [((row - q0) * 2 + half) * cap]
    int *cand_count;                 // [(row - q0) * 2 + half]; cap + 1 = overflow
    int cap;                         // per (row, half)
};

template <int MODE>
__global__ void __launch_bounds__(kFusedThreads, 1)
knn_fused_kernel(const __grid_constant__ CUtensorMap tm_qh, const __grid_constant__ CUtensorMap tm_ql,
                 const __grid_constant__ CUtensorMap tm_dh, const __grid_constant__ CUtensorMap tm_dl, const FusedArgs a)
{
    extern __shared__ unsigned char fused_raw[];
    const unsigned raw = smem_u32(fused_raw);
    const unsigned base = (raw + 1023u) & ~1023u;                // swizzle atoms want 1024-byte alignment
    unsigned char *gen = fused_raw + (base - raw);
    const int S = a.stages, KB = a.kb_per_tile;
    const int tiles_per_stage = a.a_resident ? 2 : 4;            // (B hi, B lo) or (A hi, A lo, B hi, B lo)
    const unsigned a_res = base;                                 // resident query tiles: KB x (hi, lo)
    const unsigned ring = base + (a.a_resident ? (unsigned)KB * 2u * kFTile : 0u);
    const unsigned ctl = ring + (unsigned)S * tiles_per_stage * kFTile;
    // control block: full[S], empty[S], tmem_full[2], tmem_empty[2], a_full, tmem slot; then the norm tiles
    const unsigned bar_full = ctl, bar_empty = ctl + 8u * S, bar_tfull = ctl + 16u * S, bar_tempty = bar_tfull + 16u, bar_afull = bar_tempty + 16u;
    const unsigned tmem_slot_a = bar_afull + 8u;
    const unsigned nb_a = (tmem_slot_a + 8u + 15u) & ~15u;      // |x_j|^2 of the current database tiles: [2][128] floats
    volatile unsigned *tmem_slot = reinterpret_cast<volatile unsigned *>(gen + (tmem_slot_a - base));
    float *s_nb = reinterpret_cast<float *>(gen + (nb_a - base));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = a.q0 + blockIdx.x * FBM;
    const int ntiles = (a.nd + FBN - 1) / FBN;

    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init_n(bar_full + 8u * i, 1); mbar_init_n(bar_empty + 8u * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init_n(bar_tfull + 8u * i, 1); mbar_init_n(bar_tempty + 8u * i, kEpiThreads); }
        mbar_init_n(bar_afull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tmem_slot_a) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_d = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            if (a.a_resident) {
                mbar_expect(bar_afull, (unsigned)KB * 2u * kFTile);
                for (int kb = 0; kb < KB; ++kb) {
                    tma_load_2d(a_res + (unsigned)(2 * kb) * kFTile, &tm_qh, kb * FBK, row0, bar_afull);
                    tma_load_2d(a_res + (unsigned)(2 * kb + 1) * kFTile, &tm_ql, kb * FBK, row0, bar_afull);
                }
            }
            int it = 0;
            for (int tile = 0; tile < ntiles; ++tile)
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int st = it % S;
                    mbar_wait(bar_empty + 8u * st, (unsigned)(((it / S) & 1) ^ 1));      // a fresh barrier passes parity 1
                    mbar_expect(bar_full + 8u * st, (unsigned)tiles_per_stage * kFTile);
                    const unsigned sb = ring + (unsigned)st * tiles_per_stage * kFTile;
                    tma_load_2d(sb, &tm_dh, kb * FBK, tile * FBN, bar_full + 8u * st);
                    tma_load_2d(sb + kFTile, &tm_dl, kb * FBK, tile * FBN, bar_full + 8u * st);
                    if (!a.a_resident) {
                        tma_load_2d(sb + 2 * kFTile, &tm_qh, kb * FBK, row0, bar_full + 8u * st);
                        tma_load_2d(sb + 3 * kFTile, &tm_ql, kb * FBK, row0, bar_full + 8u * st);
                    }
                }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(FBN >> 3) << 17) | ((unsigned)(FBM >> 4) << 24);
            if (a.a_resident) mbar_wait(bar_afull, 0u);
            int it = 0;
            for (int tile = 0; tile < ntiles; ++tile) {
                const int acc = tile & 1;
                mbar_wait(bar_tempty + 8u * acc, (unsigned)(((tile >> 1) & 1) ^ 1));       // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned d_tmem = tmem_d + (unsigned)(acc * FBN);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int st = it % S;
                    mbar_wait(bar_full + 8u * st, (unsigned)((it / S) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned sb = ring + (unsigned)st * tiles_per_stage * kFTile;
                    const unsigned b_hi = sb, b_lo = sb + kFTile;
                    const unsigned a_hi = a.a_resident ? a_res + (unsigned)(2 * kb) * kFTile : sb + 2 * kFTile;
                    const unsigned a_lo = a_hi + kFTile;
#pragma unroll
                    for (int ks = 0; ks < FBK / 16; ++ks) {      // K = 16 per instruction = 32 bytes along the swizzled row
                        const unsigned o = ks * 32;
                        umma_bf16(d_tmem, umma_desc_sw128(a_hi + o), umma_desc_sw128(b_hi + o), idesc, (kb | ks) ? 1u : 0u);
                        umma_bf16(d_tmem, umma_desc_sw128(a_hi + o), umma_desc_sw128(b_lo + o), idesc, 1u);
                        umma_bf16(d_tmem, umma_desc_sw128(a_lo + o), umma_desc_sw128(b_hi + o), idesc, 1u);
                    }
                    umma_commit(bar_empty + 8u * st);            // the stage is free once these MMAs have read it
                }
                umma_commit(bar_tfull + 8u * acc);               // accumulator complete -> epilogue
            }
        }
    } else {
        // ===== epilogue: thread = (query row, column half) =====
        const int wq = warp & 3;                                 // a warp may only touch TMEM lanes 32 (warp % 4) ...
        const int half = (warp - 2) >> 2;                        // warps 2-5: columns 0-63 of a tile, warps 6-9: columns 64-127
        const int r = wq * 32 + lane;
        const int et = (warp - 2) * 32 + lane;                   // 0..255 among the epilogue threads
        const int q = row0 + r;
        const bool live = q < a.q0 + a.nq;
        const float na = live ? __ldg(a.norm_q + q) : 0.f;
        const size_t slot = (size_t)(live ? q - a.q0 : 0) * 2 + half;
        float tau = -1.f;
        int cnt = 0;
        float best[kSampleR];
        if (MODE == MODE_SAMPLE) {
#pragma unroll
            for (int i = 0; i < kSampleR; ++i) best[i] = INFINITY;
            tau = live ? INFINITY : -1.f;
        } else if (live) {
            // tau_row = kSampleR-th smallest of the union of the two halves' sorted lists
            const float *l0 = a.best_in + (size_t)(q - a.q0) * 2 * kSampleR, *l1 = l0 + kSampleR;
            int i0 = 0, i1 = 0;
            float t = INFINITY;
            for (int i = 0; i < kSampleR; ++i) {
                const float x0 = l0[i0], x1 = l1[i1];
                if (x0 <= x1) { t = x0; ++i0; } else { t = x1; ++i1; }
            }
            tau = t;
            if (half == 0) a.tau_out[q - a.q0] = tau;
        }
        u64 *mybuf = MODE == MODE_EMIT ? a.cand_buf + slot * a.cap : nullptr;
        for (int tile = 0; tile < ntiles; ++tile) {
            const int acc = tile & 1;
            const int col0 = tile * FBN;
            if (et < FBN) s_nb[acc * FBN + et] = (col0 + et < a.nd) ? __ldg(a.norm_d + col0 + et) : INFINITY;     // +inf: never a candidate
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(bar_tfull + 8u * acc, (unsigned)((tile >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cb = half * 2; cb < half * 2 + 2; ++cb) {
                unsigned v[32];
                const unsigned taddr = tmem_d + ((unsigned)(wq * 32) << 16) + (unsigned)(acc * FBN + cb * 32);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                               "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                               "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                               "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const float *nbp = s_nb + acc * FBN + cb * 32;
                const float4 *nb4 = reinterpret_cast<const float4 *>(nbp);
                // Filter first, distances only for the hits: d = na + nb - 2 acc <= tau  <=>  acc >= nb / 2 + (na - tau) / 2, i.e.
                // one multiply-add and one compare per element (the rounding of the threshold moves the cut by a few ulps of
                // na + nb, which the error margin of the certificate covers, see knn_run).  0.2 % of the elements are hits, so
                // the hit path is entered per group of 4 columns (a warp has a hit in a given group 1 time out of 4).
                const float c0 = 0.5f * (na - tau);
                unsigned hits = 0u;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 nb = nb4[j4];
                    const float t0 = fmaf(0.5f, nb.x, c0), t1 = fmaf(0.5f, nb.y, c0), t2 = fmaf(0.5f, nb.z, c0), t3 = fmaf(0.5f, nb.w, c0);
                    if (MODE == MODE_SAMPLE) {
                        hits |= (__uint_as_float(v[j4 * 4 + 0]) > t0 ? 1u : 0u) << (j4 * 4 + 0);
                        hits |= (__uint_as_float(v[j4 * 4 + 1]) > t1 ? 1u : 0u) << (j4 * 4 + 1);
                        hits |= (__uint_as_float(v[j4 * 4 + 2]) > t2 ? 1u : 0u) << (j4 * 4 + 2);
                        hits |= (__uint_as_float(v[j4 * 4 + 3]) > t3 ? 1u : 0u) << (j4 * 4 + 3);
                    } else {
                        hits |= (__uint_as_float(v[j4 * 4 + 0]) >= t0 ? 1u : 0u) << (j4 * 4 + 0);
                        hits |= (__uint_as_float(v[j4 * 4 + 1]) >= t1 ? 1u : 0u) << (j4 * 4 + 1);
                        hits |= (__uint_as_float(v[j4 * 4 + 2]) >= t2 ? 1u : 0u) << (j4 * 4 + 2);
                        hits |= (__uint_as_float(v[j4 * 4 + 3]) >= t3 ? 1u : 0u) << (j4 * 4 + 3);
                    }
                }
                if (MODE == MODE_SAMPLE && !(tau < INFINITY)) hits = live ? 0xffffffffu : 0u;     // list not full yet: c0 = -inf
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    if (hits & (0xfu << (j4 * 4))) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int j = j4 * 4 + jj;
                            if (hits & (1u << j)) {
                                const float dd = fmaxf(fmaf(-2.f, __uint_as_float(v[j]), na + nbp[j]), 0.f);
                                if (MODE == MODE_SAMPLE) {
                                    if (dd < best[kSampleR - 1]) {       // the threshold shrinks while the hits of this chunk are taken
                                        best[kSampleR - 1] = dd;         // replace the largest, bubble it into place
#pragma unroll
                                        for (int i = kSampleR - 1; i > 0; --i) {
                                            const float lo = fminf(best[i - 1], best[i]), hi = fmaxf(best[i - 1], best[i]);
                                            best[i - 1] = lo; best[i] = hi;
                                        }
                                        tau = best[kSampleR - 1];
                                    }
                                } else {
                                    if (cnt < a.cap) mybuf[cnt] = make_key(dd, col0 + cb * 32 + j);
                                    ++cnt;
                                }
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(bar_tempty + 8u * acc);
        }
        if (live) {
            if (MODE == MODE_SAMPLE) {
#pragma unroll
                for (int i = 0; i < kSampleR; ++i) a.best_out[slot * kSampleR + i] = best[i];
            } else {
                a.cand_count[slot] = cnt <= a.cap ? cnt : a.cap + 1;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_d) : "memory");
}

// the C smallest of a row's emitted candidates (sorted, as knn_select_kernel leaves them) + the bound on everything else
template <int NPL>
__global__ void __launch_bounds__(256)
knn_pick_kernel(const u64 *__restrict__ cand_buf, const int *__restrict__ cand_count, const float *__restrict__ tau, int cap, int q0, int nq,
                u64 *__restrict__ cand, float *__restrict__ outside_bound, int *__restrict__ fail_rows, int *__restrict__ fail_count)
{
    constexpr int C = 32 * NPL;
    const int lane = threadIdx.x & 31;
    const int wq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wq >= nq) return;
    WarpList<NPL> list;
    list.init();
    u64 tau_key = ~0ull;
    int total = 0;
    bool overflow = false;
    for (int half = 0; half < 2; ++half) {                   // the two column halves of the fused kernel's epilogue
        const int cnt = cand_count[(size_t)wq * 2 + half];
        overflow |= cnt > cap;
        const u64 *row = cand_buf + ((size_t)wq * 2 + half) * cap;
        const int m = min(cnt, cap);
        total += m;
        for (int base = 0; base < m; base += 32) {
            const u64 key = base + lane < m ? row[base + lane] : ~0ull;
            unsigned mask = __ballot_sync(0xffffffffu, key < tau_key);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const u64 kx = __shfl_sync(0xffffffffu, key, src);
                if (kx < tau_key) {
                    list.insert(kx, lane);
                    tau_key = list.last(lane);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) cand[(size_t)wq * C + lane * NPL + i] = list.e[i];
    if (lane == 0) {
        // more than C emitted: everything outside the list is at least the list's last element; otherwise at least tau
        outside_bound[wq] = total > C ? __uint_as_float((unsigned)(tau_key >> 32)) : tau[wq];
        if (overflow) fail_rows[atomicAdd(fail_count, 1)] = q0 + wq;           // buffer overflow: exact fallback for this row
    }
}

__global__ void __launch_bounds__(256)
knn_gather_rows_kernel(const __nv_bfloat16 *__restrict__ Xh, const __nv_bfloat16 *__restrict__ Xl, const float *__restrict__ norm2,
                       const int *__restrict__ rows, int s, int dpad, __nv_bfloat16 *__restrict__ Sh, __nv_bfloat16 *__restrict__ Sl,
                       float *__restrict__ norm_s)
{
    const long long total = (long long)s * (dpad / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / (dpad / 8)), c = (int)(i % (dpad / 8));
        const size_t src = (size_t)rows[r] * dpad + c * 8, dst = (size_t)r * dpad + c * 8;
        *reinterpret_cast<uint4 *>(Sh + dst) = *reinterpret_cast<const uint4 *>(Xh + src);
        *reinterpret_cast<uint4 *>(Sl + dst) = *reinterpret_cast<const uint4 *>(Xl + src);
        if (c == 0) norm_s[r] = norm2[rows[r]];
    }
}

typedef CUresult (*TmapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D map over a row-major (rows x dpad) bf16 matrix, box = 64 columns x 128 rows, 128-byte swizzle, zero fill out of range
static int make_feature_map(CUtensorMap *map, const void *ptr, int64_t rows, int dpad)
{
    static TmapEncodeFn fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
        if (e != cudaSuccess || !f) { cudaGetLastError(); set_error("cuTensorMapEncodeTiled is not available from this driver"); return GLB_E_UNSUPPORTED; }
        fn = (TmapEncodeFn)f;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)dpad, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)dpad * 2};
    cuuint32_t box[2] = {FBK, FBM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return GLB_E_UNSUPPORTED; }
    return 0;
}

struct KnnArena {
    std::vector<void *> v;
    ~KnnArena() { for (void *p : v) dev_free(p); }
    cudaError_t alloc(void **p, size_t bytes) { cudaError_t e = dev_alloc(p, bytes ? bytes : 1); if (e == cudaSuccess) v.push_back(*p); return e; }
};

// tensor cores when d >= 64 (-DGLB_EXPERIMENT builds: GLB_KNN_TC=0 forces the fp32 SIMT distance kernel for A/B runs,
// GLB_KNN_FUSED=0 the unfused tensor-core path)
static int knn_exp_env(const char *name, int def)
{
#ifdef GLB_EXPERIMENT
    const char *e = getenv(name);
    if (e) return atoi(e);
#else
    (void)name;
#endif
    return def;
}
static bool knn_use_tc(int d) { return knn_exp_env("GLB_KNN_TC", d >= 64 ? 1 : 0) != 0; }

constexpr int64_t kFusedMinN = 16384;             // below this the distance block is small enough to go through HBM
constexpr int kFusedSample = 8192;                // sample size of the threshold pass
constexpr int kFusedCap = 512;                    // candidate buffer per (query row, column half); expected fill n R / 2s = 68 at n = 70 000

static size_t fused_smem_bytes(int kb_per_tile, int *stages, int *a_resident, int max_smem)
{
    const size_t a_bytes = (size_t)kb_per_tile * 2 * kFTile;
    *a_resident = a_bytes <= 128 * 1024 ? 1 : 0;
    const size_t per_stage = (size_t)(*a_resident ? 2 : 4) * kFTile;
    const size_t fixed = 1024 /* alignment */ + (*a_resident ? a_bytes : 0) + 2048 /* barriers, norm tiles */;
    int S = (int)(((size_t)max_smem - fixed) / per_stage);
    S = std::max(2, std::min(S, 6));
    *stages = S;
    return fixed + (size_t)S * per_stage;
}

template <int NPL>
static int knn_run(const double *d_X, int64_t n, int d, int k, long long *d_ind, double *d_dist, int *launches, int *fallback_rows,
                   cudaStream_t st, bool tc)
{
    constexpr int C = 32 * NPL;
    KnnArena A;
    const bool fused = tc && n >= kFusedMinN && knn_exp_env("GLB_KNN_FUSED", 1) != 0;
    const int dpad = fused ? (d + FBK - 1) / FBK * FBK : tc ? (d + TBK - 1) / TBK * TBK : (d + BK - 1) / BK * BK;
    const int n_pad = (int)((n + BN - 1) / BN * BN);
    // queries per block.  Unfused: QB x n_pad fp32 distances in HBM, large enough that the one-warp-per-query selection
    // fills the machine (8192 warps = 55 per SM), capped at 4 GB.  Fused: one CTA of 128 queries per SM and launch.
    int QB = 8192;
    if (fused) {
        QB = sm_count() * FBM;
    } else {
        while ((double)QB * n_pad * 4.0 > 4.0e9 && QB > 128) QB /= 2;
        if (QB > n) QB = (int)((n + BM - 1) / BM * BM);
    }
    double *mean; float *Xc, *norm2, *D = nullptr; unsigned long long *rmax2; u64 *cand; int *fail_rows, *fail_count; double *scratch;
    GLB_CUDA(A.alloc((void **)&mean, sizeof(double) * d));
    GLB_CUDA(A.alloc((void **)&Xc, sizeof(float) * (size_t)n * dpad));
    GLB_CUDA(A.alloc((void **)&norm2, sizeof(float) * (size_t)n));
    GLB_CUDA(A.alloc((void **)&rmax2, sizeof(unsigned long long)));
    if (!fused) GLB_CUDA(A.alloc((void **)&D, sizeof(float) * (size_t)QB * n_pad));
    GLB_CUDA(A.alloc((void **)&cand, sizeof(u64) * (size_t)QB * C));
    GLB_CUDA(A.alloc((void **)&fail_rows, sizeof(int) * (size_t)n));
    GLB_CUDA(A.alloc((void **)&fail_count, sizeof(int)));
    const int exact_ctas = sm_count();
    GLB_CUDA(A.alloc((void **)&scratch, sizeof(double) * (size_t)exact_ctas * n));
    GLB_CUDA(cudaMemsetAsync(rmax2, 0, sizeof(unsigned long long), st));
    GLB_CUDA(cudaMemsetAsync(fail_count, 0, sizeof(int), st));
    int nl = 0;
    knn_colsum_kernel<<<d, 256, 0, st>>>(d_X, n, d, mean); ++nl;
    knn_center_kernel<<<sm_count() * 8, 256, 0, st>>>(d_X, mean, n, d, dpad, Xc, norm2, rmax2); ++nl;
    // bound on |approximate cross term - exact| / (|x||y|):
    //   fp32 SIMT: (d + 8) roundings of 2^-24;
    //   tensor cores: 3 * 2^-18 from the hi/lo split + one 2^-23 rounding (truncation) per accumulated product, 3 per k
    double err_rel = (double)(d + 8) * 5.9604644775390625e-8;
    __nv_bfloat16 *Xh = nullptr, *Xl = nullptr;
    if (tc) {
        GLB_CUDA(A.alloc((void **)&Xh, sizeof(__nv_bfloat16) * (size_t)n * dpad));
        GLB_CUDA(A.alloc((void **)&Xl, sizeof(__nv_bfloat16) * (size_t)n * dpad));
        knn_split_kernel<<<sm_count() * 8, 256, 0, st>>>(Xc, (long long)n * dpad, Xh, Xl); ++nl;
        err_rel = 3.0 * 3.814697265625e-6 + (double)(3 * dpad + 8) * 1.1920928955078125e-7;
    }
    if (fused) {
        // ---- thresholds from a fixed pseudo-random sample of the database, then one filtered pass over all of it ----
        // sample size: the threshold is the R-th smallest of s sample distances, so its rank in the full set is ~ n R / s
        // with a relative spread of 1 / sqrt(R) = 25 %; it has to stay above the ~2k + 16 candidates the certificate wants
        // for (nearly) every row, so the expected rank is held at max(137, 8 k): s = 8192 at config 2 (k + 1 = 11),
        // 5632 at config 3 (k + 1 = 21), where 8192 sent 100-170 of the 60 000 rows to the exact fallback
        int s_n = (int)std::min<double>((double)kFusedSample, (double)n * kSampleR / std::max(137.0, 8.0 * k));
        s_n = std::max(1024, s_n / 128 * 128);
        std::vector<int> h_rows((size_t)s_n);
        {   // s_n distinct rows: a multiplicative walk through the residues mod n (an odd stride coprime to n), fixed seed
            unsigned long long x = 0x9E3779B97F4A7C15ull % (unsigned long long)n, stride = (unsigned long long)(0.6180339887 * (double)n) | 1ull;
            auto gcd = [](unsigned long long a, unsigned long long b) { while (b) { const unsigned long long t = a % b; a = b; b = t; } return a; };
            while (gcd(stride, (unsigned long long)n) != 1ull) stride += 2;
            for (int i = 0; i < s_n; ++i) { h_rows[i] = (int)x; x = (x + stride) % (unsigned long long)n; }
        }
        int *d_rows; __nv_bfloat16 *Sh, *Sl; float *norm_s, *tau, *best, *bound; u64 *cand_buf; int *cand_count;
        GLB_CUDA(A.alloc((void **)&d_rows, sizeof(int) * s_n));
        GLB_CUDA(A.alloc((void **)&Sh, sizeof(__nv_bfloat16) * (size_t)s_n * dpad));
        GLB_CUDA(A.alloc((void **)&Sl, sizeof(__nv_bfloat16) * (size_t)s_n * dpad));
        GLB_CUDA(A.alloc((void **)&norm_s, sizeof(float) * s_n));
        GLB_CUDA(A.alloc((void **)&tau, sizeof(float) * (size_t)QB));
        GLB_CUDA(A.alloc((void **)&best, sizeof(float) * (size_t)n * 2 * kSampleR));
        GLB_CUDA(A.alloc((void **)&bound, sizeof(float) * (size_t)QB));
        GLB_CUDA(A.alloc((void **)&cand_buf, sizeof(u64) * (size_t)QB * 2 * kFusedCap));
        GLB_CUDA(A.alloc((void **)&cand_count, sizeof(int) * (size_t)QB * 2));
        GLB_CUDA(cudaMemcpyAsync(d_rows, h_rows.data(), sizeof(int) * s_n, cudaMemcpyHostToDevice, st));
        knn_gather_rows_kernel<<<sm_count() * 4, 256, 0, st>>>(Xh, Xl, norm2, d_rows, s_n, dpad, Sh, Sl, norm_s); ++nl;
        CUtensorMap tm_xh, tm_xl, tm_sh, tm_sl;
        int rc;
        if ((rc = make_feature_map(&tm_xh, Xh, n, dpad)) || (rc = make_feature_map(&tm_xl, Xl, n, dpad)) ||
            (rc = make_feature_map(&tm_sh, Sh, s_n, dpad)) || (rc = make_feature_map(&tm_sl, Sl, s_n, dpad)))
            return rc;
        int dev = 0, max_smem = 0;
        GLB_CUDA(cudaGetDevice(&dev));
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        FusedArgs fa{};
        fa.norm_q = norm2; fa.kb_per_tile = dpad / FBK; fa.cap = kFusedCap;
        const size_t smem = fused_smem_bytes(fa.kb_per_tile, &fa.stages, &fa.a_resident, max_smem);
        GLB_CUDA(cudaFuncSetAttribute(knn_fused_kernel<MODE_SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GLB_CUDA(cudaFuncSetAttribute(knn_fused_kernel<MODE_EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int64_t q0 = 0; q0 < n; q0 += QB) {             // thresholds: every query against the sample
            const int nq = (int)std::min<int64_t>(QB, n - q0);
            fa.q0 = (int)q0; fa.nq = nq; fa.norm_d = norm_s; fa.nd = s_n; fa.best_out = best + (size_t)q0 * 2 * kSampleR;
            knn_fused_kernel<MODE_SAMPLE><<<(nq + FBM - 1) / FBM, kFusedThreads, smem, st>>>(tm_xh, tm_xl, tm_sh, tm_sl, fa); ++nl;
        }
        err_rel += 1.0e-6;                                   // the filter compares in fp32: a few ulps of |q|^2 + |x|^2
        for (int64_t q0 = 0; q0 < n; q0 += QB) {
            const int nq = (int)std::min<int64_t>(QB, n - q0);
            fa.q0 = (int)q0; fa.nq = nq; fa.norm_d = norm2; fa.nd = (int)n; fa.best_in = best + (size_t)q0 * 2 * kSampleR; fa.tau_out = tau;
            fa.cand_buf = cand_buf; fa.cand_count = cand_count;
            knn_fused_kernel<MODE_EMIT><<<(nq + FBM - 1) / FBM, kFusedThreads, smem, st>>>(tm_xh, tm_xl, tm_xh, tm_xl, fa); ++nl;
            knn_pick_kernel<NPL><<<(nq * 32 + 255) / 256, 256, 0, st>>>(cand_buf, cand_count, tau, kFusedCap, (int)q0, nq, cand, bound,
                                                                         fail_rows, fail_count); ++nl;
            knn_rerank_kernel<NPL><<<(nq * 32 + 255) / 256, 256, 0, st>>>(d_X, (int)n, d, norm2, rmax2, cand, (int)q0, nq, k, d_ind, d_dist,
                                                                           fail_rows, fail_count, err_rel, bound); ++nl;
        }
    } else {
        if (tc) GLB_CUDA(cudaFuncSetAttribute(knn_dist_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
        for (int64_t q0 = 0; q0 < n; q0 += QB) {
            const int nq = (int)std::min<int64_t>(QB, n - q0);
            dim3 grid((unsigned)(n_pad / BN), (unsigned)((nq + BM - 1) / BM));
            if (tc) knn_dist_tc_kernel<<<grid, kTcThreads, kTcSmemBytes, st>>>(Xh, Xl, norm2, (int)n, dpad, (int)q0, nq, D, (long long)n_pad);
            else knn_dist_kernel<<<grid, kDistThreads, 0, st>>>(Xc, norm2, (int)n, dpad, (int)q0, nq, D, (long long)n_pad);
            ++nl;
            knn_select_kernel<NPL><<<(nq * 32 + 255) / 256, 256, 0, st>>>(D, (long long)n_pad, n_pad, nq, cand); ++nl;
            knn_rerank_kernel<NPL><<<(nq * 32 + 255) / 256, 256, 0, st>>>(d_X, (int)n, d, norm2, rmax2, cand, (int)q0, nq, k, d_ind, d_dist,
                                                                           fail_rows, fail_count, err_rel, nullptr); ++nl;
        }
    }
    knn_exact_rows_kernel<<<exact_ctas, 256, 0, st>>>(d_X, (int)n, d, fail_rows, fail_count, k, scratch, d_ind, d_dist); ++nl;
    GLB_LAUNCH_CHECK();
    int h_fail = 0;
    GLB_CUDA(cudaMemcpyAsync(&h_fail, fail_count, sizeof(int), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));                      // the arena is freed on return
    if (launches) *launches += nl;
    if (fallback_rows) *fallback_rows = h_fail;
    return 0;
}

}  // namespace glb

using namespace glb;

extern "C" GLB_API int glb_knn_search(const double *d_X, int64_t n, int d, int k, int64_t *d_ind, double *d_dist, int *launches,
                                      int *fallback_rows, void *stream)
{
    GLB_CHECK_ARG(d_X && d_ind && d_dist, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31) - 256, "n out of range");
    GLB_CHECK_ARG(d > 0 && d <= 65536, "d out of range");
    GLB_CHECK_ARG(k > 0 && k <= n, "k must be in [1, n]");
    if (k > 112) { set_error("glb_knn_search: k = %d (> 112) is not supported", k); return GLB_E_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    // candidates: at least k + 8 and 1.25 k, in steps of 32; the tensor-core distances carry a ~7x wider error margin,
    // so that path keeps a list twice as long for the certificate
    const bool tc = knn_use_tc(d);
    const int kk = tc ? 2 * k + 16 : k;
    if (kk + 8 <= 32 && kk * 5 / 4 <= 32) return knn_run<1>(d_X, n, d, k, (long long *)d_ind, d_dist, launches, fallback_rows, st, tc);
    if (kk + 8 <= 64 && kk * 5 / 4 <= 64) return knn_run<2>(d_X, n, d, k, (long long *)d_ind, d_dist, launches, fallback_rows, st, tc);
    return knn_run<4>(d_X, n, d, k, (long long *)d_ind, d_dist, launches, fallback_rows, st, tc);
}

extern "C" GLB_API int glb_knn_search_host(const double *h_X, int64_t n, int d, int k, int64_t *h_ind, double *h_dist, int *launches,
                                           int *fallback_rows)
{
    GLB_CHECK_ARG(h_X && h_ind && h_dist, "null pointer");
    GLB_CHECK_ARG(n > 0 && d > 0 && k > 0 && k <= n, "bad shape");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("glb_knn_search_host: no CUDA device visible");
        return GLB_E_NOGPU;
    }
    KnnArena A;
    double *X, *dist; long long *ind;
    GLB_CUDA(A.alloc((void **)&X, sizeof(double) * (size_t)n * d));
    GLB_CUDA(A.alloc((void **)&ind, sizeof(long long) * (size_t)n * k));
    GLB_CUDA(A.alloc((void **)&dist, sizeof(double) * (size_t)n * k));
    cudaStream_t st = 0;
    GLB_CUDA(cudaMemcpyAsync(X, h_X, sizeof(double) * (size_t)n * d, cudaMemcpyHostToDevice, st));
    if (launches) *launches = 0;
    int rc = glb_knn_search(X, n, d, k, (int64_t *)ind, dist, launches, fallback_rows, st);
    if (rc) return rc;
    GLB_CUDA(cudaMemcpyAsync(h_ind, ind, sizeof(long long) * (size_t)n * k, cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaMemcpyAsync(h_dist, dist, sizeof(double) * (size_t)n * k, cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    return 0;
}
