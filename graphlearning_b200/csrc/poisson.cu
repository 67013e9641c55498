// poisson.cu - the Poisson-learning iterate u <- Db + P u on sm_100a.
//
// Replaces the hot loop of ssl.poisson._fit, gradient-descent branch
// (reference graphlearning/ssl.py:667-669; third-party arithmetic: scipy _sparsetools csr_matvecs).
//
// On a kNN graph one iteration is nnz random gathers of a 64-byte row of the label matrix (c = 10 classes).
// tools/gather_microbench.cu measured what one SM can gather from L2: 1.0 row per cycle through LDG.128
// (one L1 wavefront per row; 17 TB/s chip-wide), 3.5-5 cycles per row through the TMA unit (bulk copies,
// tile::gather4).  So the iterate is bound by L1 wavefronts, ~6800 per SM per iteration on the 70k-node
// graph = 3.5 us, and every cycle spent elsewhere (grid barriers, instruction issue, divergence) is on top.
//
// Three kernels, chosen per graph by glb_poisson_plan_create:
//   poisson_dataflow_kernel    T iterations in ONE launch, one CTA per SM, no grid barrier in the data path: every
//                              16-byte chunk of a label row carries three values plus the iteration number that
//                              produced it ("flag in data"), so a gather doubles as the synchronisation - a lane whose
//                              chunk is still from an older iteration re-polls it.  Rows of the CTA are sorted by length
//                              and stored as sliced ELL in shared memory; every warp walks its slices as ONE stream of
//                              entry pairs with 6-8 gathers per lane in flight.  Four versions of the label matrix
//                              rotate, so the first attempt of a gather may be served by L1 (under the locality ordering
//                              of reorder.cu a CTA gathers every distinct row ~2.4 times per iteration).  A gate every
//                              `gate_every` iterations (tuned per graph, usually 1) keeps the CTAs in phase.  Needs a
//                              structurally symmetric pattern (argument below).
//   poisson_persistent_kernel  T iterations in one cooperative launch with a hand-rolled grid barrier between
//                              iterations; CSR slab in shared memory.  For directed graphs (symmetrize=False).
//   poisson_step_kernel        one iteration per launch, CSR read from global memory.  For graphs too big for the
//                              shared-memory slabs; genuinely HBM/L2-gather bound.
//
// Lane mapping (all kernels): LANES lanes own one matrix row, each lane one 16-byte piece of it; a warp-wide
// LDG.128 gathers 32/LANES complete rows.  Each lane walks the nonzeros of its row UNROLL at a time: all UNROLL
// gathers are issued before the first FMA; there is no cross-lane reduction at all.
//
// Label-matrix layouts (glb_poisson_plan_ld gives the row stride):
//   plain    row-major n x ldu fp32, ldu = glb_padded_ld(c)                        (step / barrier kernels)
//   flagged  row-major n x ldu fp32, ldu = 4 * LANES; chunk q of a row = {x[3q], x[3q+1], x[3q+2], epoch}
//            (dataflow kernel; the step kernel can read and write it too, ignoring the epoch word)
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>
#include "common.cuh"

namespace glb {

constexpr int kUnroll = 8;      // nonzeros in flight per lane in the one-launch-per-iteration kernel

__device__ __forceinline__ void fma4(float4 &acc, float a, const float4 &x)
{
    acc.x = fmaf(a, x.x, acc.x);
    acc.y = fmaf(a, x.y, acc.y);
    acc.z = fmaf(a, x.z, acc.z);
    acc.w = fmaf(a, x.w, acc.w);
}

template <bool NC>
__device__ __forceinline__ float4 load_u4(const float *p)
{
    if (NC) return __ldg(reinterpret_cast<const float4 *>(p));
    return *reinterpret_cast<const float4 *>(p);   // coherent at L1 after the grid barrier's fence
}

// Nonzero j as (byte offset of row col[j] inside u, value).  Offsets are 32-bit: n * ldu * 4 < 2^32 is checked
// on the host.  The persistent kernel keeps the pairs in shared memory, precomputed once per launch.
struct CsrGlobal {
    const int *__restrict__ col;
    const float *__restrict__ val;
    unsigned row_bytes;
    __device__ __forceinline__ void get(int j, unsigned &off, float &a) const
    {
        off = (unsigned)__ldg(col + j) * row_bytes;
        a = __ldg(val + j);
    }
};
struct CsrShared {
    const int2 *cv;
    __device__ __forceinline__ void get(int j, unsigned &off, float &a) const
    {
        const int2 e = cv[j];
        off = (unsigned)e.x;
        a = __int_as_float(e.y);
    }
};

// sum_j val[j] * u[col[j], 4 columns] over the nonzeros [beg, end) of one row; ubase = u + column offset.
// Full batches of UNROLL run without predicates (all gathers issued before the first FMA); the tail is predicated.
template <bool NC, int UNROLL, typename Csr>
__device__ __forceinline__ float4 row_times_u(const Csr &csr, int beg, int end, const char *__restrict__ ubase)
{
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int j = beg;
    for (; j + UNROLL <= end; j += UNROLL) {
        unsigned off[UNROLL];
        float a[UNROLL];
        float4 x[UNROLL];
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) csr.get(j + i, off[i], a[i]);
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) x[i] = load_u4<NC>(reinterpret_cast<const float *>(ubase + off[i]));
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) fma4(acc, a[i], x[i]);
    }
    if (j < end) {
        unsigned off[UNROLL];
        float a[UNROLL];
        float4 x[UNROLL];
#pragma unroll
        for (int i = 0; i < UNROLL - 1; ++i) {
            off[i] = 0; a[i] = 0.f;
            if (j + i < end) csr.get(j + i, off[i], a[i]);
        }
#pragma unroll
        for (int i = 0; i < UNROLL - 1; ++i) {
            x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j + i < end) x[i] = load_u4<NC>(reinterpret_cast<const float *>(ubase + off[i]));
        }
#pragma unroll
        for (int i = 0; i < UNROLL - 1; ++i) fma4(acc, a[i], x[i]);      // a = 0, x = 0 past the end of the row
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------
// K1: one iteration per launch, CSR read from global memory
// ------------------------------------------------------------------------------------------------
template <int LANES, bool FLAGGED>
__global__ void __launch_bounds__(256)
poisson_step_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const float *__restrict__ val,
                    const float *__restrict__ Db, const float *__restrict__ u_in, float *__restrict__ u_out,
                    int n, int ldu)
{
    const int li = threadIdx.x % LANES;
    const long long rid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const long long nrid = ((long long)gridDim.x * blockDim.x) / LANES;
    const CsrGlobal csr{col, val, (unsigned)ldu * 4u};
    for (long long row = rid; row < n; row += nrid) {
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        for (int coff = li * 4; coff < ldu; coff += LANES * 4) {
            float4 acc = row_times_u<true, kUnroll>(csr, beg, end, reinterpret_cast<const char *>(u_in + coff));
            const size_t o = (size_t)row * ldu + coff;
            const float4 b = __ldg(reinterpret_cast<const float4 *>(Db + o));
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
            if (FLAGGED) acc.w = 0.f;                      // the epoch word of the chunk is not data
            *reinterpret_cast<float4 *>(u_out + o) = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2: persistent, T iterations per launch
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_relaxed_add(unsigned *p, unsigned v)
{
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed(unsigned *p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Watchdog of the polling loops: a poll normally ends within microseconds.  One that is still waiting after
// kPollLimit cycles (~2 s) can only mean a broken invariant (wrong buffer size, a lost CTA); it raises the flag, every
// other polling loop sees the flag within a thousand spins, the kernel drains with garbage results and
// glb_poisson_plan_check reports GLB_E_TIMEOUT instead of a hung GPU.
constexpr long long kPollLimit = 4000000000ll;
__device__ __forceinline__ bool poll_expired(unsigned *watchdog, unsigned &spins, long long &t0)
{
    if ((++spins & 1023u) != 0u) return false;
    if (t0 == 0) { t0 = clock64(); return false; }
    if (ld_relaxed(watchdog) != 0u) return true;
    if (clock64() - t0 > kPollLimit) { st_relaxed(watchdog, 1u); return true; }
    return false;
}

// Grid-wide barrier between iterations (all CTAs are co-resident: cooperative launch, one per SM).
// bar.sync orders the CTA's u stores before thread 0's release fence; the arrival is a relaxed red/st; the
// waiters poll with relaxed loads (no L1 invalidation per poll) and issue ONE acquire fence when the epoch
// is complete, which also drops this SM's stale L1 lines before the next iteration's gathers.
//   FLAGS = false: one monotone counter, spin until it reaches epoch * gridDim.x
//   FLAGS = true : one flag word per CTA (128 bytes apart); thread i spins on the flag of CTA i
constexpr int kFlagStride = 32;          // unsigned words = 128 bytes
template <bool FLAGS>
__device__ __forceinline__ void grid_barrier(unsigned *sync_words, unsigned epoch)
{
    __syncthreads();
    if (FLAGS) {
        if (threadIdx.x == 0) {
            fence_acq_rel_gpu();
            st_relaxed(sync_words + (size_t)blockIdx.x * kFlagStride, epoch);
        }
        if (threadIdx.x < gridDim.x) {
            while (ld_relaxed(sync_words + (size_t)threadIdx.x * kFlagStride) < epoch) { }
            fence_acq_rel_gpu();
        }
    } else {
        if (threadIdx.x == 0) {
            fence_acq_rel_gpu();
            red_relaxed_add(sync_words, 1u);
            while (ld_relaxed(sync_words) < epoch * gridDim.x) { }
            fence_acq_rel_gpu();
        }
    }
    __syncthreads();
}

constexpr int kLongRow = 32;     // rows with more nonzeros are split over the lane groups of a whole warp

template <int LANES, int THREADS, int UNROLL, bool FLAGS>
__global__ void __launch_bounds__(THREADS, 1)
poisson_persistent_kernel(const int *__restrict__ rowptr, const int *__restrict__ col,
                          const float *__restrict__ val, const float *__restrict__ Db, float *u0, float *u1, int n,
                          int ldu, int T, const int *__restrict__ cta_rows, int max_rows, int slab_cap,
                          unsigned *sync_words)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: (col,val) [slab_cap] int2 | rp [max_rows+1] i32 | long rows [max_rows] i32 | n_long i32 |
    //         has-source flag [max_rows] u8
    int2 *s_cv = reinterpret_cast<int2 *>(smem_raw);
    int *s_rp = reinterpret_cast<int *>(s_cv + slab_cap);
    int *s_long = s_rp + (max_rows + 1);
    int *s_nlong = s_long + max_rows;
    unsigned char *s_src = reinterpret_cast<unsigned char *>(s_nlong + 1);

    const int r0 = cta_rows[blockIdx.x];
    const int r1 = cta_rows[blockIdx.x + 1];
    const int nrows = r1 - r0;
    const int nz0 = rowptr[r0];
    const int nnz_slab = rowptr[r1] - nz0;
    if (threadIdx.x == 0) *s_nlong = 0;
    for (int i = threadIdx.x; i < nnz_slab; i += THREADS)
        s_cv[i] = make_int2((int)((unsigned)col[nz0 + i] * (unsigned)ldu * 4u), __float_as_int(val[nz0 + i]));
    for (int i = threadIdx.x; i <= nrows; i += THREADS) s_rp[i] = rowptr[r0 + i] - nz0;
    __syncthreads();
    // The Poisson source Db is zero except on the labelled rows: remember which rows of this block have
    // one instead of streaming n x ldu zeros every iteration.  Rows too long for one lane group are listed.
    for (int lr = threadIdx.x; lr < nrows; lr += THREADS) {
        bool nz = false;
        const float4 *b = reinterpret_cast<const float4 *>(Db + (size_t)(r0 + lr) * ldu);
        for (int q = 0; q < ldu / 4; ++q) {
            const float4 v = __ldg(b + q);
            nz |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);      // NaN != 0 is true
        }
        s_src[lr] = nz ? 1 : 0;
        if (s_rp[lr + 1] - s_rp[lr] > kLongRow) s_long[atomicAdd(s_nlong, 1)] = lr;
    }
    __syncthreads();
    const int n_long = *s_nlong;

    const int li = threadIdx.x % LANES;
    const int rid = threadIdx.x / LANES;
    constexpr int RPP = THREADS / LANES;            // rows per pass of the CTA
    constexpr int NG = 32 / LANES;                  // lane groups per warp
    const int lane = threadIdx.x & 31;
    const int grp = lane / LANES;
    const int warp = threadIdx.x >> 5;
    const CsrShared csr{s_cv};

    for (int t = 0; t < T; ++t) {
        const float *u_in = (t & 1) ? u1 : u0;
        float *u_out = (t & 1) ? u0 : u1;
        // phase A: one lane group per (short) row
        for (int lr = rid; lr < nrows; lr += RPP) {
            const int beg = s_rp[lr], end = s_rp[lr + 1];
            if (end - beg > kLongRow) continue;
            for (int coff = li * 4; coff < ldu; coff += LANES * 4) {
                float4 acc = row_times_u<false, UNROLL>(csr, beg, end, reinterpret_cast<const char *>(u_in + coff));
                const size_t o = (size_t)(r0 + lr) * ldu + coff;
                if (s_src[lr]) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(Db + o));
                    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
                }
                *reinterpret_cast<float4 *>(u_out + o) = acc;
            }
        }
        // phase B: one warp per long row, nonzeros split over its NG lane groups, fixed-order shuffle reduction
        for (int q = warp; q < n_long; q += THREADS / 32) {
            const int lr = s_long[q];
            const int beg = s_rp[lr], end = s_rp[lr + 1];
            const int per = (end - beg + NG - 1) / NG;
            const int gb = min(end, beg + grp * per), ge = min(end, gb + per);
            for (int coff = li * 4; coff < ldu; coff += LANES * 4) {
                float4 acc = row_times_u<false, UNROLL>(csr, gb, ge, reinterpret_cast<const char *>(u_in + coff));
#pragma unroll
                for (int off = LANES; off < 32; off <<= 1) {
                    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
                    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
                    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
                    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
                }
                if (grp == 0) {
                    const size_t o = (size_t)(r0 + lr) * ldu + coff;
                    if (s_src[lr]) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(Db + o));
                        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
                    }
                    *reinterpret_cast<float4 *>(u_out + o) = acc;
                }
            }
        }
        if (t + 1 < T) grid_barrier<FLAGS>(sync_words, (unsigned)(t + 1));
    }
}

// ------------------------------------------------------------------------------------------------
// K3: persistent, barrier-free ("flag in data"), T iterations per launch
// ------------------------------------------------------------------------------------------------
// Version v of the label matrix (v = number of iterations applied) lives in buffer v & 1 and every 16-byte
// chunk of it carries the word 1 + v.  Iteration t gathers chunks of version t, re-polling any chunk whose
// word is not 1 + t yet, and stores its own chunk of version t + 1.  Loads and stores are 16-byte single
// transactions at L2 (ld/st.relaxed.gpu, L1 bypassed), so a chunk is seen whole or not at all.
//
// Why two buffers are enough when the sparsity pattern is symmetric: chunk q of u_{v+2}[i] overwrites chunk q
// of u_v[i].  Its readers are lane q of the rows j with i in row j's pattern; by symmetry j is in row i's
// pattern, so lane q of row i gathered u_{v+1}[j] (chunk q) before it stored u_{v+2}[i], and lane q of row j
// stored u_{v+1}[j] only after its own gather of u_v[i] had returned.  Every row is owned by the same lane
// group in every iteration and a lane group finishes iteration t before it starts t + 1, so the chain holds
// chunk by chunk without any fence.  Progress: the unfinished work item with the smallest iteration number
// always has all its inputs, all CTAs are co-resident (cooperative launch), so no wait cycle can form.
constexpr int kScratchRows = 256;                // rows n .. n+255 of the label matrices are the padding targets of the slabs
constexpr int kRowSrcBit = 0x40000000;           // slot_rows: row has a nonzero source term Db
constexpr int kRing = 4;                         // versions of the label matrix in flight: version v lives in buffer v % kRing

__device__ __forceinline__ uint4 ld_chunk(const char *p)          // coherent at L2 (L1 bypassed)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// First attempt of a gather: a weak load that may be served by this SM's L1.  Chunks are self-validating (epoch word), so
// a stale line is harmless: the lane re-polls it at L2.  With kRing = 4 buffers the copy of a line that L1 may still hold
// is four iterations old, i.e. evicted long ago by the ~180 KB of label rows every iteration streams through L1, so in
// practice a first attempt either hits a line another warp of this CTA fetched in THIS iteration (under a locality
// ordering a CTA gathers every distinct row ~2.4 times per iteration) or misses and is filled from L2.
__device__ __forceinline__ uint4 ld_chunk_l1(const char *p)
{
    uint4 v;
    asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_chunk(char *p, float a, float b, float c, unsigned w)
{
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
                 "r"(__float_as_uint(c)), "r"(w) : "memory");
}

// Slot types of the sliced-ELL slabs.  NORMAL: 32/LANES rows side by side.  The other three hold ONE (part of a)
// long row whose nonzeros are dealt round-robin to the lane groups of the warp and summed with shuffles:
// LONG = whole row; OWNER/PART = a hub row split over several warps, the PART warps leave their partial sums
// (3 floats + iteration number per lane) in shared memory and the OWNER warp adds them in a fixed order.
constexpr int kSlotNormal = 0, kSlotLong = 1, kSlotPart = 2, kSlotOwner = 3;
constexpr int kLongRowDf = GLB_DATAFLOW_LONG_ROW;                   // rows with more nonzeros get a warp (or several) of their own

// Inner loop: no per-entry predicates at all.  Entries of a lane group are stored in PAIRS (one LDS.128 = two
// (offset,value) entries, 8 lane groups side by side = one 128-byte wavefront), padding entries point at the scratch
// rows behind row n-1 of the label matrix (value 0, epoch 0xffffffff: always "ready"), the epoch test is `>=` (a chunk
// of buffer t % kRing only ever holds a version congruent to t and never a later one than t, see the argument above),
// the address is one IMAD.WIDE (64-bit base + 32-bit byte offset), and a slice of width L is walked as U-entry batches
// plus 4/2/1 tails selected by the bits of L (warp-uniform), so exactly L gathers are issued per row.
__device__ __forceinline__ const char *df_addr(const char *base, unsigned off)
{
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(a) : "r"(off), "l"(base));
    return reinterpret_cast<const char *>(a);
}

struct DfRing { float *b[kRing]; };              // version v of the label matrix lives in b[v % kRing]

// Gather stream.  The round-1 kernel gave one warp ONE batch of 8 gathers in flight: wait for it (an L2 round trip of
// ~500-1000 cycles under load), consume it, only then issue the next; the LSU data pipe (one 128-byte line per cycle and
// SM, i.e. one label row per cycle) idled a third of the time.  Here every warp walks ONE contiguous
// stream of entry PAIRS (all slots of the warp back to back, slice widths padded to even only) through a ring of four
// register slots: pair p+4 is issued as soon as pair p has been validated and consumed, across slot boundaries, so 6-8
// gathers per lane are outstanding all the time and the instruction stream overlaps the memory latency.  With every
// synchronisation removed (GLB_POISSON_FREE probe of the experiment build) the data pipe runs at 88 % of its peak.
template <int RPW, bool L1F>
__device__ __forceinline__ void dfp_issue(const int4 *cv, const char *in, float (&val)[2], uint4 (&x)[2])
{
    const int4 e = cv[0];                                     // one pair = two (offset, value) entries of this lane group
    val[0] = __int_as_float(e.y); val[1] = __int_as_float(e.w);
    x[0] = L1F ? ld_chunk_l1(df_addr(in, (unsigned)e.x)) : ld_chunk(df_addr(in, (unsigned)e.x));
    x[1] = L1F ? ld_chunk_l1(df_addr(in, (unsigned)e.z)) : ld_chunk(df_addr(in, (unsigned)e.z));
}

// cv: where dfp_issue read the entries of this pair (the offsets are re-read from shared memory on the rare re-poll path
// instead of being held in registers)
__device__ __forceinline__ void dfp_consume(const int4 *cv, const char *in, unsigned expect, const float (&val)[2],
                                            uint4 (&x)[2], float &a0, float &a1, float &a2, unsigned &n_poll,
                                            unsigned &n_badbatch, unsigned *watchdog, unsigned poll_sleep)
{
    bool ok = (x[0].w >= expect) & (x[1].w >= expect);
    if (!ok) {                                           // a producer is still behind (or L1 held an old line): re-poll at L2
        const int4 e = cv[0];
        const unsigned off[2] = {(unsigned)e.x, (unsigned)e.z};
        unsigned spins = 0;
        long long t0 = 0;
        do {
            ++n_badbatch;
            if (poll_expired(watchdog, spins, t0)) break;
            if ((poll_sleep & 0xffffu) && spins > 1) __nanosleep(poll_sleep & 0xffffu);   // this warp is ahead of its producers
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (x[i].w < expect) { x[i] = ld_chunk(df_addr(in, off[i])); ++n_poll; }
            ok = (x[0].w >= expect) & (x[1].w >= expect);
        } while (!ok);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        a0 = fmaf(val[i], __uint_as_float(x[i].x), a0);
        a1 = fmaf(val[i], __uint_as_float(x[i].y), a1);
        a2 = fmaf(val[i], __uint_as_float(x[i].z), a2);
    }
}

template <int LANES, int THREADS, bool L1F>
__global__ void __launch_bounds__(THREADS, 1)
poisson_dataflow_kernel(const int2 *__restrict__ slabs, const long long *__restrict__ slab_off,
                             const int4 *__restrict__ slots, const int *__restrict__ slot_off,
                             const int *__restrict__ slot_rows, const float *__restrict__ Db, DfRing ring, int T,
                             int cap_entries, int cap_slots, int cap_parts, unsigned long long *stats,
                             unsigned *start_gate, int gate_every, unsigned *watchdog, unsigned poll_sleep)
{
    constexpr int RPW = 32 / LANES;
    constexpr int NW = THREADS / 32;
    constexpr unsigned ROWB = LANES * 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int2 *s_cv = reinterpret_cast<int2 *>(smem_raw);
    int4 *s_slot = reinterpret_cast<int4 *>(s_cv + cap_entries);
    volatile float *s_part = reinterpret_cast<volatile float *>(s_slot + cap_slots);     // [cap_parts][2][LANES][4]
    int *s_rows = reinterpret_cast<int *>(const_cast<float *>(s_part) + (size_t)cap_parts * 2 * LANES * 4);

    const long long e0 = slab_off[blockIdx.x];
    const int nent = (int)(slab_off[blockIdx.x + 1] - e0);
    const int sl0 = slot_off[blockIdx.x];
    const int nslots = slot_off[blockIdx.x + 1] - sl0;
    for (int i = threadIdx.x; i < nent; i += THREADS) s_cv[i] = slabs[e0 + i];
    for (int i = threadIdx.x; i < nslots; i += THREADS) s_slot[i] = slots[sl0 + i];
    for (int i = threadIdx.x; i < cap_parts * 2 * LANES * 4; i += THREADS) s_part[i] = 0.f;
    for (int i = threadIdx.x; i < nslots * RPW; i += THREADS) {          // which rows have a source term (see the batch kernel)
        int r = slot_rows[(size_t)sl0 * RPW + i];
        if (r >= 0) {
            bool nz = false;
            const float4 *b = reinterpret_cast<const float4 *>(Db + (size_t)r * (ROWB / 4));
            for (int q = 0; q < LANES; ++q) {
                const float4 v = __ldg(b + q);
                nz |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f);
            }
            if (nz) r |= kRowSrcBit;
        }
        s_rows[i] = r;
    }
    __syncthreads();
    if (start_gate) {                                                    // all CTAs enter iteration 0 together
        if (threadIdx.x == 0) {
            red_relaxed_add(start_gate, 1u);
            unsigned spins = 0;
            long long t0 = 0;
            while (ld_relaxed(start_gate) < gridDim.x && !poll_expired(watchdog, spins, t0)) { }
        }
        __syncthreads();
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LANES, li = lane % LANES;
    const int nsw = (nslots - warp + NW - 1) / NW;                       // slots of this warp: k * NW + warp
    int n_pairs = 0;                                                     // pairs of this warp's stream
    for (int k = 0; k < nsw; ++k) n_pairs += (s_slot[k * NW + warp].y + 1) >> 1;
    const int4 *stream = nsw > 0 ? reinterpret_cast<const int4 *>(s_cv + s_slot[warp].x) + g : nullptr;
    unsigned n_poll = 0, n_badbatch = 0;
    const long long clk0 = clock64();
#ifdef GLB_EXPERIMENT
    long long busy = 0;                                                  // GLB_POISSON_STATS: cycles this warp spent working (not at a gate)
#endif
    for (int t = 0; t < T; ++t) {
        if (start_gate && gate_every > 0 && t > 0 && t % gate_every == 0) {
            // Re-alignment gate: nobody starts iteration t before every CTA has finished iteration t - 1 (relaxed: no fence, the
            // flags keep the data correct).  ONE poller per CTA: with every warp polling the counter itself (2 368 pollers on
            // one L2 line, tried) the arrivals queue behind the polls and the iterate slows from 6.3 to 9.0 us.
            __syncthreads();
            if (threadIdx.x == 0) {
                red_relaxed_add(start_gate, 1u);
                const unsigned want = gridDim.x * (unsigned)(t / gate_every + 1);
                unsigned spins = 0;
                long long t0 = 0;
                while (ld_relaxed(start_gate) < want && !poll_expired(watchdog, spins, t0)) { }
            }
            __syncthreads();
        }
        if (nsw == 0) continue;
#ifdef GLB_EXPERIMENT
        const long long tb0 = stats ? clock64() : 0;
#endif
        const int vi = t & (kRing - 1), vo = (t + 1) & (kRing - 1);
        const char *in = reinterpret_cast<const char *>(vi == 0 ? ring.b[0] : vi == 1 ? ring.b[1] : vi == 2 ? ring.b[2] : ring.b[3]) + li * 16;
        char *out = reinterpret_cast<char *>(vo == 0 ? ring.b[0] : vo == 1 ? ring.b[1] : vo == 2 ? ring.b[2] : ring.b[3]) + li * 16;
#ifdef GLB_EXPERIMENT
        // ceiling probes (results are wrong): bit 30 of poll_sleep = every chunk counts as ready (pure gather throughput, no
        // synchronisation), bit 31 = no stores either (the label matrix stays read-only)
        const unsigned expect = (poll_sleep & 0x40000000u) ? 0u : 1u + (unsigned)t;
#else
        const unsigned expect = 1u + (unsigned)t;
#endif
        float v0[2], v1[2], v2[2], v3[2];
        uint4 x0[2], x1[2], x2[2], x3[2];
        if (0 < n_pairs) dfp_issue<RPW, L1F>(stream, in, v0, x0);
        if (1 < n_pairs) dfp_issue<RPW, L1F>(stream + RPW, in, v1, x1);
        if (2 < n_pairs) dfp_issue<RPW, L1F>(stream + 2 * RPW, in, v2, x2);
        if (3 < n_pairs) dfp_issue<RPW, L1F>(stream + 3 * RPW, in, v3, x3);
        int p = 0, k = 0;                                    // pair being consumed, slot it belongs to
        int4 sl = s_slot[warp];                              // (first entry, slice width, type | parts << 8, partial index)
        int left = (sl.y + 1) >> 1;                          // pairs of slot k not yet consumed
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        // finish slot k (sum of a long row's pieces, source term, store) and move to the next one
        auto finish = [&]() {
            const int s = k * NW + warp;
            const int type = sl.z & 0xff;
            bool store = true;
            if (type != kSlotNormal) {                       // warp-uniform: one long row dealt over the lane groups
#pragma unroll
                for (int o = LANES; o < 32; o <<= 1) {
                    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                }
                if (type == kSlotPart) {
                    if (g == 0) {
                        volatile float *pb = s_part + ((size_t)(sl.w * 2 + (t & 1)) * LANES + li) * 4;
                        pb[0] = a0; pb[1] = a1; pb[2] = a2;
                        __threadfence_block();
                        pb[3] = __uint_as_float(2u + (unsigned)t);
                    }
                    store = false;
                } else if (type == kSlotOwner) {
                    const int nparts = sl.z >> 8;
                    for (int q = 0; q < nparts; ++q) {       // partial sums of the other warps, fixed order
                        volatile float *pb = s_part + ((size_t)((sl.w + q) * 2 + (t & 1)) * LANES + li) * 4;
                        unsigned spins = 0;
                        long long tw0 = 0;
                        while (__float_as_uint(pb[3]) != 2u + (unsigned)t && !poll_expired(watchdog, spins, tw0)) { }
                        __threadfence_block();
                        a0 += pb[0]; a1 += pb[1]; a2 += pb[2];
                    }
                }
            }
            const int rinfo = s_rows[s * RPW + g];
            if (store && rinfo >= 0) {
                const unsigned row = (unsigned)(rinfo & (kRowSrcBit - 1));
                if (rinfo & kRowSrcBit) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(Db + (size_t)row * (ROWB / 4)) + li);
                    a0 += b.x; a1 += b.y; a2 += b.z;
                }
#ifdef GLB_EXPERIMENT
                if (!(poll_sleep & 0x80000000u) || a0 == 123.456f)
#endif
                st_chunk(out + (size_t)row * ROWB, a0, a1, a2, 2u + (unsigned)t);
            }
            a0 = a1 = a2 = 0.f;
            ++k;
            if (k < nsw) { sl = s_slot[k * NW + warp]; left = (sl.y + 1) >> 1; }
        };
        while (k < nsw && left == 0) finish();               // slots of empty rows
#define GLB_DFP_STEP(V, X)                                                                                              \
        {                                                                                                               \
            if (p >= n_pairs) break;                                                                                    \
            const int4 *cv = stream + (size_t)p * RPW;                                                                  \
            dfp_consume(cv, in, expect, V, X, a0, a1, a2, n_poll, n_badbatch, watchdog, poll_sleep);                    \
            if (p + 4 < n_pairs) dfp_issue<RPW, L1F>(cv + 4 * RPW, in, V, X);                                           \
            ++p;                                                                                                        \
            if (--left == 0) { finish(); while (k < nsw && left == 0) finish(); }                                       \
        }
        for (;;) {
            GLB_DFP_STEP(v0, x0)
            GLB_DFP_STEP(v1, x1)
            GLB_DFP_STEP(v2, x2)
            GLB_DFP_STEP(v3, x3)
        }
#undef GLB_DFP_STEP
#ifdef GLB_EXPERIMENT
        if (stats) busy += clock64() - tb0;
#endif
    }
    if (stats) {
        atomicAdd(stats + 0, (unsigned long long)n_poll);
        atomicAdd(stats + 1, (unsigned long long)n_badbatch);
        if (lane == 0) atomicMax(stats + 2, (unsigned long long)(clock64() - clk0));
        if (threadIdx.x == 0) atomicAdd(stats + 3, 1ull);
#ifdef GLB_EXPERIMENT
        if (lane == 0) {                                                 // busy cycles: sum and max over warps, max over CTAs of the CTA's mean
            atomicAdd(stats + 4, (unsigned long long)busy);
            atomicMax(stats + 5, (unsigned long long)busy);
            atomicAdd(stats + 6, 1ull);
        }
        __syncthreads();
        {
            __shared__ unsigned long long cta_sum, cta_max;
            if (threadIdx.x == 0) { cta_sum = 0ull; cta_max = 0ull; }
            __syncthreads();
            if (lane == 0) { atomicAdd(&cta_sum, (unsigned long long)busy); atomicMax(&cta_max, (unsigned long long)busy); }
            __syncthreads();
            if (threadIdx.x == 0) { atomicMax(stats + 7, cta_sum / NW); atomicAdd(stats + 8, cta_max); }
        }
        if (lane == 0) {                                                 // per-warp record behind the 16 summary words: busy, pairs, slots
            unsigned long long *rec = stats + 16 + ((size_t)blockIdx.x * NW + warp) * 3;
            rec[0] = (unsigned long long)busy; rec[1] = (unsigned long long)n_pairs; rec[2] = (unsigned long long)nsw;
        }
#endif
    }
}

// epoch words of the ring before a launch: version 0 in buffer 0 (word 1), nothing valid in the others (word 0: whatever
// an earlier launch left there must not pass for a version of this one)
// chunks nchunks .. nchunks + nscratch - 1 (the scratch rows): value 0, epoch 0xffffffff in every buffer
__global__ void __launch_bounds__(256) stamp_kernel(DfRing ring, long long nchunks, int nscratch)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nchunks + nscratch; i += (long long)gridDim.x * blockDim.x) {
        if (i < nchunks) {
            reinterpret_cast<unsigned *>(ring.b[0])[i * 4 + 3] = 1u;
#pragma unroll
            for (int q = 1; q < kRing; ++q) reinterpret_cast<unsigned *>(ring.b[q])[i * 4 + 3] = 0u;
        } else {
#pragma unroll
            for (int q = 0; q < kRing; ++q) reinterpret_cast<uint4 *>(ring.b[q])[i] = make_uint4(0u, 0u, 0u, 0xffffffffu);
        }
    }
}

// label matrix <-> device layout.  dst (n x ldu fp32) <- scale_row * src (n x c fp64); device row r = caller's
// row perm[r].  FLAGGED: column k lives at float (k / 3) * 4 + k % 3 of the row.
template <bool FLAGGED>
__global__ void __launch_bounds__(256)
pack_kernel(const double *__restrict__ src, const double *__restrict__ inv_scale, long long n, int c,
            float *__restrict__ dst, int ldu, const int *__restrict__ perm)
{
    const long long total = n * ldu;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ldu;
        const int p = (int)(i - r * ldu);
        const int k = FLAGGED ? ((p & 3) == 3 ? c : (p >> 2) * 3 + (p & 3)) : p;
        const long long sr = perm ? perm[r] : r;
        float v = 0.f;
        if (k < c) {
            const double s = src[sr * c + k];
            v = (float)(inv_scale ? (1.0 / inv_scale[sr]) * s : s);           // D^-1 * source, ssl.py:636
        }
        dst[i] = v;
    }
}

template <bool FLAGGED>
__global__ void __launch_bounds__(256)
unpack_kernel(const float *__restrict__ src, long long n, int c, int ldu, double *__restrict__ dst,
              const int *__restrict__ perm)
{
    const long long total = n * c;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c;
        const int k = (int)(i - r * c);
        const long long dr = perm ? perm[r] : r;
        const int p = FLAGGED ? (k / 3) * 4 + k % 3 : k;
        dst[dr * c + k] = (double)src[r * ldu + p];
    }
}

// ------------------------------------------------------------------------------------------------
// mixing vector v <- RW v (fp64) and max|v - vinf|     (ssl.py:667,669)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_max_to_global(double m, unsigned long long *out)
{
    // non-negative doubles (and NaN, which sorts above +inf) compare like their bit patterns
    for (int off = 16; off > 0; off >>= 1) {
        double o = __shfl_xor_sync(0xffffffffu, m, off);
        m = (__double_as_longlong(o) > __double_as_longlong(m)) ? o : m;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

__global__ void __launch_bounds__(256)
mixing_step_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                   const double *__restrict__ vinf, const double *__restrict__ v_in, double *__restrict__ v_out,
                   int n, unsigned long long *err_out)
{
    constexpr int G = 8;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngroups = ((long long)gridDim.x * blockDim.x) / G;
    double worst = 0.0;
    for (long long rb = 0; rb < n; rb += ngroups) {          // uniform trip count: shuffles stay converged
        const long long row = rb + gid;
        double s = 0.0;
        if (row < n) {
            const int beg = rowptr[row], end = rowptr[row + 1];
            for (int j = beg + gl; j < end; j += G) s += val[j] * v_in[col[j]];
        }
        for (int off = G / 2; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off, G);
        if (row < n && gl == 0) {
            v_out[row] = s;
            const double d = fabs(s - vinf[row]);
            // fabs(NaN) is NaN: keep it (np.max propagates NaN, ssl.py:667)
            worst = (d > worst || d != d) ? d : worst;
        }
    }
    block_max_to_global(worst, err_out);
}

// The same recurrence, `steps` steps in ONE cooperative launch: a grid barrier between steps instead of a kernel boundary
// (a step is 0.56 MB of gathers at 70k nodes - a launch per step costs more than the step).  Same per-row arithmetic in the
// same order as mixing_step_kernel (bit-identical v), v ping-pongs between v0 (even steps read it) and v1; err_out[s] = max
// |v_{s+1} - vinf|.  v is read with ld.global.cg: L1 may hold the lines of two steps ago.
// A step is a chain of dependent L2 round trips (row pointers -> entries -> v), so a lane group works on R = 8 rows at
// once, the entries of a row's next chunk are fetched while the gathers of the current one are in flight, and when the
// whole matrix is one pass of the grid (HOIST: n <= groups x R, 75 776 rows on 148 SMs) the row pointers and every row's
// first chunk of entries stay in registers across the steps: what is left per step is one gather round trip per 8
// nonzeros of a row plus the barrier (12.4 -> 4 us per step at 70k nodes).
template <bool HOIST>
__global__ void __launch_bounds__(512, 1)
mixing_persistent_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, const double *__restrict__ val,
                         const double *__restrict__ vinf, double *v0, double *v1, int n, int steps, unsigned long long *err_out,
                         unsigned *sync_words)
{
    constexpr int G = 8, R = 8;
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long ngroups = ((long long)gridDim.x * blockDim.x) / G;
    int beg[R], end[R], c0[R];
    double a0[R];
    for (int st = 0; st < steps; ++st) {
        const double *v_in = (st & 1) ? v1 : v0;
        double *v_out = (st & 1) ? v0 : v1;
        double worst = 0.0;
        for (long long rb = 0; rb < n; rb += ngroups * R) {          // uniform trip count: shuffles stay converged
            if (!HOIST || st == 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const long long row = rb + gid + r * ngroups;
                    beg[r] = end[r] = 0;
                    if (row < n) { beg[r] = rowptr[row] + gl; end[r] = rowptr[row + 1]; }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    a0[r] = 0.0; c0[r] = 0;
                    if (beg[r] < end[r]) { a0[r] = val[beg[r]]; c0[r] = col[beg[r]]; }
                }
            }
            double s[R], a[R];
            int j[R], cc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { s[r] = 0.0; j[r] = beg[r]; a[r] = a0[r]; cc[r] = c0[r]; }
            bool more;
            do {
                double x[R];
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (j[r] < end[r]) x[r] = __ldcg(v_in + cc[r]);
                more = false;
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (j[r] < end[r]) {
                        const int jn = j[r] + G;
                        double an = 0.0;
                        int cn = 0;
                        if (jn < end[r]) { an = val[jn]; cn = col[jn]; more = true; }
                        s[r] += a[r] * x[r];
                        j[r] = jn; a[r] = an; cc[r] = cn;
                    }
            } while (more);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double t = s[r];
                for (int off = G / 2; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off, G);
                const long long row = rb + gid + r * ngroups;
                if (row < n && gl == 0) {
                    v_out[row] = t;
                    const double d = fabs(t - vinf[row]);
                    worst = (d > worst || d != d) ? d : worst;
                }
            }
        }
        block_max_to_global(worst, err_out + st);
        if (st + 1 < steps) grid_barrier<false>(sync_words, (unsigned)(st + 1));
    }
}

__global__ void __launch_bounds__(256)
maxdiff_kernel(const double *__restrict__ v, const double *__restrict__ vinf, int n, unsigned long long *err_out)
{
    double worst = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double d = fabs(v[i] - vinf[i]);
        worst = (d > worst || d != d) ? d : worst;
    }
    block_max_to_global(worst, err_out);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// is (j,i) stored for every stored (i,j)?  (explicit zeros count: only the pattern matters; rows need not be sorted)
// One thread per stored entry scans row j for column i: nnz x (row length) steps, a fraction of a millisecond at 10^6
// nonzeros - the host version of round 1 took 30-40 ms on a relabelled (unsorted) matrix.
__global__ void __launch_bounds__(256)
pattern_symmetric_kernel(const int *__restrict__ rowptr, const int *__restrict__ col, int n, int *__restrict__ asym)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp; i < n; i += nwarps) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        for (int e = beg + lane; e < end; e += 32) {
            const int j = col[e];
            bool found = false;
            if (j >= 0 && j < n)
                for (int q = rowptr[j]; q < rowptr[j + 1]; ++q) found |= col[q] == (int)i;
            if (!found) *asym = 1;
        }
    }
}

template <int LANES, bool FLAGGED>
static int launch_step(const int *rp, const int *col, const float *val, const float *Db, const float *u_in,
                       float *u_out, int64_t n, int ldu, cudaStream_t st)
{
    const int threads = 256;
    const int64_t rows_per_block = threads / LANES;
    int64_t blocks = (n + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)sm_count() * 8 * 8;           // 8 resident CTAs/SM x 8 waves, then grid-stride
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    poisson_step_kernel<LANES, FLAGGED><<<(unsigned)blocks, threads, 0, st>>>(rp, col, val, Db, u_in, u_out, (int)n, ldu);
    return 0;
}

template <bool FLAGGED>
static int dispatch_step(const int *rp, const int *col, const float *val, const float *Db, const float *u_in,
                         float *u_out, int64_t n, int ldu, cudaStream_t st)
{
    switch (ldu) {
        case 4: return launch_step<1, FLAGGED>(rp, col, val, Db, u_in, u_out, n, ldu, st);
        case 8: return launch_step<2, FLAGGED>(rp, col, val, Db, u_in, u_out, n, ldu, st);
        case 16: return launch_step<4, FLAGGED>(rp, col, val, Db, u_in, u_out, n, ldu, st);
        case 32: return launch_step<8, FLAGGED>(rp, col, val, Db, u_in, u_out, n, ldu, st);
        case 64: return launch_step<16, FLAGGED>(rp, col, val, Db, u_in, u_out, n, ldu, st);
        default: return launch_step<32, FLAGGED>(rp, col, val, Db, u_in, u_out, n, ldu, st);
    }
}

static inline int float_bits(float f) { int i; memcpy(&i, &f, sizeof(i)); return i; }

static int flagged_lanes(int c)           // lanes per row of the flagged layout: 3 values per 16-byte chunk
{
    int lanes = 1;
    while (lanes * 3 < c && lanes < 32) lanes <<= 1;
    return lanes * 3 >= c ? lanes : 0;    // 0: too wide (c > 96)
}

static int stream_blocks_p(int64_t work, int threads = 256)
{
    int64_t b = (work + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace glb

using namespace glb;

struct glb_poisson_plan {
    int64_t n = 0, nnz = 0;
    int c = 0, ldu = 0, kind = GLB_POISSON_KIND_STEP;
    const int *d_rowptr = nullptr, *d_col = nullptr;       // caller-owned CSR of P
    const float *d_val = nullptr;
    int grid = 0, threads = 0;
    const void *fn = nullptr;
    size_t smem_bytes = 0;
    // barrier kernel
    int max_rows = 0, slab_cap = 0;
    unsigned *d_counter = nullptr;      // barrier words: kFlagStride * grid unsigned
    int *d_cta_rows = nullptr;          // grid + 1 row boundaries of the work-balanced partition
    // dataflow kernel: sliced-ELL slabs of every CTA, back to back
    int2 *d_slabs = nullptr;
    int4 *d_slots = nullptr;
    long long *d_slab_off = nullptr;
    int *d_slot_off = nullptr, *d_slot_rows = nullptr;
    int cap_entries = 0, cap_slots = 0, cap_parts = 0;
    double ell_fill = 0.0;              // nnz / stored entries of the slabs
    float tuned_ms[2] = {0.f, 0.f};     // AUTO: measured ms of the trial run, {dataflow, barrier}
    int tuned_gate = 32;
    float *d_ring = nullptr;            // dataflow kernel: buffers 2 and 3 of the version ring (0 and 1 are the caller's u0/u1)
    bool l1_first = true;               // dataflow kernel: first attempt of a gather through L1
    unsigned *d_gate = nullptr;         // dataflow kernel: start gate counter
    int gate_every = 32;                // dataflow kernel: iterations between re-alignment gates (tuned at plan time)
    unsigned poll_sleep = 0;            // dataflow kernel: nanoseconds a warp sleeps between re-polls of a stale batch
    bool has_long_rows = false;         // rows longer than a slice batch exist (dealt over whole warps)
    int scratch_row = 0;                // > 0: the label matrices carry that many rows behind row n-1, owned by the library
                                        // (padding targets of the slabs)
    unsigned long long *d_stats = nullptr;   // -DGLB_EXPERIMENT, GLB_POISSON_STATS=1: {re-polls, polling batches, max warp cycles, CTAs}
};

// Experiment switches exist only in -DGLB_EXPERIMENT builds (libglb200_exp.so, python -m graphlearning_b200.build --exp)
// and are read ONCE, when a plan is created; the product library has no environment lookups on the iterate path.
static int exp_env(const char *name, int def)
{
#ifdef GLB_EXPERIMENT
    const char *e = getenv(name);
    if (e) return atoi(e);
#else
    (void)name;
#endif
    return def;
}

template <int LANES>
static const void *barrier_fn(int *threads)
{
    *threads = 1024;
    return (const void *)poisson_persistent_kernel<LANES, 1024, 4, false>;
}

static const void *pick_barrier(int ldu, int *threads)
{
    switch (ldu) {
        case 4: return barrier_fn<1>(threads);
        case 8: return barrier_fn<2>(threads);
        case 16: return barrier_fn<4>(threads);
        case 32: return barrier_fn<8>(threads);
        case 64: return barrier_fn<16>(threads);
        default: return barrier_fn<32>(threads);
    }
}

// 512 threads x 8 gathers in flight per lane (16-entry batches spill at the 128-register cap, r1 visit 4)
template <int LANES>
static const void *dataflow_fn(bool l1_first, int *threads)
{
#ifdef GLB_EXPERIMENT
    // A/B switches of the experiment build: CTA size (512 measured best, profiles/r2_dataflow_pair_stream_ab.txt; 256 threads
    // with the same ring of four pairs: 10.6 us) and the
    // first gather attempt through L1 (on: 6.4 us, off: 6.9-10 us per iteration)
    const int pt = exp_env("GLB_POISSON_THREADS", 512);
    if (pt == 768) { *threads = 768; return l1_first ? (const void *)poisson_dataflow_kernel<LANES, 768, true> : (const void *)poisson_dataflow_kernel<LANES, 768, false>; }
    if (pt == 1024) { *threads = 1024; return l1_first ? (const void *)poisson_dataflow_kernel<LANES, 1024, true> : (const void *)poisson_dataflow_kernel<LANES, 1024, false>; }
    if (!l1_first) { *threads = 512; return (const void *)poisson_dataflow_kernel<LANES, 512, false>; }
#else
    (void)l1_first;
#endif
    *threads = 512;
    return (const void *)poisson_dataflow_kernel<LANES, 512, true>;
}

static const void *pick_dataflow(int lanes, bool l1_first, int *threads)
{
    switch (lanes) {
        case 1: return dataflow_fn<1>(l1_first, threads);
        case 2: return dataflow_fn<2>(l1_first, threads);
        case 4: return dataflow_fn<4>(l1_first, threads);
        case 8: return dataflow_fn<8>(l1_first, threads);
        case 16: return dataflow_fn<16>(l1_first, threads);
        default: return dataflow_fn<32>(l1_first, threads);
    }
}

// work-balanced contiguous row partition: cost(row) = nnz(row) + extra
static void balanced_bounds(const std::vector<int> &h_rp, int64_t n, int grid, double extra, std::vector<int> &bounds)
{
    bounds.assign((size_t)grid + 1, 0);
    const double total = (double)h_rp[n] + extra * (double)n;
    int b = 1;
    for (int64_t i = 0; i < n && b < grid; ++i) {
        const double pref = (double)h_rp[i + 1] + extra * (double)(i + 1);
        while (b < grid && pref >= total * b / grid) bounds[b++] = (int)(i + 1);
    }
    for (; b <= grid; ++b) bounds[b] = (int)n;
    bounds[grid] = (int)n;
}

// Sliced-ELL slabs of the dataflow kernel, built on the host from the CSR arrays of P (host copies).
struct DfSlabs {
    std::vector<int2> slab;            // entries of all CTAs back to back: (byte offset of the gathered row, value bits)
    std::vector<int4> slots;           // (first entry relative to the CTA's slab, width, type | parts << 8, partial index)
    std::vector<long long> slab_off;   // grid + 1
    std::vector<int> slot_off, slot_rows, bounds;
    int cap_entries = 0, cap_slots = 0, cap_parts = 0;
    long long wavefronts = 0, steps = 0;
    bool has_long_rows = false;
};

static void build_dataflow_slabs(const std::vector<int> &h_rp, const int *h_col, const float *h_val, int64_t n, int lanes,
                                 int grid, int nw, bool l1_first, DfSlabs &S)
{
    const int rowb = lanes * 16, rpw = 32 / lanes;
    const int64_t nnz = h_rp[n];
    std::vector<int> bounds;
    balanced_bounds(h_rp, n, grid, 2.0, bounds);
    const int part_max = kLongRowDf * rpw;            // nonzeros of one warp-wide piece of a long row (2 batches per lane group)
    struct Slot { int type, nparts, pbuf, L, cost; int rows[32]; int nz0, nz1; };
    struct CtaOut {
        std::vector<int2> slab;
        std::vector<int4> slots;
        std::vector<int> slot_rows;
        int nparts = 0, depth = 0;
        long long wavefronts = 0, steps = 0;
        bool has_long = false;
    };
    // Slabs of CTA b: short rows sorted by length, rpw per slot (few holes); long rows dealt over whole warps.
    auto build_cta = [&](int b, CtaOut &out) {
        const int r0 = bounds[b], r1 = bounds[b + 1];
        std::vector<Slot> cta_slots;
        std::vector<int> order;
        unsigned pad_next = (unsigned)b * 131u;
        auto pad_off = [&]() -> unsigned {
            // padding of the long-row slots gathers a scratch row behind row n-1 (value 0, always ready).  Through L1 one
            // row per CTA (it stays resident in that SM's L1); at L2 they are dealt round robin over the scratch rows.
            const unsigned r = l1_first ? (unsigned)(b % kScratchRows) : (pad_next++ % kScratchRows);
            return (unsigned)(n + r) * (unsigned)rowb;
        };
        for (int r = r0; r < r1; ++r) if (h_rp[r + 1] - h_rp[r] <= kLongRowDf) order.push_back(r);
        std::stable_sort(order.begin(), order.end(), [&](int a, int c2) { return h_rp[a + 1] - h_rp[a] > h_rp[c2 + 1] - h_rp[c2]; });
        for (size_t k0 = 0; k0 < order.size(); k0 += rpw) {
            cta_slots.emplace_back();
            Slot &sl = cta_slots.back();
            sl.type = kSlotNormal; sl.nparts = 0; sl.pbuf = 0; sl.nz0 = sl.nz1 = 0;
            for (int g = 0; g < 32; ++g) sl.rows[g] = g < rpw && k0 + g < order.size() ? order[k0 + g] : -1;
            sl.L = h_rp[order[k0] + 1] - h_rp[order[k0]];
            for (int g = 0; g < rpw; ++g) if (sl.rows[g] >= 0) out.wavefronts += h_rp[sl.rows[g] + 1] - h_rp[sl.rows[g]];
            out.steps += sl.L;
            sl.cost = (sl.L + 1) / 2 + 3;                    // measured: 237 cycles per pair + 840 per slot (store, bookkeeping)
        }
        // long rows: one warp-wide slot per piece of at most part_max nonzeros
        int nparts_cta = 0;
        for (int r = r0; r < r1; ++r) {
            const int len = h_rp[r + 1] - h_rp[r];
            if (len <= kLongRowDf) continue;
            out.has_long = true;
            const int m = (len + part_max - 1) / part_max;
            const int chunk = (len + m - 1) / m;
            for (int q = 0; q < m; ++q) {
                cta_slots.emplace_back();
                Slot &sl = cta_slots.back();
                sl.nparts = 0; sl.pbuf = 0;
                sl.nz0 = h_rp[r] + q * chunk;
                sl.nz1 = std::min(h_rp[r + 1], sl.nz0 + chunk);
                sl.L = (sl.nz1 - sl.nz0 + rpw - 1) / rpw;
                for (int g = 0; g < 32; ++g) sl.rows[g] = -1;
                if (q == 0) {
                    sl.type = m == 1 ? kSlotLong : kSlotOwner;
                    sl.nparts = m - 1;
                    sl.pbuf = nparts_cta;                 // partial sums nparts_cta .. nparts_cta + m - 2
                    sl.rows[0] = r;
                } else {
                    sl.type = kSlotPart;
                    sl.pbuf = nparts_cta + q - 1;
                }
                sl.cost = (sl.L + 1) / 2 + 3 + (q == 0 ? m - 1 : 0);
                out.wavefronts += sl.nz1 - sl.nz0;
                out.steps += sl.L;
            }
            nparts_cta += m - 1;
        }
        // deal the slots to the warps: PART pieces first, then OWNER/LONG, then NORMAL (a warp never waits in shared
        // memory for a piece it has not produced yet itself); inside a class longest first to the least loaded warp
        std::vector<std::vector<int>> per_warp((size_t)nw);
        std::vector<long long> load((size_t)nw, 0ll);
        std::vector<int> idx(cta_slots.size());
        for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
        auto cls = [&](int i) { const int t = cta_slots[i].type; return t == kSlotPart ? 0 : (t == kSlotNormal ? 2 : 1); };
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int c2) {
            if (cls(a) != cls(c2)) return cls(a) < cls(c2);
            return cta_slots[a].cost > cta_slots[c2].cost;
        });
        // (warps of their own for the long rows were measured slower than mixing them: 40 vs 23 us/iteration on the
        // 128-d graph, r1q probe under profiles/)
        for (int i : idx) {
            int w = 0;
            for (int q = 1; q < nw; ++q) if (load[q] < load[w]) w = q;
            per_warp[w].push_back(i);
            load[w] += cta_slots[i].cost;
        }
        size_t depth = 0;
        for (auto &v : per_warp) depth = std::max(depth, v.size());
        // The entries of one warp's slots are stored back to back (the kernel walks them as one stream of pairs);
        // the slot table is interleaved: slot k of warp w at k * nw + w.
        std::vector<int> slot_first(depth * (size_t)nw, 0);
        std::vector<int2> &slab = out.slab;
        for (int w = 0; w < nw; ++w)
            for (size_t k = 0; k < per_warp[w].size(); ++k) {
                const Slot &sl = cta_slots[per_warp[w][k]];
                slot_first[k * nw + w] = (int)slab.size();
                // entries in pairs: entry j of lane group g at (j/2 * rpw + g) * 2 + (j & 1); width rounded up to even
                const int Lst = (sl.L + 1) & ~1;
                const size_t s0 = slab.size();
                slab.resize(s0 + (size_t)Lst * rpw);
                for (size_t q = s0; q < slab.size(); ++q) slab[q] = make_int2((int)pad_off(), 0);
                for (int j = 0; j < sl.L; ++j)
                    for (int g = 0; g < rpw; ++g) {
                        int q = -1;
                        if (sl.type == kSlotNormal) {
                            const int r = sl.rows[g];
                            if (r >= 0 && j < h_rp[r + 1] - h_rp[r]) q = h_rp[r] + j;      // entries in CSR order: same sums as the step kernel
                        } else {
                            const int qq = sl.nz0 + j * rpw + g;          // round-robin over the lane groups
                            if (qq < sl.nz1) q = qq;
                        }
                        if (q < 0) continue;
                        slab[s0 + ((size_t)(j / 2) * rpw + g) * 2 + (j & 1)] = make_int2((int)((unsigned)h_col[q] * (unsigned)rowb), float_bits(h_val[q]));
                    }
            }
        for (size_t k = 0; k < depth; ++k)
            for (int w = 0; w < nw; ++w) {
                if (k >= per_warp[w].size()) {            // empty slot: nothing to gather, nothing to store
                    out.slots.push_back(make_int4(0, 0, kSlotNormal, 0));
                    for (int g = 0; g < rpw; ++g) out.slot_rows.push_back(-1);
                    continue;
                }
                const Slot &sl = cta_slots[per_warp[w][k]];
                out.slots.push_back(make_int4(slot_first[k * nw + w], sl.L, sl.type | (sl.nparts << 8), sl.pbuf));
                for (int g = 0; g < rpw; ++g) out.slot_rows.push_back(sl.rows[g]);
            }
        out.nparts = nparts_cta;
        out.depth = (int)depth;
    };
    std::vector<CtaOut> ctas((size_t)grid);
    {
        const int nthreads = std::max(1, std::min({8, (int)std::thread::hardware_concurrency(), grid}));
        std::atomic<int> next{0};
        auto worker = [&]() { for (int b; (b = next.fetch_add(1)) < grid;) build_cta(b, ctas[b]); };
        std::vector<std::thread> pool;
        for (int t = 1; t < nthreads; ++t) pool.emplace_back(worker);
        worker();
        for (auto &t : pool) t.join();
    }
    std::vector<int2> &slab = S.slab;
    std::vector<int4> &slots = S.slots;
    std::vector<long long> &slab_off = S.slab_off;
    std::vector<int> &slot_off = S.slot_off, &slot_rows = S.slot_rows;
    slab.clear(); slots.clear(); slot_rows.clear();
    slab_off.assign((size_t)grid + 1, 0);
    slot_off.assign((size_t)grid + 1, 0);
    slab.reserve((size_t)nnz + (size_t)nnz / 2 + 1024);
    int cap_entries = 0, cap_slots = 0, cap_parts = 0;
    long long wavefronts = 0, steps = 0;
    S.has_long_rows = false;
    for (int b = 0; b < grid; ++b) {
        const CtaOut &o = ctas[b];
        slab.insert(slab.end(), o.slab.begin(), o.slab.end());
        slots.insert(slots.end(), o.slots.begin(), o.slots.end());
        slot_rows.insert(slot_rows.end(), o.slot_rows.begin(), o.slot_rows.end());
        slab_off[b + 1] = (long long)slab.size();
        slot_off[b + 1] = (int)slots.size();
        cap_entries = std::max(cap_entries, (int)o.slab.size());
        cap_slots = std::max(cap_slots, o.depth * nw);
        cap_parts = std::max(cap_parts, o.nparts);
        wavefronts += o.wavefronts; steps += o.steps;
        if (o.has_long) S.has_long_rows = true;
    }
    S.wavefronts = wavefronts; S.steps = steps;
    S.cap_entries = cap_entries; S.cap_slots = cap_slots; S.cap_parts = cap_parts;
    S.bounds = bounds;
}

// Try to set the plan up for the dataflow kernel.  Returns 0 and leaves kind untouched when the graph does not
// qualify (pattern not symmetric, slabs too large for shared memory, label rows too wide).
static int plan_try_dataflow(glb_poisson_plan *p, const std::vector<int> &h_rp, int sms, int max_smem, cudaStream_t st)
{
    const int64_t n = p->n, nnz = p->nnz;
    const int lanes = flagged_lanes(p->c);
    if (!lanes) return 0;
    const int rowb = lanes * 16, rpw = 32 / lanes;
    if ((double)(n + kScratchRows) * rowb >= 4294967295.0 || n >= kRowSrcBit) return 0;
    p->l1_first = exp_env("GLB_POISSON_L1", 1) != 0;
    p->poll_sleep = (unsigned)exp_env("GLB_POISSON_SLEEP", 0) & 0xffffu;
    p->poll_sleep |= (unsigned)exp_env("GLB_POISSON_FREE", 0) << 30;        // ceiling probes, -DGLB_EXPERIMENT only
    int grid = (int)((n + 63) / 64);                  // tiny graphs: at least ~64 rows per CTA
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    if ((double)nnz * 8.0 / grid > (double)max_smem) return 0;

    PhaseTimer tm("dataflow plan");
    std::vector<int> h_col((size_t)nnz);
    std::vector<float> h_val((size_t)nnz);
    int h_asym = 0;
    {
        int *d_asym = nullptr;
        GLB_CUDA(dev_alloc(&d_asym, sizeof(int)));
        cudaMemsetAsync(d_asym, 0, sizeof(int), st);
        if (nnz) pattern_symmetric_kernel<<<sms * 8, 256, 0, st>>>(p->d_rowptr, p->d_col, (int)n, d_asym);
        cudaMemcpyAsync(&h_asym, d_asym, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (nnz) {
            cudaMemcpyAsync(h_col.data(), p->d_col, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(h_val.data(), p->d_val, sizeof(float) * (size_t)nnz, cudaMemcpyDeviceToHost, st);
        }
        const cudaError_t e = cudaStreamSynchronize(st);
        dev_free(d_asym);
        if (e != cudaSuccess) { set_error("plan_try_dataflow: %s", cudaGetErrorString(e)); return (int)e; }
    }
    tm.lap("symmetry check (device) + download col/val");
    if (h_asym) return 0;

    int threads = 0;
    const void *fn = pick_dataflow(lanes, p->l1_first, &threads);
    const int nw = threads / 32;
    DfSlabs S;
    build_dataflow_slabs(h_rp, h_col.data(), h_val.data(), n, lanes, grid, nw, p->l1_first, S);
    std::vector<int2> &slab = S.slab;
    std::vector<int4> &slots = S.slots;
    std::vector<long long> &slab_off = S.slab_off;
    std::vector<int> &slot_off = S.slot_off, &slot_rows = S.slot_rows;
    int cap_entries = S.cap_entries;
    const int cap_slots = S.cap_slots, cap_parts = S.cap_parts;
    p->has_long_rows = S.has_long_rows;
    tm.lap("slab build");
    cap_entries = (cap_entries + 1) & ~1;
    const size_t smem = (size_t)cap_entries * 8 + (size_t)cap_slots * 16 + (size_t)cap_parts * 2 * lanes * 16 + (size_t)cap_slots * rpw * 4;
    if (smem > (size_t)max_smem) return 0;
    GLB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1 || grid > per_sm * sms) return 0;

    const size_t nslab = slab.size() ? slab.size() : 1, nsl = slots.size() ? slots.size() : 1;
    GLB_CUDA(dev_alloc(&p->d_slabs, sizeof(int2) * nslab));
    GLB_CUDA(dev_alloc(&p->d_slots, sizeof(int4) * nsl));
    GLB_CUDA(dev_alloc(&p->d_slab_off, sizeof(long long) * (grid + 1)));
    GLB_CUDA(dev_alloc(&p->d_slot_off, sizeof(int) * (grid + 1)));
    GLB_CUDA(dev_alloc(&p->d_slot_rows, sizeof(int) * nsl * rpw));
    GLB_CUDA(cudaMemcpyAsync(p->d_slabs, slab.data(), sizeof(int2) * slab.size(), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(p->d_slots, slots.data(), sizeof(int4) * slots.size(), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(p->d_slab_off, slab_off.data(), sizeof(long long) * (grid + 1), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(p->d_slot_off, slot_off.data(), sizeof(int) * (grid + 1), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(p->d_slot_rows, slot_rows.data(), sizeof(int) * slot_rows.size(), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaStreamSynchronize(st));            // the staging vectors are locals
    tm.lap("upload slabs");
    p->kind = GLB_POISSON_KIND_DATAFLOW;
    p->ldu = lanes * 4;
    p->grid = grid; p->threads = threads; p->fn = fn; p->smem_bytes = smem;
    p->cap_entries = cap_entries; p->cap_slots = cap_slots; p->cap_parts = cap_parts;
    p->scratch_row = kScratchRows;
    GLB_CUDA(dev_alloc(&p->d_ring, sizeof(float) * 2 * ((size_t)n + kScratchRows) * (size_t)(lanes * 4)));
    p->ell_fill = slab.size() ? (double)nnz / (double)slab.size() : 1.0;
    GLB_CUDA(dev_alloc(&p->d_gate, 2 * sizeof(unsigned)));          // [0] gate counter, [1] watchdog flag
    GLB_CUDA(cudaMemsetAsync(p->d_gate, 0, 2 * sizeof(unsigned), st));
    if (exp_env("GLB_POISSON_STATS", 0)) GLB_CUDA(dev_alloc(&p->d_stats, (16 + (size_t)grid * 32 * 3) * sizeof(unsigned long long)));
    return 0;
}

static int plan_try_barrier(glb_poisson_plan *p, const std::vector<int> &h_rp, int sms, int max_smem, cudaStream_t st)
{
    const int64_t n = p->n, nnz = p->nnz;
    const int ldu = glb_padded_ld(p->c);
    // Upper bound for a graph that could fit the shared-memory slabs at all (8 bytes per nonzero per CTA).
    if ((double)nnz * 8.0 / sms > (double)max_smem) return 0;
    int grid = (int)((n + 127) / 128);          // tiny graphs: at least ~128 rows per CTA
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    std::vector<int> bounds;
    balanced_bounds(h_rp, n, grid, 4.0, bounds);
    int cap = 0, max_rows = 0;
    for (int b = 0; b < grid; ++b) {
        cap = std::max(cap, h_rp[bounds[b + 1]] - h_rp[bounds[b]]);
        max_rows = std::max(max_rows, bounds[b + 1] - bounds[b]);
    }
    cap = (cap + 3) & ~3;
    const size_t smem = (size_t)cap * 8 + (size_t)(2 * max_rows + 2) * 4 + (size_t)((max_rows + 15) & ~15) + 16;
    if (smem > (size_t)max_smem) return 0;
    int threads = 0;
    const void *fn = pick_barrier(ldu, &threads);
    GLB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1 || grid > per_sm * sms || grid > threads) return 0;
    GLB_CUDA(dev_alloc(&p->d_counter, sizeof(unsigned) * kFlagStride * grid));
    GLB_CUDA(dev_alloc(&p->d_cta_rows, sizeof(int) * (grid + 1)));
    GLB_CUDA(cudaMemcpyAsync(p->d_cta_rows, bounds.data(), sizeof(int) * (grid + 1), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaStreamSynchronize(st));        // bounds is a local
    p->kind = GLB_POISSON_KIND_BARRIER;
    p->ldu = ldu;
    p->grid = grid; p->max_rows = max_rows; p->slab_cap = cap; p->smem_bytes = smem;
    p->threads = threads; p->fn = fn;
    return 0;
}

// device time of a short trial run of the plan's kernel on an all-zero problem (second of two launches)
// rows x ldu floats a trial run needs for Db, u0, u1
static size_t plan_time_floats(const glb_poisson_plan *p) { return 3 * ((size_t)p->n + (size_t)p->scratch_row) * p->ldu; }

// Trial run on zeros: T iterations twice, the second one timed.  `buf` (plan_time_floats(p) floats) may be shared by
// several trials; nullptr = allocate here.
static int plan_time(glb_poisson_plan *p, float *ms, cudaStream_t st, float *shared_buf = nullptr)
{
    const int T = 32;
    const size_t rows = (size_t)p->n + (size_t)p->scratch_row;             // label matrices carry the scratch rows
    const size_t bytes = rows * p->ldu * sizeof(float);
    float *buf = shared_buf;
    if (!buf) GLB_CUDA(dev_alloc(&buf, 3 * bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = 0;
    cudaError_t ce = cudaMemsetAsync(buf, 0, 3 * bytes, st);
    for (int rep = 0; rep < 2 && rc == 0 && ce == cudaSuccess; ++rep) {
        cudaEventRecord(e0, st);
        rc = glb_poisson_iterate(p, buf, buf + rows * p->ldu, buf + 2 * rows * p->ldu, T, nullptr, nullptr, st);
        cudaEventRecord(e1, st);
    }
    if (ce == cudaSuccess) ce = cudaEventSynchronize(e1);
    if (ce == cudaSuccess && rc == 0) cudaEventElapsedTime(ms, e0, e1);
    if (ce == cudaSuccess && rc == 0) rc = glb_poisson_plan_check(p, st);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (!shared_buf) dev_free(buf);
    if (rc == 0 && ce != cudaSuccess) { set_error("plan_time: %s", cudaGetErrorString(ce)); rc = (int)ce; }
    return rc;
}

extern "C" GLB_API int glb_poisson_plan_create(glb_poisson_plan **plan, const int32_t *d_rowptr, const int32_t *d_col,
                                               const float *d_val, int64_t n, int64_t nnz, int c, int kind, void *stream)
{
    GLB_CHECK_ARG(plan && d_rowptr && (nnz == 0 || (d_col && d_val)), "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "size out of range");
    GLB_CHECK_ARG(c > 0, "c must be positive");
    GLB_CHECK_ARG(kind >= GLB_POISSON_KIND_AUTO && kind <= GLB_POISSON_KIND_DATAFLOW, "unknown kernel kind");
    GLB_CHECK_ARG((double)n * glb_padded_ld(c) * 4.0 < 4294967296.0, "label matrix larger than 4 GiB: 32-bit row offsets overflow");
    cudaStream_t st = (cudaStream_t)stream;
    glb_poisson_plan *p = new glb_poisson_plan();
    p->n = n; p->nnz = nnz; p->c = c; p->ldu = glb_padded_ld(c);
    p->d_rowptr = d_rowptr; p->d_col = d_col; p->d_val = d_val;
    struct Guard { glb_poisson_plan *p; ~Guard() { if (p) glb_poisson_plan_destroy(p); } } guard{p};

    if (kind == GLB_POISSON_KIND_AUTO) {                        // -DGLB_EXPERIMENT builds only: overrides AUTO
        const int k = exp_env("GLB_POISSON_KIND", GLB_POISSON_KIND_AUTO);
        if (k >= GLB_POISSON_KIND_STEP && k <= GLB_POISSON_KIND_DATAFLOW) kind = k;
    }
    int dev = 0, coop = 0, max_smem = 0;
    GLB_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const int sms = sm_count();
    PhaseTimer tm("plan_create");
    if (kind != GLB_POISSON_KIND_STEP && coop) {
        std::vector<int> h_rp((size_t)n + 1);
        GLB_CUDA(cudaMemcpyAsync(h_rp.data(), d_rowptr, sizeof(int) * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
        GLB_CUDA(cudaStreamSynchronize(st));
        GLB_CHECK_ARG(h_rp[0] == 0 && h_rp[n] == nnz, "rowptr does not match nnz");
        int rc = 0;
        if (kind == GLB_POISSON_KIND_AUTO || kind == GLB_POISSON_KIND_DATAFLOW)
            if ((rc = plan_try_dataflow(p, h_rp, sms, max_smem, st))) return rc;
        tm.lap("dataflow slabs");
        if (p->kind == GLB_POISSON_KIND_STEP && (kind == GLB_POISSON_KIND_AUTO || kind == GLB_POISSON_KIND_BARRIER))
            if ((rc = plan_try_barrier(p, h_rp, sms, max_smem, st))) return rc;
        const bool tune = !exp_env("GLB_POISSON_NOTUNE", 0);
        const int forced_gate = exp_env("GLB_POISSON_GATE_EVERY", -1);
        if (forced_gate >= 0) p->gate_every = p->tuned_gate = forced_gate;
        float ms_df = 0.f;
        if (p->kind == GLB_POISSON_KIND_DATAFLOW && tune && forced_gate < 0) {
            // Measure, don't guess.  Gate period of the dataflow kernel: graphs with hub rows (every hub is a meeting
            // point of hundreds of producers) run best with a gate every few iterations, hub-free graphs with rare gates.
            // (longer periods and no gate at all run 3-4x slower with the pipelined stream: profiles/r2_dataflow_gate_x_backoff_sweep.txt)
            const int cand[2] = {1, 4};
            float best = 0.f;
            float *trial = nullptr;
            GLB_CUDA(dev_alloc(&trial, plan_time_floats(p) * sizeof(float)));
            for (int i = 0; i < 2 && rc == 0; ++i) {
                p->gate_every = cand[i];
                float ms = 0.f;
                rc = plan_time(p, &ms, st, trial);
                if (rc == 0 && (i == 0 || ms < best)) { best = ms; ms_df = ms; p->tuned_gate = cand[i]; }
            }
            dev_free(trial);
            if (rc) return rc;
            p->gate_every = p->tuned_gate;
            p->tuned_ms[0] = ms_df;
            tm.lap("gate tuning");
        }
        // the barrier kernel only wins on small graphs (too few rows per SM to hide the producer-consumer latency)
        if (kind == GLB_POISSON_KIND_AUTO && p->kind == GLB_POISSON_KIND_DATAFLOW && tune && n < 64 * (int64_t)sms) {
            // ... and the dataflow kernel against the barrier kernel (small graphs: too few rows per SM to hide the
            // producer-consumer latency).  Both timed on zeros.
            glb_poisson_plan *alt = nullptr;
            rc = glb_poisson_plan_create(&alt, d_rowptr, d_col, d_val, n, nnz, c, GLB_POISSON_KIND_BARRIER, stream);
            if (rc == 0) {
                float ms_bar = 0.f;
                if (ms_df == 0.f) rc = plan_time(p, &ms_df, st);
                if (rc == 0) rc = plan_time(alt, &ms_bar, st);
                if (rc) { glb_poisson_plan_destroy(alt); return rc; }
                if (ms_bar < ms_df) {              // keep the barrier plan: swap contents, destroy the dataflow one
                    std::swap(*p, *alt);
                }
                p->tuned_ms[0] = ms_df; p->tuned_ms[1] = ms_bar;
                glb_poisson_plan_destroy(alt);
                tm.lap("barrier plan + timing");
            } else if (rc != GLB_E_UNSUPPORTED) {
                return rc;
            }
        }
    }
    if (kind != GLB_POISSON_KIND_AUTO && p->kind != kind) {
        set_error("glb_poisson_plan_create: kernel kind %d is not applicable to this graph", kind);
        return GLB_E_UNSUPPORTED;
    }
    guard.p = nullptr;
    *plan = p;
    return 0;
}

extern "C" GLB_API int glb_poisson_plan_destroy(glb_poisson_plan *plan)
{
    if (!plan) return 0;
    dev_free(plan->d_counter);  dev_free(plan->d_cta_rows);
    dev_free(plan->d_slabs);    dev_free(plan->d_slots);
    dev_free(plan->d_slab_off); dev_free(plan->d_slot_off); dev_free(plan->d_slot_rows); dev_free(plan->d_stats); dev_free(plan->d_gate);
    dev_free(plan->d_ring);
    delete plan;
    return 0;
}

extern "C" GLB_API int glb_poisson_plan_check(glb_poisson_plan *plan, void *stream)
{
    GLB_CHECK_ARG(plan, "null plan");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned flag = 0;
    if (plan->d_gate) {
        GLB_CUDA(cudaMemcpyAsync(&flag, plan->d_gate + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        GLB_CUDA(cudaStreamSynchronize(st));
        if (flag) {
            cudaMemsetAsync(plan->d_gate + 1, 0, sizeof(unsigned), st);
            set_error("glb_poisson_iterate: a polling loop of the dataflow kernel timed out (label matrices smaller than "
                      "glb_poisson_plan_rows x glb_poisson_plan_ld, or a CTA was lost); results are invalid");
            return GLB_E_TIMEOUT;
        }
    } else {
        GLB_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

extern "C" GLB_API int glb_poisson_plan_kind(const glb_poisson_plan *plan) { return plan ? plan->kind : GLB_E_INVALID; }
extern "C" GLB_API int glb_poisson_plan_ld(const glb_poisson_plan *plan) { return plan ? plan->ldu : GLB_E_INVALID; }
extern "C" GLB_API int64_t glb_poisson_plan_rows(const glb_poisson_plan *plan) { return plan ? plan->n + plan->scratch_row : GLB_E_INVALID; }
extern "C" GLB_API double glb_poisson_plan_fill(const glb_poisson_plan *plan) { return plan ? plan->ell_fill : 0.0; }
// Host-only check of the slab builder (tests, no GPU needed): builds the slabs of a c-column plan with `grid` CTAs
// of 512 threads and walks them exactly as poisson_dataflow_kernel does (pairs of one warp's slots back to back, lane
// group g of slot s owns row slot_rows[s * rpw + g], long rows summed over the lane groups and over their PART pieces),
// computing y = P x for a fixed pseudo-random x in double precision.
// out4 = {max |y - P x| / max |P x|, warp-steps per iteration, fill, rows stored other than once}
extern "C" GLB_API int glb_dataflow_slabs_check_host(const int32_t *h_rowptr, const int32_t *h_col, const float *h_val, int64_t n,
                                                     int c, int grid, double *out4)
{
    GLB_CHECK_ARG(h_rowptr && out4 && n > 0 && grid > 0 && c > 0, "bad argument");
    const int lanes = flagged_lanes(c);
    GLB_CHECK_ARG(lanes > 0, "c too wide for the dataflow kernel");
    const int rowb = lanes * 16, rpw = 32 / lanes, nw = 16;
    std::vector<int> h_rp(h_rowptr, h_rowptr + n + 1);
    DfSlabs S;
    build_dataflow_slabs(h_rp, h_col, h_val, n, lanes, grid, nw, true, S);
    std::vector<double> x((size_t)n + kScratchRows, 0.0), y((size_t)n, 0.0), ref((size_t)n, 0.0);
    unsigned long long lcg = 88172645463325252ull;
    for (int64_t i = 0; i < n; ++i) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; x[i] = (double)(lcg >> 11) / 9007199254740992.0 - 0.5; }
    std::vector<int> stored((size_t)n, 0);
    for (int b = 0; b < grid; ++b) {
        const int2 *cv = S.slab.data() + S.slab_off[b];
        const int sl0 = S.slot_off[b], nslots = S.slot_off[b + 1] - sl0;
        std::vector<double> part((size_t)S.cap_parts + 1, 0.0);
        for (int pass = 0; pass < 2; ++pass)                       // PART pieces first (the kernel's owner waits for them)
            for (int w = 0; w < nw; ++w)
                for (int k = 0; k * nw + w < nslots; ++k) {
                    const int4 sl = S.slots[sl0 + k * nw + w];
                    const int type = sl.z & 0xff;
                    if ((type == kSlotPart) != (pass == 0)) continue;
                    std::vector<double> acc((size_t)rpw, 0.0);
                    const int npairs = (sl.y + 1) >> 1;
                    for (int pp = 0; pp < npairs; ++pp)
                        for (int g = 0; g < rpw; ++g)
                            for (int h = 0; h < 2; ++h) {
                                const int2 e = cv[sl.x + ((size_t)pp * rpw + g) * 2 + h];
                                const unsigned off = (unsigned)e.x;
                                if (off % (unsigned)rowb != 0 || off / rowb >= (unsigned)(n + kScratchRows)) { out4[0] = 1e300; return 0; }
                                float v; memcpy(&v, &e.y, sizeof(v));
                                acc[g] += (double)v * x[off / rowb];
                            }
                    const int *rows = S.slot_rows.data() + (size_t)(sl0 + k * nw + w) * rpw;
                    if (type == kSlotNormal) {
                        for (int g = 0; g < rpw; ++g) if (rows[g] >= 0) { y[rows[g]] = acc[g]; ++stored[rows[g]]; }
                    } else {
                        double sum = 0.0;
                        for (int g = 0; g < rpw; ++g) sum += acc[g];
                        if (type == kSlotPart) { part[sl.w] = sum; continue; }
                        if (type == kSlotOwner) for (int q = 0; q < (sl.z >> 8); ++q) sum += part[sl.w + q];
                        if (rows[0] >= 0) { y[rows[0]] = sum; ++stored[rows[0]]; }
                    }
                }
    }
    double worst = 0.0, big = 0.0;
    long long bad = 0;
    for (int64_t i = 0; i < n; ++i) {
        double r = 0.0;
        for (int q = h_rp[i]; q < h_rp[i + 1]; ++q) r += (double)h_val[q] * x[h_col[q]];
        worst = std::max(worst, fabs(r - y[i]));
        big = std::max(big, fabs(r));
        if (stored[i] != 1) ++bad;
    }
    out4[0] = big > 0.0 ? worst / big : worst;
    out4[1] = (double)S.steps;
    out4[2] = S.slab.size() ? (double)h_rp[n] / (double)S.slab.size() : 1.0;
    out4[3] = (double)bad;
    return 0;
}

extern "C" GLB_API int glb_poisson_plan_gate(const glb_poisson_plan *plan) { return plan && plan->kind == GLB_POISSON_KIND_DATAFLOW ? plan->gate_every : 0; }

extern "C" GLB_API int glb_poisson_pack(const glb_poisson_plan *plan, const double *d_src, const double *d_deg,
                                        const int32_t *d_perm, float *d_dst, void *stream)
{
    GLB_CHECK_ARG(plan && d_src && d_dst, "null pointer");
    const int blocks = stream_blocks_p(plan->n * plan->ldu);
    if (plan->kind == GLB_POISSON_KIND_DATAFLOW)
        pack_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_src, d_deg, plan->n, plan->c, d_dst, plan->ldu, d_perm);
    else
        pack_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_src, d_deg, plan->n, plan->c, d_dst, plan->ldu, d_perm);
    GLB_LAUNCH_CHECK();
    return 0;
}

extern "C" GLB_API int glb_poisson_unpack(const glb_poisson_plan *plan, const float *d_src, const int32_t *d_perm,
                                          double *d_dst, void *stream)
{
    GLB_CHECK_ARG(plan && d_src && d_dst, "null pointer");
    const int blocks = stream_blocks_p(plan->n * plan->c);
    if (plan->kind == GLB_POISSON_KIND_DATAFLOW)
        unpack_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_src, plan->n, plan->c, plan->ldu, d_dst, d_perm);
    else
        unpack_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(d_src, plan->n, plan->c, plan->ldu, d_dst, d_perm);
    GLB_LAUNCH_CHECK();
    return 0;
}

static int plan_step(const glb_poisson_plan *p, const float *Db, const float *u_in, float *u_out, cudaStream_t st)
{
    if (p->kind == GLB_POISSON_KIND_DATAFLOW)
        return dispatch_step<true>(p->d_rowptr, p->d_col, p->d_val, Db, u_in, u_out, p->n, p->ldu, st);
    return dispatch_step<false>(p->d_rowptr, p->d_col, p->d_val, Db, u_in, u_out, p->n, p->ldu, st);
}

extern "C" GLB_API int glb_poisson_step(const glb_poisson_plan *plan, const float *d_Db, const float *d_u_in,
                                        float *d_u_out, void *stream)
{
    GLB_CHECK_ARG(plan && d_Db && d_u_in && d_u_out, "null pointer");
    GLB_CHECK_ARG(d_u_in != d_u_out, "u_in and u_out must differ");
    plan_step(plan, d_Db, d_u_in, d_u_out, (cudaStream_t)stream);
    GLB_LAUNCH_CHECK();
    return 0;
}

extern "C" GLB_API int glb_poisson_iterate(glb_poisson_plan *plan, const float *d_Db, float *d_u0, float *d_u1, int T,
                                           int *result_in_u1, int *launches, void *stream)
{
    GLB_CHECK_ARG(plan && d_Db && d_u0 && d_u1, "null pointer");
    GLB_CHECK_ARG(T >= 0, "T must be >= 0");
    GLB_CHECK_ARG(d_u0 != d_u1, "u0 and u1 must differ");
    cudaStream_t st = (cudaStream_t)stream;
    if (result_in_u1) *result_in_u1 = T & 1;
    if (T == 0) return 0;
    const int *rp = plan->d_rowptr, *col = plan->d_col;
    const float *val = plan->d_val;
    if (plan->kind == GLB_POISSON_KIND_DATAFLOW) {
        unsigned *gate = plan->d_gate;
        unsigned *watchdog = plan->d_gate + 1;                // sticky: cleared only by glb_poisson_plan_check
        GLB_CUDA(cudaMemsetAsync(gate, 0, sizeof(unsigned), st));
        int gate_every = plan->gate_every;
        // version ring: the caller's u0 / u1 are buffers 0 / 1, the plan owns buffers 2 / 3
        const size_t buf_floats = ((size_t)plan->n + (size_t)plan->scratch_row) * (size_t)plan->ldu;
        DfRing ring;
        ring.b[0] = d_u0; ring.b[1] = d_u1; ring.b[2] = plan->d_ring; ring.b[3] = plan->d_ring + buf_floats;
        stamp_kernel<<<stream_blocks_p(plan->n * (plan->ldu / 4)), 256, 0, st>>>(ring, plan->n * (plan->ldu / 4),
                                                                                 plan->scratch_row * (plan->ldu / 4));
        void *args[] = {(void *)&plan->d_slabs, (void *)&plan->d_slab_off, (void *)&plan->d_slots, (void *)&plan->d_slot_off,
                        (void *)&plan->d_slot_rows, (void *)&d_Db, (void *)&ring, (void *)&T,
                        (void *)&plan->cap_entries, (void *)&plan->cap_slots, (void *)&plan->cap_parts, (void *)&plan->d_stats,
                        (void *)&gate, (void *)&gate_every, (void *)&watchdog, (void *)&plan->poll_sleep};
        if (plan->d_stats) GLB_CUDA(cudaMemsetAsync(plan->d_stats, 0, 16 * sizeof(unsigned long long), st));
        GLB_CUDA(cudaLaunchCooperativeKernel(plan->fn, dim3(plan->grid), dim3(plan->threads), args, plan->smem_bytes, st));
        if (launches) *launches += 2;
        if ((T & (kRing - 1)) >= 2) {                         // version T sits in a plan-owned buffer: hand it to the caller
            GLB_CUDA(cudaMemcpyAsync((T & 1) ? d_u1 : d_u0, ring.b[T & (kRing - 1)], sizeof(float) * (size_t)plan->n * plan->ldu,
                                     cudaMemcpyDeviceToDevice, st));
            if (launches) *launches += 1;
        }
        if (plan->d_stats) {
            unsigned long long h[16];
            GLB_CUDA(cudaMemcpyAsync(h, plan->d_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
            GLB_CUDA(cudaStreamSynchronize(st));
            fprintf(stderr, "[glb] dataflow T=%d grid=%d: re-polls %llu (%.4f per nonzero-lane), polling batches %llu, max warp cycles/iter %.0f\n",
                    T, plan->grid, h[0], (double)h[0] / ((double)plan->nnz * (plan->ldu / 4) * T + 1), h[1], (double)h[2] / T);
#ifdef GLB_EXPERIMENT
            if (h[6] && getenv("GLB_POISSON_WSTATS_FILE")) {
                const size_t nrec = (size_t)plan->grid * (plan->threads / 32) * 3;
                std::vector<unsigned long long> rec(nrec);
                GLB_CUDA(cudaMemcpy(rec.data(), plan->d_stats + 16, nrec * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                if (FILE *f = fopen(getenv("GLB_POISSON_WSTATS_FILE"), "w")) {
                    fprintf(f, "# T=%d grid=%d warps=%d: cta warp busy_cycles pairs slots\n", T, plan->grid, plan->threads / 32);
                    for (size_t i = 0; i < nrec / 3; ++i)
                        fprintf(f, "%zu %zu %llu %llu %llu\n", i / (plan->threads / 32), i % (plan->threads / 32), rec[3 * i], rec[3 * i + 1], rec[3 * i + 2]);
                    fclose(f);
                }
            }
#endif
            if (h[6])
                fprintf(stderr, "[glb]   busy cycles per iteration: mean over warps %.0f, slowest warp %.0f, slowest CTA (mean of its warps) %.0f, "
                                "mean over CTAs of the CTA's slowest warp %.0f\n", (double)h[4] / (double)h[6] / T, (double)h[5] / T, (double)h[7] / T,
                        (double)h[8] / (double)plan->grid / T);
        }
    } else if (plan->kind == GLB_POISSON_KIND_BARRIER) {
        GLB_CUDA(cudaMemsetAsync(plan->d_counter, 0, sizeof(unsigned) * kFlagStride * plan->grid, st));
        int n = (int)plan->n, ldu = plan->ldu, max_rows = plan->max_rows, cap = plan->slab_cap;
        void *args[] = {(void *)&rp, (void *)&col, (void *)&val, (void *)&d_Db, (void *)&d_u0, (void *)&d_u1,
                        (void *)&n, (void *)&ldu, (void *)&T, (void *)&plan->d_cta_rows, (void *)&max_rows, (void *)&cap,
                        (void *)&plan->d_counter};
        GLB_CUDA(cudaLaunchCooperativeKernel(plan->fn, dim3(plan->grid), dim3(plan->threads), args,
                                             plan->smem_bytes, st));
        if (launches) *launches += 1;
    } else {
        for (int t = 0; t < T; ++t) {
            const float *in = (t & 1) ? d_u1 : d_u0;
            float *out = (t & 1) ? d_u0 : d_u1;
            plan_step(plan, d_Db, in, out, st);
        }
        GLB_LAUNCH_CHECK();
        if (launches) *launches += T;
    }
    return 0;
}

extern "C" GLB_API int glb_poisson_mixing_T(const int32_t *d_rw_rowptr, const int32_t *d_rw_col, const double *d_rw_val,
                                    const double *d_vinf, double *d_v, double *d_tmp, int64_t n, int min_iter,
                                    int max_iter, int *T_out, int *launches, void *stream)
{
    GLB_CHECK_ARG(d_rw_rowptr && d_rw_col && d_rw_val && d_vinf && d_v && d_tmp && T_out, "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31), "n out of range");
    cudaStream_t st = (cudaStream_t)stream;
    const int BATCH = 64;
    unsigned long long *d_err = nullptr;
    GLB_CUDA(dev_alloc(&d_err, sizeof(unsigned long long) * (BATCH + 1)));
    // one cooperative launch per batch when the device supports it (grid = resident CTAs of 512 threads, at most one per SM)
    int dev = 0, coop = 0, coop_grid = 0;
    unsigned *d_sync = nullptr;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (coop) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mixing_persistent_kernel<false>, 512, 0);
        coop_grid = per_sm >= 1 ? sm_count() : 0;
        if (coop_grid > (int)((n * 8 + 511) / 512)) coop_grid = (int)((n * 8 + 511) / 512);
        if (coop_grid < 1) coop = 0;
        if (coop && dev_alloc(&d_sync, sizeof(unsigned)) != cudaSuccess) { cudaGetLastError(); coop = 0; }
    }
    unsigned long long h_err[BATCH + 1];
    const double thr = 1.0 / (double)n;
    const int threads = 256;
    int blocks = ceil_div(n * 8, threads);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    int T = 0, rc = 0;
    bool done = false;
    // err_0 = max|v_0 - vinf| decides only when min_iter = 0 (the loop test is `T < min_iter or err > 1/n`)
    double err = 1.0;
    if (min_iter <= 0) {
        cudaMemsetAsync(d_err, 0, sizeof(unsigned long long) * (BATCH + 1), st);
        maxdiff_kernel<<<blocks, threads, 0, st>>>(d_v, d_vinf, (int)n, d_err);
        if (launches) *launches += 1;
        cudaMemcpyAsync(h_err, d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        memcpy(&err, &h_err[0], sizeof(double));
    }
    double *cur = d_v, *nxt = d_tmp;
    while (!done) {
        // condition of ssl.py:667 evaluated BEFORE each step, with err = max|v_T - vinf|
        if (!((T < min_iter || err > thr) && T < max_iter)) break;
        int steps = max_iter - T;
        if (steps > BATCH) steps = BATCH;
        if (T < min_iter && steps > min_iter - T) steps = min_iter - T;     // the rule cannot fire before min_iter: look there first
        cudaMemsetAsync(d_err, 0, sizeof(unsigned long long) * (BATCH + 1), st);
        if (coop) {
            cudaMemsetAsync(d_sync, 0, sizeof(unsigned), st);
            int ni = (int)n;
            void *args[] = {(void *)&d_rw_rowptr, (void *)&d_rw_col, (void *)&d_rw_val, (void *)&d_vinf, (void *)&cur, (void *)&nxt,
                            (void *)&ni, (void *)&steps, (void *)&d_err, (void *)&d_sync};
            const bool hoist = n <= (int64_t)coop_grid * (512 / 8) * 8;       // one pass of the grid: rows stay in registers
            cudaError_t le = cudaLaunchCooperativeKernel(hoist ? (const void *)mixing_persistent_kernel<true> : (const void *)mixing_persistent_kernel<false>,
                                                         dim3(coop_grid), dim3(512), args, 0, st);
            if (le != cudaSuccess) { rc = (int)le; set_error("glb_poisson_mixing_T: %s", cudaGetErrorString(le)); break; }
            if (steps & 1) { double *t = cur; cur = nxt; nxt = t; }
            if (launches) *launches += 1;
        } else {
            for (int s = 0; s < steps; ++s) {
                mixing_step_kernel<<<blocks, threads, 0, st>>>(d_rw_rowptr, d_rw_col, d_rw_val, d_vinf, cur, nxt, (int)n,
                                                               d_err + s);
                double *t = cur; cur = nxt; nxt = t;
            }
            if (launches) *launches += steps;
        }
        cudaError_t e = cudaMemcpyAsync(h_err, d_err, sizeof(unsigned long long) * steps, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = (int)e; set_error("glb_poisson_mixing_T: %s", cudaGetErrorString(e)); break; }
        for (int s = 0; s < steps; ++s) {
            T += 1;
            memcpy(&err, &h_err[s], sizeof(double));
            if (!((T < min_iter || err > thr) && T < max_iter)) { done = true; break; }
        }
    }
    dev_free(d_err);
    dev_free(d_sync);
    if (rc == 0) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { rc = (int)e; set_error("glb_poisson_mixing_T: %s", cudaGetErrorString(e)); }
    }
    *T_out = T;
    return rc;
}
