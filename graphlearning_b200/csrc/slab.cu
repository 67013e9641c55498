// slab.cu - the Poisson iterate u <- Db + P u on a ROW SLAB of P: for graphs beyond the shared-memory kernels of
// poisson.cu (BASELINE config 5: 2 M nodes) and for the row-partitioned multi-GPU run of SURVEY.md 8(e).
//
// Replaces the loop body graphlearning/ssl.py:668 (scipy csr_matvecs) for one block of rows; with several GPUs every
// rank owns one block and the label rows its neighbours need (the "halo") are PUT straight into the neighbours' label
// matrices over NVLink by the kernel that computes them - no all-gather, no pack/unpack, no second kernel.
//
// Layout of a slab (glb_slab_create, built once on the host from the slab's CSR):
//   * local index space: own rows 0..m-1, then the H halo rows (rows of other ranks that some column points to), then
//     one all-zero scratch row (padding target).  Label matrices are (m + H + 1) x ld fp32, ld = glb_padded_ld(c),
//     a row of c = 10 classes = one 64-byte piece of a 128-byte line.
//   * sliced ELL: slices of 32/LANES rows (LANES lanes own one row, 16 bytes each), rows sorted by length inside
//     windows of 256 rows so that locality survives, slice width padded to even; the slices of a window are then folded
//     (longest, shortest, second longest, ...) so that every pair of slices - and with it every warp's part of a tile -
//     holds about the same number of entries; entries (byte offset of the column's label row, fp32 value) in PAIRS,
//     interleaved so that one warp-wide 16-byte load fetches one pair of every lane group from one contiguous 128-byte
//     run.  All slices back to back = ONE stream of pairs per warp.  A tile is 8 warps x 2 slices, or x 4 when the slab
//     still has 8 tiles per CTA then (the gather pipeline of a warp drains at tile boundaries).
//   * boundary rows (rows a peer needs, or rows that read halo rows) come first, interior rows after them.
//
// Kernel (slab_step_kernel, one launch per iteration): persistent CTAs of 8 warps walk tiles (8 warps x 2 slices) in a
// static order - about half of the CTAs take the boundary tiles first so that the puts leave early (then a share of the
// interior tiles that evens out the work), the others start on interior tiles at once and never wait for a neighbour.
// Lane 0 of every warp brings the warp's part of a tile's entry stream into shared memory
// with ONE bulk copy (cp.async.bulk -> mbarrier, TMA unit, SASS UBLKCP) - the copy for the NEXT tile is issued before the
// current one is processed (two stream buffers per CTA; every warp owns a fixed region of each, because the warps of a CTA
// are not synchronised between tiles and a warp that is a tile ahead must not write over entries a neighbour still reads) -
// then the warp walks it through a ring of four pair slots - pair p+4 is issued when pair p has been consumed,
// across slice boundaries, 6-8 label-row gathers per lane in flight (the same software pipeline as
// poisson_dataflow_pipe_kernel); gathers go through L1 (a locality ordering makes neighbouring rows share most of
// their columns).  Boundary tiles wait until the neighbours' halo
// rows of this version have arrived (one flag per neighbour in this rank's memory, acquire at system scope), write every
// finished row to the local matrix AND to each peer that needs it (plain 16-byte stores to peer memory mapped through
// CUDA IPC); a CTA orders and counts the puts of all its boundary tiles with ONE system-scope fence after the last of them,
// and the last CTA to do so releases this rank's flag in every neighbour's memory.  Interior tiles never
// wait: they run while the halo rows are in flight.  Three label buffers rotate (version v in buffer v % 3): a peer may already write
// version t+2 while this rank still reads version t.
//
// Arithmetic per row: acc = 0; acc = fma(val_j, u[col_j], acc) in stored order; + Db - the same chain as
// poisson_step_kernel, so results are bitwise independent of the number of ranks.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <numeric>
#include <vector>
#include "common.cuh"

namespace glb {

constexpr int kSlabWarps = 8;                    // warps per CTA
// Slices per warp and tile (one tile = one pass of a CTA = kSlabWarps x spw slices), chosen per slab: a warp's software
// pipeline drains at every tile boundary, so big tiles pay when there are many of them per CTA (2 M rows on one GPU:
// 0.174 ms per iteration with 2, 0.159 with 4, 0.230 with 1), small ones keep the static schedule balanced when a rank
// has only a few tiles per CTA (250 k rows per rank at 8 GPUs).  -DGLB_SLAB_SPW=k in the experiment build forces k.
constexpr int kSlabSPWSmall = 2, kSlabSPWBig = 4;
constexpr int kSlabBigTilesPerCta = 8;           // use the big tiles when every CTA still gets at least this many
constexpr int kSlabWindow = 256;                 // rows are sorted by length inside windows of this many rows
constexpr int kSlabLong = 64;                    // rows with more nonzeros get a slice of their own (dealt over the lane groups)
constexpr int kSlabLongBit = 0x40000000;         // slice_rows: this slice holds ONE long row
constexpr int kMaxPeers = 16;
constexpr int kFlagWords = 64;                   // flag area at the start of every rank's region: 64 x uint32

struct SlabParams {
    const int4 *ent;                 // entry pairs, all slices back to back
    const int *slice_first;          // [nslices + 1] index of a slice's first int4
    const int *slice_rows;           // [nslices * RPW] local row (| kSlabLongBit) or -1
    const long long *send_ptr;       // [n_bnd_slices * RPW + 1] puts of the boundary rows (nullptr: none)
    const int2 *send_ent;            // {peer rank, destination row in that peer's local space}
    const unsigned char *src_flag;   // [m] row has a nonzero source term
    const float *Db;                 // m x ld
    const float *u_in;               // (m + H + 1) x ld
    float *u_out;
    float *peer_out[kMaxPeers];      // u_out of this version on every rank (nullptr: not a neighbour)
    unsigned *peer_flag[kMaxPeers];  // this rank's flag word in every neighbour's memory
    const unsigned *my_flags;        // flag words in this rank's memory, written by the neighbours
    unsigned nbr_mask;               // neighbours (bit r = rank r)
    unsigned wait_epoch;             // boundary CTAs wait for my_flags[r] >= wait_epoch (0: version 0, nothing to wait for)
    unsigned signal_epoch;           // value released into the neighbours' flags after the boundary phase
    unsigned *bnd_counter;           // finished boundary CTAs, monotone over launches
    unsigned bnd_target;             // counter value that means "all boundary CTAs of THIS launch are done"
    unsigned *err_flag;              // watchdog
    int nslices, n_bnd_tiles;
    int tile_entries;                // capacity of one WARP's region of a stream buffer, in int4
    int bnd_ctas;                    // CTAs 0 .. bnd_ctas-1 take the boundary tiles first (then their share of the interior tiles)
    int bnd_group_interior;          // interior tiles that belong to that group of CTAs
    int spw;                         // slices per warp and tile
    int exp_flags;                   // -DGLB_EXPERIMENT builds: bit 0 = skip the puts (results wrong, cost probe), bit 1 = fence per tile
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float4 ld_row_l1(const char *p)
{
    float4 v;
    asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ const char *slab_addr(const char *base, unsigned off)
{
    unsigned long long a;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(a) : "r"(off), "l"(base));
    return reinterpret_cast<const char *>(a);
}

__device__ __forceinline__ void slab_issue(const int4 *cv, const char *in, float (&val)[2], float4 (&x)[2])
{
    const int4 e = cv[0];                                    // one pair = two (offset, value) entries of this lane group
    val[0] = __int_as_float(e.y); val[1] = __int_as_float(e.w);
    x[0] = ld_row_l1(slab_addr(in, (unsigned)e.x));
    x[1] = ld_row_l1(slab_addr(in, (unsigned)e.z));
}
__device__ __forceinline__ void slab_consume(const float (&val)[2], const float4 (&x)[2], float4 &acc)
{
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        acc.x = fmaf(val[i], x[i].x, acc.x);
        acc.y = fmaf(val[i], x[i].y, acc.y);
        acc.z = fmaf(val[i], x[i].z, acc.z);
        acc.w = fmaf(val[i], x[i].w, acc.w);
    }
}

// 3 resident CTAs per SM (80 registers): with 4 the kernel spills and runs at half the speed; an L2 run-ahead prefetch of
// a tile's label rows costs more LSU slots than it saves latency (0.210 vs 0.173 ms on config 5) - both measured in round 2
template <int LANES>
__global__ void __launch_bounds__(kSlabWarps * 32, 3)
slab_step_kernel(const SlabParams p)
{
    constexpr int RPW = 32 / LANES;
    constexpr unsigned ROWB = LANES * 16;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[2][kSlabWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / LANES, li = lane % LANES;
    const int spc = kSlabWarps * p.spw;              // slices per tile
    const int ntiles = (p.nslices + spc - 1) / spc;
    if (lane == 0) {
        mbar_init(&bars[0][warp], 1);
        mbar_init(&bars[1][warp], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // this warp's part of tile `tile`: slices [s0, s1), entries [f0, f1) of the stream; one bulk copy into buffer `buf`
    auto load_tile = [&](int tile, int buf) {
        const int cta_s0 = min(tile * spc, p.nslices);
        const int s0 = min(cta_s0 + warp * p.spw, p.nslices), s1 = min(s0 + p.spw, p.nslices);
        const int f0 = p.slice_first[s0], f1 = p.slice_first[s1];
        if (f1 > f0) {
            // every warp owns a fixed region of each stream buffer: a warp that runs a tile ahead of its neighbours must not
            // write its next part over entries a slower warp is still reading (parts of consecutive tiles have different lengths)
            int4 *dst = reinterpret_cast<int4 *>(smem) + ((size_t)buf * kSlabWarps + warp) * p.tile_entries;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the warp's earlier reads of this buffer are done
            mbar_expect_tx(&bars[buf][warp], (unsigned)(f1 - f0) * 16u);
            bulk_g2s(dst, p.ent + f0, (unsigned)(f1 - f0) * 16u, &bars[buf][warp]);
        }
    };
    // step -> tile of this CTA (static schedule, -1 = done).  Group B (CTAs < bnd_ctas): boundary tiles b, b + |B|, ...,
    // then interior tiles [n_bnd, n_bnd + bnd_group_interior) with the same stride; group I: the remaining interior tiles.
    const int n_bnd = p.n_bnd_tiles, gB = p.bnd_ctas, gI = (int)gridDim.x - gB;
    auto tile_at = [&](int step) -> int {
        const int b = blockIdx.x;
        if (b < gB) {
            const int mine_bnd = b < n_bnd ? (n_bnd - b + gB - 1) / gB : 0;
            if (step < mine_bnd) return b + step * gB;
            const int t = b + (step - mine_bnd) * gB;
            return t < p.bnd_group_interior ? n_bnd + t : -1;
        }
        const int t = (b - gB) + step * gI;
        return n_bnd + p.bnd_group_interior + t < ntiles ? n_bnd + p.bnd_group_interior + t : -1;
    };
    int tile = tile_at(0), buf = 0;
    unsigned par0 = 0u, par1 = 0u;                           // phase parity of the two stream buffers' barriers
    if (lane == 0 && tile >= 0) load_tile(tile, 0);
    bool waited = p.wait_epoch == 0u;
    unsigned bnd_done = 0u;                                  // boundary tiles of this CTA whose puts have not been counted yet
    const char *in = reinterpret_cast<const char *>(p.u_in) + li * 16;
    for (int step = 0; tile >= 0; ++step, buf ^= 1) {
        __syncwarp();
        const int next = tile_at(step + 1);
        if (lane == 0 && next >= 0) load_tile(next, buf ^ 1);           // prefetch the next tile's stream
        const bool boundary = tile < p.n_bnd_tiles;
        if (boundary && !waited) {                           // the neighbours' halo rows of this version must have landed
            if (threadIdx.x < kMaxPeers && ((p.nbr_mask >> threadIdx.x) & 1u)) {
                long long t0 = 0;
                unsigned spins = 0;
                while (ld_acquire_sys(p.my_flags + threadIdx.x) < p.wait_epoch) {
                    if ((++spins & 255u) == 0u) {
                        if (t0 == 0) t0 = clock64();
                        else if (clock64() - t0 > 6000000000ll || *reinterpret_cast<volatile unsigned *>(p.err_flag)) {
                            *reinterpret_cast<volatile unsigned *>(p.err_flag) = 1u;       // a peer is gone: drain instead of hanging
                            break;
                        }
                    }
                }
            }
            __syncthreads();
            waited = true;
        }
        const int cta_s0 = min(tile * spc, p.nslices);
        const int s0 = min(cta_s0 + warp * p.spw, p.nslices), s1 = min(s0 + p.spw, p.nslices);
        const int f0 = p.slice_first[s0], f1 = p.slice_first[s1];
        if (s1 > s0) {
            if (f1 > f0) {
                mbar_wait(&bars[buf][warp], buf ? par1 : par0);
                if (buf) par1 ^= 1u; else par0 ^= 1u;
            }
            const int4 *cv0 = reinterpret_cast<const int4 *>(smem) + ((size_t)buf * kSlabWarps + warp) * p.tile_entries + g;
            const int n_pairs = (f1 - f0) / RPW;             // pairs of this warp's stream
            float v0[2], v1[2], v2[2], v3[2];
            float4 x0[2], x1[2], x2[2], x3[2];
            if (0 < n_pairs) slab_issue(cv0, in, v0, x0);
            if (1 < n_pairs) slab_issue(cv0 + RPW, in, v1, x1);
            if (2 < n_pairs) slab_issue(cv0 + 2 * RPW, in, v2, x2);
            if (3 < n_pairs) slab_issue(cv0 + 3 * RPW, in, v3, x3);
            int q = 0, s = s0;                               // pair being consumed, slice it belongs to
            int left = (p.slice_first[s0 + 1] - f0) / RPW;   // pairs of slice s not yet consumed
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            // finish slice s (sum of a long row's pieces, source term, store, puts) and move to the next one
            auto finish = [&]() {
                int rinfo = p.slice_rows[(size_t)s * RPW + g];
                const int r0info = __shfl_sync(0xffffffffu, rinfo, 0);
                if (r0info >= 0 && (r0info & kSlabLongBit)) {    // warp-uniform: one long row dealt over the lane groups
#pragma unroll
                    for (int o = LANES; o < 32; o <<= 1) {
                        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
                        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
                        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
                    }
                    rinfo = g == 0 ? (r0info & ~kSlabLongBit) : -1;
                }
                if (rinfo >= 0) {
                    const unsigned row = (unsigned)rinfo;
                    if (p.src_flag[row]) {
                        const float4 b = __ldg(reinterpret_cast<const float4 *>(p.Db + (size_t)row * (ROWB / 4)) + li);
                        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
                    }
                    *reinterpret_cast<float4 *>(reinterpret_cast<char *>(p.u_out) + (size_t)row * ROWB + li * 16) = acc;
#ifdef GLB_EXPERIMENT
                    if (boundary && p.send_ptr && !(p.exp_flags & 1)) {
#else
                    if (boundary && p.send_ptr) {            // put the row into every peer that gathers it
#endif
                        const long long q0 = p.send_ptr[(size_t)s * RPW + g], q1 = p.send_ptr[(size_t)s * RPW + g + 1];
                        for (long long k = q0; k < q1; ++k) {
                            const int2 e = p.send_ent[k];
                            *reinterpret_cast<float4 *>(reinterpret_cast<char *>(p.peer_out[e.x]) + (size_t)(unsigned)e.y * ROWB + li * 16) = acc;
                        }
                    }
                }
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
                ++s;
                if (s < s1) left = (p.slice_first[s + 1] - p.slice_first[s]) / RPW;
            };
            while (s < s1 && left == 0) finish();            // slices of empty rows
#define GLB_SLAB_STEP(V, X)                                                                          \
            {                                                                                        \
                if (q >= n_pairs) break;                                                             \
                slab_consume(V, X, acc);                                                             \
                if (q + 4 < n_pairs) slab_issue(cv0 + (size_t)(q + 4) * RPW, in, V, X);              \
                ++q;                                                                                 \
                if (--left == 0) { finish(); while (s < s1 && left == 0) finish(); }                 \
            }
            for (;;) {
                GLB_SLAB_STEP(v0, x0)
                GLB_SLAB_STEP(v1, x1)
                GLB_SLAB_STEP(v2, x2)
                GLB_SLAB_STEP(v3, x3)
            }
#undef GLB_SLAB_STEP
        }
        if (boundary && p.nbr_mask) {
            // The puts of this CTA's boundary tiles are counted ONCE, after its last boundary tile: a system-scope fence
            // waits for the NVLink acknowledgements of everything the CTA has stored to its peers (microseconds), and
            // one fence per tile put that wait between every two tiles of the boundary phase.
            ++bnd_done;
            bool flush = next < 0 || next >= p.n_bnd_tiles;
#ifdef GLB_EXPERIMENT
            flush = flush || (p.exp_flags & 2);
#endif
            if (flush) {
                __syncthreads();                             // every put of these tiles is issued ...
                if (threadIdx.x == 0) {
                    __threadfence_system();                  // ... and ordered before the count
                    const unsigned done = atomicAdd(p.bnd_counter, bnd_done) + bnd_done;
                    if (done == p.bnd_target) {              // last boundary tile of this launch: release the neighbours
                        __threadfence_system();
                        for (int r = 0; r < kMaxPeers; ++r)
                            if ((p.nbr_mask >> r) & 1u) st_release_sys(p.peer_flag[r], p.signal_epoch);
                    }
                }
                bnd_done = 0u;
            }
        }
        tile = next;
    }
}

__global__ void __launch_bounds__(256) slab_mark_sources_kernel(const float *__restrict__ Db, long long m, int ld, unsigned char *__restrict__ flag)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < m; r += (long long)gridDim.x * blockDim.x) {
        bool nz = false;
        for (int k = 0; k < ld; ++k) nz |= Db[r * ld + k] != 0.f;             // NaN != 0 is true
        flag[r] = nz ? 1 : 0;
    }
}

// rows [0, m) of a label buffer (plain layout) <-> m x c float64; scale: dst = src / deg (Db = D^-1 source, ssl.py:636)
__global__ void __launch_bounds__(256) slab_pack_kernel(const double *__restrict__ src, const double *__restrict__ deg, long long m, int c,
                                                        int ld, float *__restrict__ dst)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m * ld; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / ld;
        const int k = (int)(i - r * ld);
        float v = 0.f;
        if (k < c) v = (float)(deg ? (1.0 / deg[r]) * src[r * c + k] : src[r * c + k]);
        dst[i] = v;
    }
}
__global__ void __launch_bounds__(256) slab_unpack_kernel(const float *__restrict__ src, long long m, int c, int ld, double *__restrict__ dst)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m * c; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / c;
        dst[i] = (double)src[r * ld + (i - r * c)];
    }
}

static inline int slab_float_bits(float f) { int i; memcpy(&i, &f, sizeof(i)); return i; }

}  // namespace glb

using namespace glb;

struct glb_slab {
    int64_t m = 0, rows_total = 0, nnz = 0;
    int c = 0, ld = 0, lanes = 0, rpw = 0;
    int nslices = 0, n_bnd_slices = 0, grid = 0, tile_entries = 0, spw = 2;
    size_t smem_bytes = 0;
    double fill = 1.0;
    int4 *d_ent = nullptr;
    int *d_slice_first = nullptr, *d_slice_rows = nullptr;
    long long *d_send_ptr = nullptr;
    int2 *d_send_ent = nullptr;
    unsigned char *d_src_flag = nullptr;
    unsigned *d_sync = nullptr;          // [0] boundary counter, [1] watchdog flag
    // multi-GPU attachment
    int rank = 0, world = 1;
    unsigned nbr_mask = 0;
    char *region[kMaxPeers] = {nullptr};
    int64_t region_rows[kMaxPeers] = {0};
    unsigned epoch = 0, launches = 0;
    const void *fn = nullptr;
};

template <int LANES>
static const void *slab_fn() { return (const void *)slab_step_kernel<LANES>; }

static const void *slab_pick(int lanes)
{
    switch (lanes) {
        case 1: return slab_fn<1>();
        case 2: return slab_fn<2>();
        case 4: return slab_fn<4>();
        case 8: return slab_fn<8>();
        case 16: return slab_fn<16>();
        default: return slab_fn<32>();
    }
}

// Everything glb_slab_create builds on the host: the sliced-ELL entry stream in tile / warp / slice order and the puts of
// the boundary rows.  Pure host code (glb_slab_check_host walks the result without a GPU).
struct SlabHost {
    int ld = 0, lanes = 0, rpw = 0, spw = 0, nslices = 0, n_bnd_slices = 0;
    int64_t rows_total = 0, nnz = 0, stored = 0;
    size_t max_warp_bytes = 0;                   // longest part of one warp = size of a warp's region of a stream buffer
    std::vector<int> slice_first, slice_rows;
    std::vector<int4> ent;
    std::vector<long long> send_ptr;
    std::vector<int2> send_ent;
};

static int slab_build_host(SlabHost &H, const int32_t *h_rowptr, const int32_t *h_col, const float *h_val, int64_t m, int64_t n_halo,
                           int c, const uint8_t *h_boundary, const int64_t *h_send_ptr, const int32_t *h_send_peer,
                           const int32_t *h_send_dst, int sms)
{
    GLB_CHECK_ARG(h_rowptr && m > 0 && n_halo >= 0 && c > 0 && sms > 0, "bad argument");
    const int64_t nnz = h_rowptr[m];
    GLB_CHECK_ARG(nnz == 0 || (h_col && h_val), "null pointer");
    const int ld = glb_padded_ld(c);
    GLB_CHECK_ARG(ld <= 128, "c > 128 is not supported by the slab kernel");
    const int lanes = ld / 4, rpw = 32 / lanes;
    const unsigned rowb = (unsigned)lanes * 16u;
    const int64_t rows_total = m + n_halo + 1;
    GLB_CHECK_ARG((double)rows_total * rowb < 4294967296.0, "label matrix of this slab exceeds 4 GiB: 32-bit row offsets overflow");
    PhaseTimer tm("slab_build_host");

    int spw = kSlabSPWSmall;
    {   // big tiles when the slab has many tiles per CTA and two stream buffers of a big tile leave room for three CTAs per SM
        const double tiles_big = (double)m / rpw / (kSlabWarps * kSlabSPWBig);
        const double bytes_big = 2.0 * 1.3 * ((double)nnz / (double)m) * rpw * kSlabWarps * kSlabSPWBig * 8.0;
        if (tiles_big >= (double)kSlabBigTilesPerCta * 3 * sms && bytes_big <= 72.0 * 1024) spw = kSlabSPWBig;
#ifdef GLB_SLAB_SPW
        spw = GLB_SLAB_SPW;
#endif
    }
    const int spc = kSlabWarps * spw;                        // slices per tile of THIS slab
    // ---- row order: boundary rows first; inside each group long rows, then windows of rows sorted by length --------------
    struct Slice { int L; int rows[32]; bool is_long; };
    std::vector<Slice> slices;
    std::vector<int> group, win;
    int n_bnd_slices = 0;
    int64_t stored = 0;
    for (int pass = 0; pass < 2; ++pass) {
        group.clear();
        for (int64_t r = 0; r < m; ++r) {
            const bool b = h_boundary && h_boundary[r];
            if (b == (pass == 0)) group.push_back((int)r);
        }
        for (int r : group)
            if (h_rowptr[r + 1] - h_rowptr[r] > kSlabLong) {
                Slice sl{};
                sl.is_long = true;
                sl.L = (h_rowptr[r + 1] - h_rowptr[r] + rpw - 1) / rpw;
                for (int q = 0; q < rpw; ++q) sl.rows[q] = -1;
                sl.rows[0] = r;
                slices.push_back(sl);
            }
        for (size_t w0 = 0; w0 < group.size(); w0 += kSlabWindow) {
            win.clear();
            for (size_t k = w0; k < std::min(group.size(), w0 + (size_t)kSlabWindow); ++k)
                if (h_rowptr[group[k] + 1] - h_rowptr[group[k]] <= kSlabLong) win.push_back(group[k]);
            std::stable_sort(win.begin(), win.end(), [&](int a, int b) { return h_rowptr[a + 1] - h_rowptr[a] > h_rowptr[b + 1] - h_rowptr[b]; });
            const size_t first_of_window = slices.size();
            for (size_t k0 = 0; k0 < win.size(); k0 += rpw) {
                Slice sl{};
                sl.is_long = false;
                sl.L = h_rowptr[win[k0] + 1] - h_rowptr[win[k0]];
                for (int q = 0; q < rpw; ++q) sl.rows[q] = k0 + q < win.size() ? win[k0 + q] : -1;
                slices.push_back(sl);
            }
            // The window's slices come out sorted by length.  Fold them - longest, shortest, second longest, second
            // shortest, ... - so that every run of an even number of slices holds about the same number of entries: the
            // parts of the warps of a tile are balanced (a warp that always got the longest slices would finish last, and
            // the longest part sizes every warp's region of the stream buffers).
            {
                const size_t cnt = slices.size() - first_of_window;
                std::vector<Slice> sorted(slices.begin() + first_of_window, slices.end());
                for (size_t q = 0; q < cnt; ++q) slices[first_of_window + q] = (q & 1) ? sorted[cnt - 1 - q / 2] : sorted[q / 2];
            }
        }
        if (pass == 0) {                                     // boundary slices fill whole tiles
            while (slices.size() % spc) {
                Slice sl{};
                for (int q = 0; q < rpw; ++q) sl.rows[q] = -1;
                slices.push_back(sl);
            }
            n_bnd_slices = (int)slices.size();
        }
    }
    const int nslices = (int)slices.size();
    std::vector<int> slice_first((size_t)nslices + 1, 0), slice_rows((size_t)nslices * rpw, -1);
    for (int s = 0; s < nslices; ++s) {
        const int Lst = (slices[s].L + 1) / 2 * 2;
        slice_first[s + 1] = slice_first[s] + Lst / 2 * rpw;
        stored += (int64_t)Lst * rpw;
        GLB_CHECK_ARG(slice_first[s + 1] >= slice_first[s], "slab too large (entry index overflows 32 bits)");
    }
    const unsigned pad = (unsigned)(rows_total - 1) * rowb;       // the all-zero scratch row
    std::vector<int4> ent((size_t)slice_first[nslices], make_int4((int)pad, 0, (int)pad, 0));
    int2 *e2 = reinterpret_cast<int2 *>(ent.data());
    size_t max_cta = 0;                                      // longest part of one warp, in bytes
    for (int s = 0; s < nslices; ++s) {
        const Slice &sl = slices[s];
        const size_t base = (size_t)slice_first[s] * 2;          // in entries
        if (sl.is_long) {
            const int r = sl.rows[0];
            slice_rows[(size_t)s * rpw] = r | kSlabLongBit;
            const int len = h_rowptr[r + 1] - h_rowptr[r];
            for (int q = 0; q < len; ++q) {                      // round-robin over the lane groups
                const int j = q / rpw, g = q % rpw;
                const int64_t src = (int64_t)h_rowptr[r] + q;
                e2[base + ((size_t)(j / 2) * rpw + g) * 2 + (j & 1)] = make_int2((int)((unsigned)h_col[src] * rowb), slab_float_bits(h_val[src]));
            }
        } else {
            for (int g = 0; g < rpw; ++g) {
                const int r = sl.rows[g];
                slice_rows[(size_t)s * rpw + g] = r;
                if (r < 0) continue;
                const int len = h_rowptr[r + 1] - h_rowptr[r];
                for (int j = 0; j < len; ++j) {
                    const int64_t src = (int64_t)h_rowptr[r] + j;
                    GLB_CHECK_ARG(h_col[src] >= 0 && h_col[src] < m + n_halo, "column index outside the slab's local index space");
                    e2[base + ((size_t)(j / 2) * rpw + g) * 2 + (j & 1)] = make_int2((int)((unsigned)h_col[src] * rowb), slab_float_bits(h_val[src]));
                }
            }
        }
        if (s % spw == spw - 1 || s == nslices - 1) {            // end of one warp's part of a tile
            const int w0 = s / spw * spw;
            max_cta = std::max(max_cta, (size_t)(slice_first[s + 1] - slice_first[w0]) * 16);
        }
    }
    tm.lap("sliced-ELL build");
    // ---- puts of the boundary rows -----------------------------------------------------------------------------------------
    std::vector<long long> send_ptr;
    std::vector<int2> send_ent;
    if (h_send_ptr && n_bnd_slices > 0) {
        send_ptr.assign((size_t)n_bnd_slices * rpw + 1, 0);
        for (int s = 0; s < n_bnd_slices; ++s)
            for (int g = 0; g < rpw; ++g) {
                int r = slice_rows[(size_t)s * rpw + g];
                if (r >= 0) {
                    r &= ~kSlabLongBit;
                    for (int64_t q = h_send_ptr[r]; q < h_send_ptr[r + 1]; ++q) {
                        GLB_CHECK_ARG(h_send_peer[q] >= 0 && h_send_peer[q] < kMaxPeers && h_send_dst[q] >= 0, "bad put entry");
                        send_ent.push_back(make_int2(h_send_peer[q], h_send_dst[q]));
                    }
                }
                send_ptr[(size_t)s * rpw + g + 1] = (long long)send_ent.size();
            }
        for (int64_t r = 0; r < m; ++r)
            GLB_CHECK_ARG(h_send_ptr[r + 1] == h_send_ptr[r] || (h_boundary && h_boundary[r]), "a row with puts must be flagged as boundary");
    }

    H.ld = ld; H.lanes = lanes; H.rpw = rpw; H.spw = spw; H.nslices = nslices; H.n_bnd_slices = n_bnd_slices;
    H.rows_total = rows_total; H.nnz = nnz; H.stored = stored; H.max_warp_bytes = max_cta;
    H.slice_first.swap(slice_first); H.slice_rows.swap(slice_rows); H.ent.swap(ent);
    H.send_ptr.swap(send_ptr); H.send_ent.swap(send_ent);
    return 0;
}

extern "C" GLB_API int glb_slab_create(glb_slab **out, const int32_t *h_rowptr, const int32_t *h_col, const float *h_val,
                                       int64_t m, int64_t n_halo, int c, const uint8_t *h_boundary, const int64_t *h_send_ptr,
                                       const int32_t *h_send_peer, const int32_t *h_send_dst, void *stream)
{
    GLB_CHECK_ARG(out, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    PhaseTimer tm("slab_create");
    SlabHost H;
    int rc = slab_build_host(H, h_rowptr, h_col, h_val, m, n_halo, c, h_boundary, h_send_ptr, h_send_peer, h_send_dst, sm_count());
    if (rc) return rc;
    const int ld = H.ld, lanes = H.lanes, rpw = H.rpw, spw = H.spw, nslices = H.nslices, n_bnd_slices = H.n_bnd_slices;
    const int spc = kSlabWarps * spw;
    const int64_t rows_total = H.rows_total, nnz = H.nnz, stored = H.stored;
    const size_t max_cta = H.max_warp_bytes;
    const std::vector<int> &slice_first = H.slice_first, &slice_rows = H.slice_rows;
    const std::vector<int4> &ent = H.ent;
    const std::vector<long long> &send_ptr = H.send_ptr;
    const std::vector<int2> &send_ent = H.send_ent;

    int dev = 0, max_smem = 0;
    GLB_CUDA(cudaGetDevice(&dev));
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (2 * kSlabWarps * max_cta + 1024 > (size_t)max_smem) {
        set_error("glb_slab_create: the entry streams of a tile need 2 x %d x %zu bytes of shared memory (a row is too long for the slab kernel)", kSlabWarps, max_cta);
        return GLB_E_UNSUPPORTED;
    }
    glb_slab *s = new glb_slab();
    struct Guard { glb_slab *s; ~Guard() { if (s) glb_slab_destroy(s); } } guard{s};
    s->m = m; s->rows_total = rows_total; s->nnz = nnz; s->c = c; s->ld = ld; s->lanes = lanes; s->rpw = rpw;
    s->nslices = nslices; s->n_bnd_slices = n_bnd_slices; s->spw = spw;
    s->tile_entries = (int)(std::max<size_t>(max_cta, 16) / 16);
    s->smem_bytes = 2 * (size_t)kSlabWarps * s->tile_entries * 16;
    s->fill = stored ? (double)nnz / (double)stored : 1.0;
    s->fn = slab_pick(lanes);
    GLB_CUDA(cudaFuncSetAttribute(s->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    {   // shared-memory carve-out for three resident CTAs (80 registers x 256 threads); the rest of the 228 KB stays L1 for the label-row gathers
        const size_t want = std::min<size_t>(3 * (s->smem_bytes + 1024), (size_t)max_smem);
        cudaFuncSetAttribute(s->fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)std::min<size_t>(100, (want * 100 + max_smem - 1) / max_smem));
        int per_sm = 0;
        GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s->fn, kSlabWarps * 32, s->smem_bytes));
        const int ntiles = (nslices + spc - 1) / spc;
        s->grid = std::max(1, std::min(ntiles, std::max(1, per_sm) * sm_count()));        // persistent: every CTA walks tiles b, b + grid, ...
    }
    GLB_CUDA(dev_alloc(&s->d_ent, sizeof(int4) * std::max<size_t>(ent.size(), 1)));
    GLB_CUDA(dev_alloc(&s->d_slice_first, sizeof(int) * (nslices + 1)));
    GLB_CUDA(dev_alloc(&s->d_slice_rows, sizeof(int) * std::max<size_t>(slice_rows.size(), 1)));
    GLB_CUDA(dev_alloc(&s->d_src_flag, (size_t)m));
    GLB_CUDA(dev_alloc(&s->d_sync, 2 * sizeof(unsigned)));
    GLB_CUDA(cudaMemsetAsync(s->d_sync, 0, 2 * sizeof(unsigned), st));
    GLB_CUDA(cudaMemcpyAsync(s->d_ent, ent.data(), sizeof(int4) * ent.size(), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(s->d_slice_first, slice_first.data(), sizeof(int) * (nslices + 1), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(s->d_slice_rows, slice_rows.data(), sizeof(int) * slice_rows.size(), cudaMemcpyHostToDevice, st));
    if (!send_ptr.empty()) {
        GLB_CUDA(dev_alloc(&s->d_send_ptr, sizeof(long long) * send_ptr.size()));
        GLB_CUDA(dev_alloc(&s->d_send_ent, sizeof(int2) * std::max<size_t>(send_ent.size(), 1)));
        GLB_CUDA(cudaMemcpyAsync(s->d_send_ptr, send_ptr.data(), sizeof(long long) * send_ptr.size(), cudaMemcpyHostToDevice, st));
        GLB_CUDA(cudaMemcpyAsync(s->d_send_ent, send_ent.data(), sizeof(int2) * send_ent.size(), cudaMemcpyHostToDevice, st));
    }
    GLB_CUDA(cudaStreamSynchronize(st));                     // the staging vectors are locals
    tm.lap("upload");
    guard.s = nullptr;
    *out = s;
    return 0;
}

// Host-only self-check of the slab builder (no GPU): the stream is walked tile by tile, warp by warp, slice by slice as
// slab_step_kernel walks it, y = P x in double precision.
extern "C" GLB_API int glb_slab_check_host(const int32_t *h_rowptr, const int32_t *h_col, const float *h_val, int64_t m, int64_t n_halo,
                                           int c, const uint8_t *h_boundary, int sms, double *out8)
{
    GLB_CHECK_ARG(out8, "null pointer");
    SlabHost H;
    int rc = slab_build_host(H, h_rowptr, h_col, h_val, m, n_halo, c, h_boundary, nullptr, nullptr, nullptr, sms);
    if (rc) return rc;
    const int rpw = H.rpw, spw = H.spw, spc = kSlabWarps * spw;
    const unsigned rowb = (unsigned)H.lanes * 16u, pad = (unsigned)(H.rows_total - 1) * rowb;
    std::vector<double> x((size_t)H.rows_total, 0.0), y((size_t)m, 0.0), ref((size_t)m, 0.0);
    for (int64_t i = 0; i + 1 < H.rows_total; ++i) x[(size_t)i] = 0.25 + (double)((i * 2654435761ull) % 1000) / 1000.0;   // scratch row stays 0
    std::vector<int> seen((size_t)m, 0);
    const int2 *e2 = reinterpret_cast<const int2 *>(H.ent.data());
    const int ntiles = (H.nslices + spc - 1) / spc;
    int64_t bad = 0, misplaced = 0;
    size_t longest = 0, total_parts = 0, nparts = 0;
    for (int t = 0; t < ntiles; ++t)
        for (int w = 0; w < kSlabWarps; ++w) {
            const int s0 = std::min(std::min(t * spc, H.nslices) + w * spw, H.nslices), s1 = std::min(s0 + spw, H.nslices);
            const size_t part = (size_t)(H.slice_first[s1] - H.slice_first[s0]) * 16;
            if (part > H.max_warp_bytes) ++bad;                                  // would overflow the warp's region of the stream buffer
            longest = std::max(longest, part);
            if (s1 > s0) { total_parts += part; ++nparts; }
            for (int sl = s0; sl < s1; ++sl) {
                const size_t base = (size_t)H.slice_first[sl] * 2;
                const int Lst = (H.slice_first[sl + 1] - H.slice_first[sl]) * 2 / rpw;
                const int r0 = H.slice_rows[(size_t)sl * rpw];
                const bool is_long = r0 >= 0 && (r0 & kSlabLongBit);
                for (int g = 0; g < rpw; ++g) {
                    int row = is_long ? (r0 & ~kSlabLongBit) : H.slice_rows[(size_t)sl * rpw + g];
                    if (!is_long || g == 0) {
                        if (row >= 0) {
                            seen[(size_t)row] += 1;
                            if (h_boundary && (bool)h_boundary[row] != (sl < H.n_bnd_slices)) ++misplaced;
                        }
                    }
                    for (int j = 0; j < Lst; ++j) {
                        const int2 e = e2[base + ((size_t)(j / 2) * rpw + g) * 2 + (j & 1)];
                        float v;
                        memcpy(&v, &e.y, sizeof(v));
                        if (row < 0) { if ((unsigned)e.x != pad || v != 0.f) ++bad; continue; }     // padding lane group
                        if ((unsigned)e.x % rowb) { ++bad; continue; }
                        y[(size_t)row] += (double)v * x[(size_t)((unsigned)e.x / rowb)];
                    }
                }
            }
        }
    double worst = 0.0, big = 0.0;
    for (int64_t r = 0; r < m; ++r) {
        double acc = 0.0;
        for (int j = h_rowptr[r]; j < h_rowptr[r + 1]; ++j) acc += (double)h_val[j] * x[(size_t)h_col[j]];
        ref[(size_t)r] = acc;
        big = std::max(big, fabs(acc));
        worst = std::max(worst, fabs(acc - y[(size_t)r]));
        if (seen[(size_t)r] != 1) ++bad;
    }
    out8[0] = big > 0.0 ? worst / big : worst;
    out8[1] = (double)bad;
    out8[2] = (double)misplaced;
    out8[3] = H.stored ? (double)H.nnz / (double)H.stored : 1.0;
    out8[4] = (double)spc;
    out8[5] = (double)H.max_warp_bytes;
    out8[6] = nparts ? (double)longest / ((double)total_parts / (double)nparts) : 1.0;
    out8[7] = (double)ntiles;
    return 0;
}

extern "C" GLB_API int glb_slab_destroy(glb_slab *s)
{
    if (!s) return 0;
    dev_free(s->d_ent); dev_free(s->d_slice_first); dev_free(s->d_slice_rows); dev_free(s->d_send_ptr); dev_free(s->d_send_ent);
    dev_free(s->d_src_flag); dev_free(s->d_sync);
    delete s;
    return 0;
}

extern "C" GLB_API int64_t glb_slab_rows(const glb_slab *s) { return s ? s->rows_total : GLB_E_INVALID; }
extern "C" GLB_API int glb_slab_ld(const glb_slab *s) { return s ? s->ld : GLB_E_INVALID; }
extern "C" GLB_API double glb_slab_fill(const glb_slab *s) { return s ? s->fill : 0.0; }
extern "C" GLB_API int glb_slab_tile_slices(const glb_slab *s) { return s ? kSlabWarps * s->spw : GLB_E_INVALID; }
extern "C" GLB_API int64_t glb_slab_region_bytes(const glb_slab *s)
{
    return s ? (int64_t)(kFlagWords * sizeof(unsigned)) + 3 * s->rows_total * s->ld * (int64_t)sizeof(float) : GLB_E_INVALID;
}

static float *slab_buffer(const glb_slab *s, int rank, int v)
{
    return reinterpret_cast<float *>(s->region[rank] + kFlagWords * sizeof(unsigned)) + (size_t)v * s->region_rows[rank] * s->ld;
}

extern "C" GLB_API int glb_slab_attach(glb_slab *s, int rank, int world, void *const *region_base, const int64_t *region_rows,
                                       uint32_t neighbour_mask)
{
    GLB_CHECK_ARG(s && region_base && region_rows, "null pointer");
    GLB_CHECK_ARG(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "rank/world out of range");
    GLB_CHECK_ARG(region_base[rank] && region_rows[rank] == s->rows_total, "own region missing or of the wrong size");
    GLB_CHECK_ARG(!((neighbour_mask >> rank) & 1u), "a rank is not its own neighbour");
    for (int r = 0; r < world; ++r) {
        GLB_CHECK_ARG(!((neighbour_mask >> r) & 1u) || region_base[r], "a neighbour's region is not mapped");
        s->region[r] = (char *)region_base[r];
        s->region_rows[r] = region_rows[r];
    }
    s->rank = rank; s->world = world; s->nbr_mask = neighbour_mask;
    return 0;
}

extern "C" GLB_API int glb_slab_buffer(const glb_slab *s, int v, float **d_buf)
{
    GLB_CHECK_ARG(s && d_buf && v >= 0 && v < 3 && s->region[s->rank], "bad argument (attach the slab first)");
    *d_buf = slab_buffer(s, s->rank, v);
    return 0;
}

extern "C" GLB_API int glb_slab_pack(const glb_slab *s, const double *d_src, const double *d_deg, float *d_dst, void *stream)
{
    GLB_CHECK_ARG(s && d_src && d_dst, "null pointer");
    const int blocks = (int)std::min<int64_t>((s->m * s->ld + 255) / 256, (int64_t)sm_count() * 16);
    slab_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_src, d_deg, s->m, s->c, s->ld, d_dst);
    GLB_LAUNCH_CHECK();
    return 0;
}

extern "C" GLB_API int glb_slab_unpack(const glb_slab *s, int v, double *d_dst, void *stream)
{
    GLB_CHECK_ARG(s && d_dst && v >= 0 && v < 3 && s->region[s->rank], "bad argument (attach the slab first)");
    const int blocks = (int)std::min<int64_t>((s->m * s->c + 255) / 256, (int64_t)sm_count() * 16);
    slab_unpack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(slab_buffer(s, s->rank, v), s->m, s->c, s->ld, d_dst);
    GLB_LAUNCH_CHECK();
    return 0;
}

extern "C" GLB_API int glb_slab_iterate(glb_slab *s, const float *d_Db, int T, int *result_buffer, int *launches, void *stream)
{
    GLB_CHECK_ARG(s && d_Db && T >= 0, "bad argument");
    GLB_CHECK_ARG(s->region[s->rank], "attach the slab first (glb_slab_attach)");
    cudaStream_t st = (cudaStream_t)stream;
    if (result_buffer) *result_buffer = T % 3;
    if (T == 0) return 0;
    const int blocks = (int)std::min<int64_t>((s->m + 255) / 256, (int64_t)sm_count() * 16);
    slab_mark_sources_kernel<<<blocks, 256, 0, st>>>(d_Db, s->m, s->ld, s->d_src_flag);
    SlabParams p{};
    p.ent = s->d_ent; p.slice_first = s->d_slice_first; p.slice_rows = s->d_slice_rows;
    p.send_ptr = s->nbr_mask ? s->d_send_ptr : nullptr; p.send_ent = s->d_send_ent;
    p.src_flag = s->d_src_flag; p.Db = d_Db;
    p.my_flags = reinterpret_cast<const unsigned *>(s->region[s->rank]);
    p.nbr_mask = s->nbr_mask;
    p.bnd_counter = s->d_sync; p.err_flag = s->d_sync + 1;
    const int spc = kSlabWarps * s->spw;
    p.spw = s->spw;
    p.nslices = s->nslices; p.n_bnd_tiles = s->n_bnd_slices / spc; p.tile_entries = s->tile_entries;
    {   // half of the CTAs start on the boundary tiles (the puts leave in the first half of the launch), the rest on interior
        // tiles; the boundary group then takes as many interior tiles as evens out the number of tiles per CTA
        const int ntiles = (s->nslices + spc - 1) / spc, G = s->grid;
        const double bfrac = ntiles ? (double)p.n_bnd_tiles / ntiles : 0.0;
        p.bnd_ctas = p.n_bnd_tiles == 0 ? 0 : (p.n_bnd_tiles >= ntiles || G < 2) ? G : std::min(G - 1, std::max(1, (int)(G * std::max(0.5, bfrac) + 0.5)));
        const long long quota = ((long long)ntiles + G - 1) / G;                 // tiles per CTA
        const long long share = quota * p.bnd_ctas - p.n_bnd_tiles;              // interior tiles of the boundary group
        p.bnd_group_interior = (int)std::max<long long>(0, std::min<long long>(share, ntiles - p.n_bnd_tiles));
#ifdef GLB_EXPERIMENT
        if (const char *e = getenv("GLB_SLAB_BND_FRAC")) {                      // share of the CTAs that start on the boundary tiles
            const double f = atof(e);
            if (p.n_bnd_tiles > 0 && p.n_bnd_tiles < ntiles && G >= 2) {
                p.bnd_ctas = std::min(G - (f >= 1.0 ? 0 : 1), std::max(1, (int)(G * f + 0.5)));
                const long long share2 = quota * p.bnd_ctas - p.n_bnd_tiles;
                p.bnd_group_interior = (int)std::max<long long>(0, std::min<long long>(share2, ntiles - p.n_bnd_tiles));
            }
        }
        if (const char *e = getenv("GLB_SLAB_EXP")) p.exp_flags = atoi(e);
#endif
        if (p.bnd_ctas >= G) p.bnd_group_interior = ntiles - p.n_bnd_tiles;     // no interior group: the boundary group does everything
    }
    for (int r = 0; r < s->world; ++r)
        if ((s->nbr_mask >> r) & 1u) p.peer_flag[r] = reinterpret_cast<unsigned *>(s->region[r]) + s->rank;
    for (int t = 0; t < T; ++t) {
        const int vi = t % 3, vo = (t + 1) % 3;
        p.u_in = slab_buffer(s, s->rank, vi);
        p.u_out = slab_buffer(s, s->rank, vo);
        for (int r = 0; r < s->world; ++r)
            if ((s->nbr_mask >> r) & 1u) p.peer_out[r] = slab_buffer(s, r, vo);
        p.wait_epoch = t == 0 ? 0u : s->epoch;               // version 0 is the caller's (reset + barrier before the run)
        p.signal_epoch = ++s->epoch;
        s->launches += 1;
        p.bnd_target = (unsigned)p.n_bnd_tiles * s->launches;
        void *args[] = {(void *)&p};
        GLB_CUDA(cudaLaunchKernel(s->fn, dim3(s->grid), dim3(kSlabWarps * 32), args, s->smem_bytes, st));
    }
    if (launches) *launches += T + 1;
    return 0;
}

// Zero all three label buffers (own rows, halo rows, scratch row).  Call on every rank, then synchronise the ranks
// (stream + process barrier) before glb_slab_iterate: version 0 of a run is u = 0 everywhere (ssl.py:638).
extern "C" GLB_API int glb_slab_reset(glb_slab *s, void *stream)
{
    GLB_CHECK_ARG(s && s->region[s->rank], "attach the slab first");
    GLB_CUDA(cudaMemsetAsync(slab_buffer(s, s->rank, 0), 0, 3 * (size_t)s->rows_total * s->ld * sizeof(float), (cudaStream_t)stream));
    return 0;
}

extern "C" GLB_API int glb_slab_check(glb_slab *s, void *stream)
{
    GLB_CHECK_ARG(s, "null slab");
    unsigned flag = 0;
    GLB_CUDA(cudaMemcpyAsync(&flag, s->d_sync + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    GLB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (flag) {
        cudaMemsetAsync(s->d_sync + 1, 0, sizeof(unsigned), (cudaStream_t)stream);
        set_error("glb_slab_iterate: a boundary CTA waited ~3 s for a neighbour's halo rows (a peer rank is gone or not iterating); results are invalid");
        return GLB_E_TIMEOUT;
    }
    return 0;
}

// ---- peer-mapped memory (CUDA IPC) -----------------------------------------------------------------------------------------
// One region per rank: 64 flag words + three label buffers.  The owner allocates it and hands the 64-byte handle to its
// peers (any byte transport: torch.distributed all_gather in distributed.py); a peer maps it and gets a pointer it can
// store to from kernels (NVLink P2P).
extern "C" GLB_API int glb_ipc_alloc(int64_t bytes, void **d_ptr, void *handle64)
{
    GLB_CHECK_ARG(bytes > 0 && d_ptr && handle64, "bad argument");
    void *p = nullptr;
    GLB_CUDA(cudaMalloc(&p, (size_t)bytes));
    GLB_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); set_error("glb_ipc_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); return (int)e; }
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, sizeof(h));
    *d_ptr = p;
    return 0;
}
extern "C" GLB_API int glb_ipc_open(const void *handle64, void **d_ptr)
{
    GLB_CHECK_ARG(handle64 && d_ptr, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    GLB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" GLB_API int glb_ipc_close(void *d_ptr) { if (d_ptr) GLB_CUDA(cudaIpcCloseMemHandle(d_ptr)); return 0; }
extern "C" GLB_API int glb_ipc_free(void *d_ptr) { if (d_ptr) GLB_CUDA(cudaFree(d_ptr)); return 0; }
