// knn_select.cuh - running top-C selection shared by the kNN distance kernels (SIMT and tcgen05).
//
// One thread owns one query row.  A candidate is kept when its approximate squared distance is below the
// row's threshold tau (= the C-th best seen at the last compaction, +inf before).  Kept candidates are
// appended to the row's shared-memory buffer of CAP = 2C packed keys
//      key = (float bits of d2) << 32 | candidate index        (d2 >= 0, so integer order == (d2, index) order)
// and when a buffer is nearly full the owning warp sorts it cooperatively (bitonic network in shared
// memory) and keeps the best C.  After warm-up inserts are rare (~C ln(n/C) per row in total), so the scan
// costs one compare per distance.
#pragma once
#include <stdint.h>

namespace glb {

typedef unsigned long long u64;

__device__ __forceinline__ u64 knn_key(float d2, int idx)
{
    return ((u64)__float_as_uint(fmaxf(d2, 0.f)) << 32) | (u64)(unsigned)idx;
}
__device__ __forceinline__ float knn_key_d2(u64 key) { return __uint_as_float((unsigned)(key >> 32)); }
__device__ __forceinline__ int knn_key_idx(u64 key) { return (int)(unsigned)(key & 0xffffffffull); }

// ascending bitonic sort of N (power of two) keys in shared memory by one warp
template <int N>
__device__ __forceinline__ void warp_bitonic_sort(u64 *buf, int lane)
{
#pragma unroll 1
    for (int k = 2; k <= N; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < N / 2; t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const bool up = (i & k) == 0;
                const u64 a = buf[i], b = buf[p];
                if ((a > b) == up) { buf[i] = b; buf[p] = a; }
            }
            __syncwarp();
        }
    }
}

template <int C>
struct KnnSelect {
    static constexpr int CAP = 2 * C;
    static constexpr int STRIDE = CAP + 1;          // u64 words per row; +1 spreads rows over banks
    static constexpr int MARGIN = 8;                // inserts allowed between two fullness checks

    // append (called by the row's own thread)
    static __device__ __forceinline__ void push(u64 *rowbuf, int &cnt, float d2, int idx)
    {
        rowbuf[cnt] = knn_key(d2, idx);
        ++cnt;
    }

    // warp-uniform call: compacts every row of this warp whose bit is set in `rows` (lane r <-> row r of the warp)
    static __device__ __forceinline__ void compact(u64 *warp_rows, unsigned rows, int &cnt, float &tau, int lane)
    {
        while (rows) {
            const int r = __ffs(rows) - 1;
            rows &= rows - 1;
            const int rcnt = __shfl_sync(0xffffffffu, cnt, r);
            u64 *rb = warp_rows + (size_t)r * STRIDE;
            for (int i = rcnt + lane; i < CAP; i += 32) rb[i] = ~0ull;
            __syncwarp();
            warp_bitonic_sort<CAP>(rb, lane);
            if (lane == r) {
                if (rcnt >= C) tau = knn_key_d2(rb[C - 1]);
                cnt = min(rcnt, C);
            }
        }
        __syncwarp();
    }

    // warp-uniform: after pushes, compact the rows that could overflow within the next MARGIN pushes
    static __device__ __forceinline__ void maybe_compact(u64 *warp_rows, int &cnt, float &tau, int lane)
    {
        const unsigned full = __ballot_sync(0xffffffffu, cnt > CAP - MARGIN);
        if (full) compact(warp_rows, full, cnt, tau, lane);
    }
};

}  // namespace glb
