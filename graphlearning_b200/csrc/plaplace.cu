// plaplace.cu - the p-Laplace / AMLE neighbour sweeps of the reference's C extension on sm_100a.
//
// Replaces c_code/lp_iterate.cpp (reached through cextensions.lp_iterate / lip_iterate,
// c_code/cextensions.cpp:19-107, from graph.plaplace / graph.amle, graphlearning/graph.py:1177-1332):
//   lp_iterate_main            :35-125   Jacobi sweeps of an upper and a lower barrier function
//   lip_iterate_main           :129-187  Gauss-Seidel sweeps, game-theoretic p-Laplacian with 0/1 weights in L_inf
//   lip_iterate_weighted_main  :190-259  Gauss-Seidel sweeps, weighted infinity Laplacian by 30-step bisection
//
// All arithmetic is fp64 with explicit round-to-nearest intrinsics (no FMA contraction) and every per-row sum
// runs in stored order, so the iterates are BIT-IDENTICAL to the reference's (built without -ffast-math).
//
// Jacobi (lp): one persistent cooperative launch, one thread per row and round, (upper, lower) interleaved as one
// 16-byte cell so a neighbour costs one gather for both functions; one grid barrier per sweep which also carries
// the sweep's error (max(uu - ul)) for the reference's stopping rule `err < tol && it > 10`.  The reference swaps
// its buffer pointers after every sweep (:116-123), so the caller's arrays hold the result of the last ODD sweep;
// the two device buffers play exactly the same roles and buffer 0 is what is returned.
//
// Gauss-Seidel (lip): the reference updates u in place in row order, so row i sees the NEW value of neighbours j < i
// and the OLD value of neighbours j >= i.  That is a dependency DAG (depth 37 on the 70k-node benchmark graph), not
// a sequential chain.  The kernel runs it as a dataflow: every value lives in a 16-byte cell {value, version}
// written and read as ONE 16-byte transaction at L2; sweep t reads neighbours j >= i from buffer t&1 (complete
// since the barrier that ended sweep t-1) and neighbours j < i from buffer (t+1)&1, re-polling a cell until its
// version is t+1.  Threads own rows in increasing order, so the smallest unfinished (sweep,row) never waits:
// no deadlock (all CTAs co-resident: cooperative launch).  One grid barrier per sweep carries the error for the
// stopping rule `err < tol && it > 20`.  Result: exactly the reference's sequential sweep, at any sweep count.
#include <math.h>
#include <stdlib.h>
#include <vector>
#include "common.cuh"

namespace glb {
namespace {

struct __align__(16) Cell { double a; unsigned long long b; };     // lip: {u, version};  lp: {uu, bits(ul)}

__device__ __forceinline__ Cell ld_cell(const Cell *p)
{
    Cell c;
    unsigned long long x;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(x), "=l"(c.b) : "l"(p) : "memory");
    c.a = __longlong_as_double((long long)x);
    return c;
}
__device__ __forceinline__ void st_cell(Cell *p, double a, unsigned long long b)
{
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"((unsigned long long)__double_as_longlong(a)), "l"(b) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Watchdog of the waiting loops (a wait normally ends within microseconds): after ~2 s without progress the flag is
// raised, every other loop sees it within a thousand spins, the kernel drains and the host entry point reports
// GLB_E_TIMEOUT instead of leaving a hung GPU.  counter[1] is the flag (counter[0] the barrier counter).
constexpr long long kWaitLimit = 4000000000ll;
__device__ __forceinline__ bool wait_expired(unsigned *flag, unsigned &spins, long long &t0)
{
    if ((++spins & 1023u) != 0u) return false;
    if (t0 == 0) { t0 = clock64(); return false; }
    if (ld_relaxed_u32(flag) != 0u) return true;
    if (clock64() - t0 > kWaitLimit) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory"); return true; }
    return false;
}

// Grid barrier that also returns the maximum of one non-negative double per thread (the sweep's error).
// slots[3]: sweep t accumulates into slots[t % 3]; slot (t+1) % 3 is cleared during sweep t (its last readers left
// it before the barrier that ended sweep t-1).
__device__ __forceinline__ double barrier_max(double mine, unsigned long long *slots, unsigned *counter, unsigned t)
{
    __shared__ unsigned long long s_max;
    unsigned long long bits = (unsigned long long)__double_as_longlong(mine);    // mine >= 0: bit order = value order
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
    }
    if (threadIdx.x == 0) s_max = 0ull;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && bits) atomicMax(&s_max, bits);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) slots[(t + 1) % 3] = 0ull;
        if (s_max) atomicMax(slots + t % 3, s_max);
        fence_gpu();
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(counter), "r"(1u) : "memory");
        const unsigned want = (t + 1) * gridDim.x;
        unsigned spins = 0;
        long long t0 = 0;
        while (ld_relaxed_u32(counter) < want && !wait_expired(counter + 1, spins, t0)) { }
        fence_gpu();
        s_max = ld_relaxed_u64(slots + t % 3);
    }
    __syncthreads();
    const double r = __longlong_as_double((long long)s_max);
    __syncthreads();                                   // s_max is rewritten by the next call
    return r;
}

// row offsets from the row-sorted COO row index: start[i] = first k with row[k] >= i   (lp_iterate.cpp:48-57)
__global__ void __launch_bounds__(256) row_start_kernel(const int *__restrict__ row, int M, int n, int *__restrict__ start)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        int lo = 0, hi = M;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(row + mid) < i) lo = mid + 1; else hi = mid;
        }
        start[i] = lo;
    }
}

// status[0] |= 1 when the row index is not sorted / out of range, |= 2 when a neighbour index is out of range
__global__ void __launch_bounds__(256) check_coo_kernel(const int *__restrict__ row, const int *__restrict__ nbr, int M, int n, int *status)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < M; k += gridDim.x * blockDim.x) {
        const int r = row[k];
        if (r < 0 || r >= n || (k > 0 && row[k - 1] > r)) atomicOr(status, 1);
        const int j = nbr[k];
        if (j < 0 || j >= n) atomicOr(status, 2);
    }
}

__global__ void __launch_bounds__(256) max_weight_kernel(const double *__restrict__ W, int M, unsigned long long *out)
{
    double mx = 0.0;                                                       // maxWGT starts at 0 (:60-62)
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < M; k += gridDim.x * blockDim.x) {
        const double w = W[k];
        if (w > mx) mx = w;
    }
    if (mx > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(mx));
}

// ------------------------------------------------------------------------------------------------------------
// Jacobi: lp_iterate_main
// ------------------------------------------------------------------------------------------------------------
struct LpArgs {
    const int *start, *nbr;
    const double *W;
    const int *lab;                 // lab[i] = index into labval or -1
    const double *labval;
    Cell *c0, *c1;
    double *invdeg;
    unsigned long long *slots;      // [3] error slots + [1] max weight at slots[3]
    unsigned *counter;
    int *sweeps;
    int n, T;
    double p, tol;
};

__global__ void __launch_bounds__(256) lp_jacobi_kernel(LpArgs A)
{
    const int NT = gridDim.x * blockDim.x, gt = blockIdx.x * blockDim.x + threadIdx.x;
    const double alpha = __ddiv_rn(1.0, A.p);
    const double delta = __dsub_rn(1.0, __ddiv_rn(2.0, A.p));
    double dt = __ddiv_rn(0.9, __dadd_rn(alpha, __dmul_rn(2.0, delta)));
    dt = __ddiv_rn(dt, __longlong_as_double((long long)A.slots[3]));
    for (int i = gt; i < A.n; i += NT) {                                   // invdeg[i] = alpha / sum_j W_ij (:47-58)
        double s = 0.0;
        for (int k = A.start[i]; k < A.start[i + 1]; ++k) s = __dadd_rn(s, A.W[k]);
        A.invdeg[i] = __ddiv_rn(alpha, s);
    }
    int it = 0, done = 0;
    for (; it < A.T; ++it) {
        const Cell *cur = (it & 1) ? A.c1 : A.c0;
        Cell *nxt = (it & 1) ? A.c0 : A.c1;
        double err = 0.0;
        for (int i = gt; i < A.n; i += NT) {
            const Cell me = ld_cell(cur + i);
            const double uu = me.a, ul = __longlong_as_double((long long)me.b);
            double minu = 0.0, maxu = 0.0, sumu = 0.0, minl = 0.0, maxl = 0.0, suml = 0.0;
            const int e = A.start[i + 1];
            for (int k0 = A.start[i]; k0 < e; k0 += 8) {                   // 8 gathers in flight, consumed in stored order
                Cell c[8];
                double w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (k0 + q < e) { c[q] = ld_cell(cur + __ldg(A.nbr + k0 + q)); w[q] = __ldg(A.W + k0 + q); }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (k0 + q >= e) continue;
                    const double du = __dmul_rn(w[q], __dsub_rn(c[q].a, uu));
                    const double dl = __dmul_rn(w[q], __dsub_rn(__longlong_as_double((long long)c[q].b), ul));
                    minu = du < minu ? du : minu;  maxu = du > maxu ? du : maxu;  sumu = __dadd_rn(sumu, du);
                    minl = dl < minl ? dl : minl;  maxl = dl > maxl ? dl : maxl;  suml = __dadd_rn(suml, dl);
                }
            }
            const double idg = A.invdeg[i];
            double vu = __dadd_rn(uu, __dmul_rn(dt, __dadd_rn(__dmul_rn(idg, sumu), __dmul_rn(delta, __dadd_rn(minu, maxu)))));
            double vl = __dadd_rn(ul, __dmul_rn(dt, __dadd_rn(__dmul_rn(idg, suml), __dmul_rn(delta, __dadd_rn(minl, maxl)))));
            const double d = __dsub_rn(uu, ul);
            if (d > err) err = d;
            const int l = A.lab[i];
            if (l >= 0) vu = vl = A.labval[l];                             // Dirichlet values into the new buffer (:106-110)
            st_cell(nxt + i, vu, (unsigned long long)__double_as_longlong(vl));
        }
        const double gerr = barrier_max(err, A.slots, A.counter, (unsigned)it);
        if (gerr < A.tol && it > 10) { done = it + 1; break; }
    }
    if (gt == 0) *A.sweeps = done ? done : it;
}

__global__ void __launch_bounds__(256) lp_pack_kernel(const double *uu, const double *ul, Cell *c0, Cell *c1, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        c0[i].a = uu[i]; c0[i].b = (unsigned long long)__double_as_longlong(ul[i]);
        c1[i].a = 0.0;   c1[i].b = 0ull;                                   // the reference's vu, vl start at 0 (:69-70)
    }
}
__global__ void __launch_bounds__(256) lp_unpack_kernel(const Cell *c0, double *uu, double *ul, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uu[i] = c0[i].a; ul[i] = __longlong_as_double((long long)c0[i].b);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Gauss-Seidel dataflow: lip_iterate_main / lip_iterate_weighted_main
// ------------------------------------------------------------------------------------------------------------
struct LipArgs {
    const int *start, *nbr;
    const double *W;
    const int *lab;
    const int *order, *crit;        // level schedule: row of sorted position p, and the row's latest-level producer (or -1)
    Cell *c0, *c1;
    double *u_out;
    unsigned long long *slots;
    unsigned *counter;
    int *sweeps;
    int n, M, T, n_active;
    int use_crit;                   // poll the latest-level producer alone before gathering the whole row
    double tol, alpha, beta;
};

constexpr unsigned long long kVerFixed = ~0ull;      // Dirichlet rows: valid in every sweep
constexpr int kLipBatch = 16;                        // neighbour cells in flight per retry (most rows: one round trip)
constexpr int kLipCap = 32;                          // neighbour values kept in local memory for the bisection

// One row per lane and round, rows dealt to lanes in LEVEL order (level = longest chain of same-sweep producers, computed
// on the host per call, a warp never mixes levels): the lanes of a warp become ready together, first consume their
// neighbours and then run the update - for AMLE the 30-step bisection - in lockstep; a waiting lane polls only its
// latest-level producer.  A lane must never spin on a cell: its producer may be another lane of the same warp
// (a lower-numbered neighbour in the same round), and a lane that leaves a spin loop waits at the loop's
// reconvergence point for the lanes still inside.  So the warp runs a retry loop instead: per pass every pending lane
// consumes, IN STORED ORDER, as many of its neighbours as are ready (16 gathers in flight), keeps its running
// (min, max, sum, degree) in registers, and finishes the row once all neighbours have been consumed.
template <bool WEIGHTED, bool LOCKSTEP>
__global__ void __launch_bounds__(256, 2) lip_gauss_seidel_kernel(LipArgs A)
{
    const int NT = gridDim.x * blockDim.x, gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    int it = 0, done = 0;
    for (; it < A.T; ++it) {
        const Cell *cur = (it & 1) ? A.c1 : A.c0;
        Cell *nxt = (it & 1) ? A.c0 : A.c1;
        const unsigned long long want = (unsigned long long)it + 1ull;
        double err = 0.0;
        for (int p0 = gt - lane; p0 < A.n_active; p0 += NT) {                // warp-uniform trip count
            // the schedule pads every level to a multiple of 32 (-1 entries): a warp holds rows of ONE level, which
            // have no dependencies among themselves
            const int i = p0 + lane < A.n_active ? __ldg(A.order + p0 + lane) : -1;
            const bool active = i >= 0;
            bool pending = active;
            int s = 0, L = 0, k = 0, crit = -1;
            double uv[kLipCap], wv[kLipCap];
            double minu = 0.0, maxu = 0.0, sumu = 0.0, deg = 0.0, uold = 0.0;
            bool empty_row = false;
            if (active) {
                s = A.start[i];
                L = A.start[i + 1] - s;
                crit = A.use_crit ? __ldg(A.crit + i) : -1;
                uold = ld_cell(cur + i).a;
                if (L == 0) {
                    // the reference reads u[I[start[i]]] (the next row's first neighbour) even for an empty row (:163, :223):
                    // treat that entry as the row's only "neighbour" for min/max and skip the sums
                    empty_row = true;
                    if (s < A.M) L = 1; else { minu = maxu = qnan; }
                }
            }
            auto update_row = [&]() {

                double ne;
                if (!WEIGHTED) {
                    ne = __dadd_rn(__ddiv_rn(__dmul_rn(A.alpha, sumu), deg), __ddiv_rn(__dmul_rn(A.beta, __dadd_rn(minu, maxu)), 2.0));
                } else {
                    double a = minu, b = maxu;
                    const int Lr = empty_row ? 0 : L;
                    const int Lc = Lr < kLipCap ? Lr : kLipCap;
                    for (int r = 0; r < 30; ++r) {                         // bisection on min_j w(t-u_j) + max_j w(t-u_j) (:229-243)
                        const double tm = __ddiv_rn(__dadd_rn(a, b), 2.0);
                        double minw = 0.0, maxw = 0.0;
                        for (int kk = 0; kk < Lc; ++kk) {
                            const double d = __dmul_rn(wv[kk], __dsub_rn(tm, uv[kk]));
                            minw = d < minw ? d : minw;
                            maxw = d > maxw ? d : maxw;
                        }
                        for (int kk = kLipCap; kk < Lr; ++kk) {            // beyond the local cache: final for this sweep, read again
                            const int j = __ldg(A.nbr + s + kk);
                            const double d = __dmul_rn(__ldg(A.W + s + kk), __dsub_rn(tm, ld_cell((j < i ? nxt : cur) + j).a));
                            minw = d < minw ? d : minw;
                            maxw = d > maxw ? d : maxw;
                        }
                        if (__dadd_rn(minw, maxw) > 0.0) b = tm; else a = tm;
                    }
                    ne = __ddiv_rn(__dadd_rn(a, b), 2.0);
                }
                double d = __dsub_rn(uold, ne);
                d = d < 0.0 ? -d : d;
                if (d > err) err = d;
                st_cell(nxt + i, ne, want);
            };
            // phase 1: consume the neighbours in stored order as they become ready (retry loop, never a spin)
            unsigned spins = 0;
            long long tw0 = 0;
            while (__any_sync(0xffffffffu, pending)) {
                if (!pending) continue;
                if (wait_expired(A.counter + 1, spins, tw0)) { pending = false; continue; }
                if (crit >= 0) {                                  // one cheap poll on the producer expected last
                    if (ld_cell(nxt + crit).b < want) continue;
                    crit = -1;
                }
                int jj[kLipBatch];
                Cell c[kLipBatch];
#pragma unroll
                for (int q = 0; q < kLipBatch; ++q) {
                    jj[q] = k + q < L ? __ldg(A.nbr + s + k + q) : -1;
                    if (jj[q] >= 0) c[q] = ld_cell((jj[q] < i ? nxt : cur) + jj[q]);
                }
                bool stop = false;
#pragma unroll
                for (int q = 0; q < kLipBatch; ++q) {
                    if (stop || jj[q] < 0) continue;
                    if (jj[q] < i && c[q].b < want) { stop = true; continue; }   // producer not there yet: retry from here
                    const double v = c[q].a;
                    if (k == 0) { minu = v; maxu = v; }
                    if (!empty_row) {
                        const double w = __ldg(A.W + s + k);
                        if (!WEIGHTED) {
                            sumu = __dadd_rn(sumu, __dmul_rn(w, v));
                            deg = __dadd_rn(deg, w);
                        } else if (k < kLipCap) {
                            uv[k] = v; wv[k] = w;
                        }
                    }
                    minu = v < minu ? v : minu;
                    maxu = v > maxu ? v : maxu;
                    ++k;
                }
                if (k < L) continue;
                pending = false;
                if (!LOCKSTEP) update_row();                      // store at once: the next level is waiting for it
            }
            // lockstep variant: the whole warp updates its rows together (one pass through the 30-step bisection)
            if (LOCKSTEP && active) update_row();
        }
        const double gerr = barrier_max(err, A.slots, A.counter, (unsigned)it);
        if (gerr < A.tol && it > 20) { done = it + 1; break; }
    }
    const int nsweeps = done ? done : it;
    const Cell *fin = (nsweeps & 1) ? A.c1 : A.c0;      // sweep t writes buffer (t+1)&1
    for (int i = gt; i < A.n; i += NT) A.u_out[i] = ld_cell(fin + i).a;
    if (gt == 0) *A.sweeps = nsweeps;
}

__global__ void __launch_bounds__(256) lip_pack_kernel(const double *u, const int *lab, const double *labval, Cell *c0, Cell *c1, int n)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int l = lab[i];
        if (l >= 0) {
            c0[i].a = labval[l]; c0[i].b = kVerFixed;
            c1[i].a = labval[l]; c1[i].b = kVerFixed;
        } else {
            c0[i].a = u[i]; c0[i].b = 0ull;
            c1[i].a = u[i]; c1[i].b = 0ull;
        }
    }
}


// ------------------------------------------------------------------------------------------------------------
// Batched Gauss-Seidel: c right-hand sides (the one-vs-rest classes of ssl.plaplace / ssl.amle) in one launch
// ------------------------------------------------------------------------------------------------------------
// The classes share the graph AND the set of Dirichlet rows, so they share the dependency DAG: one thread per (row, class),
// cells at index row * c + class, and every latency hop of the sweep carries c values instead of one.  Each class keeps
// the reference's own stopping rule: the per-sweep barrier carries one error per class, a class whose error drops below
// tol (after sweep 20) is frozen at that sweep - exactly where the reference's separate run would have stopped - while
// the others go on.  Results per class are bit-identical to c separate calls.
struct LipMultiArgs {
    const int *start, *nbr;
    const double *W;
    const int *lab, *order;
    Cell *c0, *c1;
    double *u_out;
    unsigned long long *errs;       // [3][c] per-sweep, per-class error bits
    unsigned *counter;              // [0] barrier, [1] watchdog
    int *sweeps;                    // [c]
    int n, M, T, n_active, c;
    double tol, alpha, beta;
};

constexpr int kMaxClasses = 32;

template <bool WEIGHTED>
__global__ void __launch_bounds__(256, 2) lip_multi_kernel(LipMultiArgs A)
{
    __shared__ unsigned long long s_err[kMaxClasses];
    const int c = A.c;
    const int NT = (gridDim.x * blockDim.x) / (32 * c) * (32 * c);          // stride: a thread keeps its class in every round
    const int gt = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool worker = gt < NT;                                            // threads beyond the stride only take part in barriers
    const int cls = gt % c;
    const long long total = (long long)A.n_active * c;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    unsigned stopped_mask = 0u;                                             // classes frozen so far (same in every thread)
    int my_stop = 0;                                                        // sweeps executed by this thread's class (0 = still running)
    int it = 0;
    for (; it < A.T && stopped_mask != (c == 32 ? 0xffffffffu : ((1u << c) - 1u)); ++it) {
        const Cell *cur = (it & 1) ? A.c1 : A.c0;
        Cell *nxt = (it & 1) ? A.c0 : A.c1;
        const unsigned long long want = (unsigned long long)it + 1ull;
        double err = 0.0;
        const bool running = worker && !my_stop;
        for (long long p0 = gt - lane; p0 < total; p0 += NT) {              // warp-uniform trip count
            const long long idx = p0 + lane;
            bool pending = running && idx < total;
            const int i = pending ? __ldg(A.order + (int)(idx / c)) : 0;
            int s = 0, L = 0, k = 0;
            double uv[kLipCap], wv[kLipCap];
            double minu = 0.0, maxu = 0.0, sumu = 0.0, deg = 0.0, uold = 0.0;
            bool empty_row = false;
            if (pending) {
                s = A.start[i];
                L = A.start[i + 1] - s;
                uold = ld_cell(cur + (size_t)i * c + cls).a;
                if (L == 0) {
                    empty_row = true;
                    if (s < A.M) L = 1; else { minu = maxu = qnan; }
                }
            }
            unsigned spins = 0;
            long long tw0 = 0;
            while (__any_sync(0xffffffffu, pending)) {
                if (!pending) continue;
                if (wait_expired(A.counter + 1, spins, tw0)) { pending = false; continue; }
                int jj[kLipBatch];
                Cell cc[kLipBatch];
#pragma unroll
                for (int q = 0; q < kLipBatch; ++q) {
                    jj[q] = k + q < L ? __ldg(A.nbr + s + k + q) : -1;
                    if (jj[q] >= 0) cc[q] = ld_cell((jj[q] < i ? nxt : cur) + (size_t)jj[q] * c + cls);
                }
                bool halt = false;
#pragma unroll
                for (int q = 0; q < kLipBatch; ++q) {
                    if (halt || jj[q] < 0) continue;
                    if (jj[q] < i && cc[q].b < want) { halt = true; continue; }
                    const double v = cc[q].a;
                    if (k == 0) { minu = v; maxu = v; }
                    if (!empty_row) {
                        const double w = __ldg(A.W + s + k);
                        if (!WEIGHTED) {
                            sumu = __dadd_rn(sumu, __dmul_rn(w, v));
                            deg = __dadd_rn(deg, w);
                        } else if (k < kLipCap) {
                            uv[k] = v; wv[k] = w;
                        }
                    }
                    minu = v < minu ? v : minu;
                    maxu = v > maxu ? v : maxu;
                    ++k;
                }
                if (k < L) continue;
                pending = false;
                double ne;
                if (!WEIGHTED) {
                    ne = __dadd_rn(__ddiv_rn(__dmul_rn(A.alpha, sumu), deg), __ddiv_rn(__dmul_rn(A.beta, __dadd_rn(minu, maxu)), 2.0));
                } else {
                    double a = minu, b = maxu;
                    const int Lr = empty_row ? 0 : L;
                    const int Lc = Lr < kLipCap ? Lr : kLipCap;
                    for (int r = 0; r < 30; ++r) {
                        const double tm = __ddiv_rn(__dadd_rn(a, b), 2.0);
                        double minw = 0.0, maxw = 0.0;
                        for (int kk = 0; kk < Lc; ++kk) {
                            const double d = __dmul_rn(wv[kk], __dsub_rn(tm, uv[kk]));
                            minw = d < minw ? d : minw;
                            maxw = d > maxw ? d : maxw;
                        }
                        for (int kk = kLipCap; kk < Lr; ++kk) {
                            const int j = __ldg(A.nbr + s + kk);
                            const double d = __dmul_rn(__ldg(A.W + s + kk), __dsub_rn(tm, ld_cell((j < i ? nxt : cur) + (size_t)j * c + cls).a));
                            minw = d < minw ? d : minw;
                            maxw = d > maxw ? d : maxw;
                        }
                        if (__dadd_rn(minw, maxw) > 0.0) b = tm; else a = tm;
                    }
                    ne = __ddiv_rn(__dadd_rn(a, b), 2.0);
                }
                double d = __dsub_rn(uold, ne);
                d = d < 0.0 ? -d : d;
                if (d > err) err = d;
                st_cell(nxt + (size_t)i * c + cls, ne, want);
            }
        }
        // barrier that carries one error per class: block-level max in shared memory, one atomic per class and block
        if (threadIdx.x < kMaxClasses) s_err[threadIdx.x] = 0ull;
        __syncthreads();
        if (running && err > 0.0) atomicMax(&s_err[cls], (unsigned long long)__double_as_longlong(err));
        __syncthreads();
        unsigned long long *slot = A.errs + (size_t)(it % 3) * c;
        if ((int)threadIdx.x < c) {
            if (blockIdx.x == 0) A.errs[(size_t)((it + 1) % 3) * c + threadIdx.x] = 0ull;
            if (s_err[threadIdx.x]) atomicMax(slot + threadIdx.x, s_err[threadIdx.x]);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_gpu();
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(A.counter), "r"(1u) : "memory");
            const unsigned wantc = (unsigned)(it + 1) * gridDim.x;
            unsigned spins = 0;
            long long t0 = 0;
            while (ld_relaxed_u32(A.counter) < wantc && !wait_expired(A.counter + 1, spins, t0)) { }
            fence_gpu();
        }
        __syncthreads();
        for (int k2 = 0; k2 < c; ++k2) {                                   // every thread takes the same decisions
            if (stopped_mask & (1u << k2)) continue;
            const double e = __longlong_as_double((long long)ld_relaxed_u64(slot + k2));
            if (e < A.tol && it > 20) {
                stopped_mask |= 1u << k2;
                if (k2 == cls) my_stop = it + 1;
            }
        }
        __syncthreads();
    }
    const int nsweeps = my_stop ? my_stop : it;                             // `it` = T, or the sweep at which the last class stopped
    if (worker) {
        const Cell *fin = (nsweeps & 1) ? A.c1 : A.c0;
        for (long long idx = gt; idx < (long long)A.n * c; idx += NT) A.u_out[idx] = ld_cell(fin + idx).a;   // idx % c == cls
        if (gt < c) A.sweeps[gt] = nsweeps;
    }
}

__global__ void __launch_bounds__(256)
lip_multi_pack_kernel(const double *u, const int *lab, const double *labval, Cell *c0, Cell *c1, int n, int c)
{
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)n * c; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / c), k = (int)(idx % c);
        const int l = lab[i];
        if (l >= 0) {
            c0[idx].a = labval[(size_t)l * c + k]; c0[idx].b = kVerFixed;
            c1[idx].a = labval[(size_t)l * c + k]; c1[idx].b = kVerFixed;
        } else {
            c0[idx].a = u[idx]; c0[idx].b = 0ull;
            c1[idx].a = u[idx]; c1[idx].b = 0ull;
        }
    }
}

struct Arena {
    std::vector<void *> ptrs;
    ~Arena() { for (void *p : ptrs) dev_free(p); }
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

int coop_grid(const void *fn, int threads, int *grid)
{
    int per_sm = 0;
    GLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, 0));
    if (per_sm < 1) { set_error("kernel does not fit on an SM"); return GLB_E_INVALID; }
    *grid = per_sm * sm_count();
    return 0;
}

// label map: lab[i] = LAST j with ind[j] == i (the reference's loops assign in order, the last one wins), -1 otherwise
int build_labels(const int32_t *h_ind, int n, int m, std::vector<int> &lab)
{
    lab.assign((size_t)n, -1);
    for (int j = 0; j < m; ++j) {
        int i = h_ind[j];
        if (i < 0) i += n;                                        // numpy-style negative indices never reach the C code; be lenient
        if (i < 0 || i >= n) { set_error("boundary index %d out of range", h_ind[j]); return GLB_E_INVALID; }
        lab[i] = j;
    }
    return 0;
}

struct Common {
    Arena A;
    int *start = nullptr, *nbr = nullptr, *row = nullptr, *lab = nullptr, *status = nullptr, *sweeps = nullptr;
    double *W = nullptr, *labval = nullptr;
    Cell *c0 = nullptr, *c1 = nullptr;
    unsigned long long *slots = nullptr;
    unsigned *counter = nullptr;
};

int upload_common(Common &C, const int32_t *h_nbr, const int32_t *h_row, const double *h_w, const int32_t *h_ind,
                  const double *h_val, int n, int M, int m, cudaStream_t st, int *nl, std::vector<int> &lab)
{
    int rc = build_labels(h_ind, n, m, lab);
    if (rc) return rc;
    GLB_CUDA(C.A.alloc(&C.start, (size_t)n + 1)); GLB_CUDA(C.A.alloc(&C.nbr, (size_t)M)); GLB_CUDA(C.A.alloc(&C.row, (size_t)M));
    GLB_CUDA(C.A.alloc(&C.W, (size_t)M));         GLB_CUDA(C.A.alloc(&C.lab, (size_t)n)); GLB_CUDA(C.A.alloc(&C.labval, (size_t)m));
    GLB_CUDA(C.A.alloc(&C.c0, (size_t)n));        GLB_CUDA(C.A.alloc(&C.c1, (size_t)n));
    GLB_CUDA(C.A.alloc(&C.slots, 4));             GLB_CUDA(C.A.alloc(&C.counter, 2));
    GLB_CUDA(C.A.alloc(&C.status, 1));            GLB_CUDA(C.A.alloc(&C.sweeps, 1));
    GLB_CUDA(cudaMemcpyAsync(C.nbr, h_nbr, (size_t)M * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(C.row, h_row, (size_t)M * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(C.W, h_w, (size_t)M * sizeof(double), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(C.lab, lab.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    if (m) GLB_CUDA(cudaMemcpyAsync(C.labval, h_val, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemsetAsync(C.slots, 0, 4 * sizeof(unsigned long long), st));
    GLB_CUDA(cudaMemsetAsync(C.counter, 0, 2 * sizeof(unsigned), st));
    GLB_CUDA(cudaMemsetAsync(C.status, 0, sizeof(int), st));
    GLB_CUDA(cudaMemsetAsync(C.sweeps, 0, sizeof(int), st));
    const int gb = sm_count() * 4;
    if (M) check_coo_kernel<<<gb, 256, 0, st>>>(C.row, C.nbr, M, n, C.status);
    row_start_kernel<<<gb, 256, 0, st>>>(C.row, M, n, C.start);
    if (M) max_weight_kernel<<<gb, 256, 0, st>>>(C.W, M, C.slots + 3);
    *nl += 3;
    GLB_LAUNCH_CHECK();
    int status = 0;
    GLB_CUDA(cudaMemcpyAsync(&status, C.status, sizeof(int), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    if (status & 1) { set_error("row index array must be sorted ascending with entries in [0,n) (graph.__ccode_init__ order)"); return GLB_E_INVALID; }
    if (status & 2) { set_error("neighbour index out of range"); return GLB_E_INVALID; }
    return 0;
}

// Level schedule of one Gauss-Seidel sweep (arrays validated by upload_common before this runs): level[i] = 1 + the
// largest level among the unlabelled neighbours j < i of the unlabelled row i (those are the values row i must wait for
// inside a sweep), rows sorted by (level, index), every level padded with -1 to a multiple of 32; crit[i] = a neighbour attaining that maximum, -1 without producers.
void level_schedule(const int32_t *h_nbr, const int32_t *h_row, const std::vector<int> &lab, int n, int M,
                    std::vector<int> &order, std::vector<int> &crit, int *depth)
{
    std::vector<int> start((size_t)n + 1, 0), level((size_t)n, -1);
    for (int k = 0; k < M; ++k) ++start[(size_t)h_row[k] + 1];
    for (int i = 0; i < n; ++i) start[i + 1] += start[i];
    crit.assign((size_t)n, -1);
    int maxl = -1;
    std::vector<int> count;
    for (int i = 0; i < n; ++i) {
        if (lab[i] >= 0) continue;
        int s = start[i], e = start[i + 1];
        if (e == s && s < M) e = s + 1;                        // the empty row's stray read of the next entry
        int l = -1, c = -1;
        for (int k = s; k < e; ++k) {
            const int j = h_nbr[k];
            if (j < i && lab[j] < 0 && level[j] > l) { l = level[j]; c = j; }
        }
        level[i] = l + 1;
        crit[i] = c;
        if (level[i] > maxl) { maxl = level[i]; count.resize((size_t)maxl + 1, 0); }
        ++count[level[i]];
    }
    std::vector<int> pos((size_t)maxl + 2, 0);
    for (int l = 0; l <= maxl; ++l) pos[l + 1] = pos[l] + ((count[l] + 31) & ~31);      // a warp never straddles two levels
    order.assign((size_t)pos[maxl + 1], -1);
    for (int i = 0; i < n; ++i)
        if (level[i] >= 0) order[pos[level[i]]++] = i;
    *depth = maxl + 1;
}

int no_gpu()
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device visible (there is no CPU fallback)");
        return GLB_E_NOGPU;
    }
    return 0;
}

}  // namespace
}  // namespace glb

using namespace glb;

extern "C" GLB_API int glb_lp_iterate_host(double *h_uu, double *h_ul, const int32_t *h_nbr, const int32_t *h_row,
                                           const double *h_w, const int32_t *h_ind, const double *h_val, double p, int T,
                                           double tol, int n, int M, int m, int *sweeps, int *launches)
{
    GLB_CHECK_ARG(h_uu && h_ul && (M == 0 || (h_nbr && h_row && h_w)) && (m == 0 || (h_ind && h_val)), "null pointer");
    GLB_CHECK_ARG(n > 0 && M >= 0 && m >= 0 && T >= 0, "size out of range");
    int rc = no_gpu();
    if (rc) return rc;
    cudaStream_t st = 0;
    int nl = 0;
    Common C;
    std::vector<int> lab;
    if ((rc = upload_common(C, h_nbr, h_row, h_w, h_ind, h_val, n, M, m, st, &nl, lab))) return rc;
    double *uu, *ul, *invdeg;
    GLB_CUDA(C.A.alloc(&uu, (size_t)n)); GLB_CUDA(C.A.alloc(&ul, (size_t)n)); GLB_CUDA(C.A.alloc(&invdeg, (size_t)n));
    GLB_CUDA(cudaMemcpyAsync(uu, h_uu, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(ul, h_ul, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    const int gb = sm_count() * 4;
    lp_pack_kernel<<<gb, 256, 0, st>>>(uu, ul, C.c0, C.c1, n);
    LpArgs A{C.start, C.nbr, C.W, C.lab, C.labval, C.c0, C.c1, invdeg, C.slots, C.counter, C.sweeps, n, T, p, tol};
    int grid = 0;
    if ((rc = coop_grid((const void *)lp_jacobi_kernel, 256, &grid))) return rc;
    grid = std::min(grid, std::max(1, ceil_div(n, 256)));
    void *args[] = {&A};
    GLB_CUDA(cudaLaunchCooperativeKernel((const void *)lp_jacobi_kernel, dim3(grid), dim3(256), args, 0, st));
    lp_unpack_kernel<<<gb, 256, 0, st>>>(C.c0, uu, ul, n);
    nl += 3;
    GLB_LAUNCH_CHECK();
    int sw = 0;
    GLB_CUDA(cudaMemcpyAsync(h_uu, uu, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaMemcpyAsync(h_ul, ul, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaMemcpyAsync(&sw, C.sweeps, sizeof(int), cudaMemcpyDeviceToHost, st));
    unsigned timed_out = 0;
    GLB_CUDA(cudaMemcpyAsync(&timed_out, C.counter + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    if (timed_out) { set_error("%s: a waiting loop of the sweep kernel hit its watchdog; results are invalid", __func__); return GLB_E_TIMEOUT; }
    if (sweeps) *sweeps = sw;
    if (launches) *launches = nl;
    return 0;
}

extern "C" GLB_API int glb_lip_iterate_host(double *h_u, const int32_t *h_nbr, const int32_t *h_row, const double *h_w,
                                            const int32_t *h_ind, const double *h_val, int T, double tol, int weighted,
                                            double alpha, double beta, int n, int M, int m, int *sweeps, int *launches)
{
    GLB_CHECK_ARG(h_u && (M == 0 || (h_nbr && h_row && h_w)) && (m == 0 || (h_ind && h_val)), "null pointer");
    GLB_CHECK_ARG(n > 0 && M >= 0 && m >= 0 && T >= 0, "size out of range");
    int rc = no_gpu();
    if (rc) return rc;
    cudaStream_t st = 0;
    int nl = 0;
    Common C;
    std::vector<int> lab, order, crit;
    if ((rc = upload_common(C, h_nbr, h_row, h_w, h_ind, h_val, n, M, m, st, &nl, lab))) return rc;
    // Schedule (the alternatives - sweeps overlapping through a ring of version buffers, per-level completion counters - were
    // bit-identical and not faster, profiles/r1_lip_schedules.txt; they are gone from the product):
    //   AMLE (weighted): rows dealt to warps by level of the dependency DAG, a waiting lane polls only its latest-level
    //                    producer, the warp runs the 30-step bisection in lockstep with all 32 lanes busy;
    //   unweighted:      rows in natural order, every lane gathers and stores on its own - the update is a handful of
    //                    flops, so the shortest path from "last neighbour ready" to "value published" wins.
    const bool by_level = weighted != 0;
    int depth = 0;
    level_schedule(h_nbr, h_row, lab, n, M, order, crit, &depth);
    if (!by_level) {
        order.clear();
        for (int i = 0; i < n; ++i) if (lab[i] < 0) order.push_back(i);
    }
    double *u, *u_out;
    int *d_order, *d_crit;
    GLB_CUDA(C.A.alloc(&u, (size_t)n)); GLB_CUDA(C.A.alloc(&u_out, (size_t)n));
    GLB_CUDA(C.A.alloc(&d_order, order.size())); GLB_CUDA(C.A.alloc(&d_crit, (size_t)n));
    if (!order.empty()) GLB_CUDA(cudaMemcpyAsync(d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(d_crit, crit.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(u, h_u, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
    const int gb = sm_count() * 4;
    lip_pack_kernel<<<gb, 256, 0, st>>>(u, C.lab, C.labval, C.c0, C.c1, n);
    LipArgs A{C.start, C.nbr, C.W, C.lab, d_order, d_crit, C.c0, C.c1, u_out, C.slots, C.counter, C.sweeps, n, M, T,
              (int)order.size(), by_level ? 1 : 0, tol, alpha, beta};
    const void *fn = weighted ? (const void *)lip_gauss_seidel_kernel<true, true> : (const void *)lip_gauss_seidel_kernel<false, false>;
    int grid = 0;
    if ((rc = coop_grid(fn, 256, &grid))) return rc;
    grid = std::min(grid, std::max(1, ceil_div(n, 256)));
    void *args[] = {&A};
    GLB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, st));
    nl += 2;
    GLB_LAUNCH_CHECK();
    int sw = 0;
    GLB_CUDA(cudaMemcpyAsync(h_u, u_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaMemcpyAsync(&sw, C.sweeps, sizeof(int), cudaMemcpyDeviceToHost, st));
    unsigned timed_out = 0;
    GLB_CUDA(cudaMemcpyAsync(&timed_out, C.counter + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    if (timed_out) { set_error("%s: a waiting loop of the sweep kernel hit its watchdog; results are invalid", __func__); return GLB_E_TIMEOUT; }
    if (sweeps) *sweeps = sw;
    if (launches) *launches = nl;
    return 0;
}

extern "C" GLB_API int glb_lip_iterate_multi_host(double *h_u, const int32_t *h_nbr, const int32_t *h_row, const double *h_w,
                                                  const int32_t *h_ind, const double *h_val, int T, double tol, int weighted,
                                                  double alpha, double beta, int n, int M, int m, int c, int *sweeps, int *launches)
{
    GLB_CHECK_ARG(h_u && (M == 0 || (h_nbr && h_row && h_w)) && (m == 0 || (h_ind && h_val)), "null pointer");
    GLB_CHECK_ARG(n > 0 && M >= 0 && m >= 0 && T >= 0, "size out of range");
    GLB_CHECK_ARG(c >= 1 && c <= kMaxClasses && (long long)n * c < (1ll << 31), "1 <= c <= 32 right-hand sides");
    int rc = no_gpu();
    if (rc) return rc;
    cudaStream_t st = 0;
    int nl = 0;
    Common C;
    std::vector<int> lab, order;
    // upload_common allocates single-class cells and uploads the first column of val; the per-class arrays follow here
    if ((rc = upload_common(C, h_nbr, h_row, h_w, h_ind, h_val, n, M, 0, st, &nl, lab))) return rc;
    if ((rc = build_labels(h_ind, n, m, lab))) return rc;
    for (int i = 0; i < n; ++i) if (lab[i] < 0) order.push_back(i);
    double *u, *u_out, *labval;
    int *d_order, *d_lab, *d_sweeps;
    Cell *c0, *c1;
    unsigned long long *errs;
    const size_t nc = (size_t)n * c;
    GLB_CUDA(C.A.alloc(&u, nc));            GLB_CUDA(C.A.alloc(&u_out, nc));       GLB_CUDA(C.A.alloc(&labval, (size_t)m * c));
    GLB_CUDA(C.A.alloc(&d_order, order.size())); GLB_CUDA(C.A.alloc(&d_lab, (size_t)n)); GLB_CUDA(C.A.alloc(&d_sweeps, (size_t)c));
    GLB_CUDA(C.A.alloc(&c0, nc));           GLB_CUDA(C.A.alloc(&c1, nc));          GLB_CUDA(C.A.alloc(&errs, (size_t)3 * c));
    GLB_CUDA(cudaMemcpyAsync(u, h_u, nc * sizeof(double), cudaMemcpyHostToDevice, st));
    if (m) GLB_CUDA(cudaMemcpyAsync(labval, h_val, (size_t)m * c * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!order.empty()) GLB_CUDA(cudaMemcpyAsync(d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemcpyAsync(d_lab, lab.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
    GLB_CUDA(cudaMemsetAsync(errs, 0, (size_t)3 * c * sizeof(unsigned long long), st));
    GLB_CUDA(cudaMemsetAsync(d_sweeps, 0, (size_t)c * sizeof(int), st));
    const int gb = sm_count() * 4;
    lip_multi_pack_kernel<<<gb, 256, 0, st>>>(u, d_lab, labval, c0, c1, n, c);
    LipMultiArgs A{C.start, C.nbr, C.W, d_lab, d_order, c0, c1, u_out, errs, C.counter, d_sweeps, n, M, T, (int)order.size(), c,
                   tol, alpha, beta};
    const void *fn = weighted ? (const void *)lip_multi_kernel<true> : (const void *)lip_multi_kernel<false>;
    int grid = 0;
    if ((rc = coop_grid(fn, 256, &grid))) return rc;
    grid = std::min(grid, std::max(ceil_div(32 * c, 256) + 1, ceil_div((int64_t)n * c, 256)));
    void *args[] = {&A};
    GLB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, 0, st));
    nl += 2;
    GLB_LAUNCH_CHECK();
    std::vector<int> sw((size_t)c, 0);
    GLB_CUDA(cudaMemcpyAsync(h_u, u_out, nc * sizeof(double), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaMemcpyAsync(sw.data(), d_sweeps, (size_t)c * sizeof(int), cudaMemcpyDeviceToHost, st));
    unsigned timed_out = 0;
    GLB_CUDA(cudaMemcpyAsync(&timed_out, C.counter + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    GLB_CUDA(cudaStreamSynchronize(st));
    if (timed_out) { set_error("%s: a waiting loop of the sweep kernel hit its watchdog; results are invalid", __func__); return GLB_E_TIMEOUT; }
    if (sweeps) for (int k = 0; k < c; ++k) sweeps[k] = sw[k];
    if (launches) *launches = nl;
    return 0;
}
