// reorder.cu - locality ordering of the graph and the device-side relabelling of CSR / label matrices.
//
// Not in the reference: scipy's csr_matvecs walks rows in storage order on one core and does not care about
// node numbering.  On the GPU every iteration gathers nnz rows of u; with a bandwidth-reducing ordering
// (reverse Cuthill-McKee) the rows gathered by one CTA overlap ~2.5-4x, so most gathers hit that SM's L1
// instead of L2.  The ordering is structural preprocessing (integers only, no arithmetic of the path):
// it runs once per graph on the host from the CSR pattern the caller already holds there; all relabelling of
// matrices and label matrices is done on the device.  Results are returned in the caller's numbering.
#include <algorithm>
#include <atomic>
#include <numeric>
#include <thread>
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
    std::vector<int> deg((size_t)n), by_deg((size_t)n), order;
    std::vector<char> seen((size_t)n, 0);
    order.reserve((size_t)n);
    for (int64_t i = 0; i < n; ++i) deg[i] = rp[i + 1] - rp[i];
    std::iota(by_deg.begin(), by_deg.end(), 0);
    std::stable_sort(by_deg.begin(), by_deg.end(), [&](int a, int b) { return deg[a] < deg[b]; });
    std::vector<int> nb;
    for (int64_t s = 0; s < n; ++s) {
        const int start = by_deg[s];
        if (seen[start]) continue;
        seen[start] = 1;
        size_t head = order.size();
        order.push_back(start);
        while (head < order.size()) {
            const int v = order[head++];
            nb.clear();
            for (int j = rp[v]; j < rp[v + 1]; ++j) {
                const int w = col[j];
                if (w >= 0 && w < n && !seen[w]) { seen[w] = 1; nb.push_back(w); }
            }
            std::sort(nb.begin(), nb.end(), [&](int a, int b) { return deg[a] != deg[b] ? deg[a] < deg[b] : a < b; });
            order.insert(order.end(), nb.begin(), nb.end());
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

// ------------------------------------------------------------------------------------------------
// Octet ordering for the dataflow iterate (poisson.cu).
//
// The dataflow kernel is bound by L1 wavefronts: one warp-wide gather instruction touches 32/LANES rows of the label
// matrix and costs one wavefront per DISTINCT 128-byte line.  The 32/LANES matrix rows that one warp processes side by
// side (a "slot") therefore should have as many neighbours in common as possible, and neighbours that are needed
// together should share a line.  This ordering makes aligned groups of 2, 4 and 8 consecutive rows out of nodes with
// many common neighbours: three rounds of heavy-edge matching on the shared-neighbour counts |N(i) & N(j)| of adjacent
// nodes (the weights of a coarser level are the sums over the pairs of members), then the octets are put in reverse
// Cuthill-McKee order of the octet graph so that a CTA's row block stays compact for L1/L2 locality.  Rows longer than
// kLongRowDf nonzeros ("long" rows, processed warp-wide by the kernel) take no part: they are spread evenly between the
// octets, so that counting SHORT rows every octet starts at a multiple of 8.  Groups that stayed incomplete go last.
// Integers only; no arithmetic of the path.  h_perm[new] = old.
namespace glb {

// fn(begin, end) over [0, n) in chunks on up to 8 host threads
template <typename F>
static void parallel_chunks(int n, int chunk, F &&fn)
{
    const int nchunks = (n + chunk - 1) / chunk;
    const int nt = std::max(1, std::min({8, (int)std::thread::hardware_concurrency(), nchunks}));
    std::atomic<int> next{0};
    auto worker = [&]() { for (int c; (c = next.fetch_add(1)) < nchunks;) fn(c * chunk, std::min(n, (c + 1) * chunk)); };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
}

struct WGraph {                       // symmetric weighted graph, CSR
    std::vector<int> rp, col;
    std::vector<float> w;
    int n() const { return (int)rp.size() - 1; }
};

// heavy-edge matching; gid[i] = group of node i, returns the number of groups.  Nodes in order of increasing degree,
// each takes its heaviest unmatched neighbour; leftovers are paired in index order.
static int match_level(const WGraph &G, std::vector<int> &gid)
{
    const int n = G.n();
    std::vector<int> mate((size_t)n, -1), order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return G.rp[a + 1] - G.rp[a] < G.rp[b + 1] - G.rp[b]; });
    for (int i : order) {
        if (mate[i] >= 0) continue;
        int best = -1;
        float bw = -1.f;
        for (int p = G.rp[i]; p < G.rp[i + 1]; ++p) {
            const int j = G.col[p];
            if (j != i && mate[j] < 0 && (G.w[p] > bw || (G.w[p] == bw && j < best))) { best = j; bw = G.w[p]; }
        }
        if (best >= 0) { mate[i] = best; mate[best] = i; }
    }
    int prev = -1;
    for (int i = 0; i < n; ++i) {
        if (mate[i] >= 0) continue;
        if (prev < 0) { prev = i; } else { mate[i] = prev; mate[prev] = i; prev = -1; }
    }
    gid.assign((size_t)n, -1);
    int g = 0;
    for (int i = 0; i < n; ++i) {
        if (gid[i] >= 0) continue;
        gid[i] = g;
        if (mate[i] >= 0) gid[mate[i]] = g;
        ++g;
    }
    return g;
}

static void coarsen(const WGraph &G, const std::vector<int> &gid, int ng, WGraph &C)
{
    const int n = G.n();
    std::vector<int> mrp((size_t)ng + 1, 0), mem((size_t)n);
    for (int i = 0; i < n; ++i) ++mrp[gid[i] + 1];
    for (int g = 0; g < ng; ++g) mrp[g + 1] += mrp[g];
    {
        std::vector<int> fill(mrp.begin(), mrp.end() - 1);
        for (int i = 0; i < n; ++i) mem[fill[gid[i]]++] = i;
    }
    // upper bound of the merged list of group g = sum of its members' degrees; merged in place in that window
    std::vector<long long> win((size_t)ng + 1, 0);
    for (int g = 0; g < ng; ++g) {
        long long d = 0;
        for (int q = mrp[g]; q < mrp[g + 1]; ++q) d += G.rp[mem[q] + 1] - G.rp[mem[q]];
        win[g + 1] = win[g] + d;
    }
    std::vector<int> tcol((size_t)win[ng]);
    std::vector<float> tw((size_t)win[ng]);
    std::vector<int> clen((size_t)ng);
    parallel_chunks(ng, 1024, [&](int g0, int g1) {
        std::vector<std::pair<int, float>> acc;
        for (int g = g0; g < g1; ++g) {
            acc.clear();
            for (int q = mrp[g]; q < mrp[g + 1]; ++q) {
                const int i = mem[q];
                for (int p = G.rp[i]; p < G.rp[i + 1]; ++p) {
                    const int h = gid[G.col[p]];
                    if (h != g) acc.emplace_back(h, G.w[p]);
                }
            }
            std::sort(acc.begin(), acc.end(), [](const std::pair<int, float> &a, const std::pair<int, float> &b) { return a.first < b.first; });
            long long o = win[g];
            for (size_t q = 0; q < acc.size();) {
                size_t e = q;
                float s = 0.f;
                while (e < acc.size() && acc[e].first == acc[q].first) s += acc[e++].second;
                tcol[o] = acc[q].first; tw[o] = s; ++o;
                q = e;
            }
            clen[g] = (int)(o - win[g]);
        }
    });
    C.rp.assign((size_t)ng + 1, 0);
    for (int g = 0; g < ng; ++g) C.rp[g + 1] = C.rp[g] + clen[g];
    C.col.resize((size_t)C.rp[ng]); C.w.resize((size_t)C.rp[ng]);
    for (int g = 0; g < ng; ++g) {
        std::copy(tcol.begin() + win[g], tcol.begin() + win[g] + clen[g], C.col.begin() + C.rp[g]);
        std::copy(tw.begin() + win[g], tw.begin() + win[g] + clen[g], C.w.begin() + C.rp[g]);
    }
}

}  // namespace glb

extern "C" GLB_API int glb_octet_order_host(const int32_t *h_rowptr, const int32_t *h_col, int64_t n, int32_t *h_perm)
{
    GLB_CHECK_ARG(h_rowptr && h_perm && (h_col || h_rowptr[n] == 0), "null pointer");
    GLB_CHECK_ARG(n > 0 && n < (1ll << 31), "n out of range");
    const int N = (int)n;
    PhaseTimer tm("octet order");
    // symmetric simple adjacency of the SHORT rows (sorted, unique, no self loops, no long rows)
    std::vector<char> is_short((size_t)N);
    std::vector<int> sid((size_t)N, -1), short_nodes, long_nodes;
    for (int i = 0; i < N; ++i) {
        is_short[i] = h_rowptr[i + 1] - h_rowptr[i] <= GLB_DATAFLOW_LONG_ROW;
        if (is_short[i]) { sid[i] = (int)short_nodes.size(); short_nodes.push_back(i); } else long_nodes.push_back(i);
    }
    const int ns = (int)short_nodes.size();
    WGraph G;
    {
        std::vector<int> cnt((size_t)ns + 1, 0);
        auto each_edge = [&](auto &&f) {
            for (int i = 0; i < N; ++i) {
                if (!is_short[i]) continue;
                for (int p = h_rowptr[i]; p < h_rowptr[i + 1]; ++p) {
                    const int j = h_col[p];
                    if (j < 0 || j >= N || j == i || !is_short[j]) continue;
                    f(sid[i], sid[j]);
                }
            }
        };
        each_edge([&](int a, int b) { ++cnt[a + 1]; ++cnt[b + 1]; });
        for (int i = 0; i < ns; ++i) cnt[i + 1] += cnt[i];
        std::vector<int> adj((size_t)cnt[ns]), fill(cnt.begin(), cnt.end() - 1);
        each_edge([&](int a, int b) { adj[fill[a]++] = b; adj[fill[b]++] = a; });
        G.rp.assign((size_t)ns + 1, 0);
        G.col.reserve(adj.size() / 2 + 16);
        std::vector<int> ulen((size_t)ns);
        parallel_chunks(ns, 2048, [&](int i0, int i1) {
            for (int i = i0; i < i1; ++i) {
                std::sort(adj.begin() + cnt[i], adj.begin() + cnt[i + 1]);
                ulen[i] = (int)(std::unique(adj.begin() + cnt[i], adj.begin() + cnt[i + 1]) - (adj.begin() + cnt[i]));
            }
        });
        for (int i = 0; i < ns; ++i) {
            G.col.insert(G.col.end(), adj.begin() + cnt[i], adj.begin() + cnt[i] + ulen[i]);
            G.rp[i + 1] = (int)G.col.size();
        }
        // weight of edge (i,j) = 1 + number of common neighbours (sorted-list intersection)
        G.w.resize(G.col.size());
        parallel_chunks(ns, 2048, [&](int i0, int i1) {
        for (int i = i0; i < i1; ++i)
            for (int p = G.rp[i]; p < G.rp[i + 1]; ++p) {
                const int j = G.col[p];
                int a = G.rp[i], b = G.rp[j], c = 0;
                const int ae = G.rp[i + 1], be = G.rp[j + 1];
                while (a < ae && b < be) {
                    const int x = G.col[a], y = G.col[b];
                    c += x == y; a += x <= y; b += y <= x;
                }
                G.w[p] = 1.f + (float)c;
            }
        });
    }
    tm.lap("adjacency + shared-neighbour weights");
    // three levels of matching: groups of up to 2, 4, 8 short rows
    std::vector<std::vector<int>> members((size_t)ns);
    for (int i = 0; i < ns; ++i) members[i].push_back(short_nodes[i]);
    WGraph C;
    for (int lvl = 0; lvl < 3; ++lvl) {
        std::vector<int> gid;
        const int ng = match_level(G, gid);
        coarsen(G, gid, ng, C);
        std::vector<std::vector<int>> nm((size_t)ng);
        for (size_t i = 0; i < members.size(); ++i) nm[gid[i]].insert(nm[gid[i]].end(), members[i].begin(), members[i].end());
        members.swap(nm);
        std::swap(G, C);
        tm.lap("matching level");
    }
    // octets in reverse Cuthill-McKee order of the octet graph; complete octets first
    std::vector<int> gorder;
    if (G.n() > 0) rcm_order(G.rp.data(), G.col.data(), G.n(), gorder);
    std::vector<int> full, rest;
    for (int g : gorder) (members[g].size() == 8 ? full : rest).push_back(g);
    size_t at = 0, next_long = 0;
    const size_t nl = long_nodes.size(), nf = full.size();
    for (size_t k = 0; k < nf; ++k) {
        for (int v : members[full[k]]) h_perm[at++] = v;
        // long rows spread evenly between the octets
        const size_t upto = nf ? (k + 1) * nl / nf : nl;
        while (next_long < upto) h_perm[at++] = long_nodes[next_long++];
    }
    while (next_long < nl) h_perm[at++] = long_nodes[next_long++];
    for (int g : rest) for (int v : members[g]) h_perm[at++] = v;
    tm.lap("octet RCM + output");
    if ((int64_t)at != n) { set_error("glb_octet_order_host: internal error (%zu of %lld rows placed)", at, (long long)n); return GLB_E_INVALID; }
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
