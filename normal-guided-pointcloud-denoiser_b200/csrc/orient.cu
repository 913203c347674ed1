// Consistent normal orientation on the GPU: GraphBuilder.flipNormals (GraphBuilder.py:129-209).
//   cost of an edge   1 - |n_u . n_v|                                    (calculateEdgeCost, :135-146)
//   spanning tree     minimum over that cost                               (calculateUndirectedMST, :148-175: Kruskal)
//   propagation       from the top-most point (max z, made to point up), a child is flipped when
//                     n_parent . n_child < cos(7 pi / 12)                 (flipNormalsWithMST, :177-209: DFS)
// The reference's Kruskal is an O(N^2) Python loop and its DFS is recursive; here the tree is built by Boruvka rounds
// (every component picks its cheapest outgoing edge, components merge, O(log N) rounds of edge-parallel kernels) and the
// signs are propagated breadth-first with a device-side frontier (a tree has one path from the root to a node, so DFS and
// BFS give the same signs).  Equal costs are frequent (flat regions: cost exactly 0): edges are ordered by
// (cost, lower endpoint, higher endpoint), a strict total order on undirected edges, which makes the tree unique and the
// result independent of scheduling.  (The reference's argsort is not stable, so its tree is one of the minimum trees; the
// orientations agree except where the trees differ across a crease -- see tests.)
#include "grid.cuh"
#include "../../include/ngpd.h"

namespace ngpd {

constexpr unsigned long long OR_NONE = ~0ull;

__global__ void __launch_bounds__(256) or_cost_kernel(Packed3 nrm, const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t e,
                                                      uint32_t* __restrict__ cost) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    const float c = 1.0f - fabsf(dot3(nrm((int64_t)src[i]), nrm((int64_t)dst[i])));
    // costs are >= 0 up to rounding (-1e-7 for nearly parallel unit normals): order-preserving unsigned image of the float
    const uint32_t b = __float_as_uint(c);
    cost[i] = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(256) or_init_kernel(int64_t n, int32_t* __restrict__ comp) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) comp[i] = (int32_t)i;
}
__global__ void __launch_bounds__(256) or_reset_kernel(int64_t n, uint32_t* __restrict__ best_cost, unsigned long long* __restrict__ best_pair) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { best_cost[i] = 0xffffffffu; best_pair[i] = OR_NONE; }
}
// cheapest outgoing edge of every component: first the cost ...
__global__ void __launch_bounds__(256) or_min_cost_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, const uint32_t* __restrict__ cost,
                                                          const int32_t* __restrict__ comp, int64_t e, uint32_t* __restrict__ best_cost) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    const int cu = comp[src[i]], cv = comp[dst[i]];
    if (cu == cv) return;
    const uint32_t c = cost[i];
    if (c < best_cost[cu]) atomicMin(best_cost + cu, c);
    if (c < best_cost[cv]) atomicMin(best_cost + cv, c);
}
// ... then, among the edges of that cost, the smallest (lower endpoint, higher endpoint)
__global__ void __launch_bounds__(256) or_min_pair_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ dst, const uint32_t* __restrict__ cost,
                                                          const int32_t* __restrict__ comp, int64_t e, const uint32_t* __restrict__ best_cost,
                                                          unsigned long long* __restrict__ best_pair) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    const int u = src[i], v = dst[i];
    const int cu = comp[u], cv = comp[v];
    if (cu == cv) return;
    const uint32_t c = cost[i];
    const unsigned long long pair = ((unsigned long long)(unsigned)min(u, v) << 32) | (unsigned)max(u, v);
    if (c == best_cost[cu] && pair < best_pair[cu]) atomicMin(best_pair + cu, pair);
    if (c == best_cost[cv] && pair < best_pair[cv]) atomicMin(best_pair + cv, pair);
}
// every component root hooks onto the component at the other end of its edge
__global__ void __launch_bounds__(256) or_hook_kernel(int64_t n, const int32_t* __restrict__ comp, const unsigned long long* __restrict__ best_pair,
                                                      int32_t* __restrict__ parent) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int p = (int)r;
    if (comp[r] == r && best_pair[r] != OR_NONE) {
        const int lo = (int)(best_pair[r] >> 32), hi = (int)(best_pair[r] & 0xffffffffu);
        p = comp[lo] == (int)r ? comp[hi] : comp[lo];
    }
    parent[r] = p;
}
// two components that chose the same edge point at each other: the smaller id stays a root.  Every other hook adds its
// edge to the tree.  (A strict total order on the edges rules out longer cycles.)
__global__ void __launch_bounds__(256) or_break_kernel(int64_t n, const int32_t* __restrict__ comp, const int32_t* __restrict__ parent,
                                                       const unsigned long long* __restrict__ best_pair, int32_t* __restrict__ parent_out,
                                                       int2* __restrict__ tree, int32_t* __restrict__ tree_count) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int p = parent[r];
    if (comp[r] == r && p != (int)r) {
        if (parent[p] == (int)r && (int)r < p) p = (int)r;
        else {
            const int at = atomicAdd(tree_count, 1);
            tree[at] = make_int2((int)(best_pair[r] >> 32), (int)(best_pair[r] & 0xffffffffu));
        }
    }
    parent_out[r] = p;
}
__global__ void __launch_bounds__(256) or_flatten_kernel(int64_t n, const int32_t* __restrict__ parent, int32_t* __restrict__ comp) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int r = comp[v];
    while (parent[r] != r) r = parent[r];
    comp[v] = r;
}

// tree adjacency in CSR form
__global__ void __launch_bounds__(256) or_degree_kernel(const int2* __restrict__ tree, int t, int32_t* __restrict__ deg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t) return;
    atomicAdd(deg + tree[i].x, 1);
    atomicAdd(deg + tree[i].y, 1);
}
__global__ void __launch_bounds__(256) or_fill_kernel(const int2* __restrict__ tree, int t, const int32_t* __restrict__ start, int32_t* __restrict__ cursor,
                                                      int32_t* __restrict__ adj) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t) return;
    const int a = tree[i].x, b = tree[i].y;
    adj[start[a] + atomicAdd(cursor + a, 1)] = b;
    adj[start[b] + atomicAdd(cursor + b, 1)] = a;
}

// root = first point of maximal z (torch.argmax, :205), made to point up (:206-207)
__device__ __forceinline__ int or_ordered(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__global__ void __launch_bounds__(256) or_zmax_kernel(const float* __restrict__ pos, int64_t n, int* __restrict__ zmax) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const int z = or_ordered(pos[3 * i + 2]); if (z > *zmax) atomicMax(zmax, z); }
}
__global__ void __launch_bounds__(256) or_root_kernel(const float* __restrict__ pos, int64_t n, const int* __restrict__ zmax, int* __restrict__ root) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && or_ordered(pos[3 * i + 2]) == *zmax) atomicMin(root, (int)i);
}
__global__ void or_seed_kernel(float* __restrict__ nrm, const int* __restrict__ root, uint8_t* __restrict__ seen, int32_t* __restrict__ frontier,
                               int32_t* __restrict__ counts) {
    const int r = *root;
    if (nrm[3 * r + 2] < 0.0f) { nrm[3 * r] = -nrm[3 * r]; nrm[3 * r + 1] = -nrm[3 * r + 1]; nrm[3 * r + 2] = -nrm[3 * r + 2]; }
    seen[r] = 1;
    frontier[0] = r;
    counts[0] = 1;
}
// one breadth-first level: the children of the frontier get their sign from their (final) parent
__global__ void __launch_bounds__(128) or_level_kernel(float* __restrict__ nrm, const int32_t* __restrict__ start, const int32_t* __restrict__ adj,
                                                       uint8_t* __restrict__ seen, const int32_t* __restrict__ frontier, const int32_t* __restrict__ count_in,
                                                       int32_t* __restrict__ next, int32_t* __restrict__ count_out, float thr) {
    const int cnt = *count_in;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const int p = frontier[i];
        const V3 np = v3(nrm[3 * p], nrm[3 * p + 1], nrm[3 * p + 2]);
        for (int a = start[p]; a < start[p + 1]; ++a) {
            const int c = adj[a];
            if (seen[c]) continue;                      // the parent; a tree node is discovered exactly once
            seen[c] = 1;
            const V3 nc = v3(nrm[3 * c], nrm[3 * c + 1], nrm[3 * c + 2]);
            if (dot3(np, nc) < thr) { nrm[3 * c] = -nc.x; nrm[3 * c + 1] = -nc.y; nrm[3 * c + 2] = -nc.z; }
            next[atomicAdd(count_out, 1)] = c;
        }
    }
}

}  // namespace ngpd

using namespace ngpd;

extern "C" __attribute__((visibility("default"))) int ngpd_orient_normals(const float* pos, float* nrm, int64_t n, const int32_t* edge_src,
                                                                      const int32_t* edge_dst, int64_t e, float flip_threshold, int32_t* info_out_host,
                                                                      void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(pos && nrm && (e == 0 || (edge_src && edge_dst)), "ngpd_orient_normals: NULL argument");
    NGPD_REQUIRE(n < (1ll << 31) && e < (1ll << 40), "ngpd_orient_normals: too many points");
    if (info_out_host) info_out_host[0] = info_out_host[1] = info_out_host[2] = 0;
    if (n <= 0) return 0;
    const unsigned bn = (unsigned)cdiv(n, 256), be = (unsigned)cdiv(std::max<int64_t>(e, 1), 256);
    uint32_t *cost = nullptr, *best_cost = nullptr;
    unsigned long long* best_pair = nullptr;
    int32_t *comp = nullptr, *parent = nullptr, *parent2 = nullptr, *tree_count = nullptr;
    int2* tree = nullptr;
    NGPD_CUDA_OK(cudaMallocAsync(&cost, std::max<int64_t>(e, 1) * sizeof(uint32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&best_cost, n * sizeof(uint32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&best_pair, n * sizeof(unsigned long long), st));
    NGPD_CUDA_OK(cudaMallocAsync(&comp, n * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&parent, n * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&parent2, n * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&tree, n * sizeof(int2), st));
    NGPD_CUDA_OK(cudaMallocAsync(&tree_count, sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMemsetAsync(tree_count, 0, sizeof(int32_t), st));
    if (e > 0) or_cost_kernel<<<be, 256, 0, st>>>(Packed3{nrm}, edge_src, edge_dst, e, cost);
    or_init_kernel<<<bn, 256, 0, st>>>(n, comp);
    // ---- Boruvka rounds: the number of components at least halves each time
    int t = 0, rounds = 0;
    while (e > 0) {
        or_reset_kernel<<<bn, 256, 0, st>>>(n, best_cost, best_pair);
        or_min_cost_kernel<<<be, 256, 0, st>>>(edge_src, edge_dst, cost, comp, e, best_cost);
        or_min_pair_kernel<<<be, 256, 0, st>>>(edge_src, edge_dst, cost, comp, e, best_cost, best_pair);
        or_hook_kernel<<<bn, 256, 0, st>>>(n, comp, best_pair, parent);
        or_break_kernel<<<bn, 256, 0, st>>>(n, comp, parent, best_pair, parent2, tree, tree_count);
        or_flatten_kernel<<<bn, 256, 0, st>>>(n, parent2, comp);
        int now = 0;
        NGPD_CUDA_OK(cudaMemcpyAsync(&now, tree_count, sizeof(int), cudaMemcpyDeviceToHost, st));
        NGPD_CUDA_OK(cudaStreamSynchronize(st));
        ++rounds;
        if (now == t || rounds > 64) break;             // nothing merged: every component is complete
        t = now;
    }
    NGPD_CUDA_OK(cudaGetLastError());
    // ---- tree adjacency
    int32_t *start = nullptr, *cursor = nullptr, *adj = nullptr, *frontier[2] = {nullptr, nullptr}, *counts = nullptr;
    int *zr = nullptr;
    uint8_t* seen = nullptr;
    constexpr int BATCH = 64;
    NGPD_CUDA_OK(cudaMallocAsync(&start, (n + 1) * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&cursor, n * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&adj, std::max(2 * t, 1) * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&frontier[0], n * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&frontier[1], n * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&counts, (BATCH + 1) * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMallocAsync(&zr, 2 * sizeof(int), st));
    NGPD_CUDA_OK(cudaMallocAsync(&seen, n, st));
    NGPD_CUDA_OK(cudaMemsetAsync(start, 0, (n + 1) * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMemsetAsync(cursor, 0, n * sizeof(int32_t), st));
    NGPD_CUDA_OK(cudaMemsetAsync(seen, 0, n, st));
    if (t > 0) or_degree_kernel<<<(unsigned)cdiv(t, 256), 256, 0, st>>>(tree, t, start);
    int rc = exclusive_scan_i32(start, n + 1, nullptr, st);
    if (rc) return rc;
    if (t > 0) or_fill_kernel<<<(unsigned)cdiv(t, 256), 256, 0, st>>>(tree, t, start, cursor, adj);
    // ---- root and breadth-first propagation, BATCH levels per host round trip
    const int init[2] = {(int)0x80000000, 0x7fffffff};
    NGPD_CUDA_OK(cudaMemcpyAsync(zr, init, sizeof(init), cudaMemcpyHostToDevice, st));
    or_zmax_kernel<<<bn, 256, 0, st>>>(pos, n, zr);
    or_root_kernel<<<bn, 256, 0, st>>>(pos, n, zr, zr + 1);
    NGPD_CUDA_OK(cudaMemsetAsync(counts, 0, (BATCH + 1) * sizeof(int32_t), st));
    or_seed_kernel<<<1, 1, 0, st>>>(nrm, zr + 1, seen, frontier[0], counts);
    int levels = 0, cur = 0;
    const int lb = (int)std::min<int64_t>(cdiv(n, 128), (int64_t)num_sms() * 8);
    for (;;) {
        for (int l = 0; l < BATCH; ++l) {
            or_level_kernel<<<lb, 128, 0, st>>>(nrm, start, adj, seen, frontier[cur], counts + l, frontier[cur ^ 1], counts + l + 1, flip_threshold);
            cur ^= 1;
        }
        int32_t h[BATCH + 1];
        NGPD_CUDA_OK(cudaMemcpyAsync(h, counts, sizeof(h), cudaMemcpyDeviceToHost, st));
        NGPD_CUDA_OK(cudaStreamSynchronize(st));
        for (int l = 1; l <= BATCH && h[l] > 0; ++l) ++levels;
        if (h[BATCH] == 0) break;
        // carry the last level's count over into a fresh batch
        NGPD_CUDA_OK(cudaMemsetAsync(counts, 0, (BATCH + 1) * sizeof(int32_t), st));
        NGPD_CUDA_OK(cudaMemcpyAsync(counts, &h[BATCH], sizeof(int32_t), cudaMemcpyHostToDevice, st));
        NGPD_CUDA_OK(cudaStreamSynchronize(st));
    }
    NGPD_CUDA_OK(cudaGetLastError());
    if (info_out_host) { info_out_host[0] = (int32_t)(n - t); info_out_host[1] = levels; info_out_host[2] = rounds; }
    void* bufs[] = {cost, best_cost, best_pair, comp, parent, parent2, tree, tree_count, start, cursor, adj, frontier[0], frontier[1], counts, zr, seen};
    for (void* b : bufs) NGPD_CUDA_OK(cudaFreeAsync(b, st));
    return 0;
}
