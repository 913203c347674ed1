#include <algorithm>
#include "knn_stream.cuh"
#include "../../include/ngpd.h"

namespace ngpd {

// out_mode 0: original indices (public ABI), 1: tree positions (fused session)
template <int K>
__global__ void __launch_bounds__(128) knn_kernel(GridView g, const float* __restrict__ query, const int32_t* __restrict__ qorder,
                                                  int64_t m, int k, int skip_self, int32_t* __restrict__ idx_out,
                                                  float* __restrict__ d2_out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    int64_t qi = qorder ? (int64_t)qorder[t] : t;
    double qx = (double)__ldg(query + 3 * qi), qy = (double)__ldg(query + 3 * qi + 1), qz = (double)__ldg(query + 3 * qi + 2);
    TopK<K> top;
    top.init();
    knn_search<K>(top, g, qx, qy, qz, skip_self ? (int)qi : -1);
    int32_t* row = idx_out + qi * k;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        if (a < k) {
            int j = top.id[a];
            row[a] = j >= 0 ? __float_as_int(__ldg(&g.pts[j].w)) : g.n;
            if (d2_out) d2_out[qi * k + a] = j >= 0 ? (float)top.d[a] : INFINITY;
        }
    }
}

// fast path, tier 1: per-lane candidate streams over the 3x3x3 block + network selection (knn_stream.cuh); queries it
// cannot settle go to `fix_list`
template <int K, int R>
__device__ __forceinline__ void knn_fast_body(KsShared<R>& sm, const GridView& g, const float* __restrict__ query, const int32_t* __restrict__ qorder,
                                              int64_t t, bool active, int k, int skip_self, int32_t* __restrict__ idx_out,
                                              float* __restrict__ d2_out, int32_t* __restrict__ fail_list, int32_t* __restrict__ fail_count) {
    int64_t qi = active ? (qorder ? (int64_t)qorder[t] : t) : 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) { qx = __ldg(query + 3 * qi); qy = __ldg(query + 3 * qi + 1); qz = __ldg(query + 3 * qi + 2); }
    KsTop<K> top;
    bool ok = skip_self ? knn_stream<K, K, R, true>(top, sm, g, qx, qy, qz, active, (int)qi, INFINITY)
                        : knn_stream<K, K, R, false>(top, sm, g, qx, qy, qz, active, -1, INFINITY);
    double ex[K];
    ks_finalize<K, K>(top, g.pts, qx, qy, qz, ex);
    if (active && ok) {
        int32_t* row = idx_out + qi * k;
#pragma unroll
        for (int a = 0; a < K; ++a)
            if (a < k) {
                row[a] = __float_as_int(__ldg(&g.pts[max(top.id[a], 0)].w));
                if (d2_out) d2_out[qi * k + a] = (float)ex[a];
            }
    }
    fix_append(active && !ok, (int)t, fail_list, fail_count);
}

template <int K>
__global__ void __launch_bounds__(KsCfg<1>::THREADS, K <= 16 ? 5 : 1) knn_fast_kernel(GridView g, const float* __restrict__ query, const int32_t* __restrict__ qorder,
                                                                     int64_t m, int k, int skip_self, int32_t* __restrict__ idx_out,
                                                                     float* __restrict__ d2_out, int32_t* __restrict__ fail_list,
                                                                     int32_t* __restrict__ fail_count) {
    __shared__ KsShared<1> sm;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    knn_fast_body<K, 1>(sm, g, query, qorder, t, t < m, k, skip_self, idx_out, d2_out, fail_list, fail_count);
}

// tier 2: the queries tier 1 listed, over the 5x5x5 block
template <int K>
__global__ void __launch_bounds__(KsCfg<2>::THREADS) knn_wide_kernel(GridView g, const float* __restrict__ query, const int32_t* __restrict__ qorder,
                                                                     int k, int skip_self, int32_t* __restrict__ idx_out, float* __restrict__ d2_out,
                                                                     const int32_t* __restrict__ todo_list, const int32_t* __restrict__ todo_count,
                                                                     int32_t* __restrict__ fail_list, int32_t* __restrict__ fail_count) {
    __shared__ KsShared<2> sm;
    const int cnt = *todo_count;
    for (int base = blockIdx.x * blockDim.x; base < cnt; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        const bool active = i < cnt;
        knn_fast_body<K, 2>(sm, g, query, qorder, active ? (int64_t)todo_list[i] : 0, active, k, skip_self, idx_out, d2_out, fail_list, fail_count);
    }
}

// exact shell search for the listed queries
template <int K>
__global__ void __launch_bounds__(128) knn_fix_kernel(GridView g, const float* __restrict__ query, const int32_t* __restrict__ qorder,
                                                      int k, int skip_self, int32_t* __restrict__ idx_out, float* __restrict__ d2_out,
                                                      const int32_t* __restrict__ fix_list, const int32_t* __restrict__ fix_count) {
    const int cnt = *fix_count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        int64_t t = fix_list[i];
        int64_t qi = qorder ? (int64_t)qorder[t] : t;
        double qx = (double)__ldg(query + 3 * qi), qy = (double)__ldg(query + 3 * qi + 1), qz = (double)__ldg(query + 3 * qi + 2);
        TopK<K> top;
        top.init();
        knn_search<K>(top, g, qx, qy, qz, skip_self ? (int)qi : -1);
        int32_t* row = idx_out + qi * k;
#pragma unroll
        for (int a = 0; a < K; ++a)
            if (a < k) {
                int j = top.id[a];
                row[a] = j >= 0 ? __float_as_int(__ldg(&g.pts[j].w)) : g.n;
                if (d2_out) d2_out[qi * k + a] = j >= 0 ? (float)top.d[a] : INFINITY;
            }
    }
}

// REDUCE: the block also leaves {sum d2, sum d, max d2, rows} in part[4 * blockIdx.x ..] (fp64; added up in a fixed order afterwards)
template <bool REDUCE>
__global__ void __launch_bounds__(128) nn_sqdist_kernel(GridView g, const float* __restrict__ query, const int32_t* __restrict__ qorder,
                                                        int64_t m, float* __restrict__ d2_out, int32_t* __restrict__ idx_out,
                                                        double* __restrict__ part) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float d2 = 0.0f;
    const bool active = t < m;
    if (active) {
        int64_t qi = qorder ? (int64_t)qorder[t] : t;
        float fx = __ldg(query + 3 * qi), fy = __ldg(query + 3 * qi + 1), fz = __ldg(query + 3 * qi + 2);
        TopK<1> top;
        top.init();
        knn_search<1>(top, g, (double)fx, (double)fy, (double)fz, -1);
        int j = top.id[0];
        float4 p = __ldg(g.pts + (j >= 0 ? j : 0));
        // fp32 recomputation, as the reference does after its 1-NN lookup (Utils.py:262-263)
        float dx = p.x - fx, dy = p.y - fy, dz = p.z - fz;
        d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2_out) d2_out[qi] = d2;
        if (idx_out) idx_out[qi] = j >= 0 ? __float_as_int(p.w) : g.n;
    }
    if (REDUCE) {
        double s2 = active ? (double)d2 : 0.0, s1 = active ? (double)sqrtf(d2) : 0.0, mx = s2, cnt = active ? 1.0 : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s2 += __shfl_xor_sync(0xffffffffu, s2, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        __shared__ double red[4][4];
        const int w = threadIdx.x >> 5;
        if ((threadIdx.x & 31) == 0) { red[w][0] = s2; red[w][1] = s1; red[w][2] = mx; red[w][3] = cnt; }
        __syncthreads();
        if (threadIdx.x < 4) {
            const int c = threadIdx.x;
            double v = red[0][c];
            for (int q = 1; q < 4; ++q) v = c == 2 ? fmax(v, red[q][c]) : v + red[q][c];
            part[(int64_t)blockIdx.x * 4 + c] = v;
        }
    }
}
// one block: thread t adds blocks t, t + 256, ... in order, then the 256 partial results are added in thread order
__global__ void __launch_bounds__(256) nn_reduce_kernel(const double* __restrict__ part, int64_t blocks, double* __restrict__ acc4) {
    __shared__ double sh[256][4];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t b = threadIdx.x; b < blocks; b += 256) {
        v[0] += part[b * 4]; v[1] += part[b * 4 + 1]; v[2] = fmax(v[2], part[b * 4 + 2]); v[3] += part[b * 4 + 3];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) sh[threadIdx.x][c] = v[c];
    __syncthreads();
    if (threadIdx.x < 4) {
        const int c = threadIdx.x;
        double t = sh[0][c];
        for (int q = 1; q < 256; ++q) t = c == 2 ? fmax(t, sh[q][c]) : t + sh[q][c];
        acc4[c] = t;
    }
}

__global__ void __launch_bounds__(256) order_from_vals_kernel(const uint32_t* __restrict__ vals, int64_t m, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = (int32_t)vals[i];
}
__global__ void __launch_bounds__(256) order_from_pts_kernel(const float4* __restrict__ pts, int64_t m, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = __float_as_int(pts[i].w);
}

// visiting order of the queries: tree order when query i is tree point i, otherwise sorted by the
// queries' own cell keys; nullptr (identity) for small or already coherent batches
static int make_query_order(const ngpd_grid* G, const float* query, int64_t m, int flags, cudaStream_t stream, StreamBuf<int32_t>& order) {
    if ((flags & NGPD_KNN_COHERENT) || m < 2048) return 0;       // order stays null = identity
    NGPD_CUDA_OK(order.alloc(m));
    if ((flags & NGPD_KNN_QUERY_IS_TREE) && m == G->n) {
        order_from_pts_kernel<<<(unsigned)cdiv(m, 256), 256, 0, stream>>>(G->pts, m, order);
        NGPD_CUDA_OK(cudaGetLastError());
        return 0;
    }
    StreamBuf<uint64_t> keys(stream), keys2(stream);
    StreamBuf<uint32_t> vals(stream), vals2(stream);
    NGPD_CUDA_OK(keys.alloc(m)); NGPD_CUDA_OK(keys2.alloc(m));
    NGPD_CUDA_OK(vals.alloc(m)); NGPD_CUDA_OK(vals2.alloc(m));
    int rc = point_keys(G->v, query, m, keys, vals, stream);
    if (rc) return rc;
    bool in_tmp = false;
    rc = radix_sort_pairs(keys, vals, keys2, vals2, m, key_bits(G->v), stream, &in_tmp);
    if (rc) return rc;
    order_from_vals_kernel<<<(unsigned)cdiv(m, 256), 256, 0, stream>>>(in_tmp ? vals2.p : vals.p, m, order);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int K>
static void launch_knn(const ngpd_grid* G, const float* query, const int32_t* order, int64_t m, int k, int skip, int32_t* idx, float* d2, cudaStream_t s) {
    knn_kernel<K><<<(unsigned)cdiv(m, 128), 128, 0, s>>>(G->v, query, order, m, k, skip, idx, d2);
}

template <int K>
static int launch_knn_fast(const ngpd_grid* G, const float* query, const int32_t* order, int64_t m, int k, int skip, int32_t* idx, float* d2,
                           cudaStream_t s) {
    // two fail lists (tier 1 -> tier 2 -> exact search) and their counters
    StreamBuf<int32_t> fix(s);
    NGPD_CUDA_OK(fix.alloc((size_t)(2 * m + 2)));
    int32_t *list1 = fix.p, *list2 = fix.p + m, *cnt1 = fix.p + 2 * m, *cnt2 = fix.p + 2 * m + 1;
    NGPD_CUDA_OK(cudaMemsetAsync(cnt1, 0, 2 * sizeof(int32_t), s));
    knn_fast_kernel<K><<<(unsigned)cdiv(m, KsCfg<1>::THREADS), KsCfg<1>::THREADS, 0, s>>>(G->v, query, order, m, k, skip, idx, d2, list1, cnt1);
    int wide = (int)std::min<int64_t>(cdiv(m, KsCfg<2>::THREADS), (int64_t)num_sms() * 16);
    knn_wide_kernel<K><<<wide, KsCfg<2>::THREADS, 0, s>>>(G->v, query, order, k, skip, idx, d2, list1, cnt1, list2, cnt2);
    int blocks = (int)std::min<int64_t>(cdiv(m, 128), (int64_t)num_sms() * 8);
    knn_fix_kernel<K><<<blocks, 128, 0, s>>>(G->v, query, order, k, skip, idx, d2, list2, cnt2);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ngpd

using namespace ngpd;

extern "C" __attribute__((visibility("default"))) int ngpd_knn(const ngpd_grid_t* G, const float* query, int64_t m, int k, int flags,
                        int32_t* idx_out, float* d2_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    NGPD_REQUIRE(G && query && idx_out, "ngpd_knn: NULL argument");
    NGPD_REQUIRE(k >= 1 && k <= 64, "ngpd_knn: k must be in [1, 64]");
    if (m <= 0) return 0;
    StreamBuf<int32_t> order(stream);
    int rc = make_query_order(G, query, m, flags, stream, order);
    if (rc) return rc;
    int skip = (flags & NGPD_KNN_SKIP_SELF) ? 1 : 0;
    const bool exact_only = (flags & NGPD_KNN_EXACT_ONLY) != 0;
    if (k <= 1) launch_knn<1>(G, query, order, m, k, skip, idx_out, d2_out, stream);
    else if (k <= 4) launch_knn<4>(G, query, order, m, k, skip, idx_out, d2_out, stream);
    else if (k <= 8) { if (exact_only) launch_knn<8>(G, query, order, m, k, skip, idx_out, d2_out, stream); else rc = launch_knn_fast<8>(G, query, order, m, k, skip, idx_out, d2_out, stream); }
    else if (k <= 16) { if (exact_only) launch_knn<16>(G, query, order, m, k, skip, idx_out, d2_out, stream); else rc = launch_knn_fast<16>(G, query, order, m, k, skip, idx_out, d2_out, stream); }
    else if (k <= 32) { if (exact_only) launch_knn<32>(G, query, order, m, k, skip, idx_out, d2_out, stream); else rc = launch_knn_fast<32>(G, query, order, m, k, skip, idx_out, d2_out, stream); }
    else { if (exact_only) launch_knn<64>(G, query, order, m, k, skip, idx_out, d2_out, stream); else rc = launch_knn_fast<64>(G, query, order, m, k, skip, idx_out, d2_out, stream); }
    if (rc) return rc;
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

static int nn_sqdist_impl(const ngpd_grid_t* G, const float* query, int64_t m, int flags, float* d2_out, int32_t* idx_out, double* acc4_out,
                          cudaStream_t stream) {
    StreamBuf<int32_t> order(stream);
    int rc = make_query_order(G, query, m, flags, stream, order);
    if (rc) return rc;
    const int64_t blocks = cdiv(m, 128);
    if (acc4_out) {
        StreamBuf<double> part(stream);
        NGPD_CUDA_OK(part.alloc((size_t)blocks * 4));
        nn_sqdist_kernel<true><<<(unsigned)blocks, 128, 0, stream>>>(G->v, query, order, m, d2_out, idx_out, part);
        nn_reduce_kernel<<<1, 256, 0, stream>>>(part, blocks, acc4_out);
    } else {
        nn_sqdist_kernel<false><<<(unsigned)blocks, 128, 0, stream>>>(G->v, query, order, m, d2_out, idx_out, nullptr);
    }
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_nn_sqdist(const ngpd_grid_t* G, const float* query, int64_t m, int flags,
                              float* d2_out, int32_t* idx_out, void* stream_) {
    NGPD_REQUIRE(G && query, "ngpd_nn_sqdist: NULL argument");
    if (m <= 0) return 0;
    return nn_sqdist_impl(G, query, m, flags, d2_out, idx_out, nullptr, (cudaStream_t)stream_);
}

// the same pass with the reduction the metrics' callers apply fused in (Utils.py:253-295: .mean() of the squared distances for
// Chamfer / sCD, .max() of the distances for Hausdorff, mean distance / diagonal for PaperDistance): acc4_out (device, 4 doubles)
// = {sum d2, sum d, max d2, rows}, added up in a fixed order (reproducible); d2_out / idx_out are optional here
extern "C" __attribute__((visibility("default"))) int ngpd_nn_sqdist_reduce(const ngpd_grid_t* G, const float* query, int64_t m, int flags,
                                     float* d2_out, int32_t* idx_out, double* acc4_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    NGPD_REQUIRE(G && query && acc4_out, "ngpd_nn_sqdist_reduce: NULL argument");
    if (m <= 0) { NGPD_CUDA_OK(cudaMemsetAsync(acc4_out, 0, 4 * sizeof(double), stream)); return 0; }
    return nn_sqdist_impl(G, query, m, flags, d2_out, idx_out, acc4_out, stream);
}
