// Kernel group 1 of the north star: spatial index construction (bbox, cell keys, LSD radix sort with
// warp-level digit ranking, prefix scans, brick tables).  All kernels are hand-written; no CUB/Thrust.
#include <cstdarg>
#include <cstring>
#include <cfloat>
#include <vector>
#include "grid.cuh"
#include "../../include/ngpd.h"

namespace ngpd {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

// ------------------------------------------------------------------------------------------------
// bounding box
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
static inline float ordered_float(int i) {
    int j = i >= 0 ? i : i ^ 0x7fffffff;
    float f;
    memcpy(&f, &j, 4);
    return f;
}

__global__ void __launch_bounds__(256) bbox_kernel(const float* __restrict__ pos, int64_t n, int* __restrict__ out6) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool bad = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __ldg(pos + 3 * i + c);
            bad |= !isfinite(v);
            lo[c] = fminf(lo[c], v);
            hi[c] = fmaxf(hi[c], v);
        }
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(out6 + 6, 1);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            atomicMin(out6 + c, float_ordered(lo[c]));
            atomicMax(out6 + 3 + c, float_ordered(hi[c]));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan (int32), three-phase, recursive on the block sums
// ------------------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256, SC_ITEMS = 16, SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& block_total) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < (SC_THREADS / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane < (SC_THREADS / 32)) warp_sums[lane] = s;
    }
    __syncthreads();
    int base = w > 0 ? warp_sums[w - 1] : 0;
    block_total = warp_sums[SC_THREADS / 32 - 1];
    return base + inc - v;
}

__global__ void __launch_bounds__(SC_THREADS) scan_reduce_kernel(const int* __restrict__ data, int64_t n, int* __restrict__ block_sums) {
    __shared__ int ws[SC_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    int s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i)
        if (base + i < n) s += data[base + i];
    int total;
    block_exclusive_scan(s, ws, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SC_THREADS) scan_apply_kernel(int* __restrict__ data, int64_t n, const int* __restrict__ block_offsets) {
    __shared__ int ws[SC_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
    int v[SC_ITEMS];
    int s = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        v[i] = (base + i < n) ? data[base + i] : 0;
        s += v[i];
    }
    int total;
    int run = block_exclusive_scan(s, ws, total) + (block_offsets ? block_offsets[blockIdx.x] : 0);
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
}

// in-place exclusive scan; when total_out_host != nullptr the grand total is copied back (synchronises)
int exclusive_scan_i32(int* data, int64_t n, int* total_out_host, cudaStream_t stream) {
    if (n <= 0) {
        if (total_out_host) *total_out_host = 0;
        return 0;
    }
    int last = 0;
    if (total_out_host) NGPD_CUDA_OK(cudaMemcpyAsync(&last, data + (n - 1), 4, cudaMemcpyDeviceToHost, stream));
    int64_t nb = cdiv(n, SC_TILE);
    if (nb == 1) {
        scan_apply_kernel<<<1, SC_THREADS, 0, stream>>>(data, n, nullptr);
    } else {
        int* sums = nullptr;
        NGPD_CUDA_OK(cudaMallocAsync(&sums, nb * sizeof(int), stream));
        scan_reduce_kernel<<<(unsigned)nb, SC_THREADS, 0, stream>>>(data, n, sums);
        int rc = exclusive_scan_i32(sums, nb, nullptr, stream);
        if (rc) return rc;
        scan_apply_kernel<<<(unsigned)nb, SC_THREADS, 0, stream>>>(data, n, sums);
        NGPD_CUDA_OK(cudaFreeAsync(sums, stream));
    }
    NGPD_CUDA_OK(cudaGetLastError());
    if (total_out_host) {
        int ex = 0;
        NGPD_CUDA_OK(cudaMemcpyAsync(&ex, data + (n - 1), 4, cudaMemcpyDeviceToHost, stream));
        NGPD_CUDA_OK(cudaStreamSynchronize(stream));
        *total_out_host = ex + last;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass.
//   count:   per-block digit histogram, stored digit-major so one scan yields global offsets
//   scatter: stable ranking inside the block with warp match-any (one shared counter row per warp),
//            then direct scatter.  Stability: rounds, lanes, warps and blocks are all visited in
//            input order.
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32, RS_ITEMS = 16, RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS) rs_count_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                              int* __restrict__ counts, int nblocks) {
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int i = 0; i < RS_ITEMS; ++i) {
        int64_t idx = base + i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&hist[(int)((keys[idx] >> shift) & 255)], 1);
    }
    __syncthreads();
    counts[(int64_t)threadIdx.x * nblocks + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                int64_t n, int shift, const int* __restrict__ offsets, int nblocks) {
    __shared__ int cnt[RS_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    uint64_t k[RS_ITEMS];
    uint32_t v[RS_ITEMS];
    int rank[RS_ITEMS];
    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (32 * RS_ITEMS);
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t idx = wbase + r * 32 + lane;
        bool valid = idx < n;
        k[r] = valid ? keys_in[idx] : 0;
        v[r] = valid ? vals_in[idx] : 0;
        int d = valid ? (int)((k[r] >> shift) & 255) : 256;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        int b = 0;
        if (lane == leader && valid) {
            b = cnt[w][d];
            cnt[w][d] = b + __popc(peers);
        }
        b = __shfl_sync(0xffffffffu, b, leader);
        rank[r] = b + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    {
        int d = threadIdx.x;
        int run = offsets[(int64_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) {
            int c = cnt[ww][d];
            cnt[ww][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t idx = wbase + r * 32 + lane;
        if (idx < n) {
            int d = (int)((k[r] >> shift) & 255);
            int64_t p = (int64_t)cnt[w][d] + rank[r];
            keys_out[p] = k[r];
            vals_out[p] = v[r];
        }
    }
}

int radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp, int64_t n,
                     int significant_bits, cudaStream_t stream, bool* result_in_tmp) {
    *result_in_tmp = false;
    if (n <= 1) return 0;
    int nblocks = (int)cdiv(n, RS_TILE);
    int* counts = nullptr;
    NGPD_CUDA_OK(cudaMallocAsync(&counts, (size_t)256 * nblocks * sizeof(int), stream));
    uint64_t *ki = keys, *ko = keys_tmp;
    uint32_t *vi = vals, *vo = vals_tmp;
    for (int shift = 0; shift < significant_bits; shift += 8) {
        rs_count_kernel<<<nblocks, RS_THREADS, 0, stream>>>(ki, n, shift, counts, nblocks);
        int rc = exclusive_scan_i32(counts, (int64_t)256 * nblocks, nullptr, stream);
        if (rc) return rc;
        rs_scatter_kernel<<<nblocks, RS_THREADS, 0, stream>>>(ki, vi, ko, vo, n, shift, counts, nblocks);
        NGPD_CUDA_OK(cudaGetLastError());
        uint64_t* tk = ki; ki = ko; ko = tk;
        uint32_t* tv = vi; vi = vo; vo = tv;
        *result_in_tmp = !*result_in_tmp;
    }
    NGPD_CUDA_OK(cudaFreeAsync(counts, stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// keys, gather, brick tables
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) keys_kernel(GridView g, const float* __restrict__ pos, int64_t n,
                                                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = cell_of((double)__ldg(pos + 3 * i), g.ox, g.inv_h, g.nx);
    int cy = cell_of((double)__ldg(pos + 3 * i + 1), g.oy, g.inv_h, g.ny);
    int cz = cell_of((double)__ldg(pos + 3 * i + 2), g.oz, g.inv_h, g.nz);
    keys[i] = cell_key(cx, cy, cz);
    vals[i] = (uint32_t)i;
}

int point_keys(const GridView& g, const float* pos, int64_t n, uint64_t* keys, uint32_t* vals, cudaStream_t stream) {
    if (n > 0) keys_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(g, pos, n, keys, vals);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

int key_bits(const GridView& g) {
    int m = g.tbx > g.tby ? g.tbx : g.tby;
    m = m > g.tbz ? m : g.tbz;
    int b = 0;
    while ((1 << b) < m) ++b;
    return 9 + 3 * b;
}

__global__ void __launch_bounds__(256) gather_points_kernel(const float* __restrict__ pos, const uint32_t* __restrict__ vals,
                                                            const uint64_t* __restrict__ keys, int64_t n,
                                                            float4* __restrict__ pts, int* __restrict__ brick_flag) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint32_t o = vals[s];
    pts[s] = make_float4(__ldg(pos + 3 * (int64_t)o), __ldg(pos + 3 * (int64_t)o + 1), __ldg(pos + 3 * (int64_t)o + 2), __int_as_float((int)o));
    brick_flag[s] = (s == 0 || (keys[s] >> 9) != (keys[s - 1] >> 9)) ? 1 : 0;
}

// brick_flag has been exclusive-scanned: brick id of sorted point s = scan[s] (+1 if s starts a brick... see below)
__global__ void __launch_bounds__(256) brick_starts_kernel(const uint64_t* __restrict__ keys, const int* __restrict__ scan, int64_t n,
                                                           int* __restrict__ brick_start, int nbricks, GridView g, int* __restrict__ top) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    bool first = (s == 0) || (keys[s] >> 9) != (keys[s - 1] >> 9);
    if (first) {
        int b = scan[s];   // exclusive scan of the flags: number of brick starts before s
        brick_start[b] = (int)s;
        uint64_t bk = keys[s] >> 9;
        int bx = (int)compact3(bk), by = (int)compact3(bk >> 1), bz = (int)compact3(bk >> 2);
        top[((int64_t)bz * g.tby + by) * g.tbx + bx] = b;
    }
    if (s == n - 1) brick_start[nbricks] = (int)n;
}

__global__ void __launch_bounds__(128) fine_table_kernel(const uint64_t* __restrict__ keys, const int* __restrict__ brick_start,
                                                         int nbricks, int* __restrict__ fine, unsigned long long* __restrict__ occupied) {
    int b = blockIdx.x;
    int lo0 = brick_start[b], hi0 = brick_start[b + 1];
    uint64_t bk = keys[lo0] >> 9;
    __shared__ int tab[513];
    for (int c = threadIdx.x; c <= 512; c += blockDim.x) {
        int lo = lo0, hi = hi0;
        if (c < 512) {
            uint64_t want = (bk << 9) | (uint64_t)c;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (keys[mid] < want) lo = mid + 1; else hi = mid;
            }
        } else {
            lo = hi0;
        }
        tab[c] = lo;
        fine[(int64_t)b * 513 + c] = lo;
    }
    __syncthreads();
    int occ = 0;
    for (int c = threadIdx.x; c < 512; c += blockDim.x) occ += tab[c + 1] > tab[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) occ += __shfl_xor_sync(0xffffffffu, occ, o);
    if ((threadIdx.x & 31) == 0 && occ) atomicAdd(occupied, (unsigned long long)occ);
}

static void destroy_buffers(ngpd_grid* g) {
    if (g->pts) cudaFree(g->pts);
    if (g->top) cudaFree(g->top);
    if (g->fine) cudaFree(g->fine);
    g->pts = nullptr; g->top = nullptr; g->fine = nullptr;
}

static int build_once(ngpd_grid* G, const float* pos, int64_t n, double h, cudaStream_t stream) {
    destroy_buffers(G);
    GridView& v = G->v;
    const float* bb = G->bbox;
    v.ox = bb[0]; v.oy = bb[1]; v.oz = bb[2];
    v.h = h; v.inv_h = 1.0 / h;
    auto dim = [&](float lo, float hi) {
        double e = ((double)hi - (double)lo) * v.inv_h;
        int64_t c = (int64_t)floor(e) + 1;
        return (int)(c < 1 ? 1 : c);
    };
    v.nx = dim(bb[0], bb[3]); v.ny = dim(bb[1], bb[4]); v.nz = dim(bb[2], bb[5]);
    v.tbx = (v.nx + 7) / 8; v.tby = (v.ny + 7) / 8; v.tbz = (v.nz + 7) / 8;
    v.n = (int)n;
    int64_t ntop = (int64_t)v.tbx * v.tby * v.tbz;
    NGPD_CUDA_OK(cudaMalloc(&G->top, ntop * sizeof(int)));
    NGPD_CUDA_OK(cudaMemsetAsync(G->top, 0xff, ntop * sizeof(int), stream));
    // padding entries: the streaming k-NN kernels load up to KS_PAD candidates per step and mask the overrun
    NGPD_CUDA_OK(cudaMalloc(&G->pts, ((size_t)n + KS_PAD) * sizeof(float4)));
    NGPD_CUDA_OK(cudaMemsetAsync(G->pts + n, 0, KS_PAD * sizeof(float4), stream));
    uint64_t *keys = nullptr, *keys2 = nullptr;
    uint32_t *vals = nullptr, *vals2 = nullptr;
    int *flag = nullptr, *brick_start = nullptr;
    unsigned long long* occ = nullptr;
    NGPD_CUDA_OK(cudaMalloc(&keys, n * 8)); NGPD_CUDA_OK(cudaMalloc(&keys2, n * 8));
    NGPD_CUDA_OK(cudaMalloc(&vals, n * 4)); NGPD_CUDA_OK(cudaMalloc(&vals2, n * 4));
    NGPD_CUDA_OK(cudaMalloc(&flag, n * 4));
    NGPD_CUDA_OK(cudaMalloc(&occ, 8));
    NGPD_CUDA_OK(cudaMemsetAsync(occ, 0, 8, stream));
    int rc = point_keys(v, pos, n, keys, vals, stream);
    if (rc) return rc;
    bool in_tmp = false;
    rc = radix_sort_pairs(keys, vals, keys2, vals2, n, key_bits(v), stream, &in_tmp);
    if (rc) return rc;
    const uint64_t* sk = in_tmp ? keys2 : keys;
    const uint32_t* sv = in_tmp ? vals2 : vals;
    unsigned nb = (unsigned)cdiv(n, 256);
    gather_points_kernel<<<nb, 256, 0, stream>>>(pos, sv, sk, n, G->pts, flag);
    int nbricks = 0;
    rc = exclusive_scan_i32(flag, n, &nbricks, stream);
    if (rc) return rc;
    G->nbricks = nbricks;
    NGPD_CUDA_OK(cudaMalloc(&brick_start, ((size_t)nbricks + 1) * sizeof(int)));
    NGPD_CUDA_OK(cudaMalloc(&G->fine, (size_t)nbricks * 513 * sizeof(int)));
    brick_starts_kernel<<<nb, 256, 0, stream>>>(sk, flag, n, brick_start, nbricks, v, G->top);
    fine_table_kernel<<<nbricks, 128, 0, stream>>>(sk, brick_start, nbricks, G->fine, occ);
    NGPD_CUDA_OK(cudaGetLastError());
    unsigned long long occ_h = 0;
    NGPD_CUDA_OK(cudaMemcpyAsync(&occ_h, occ, 8, cudaMemcpyDeviceToHost, stream));
    NGPD_CUDA_OK(cudaStreamSynchronize(stream));
    G->occupied_cells = (int64_t)occ_h;
    cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(vals2); cudaFree(flag); cudaFree(brick_start); cudaFree(occ);
    v.pts = G->pts; v.top = G->top; v.fine = G->fine;
    G->bytes = (int64_t)n * 16 + ntop * 4 + (int64_t)nbricks * 513 * 4;
    return 0;
}

}  // namespace ngpd

using namespace ngpd;

extern "C" __attribute__((visibility("default"))) const char* ngpd_last_error(void) { return ngpd::get_error(); }
extern "C" __attribute__((visibility("default"))) int ngpd_version(void) { return 100; }

// The library's scratch memory is stream-ordered (cudaMallocAsync).  By default the driver's pool hands freed memory
// back to the OS at every synchronisation, so each call would pay for mapping it again; keep it cached instead.
static void keep_scratch_pool_cached() {
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    done[dev] = true;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
}

extern "C" __attribute__((visibility("default"))) int ngpd_grid_create(const float* pos, int64_t n, float cell_size, int k_hint, void* stream_, ngpd_grid_t** out) {
    cudaStream_t stream = (cudaStream_t)stream_;
    keep_scratch_pool_cached();
    NGPD_REQUIRE(out != nullptr, "ngpd_grid_create: out is NULL");
    *out = nullptr;
    NGPD_REQUIRE(pos != nullptr && n > 0, "ngpd_grid_create: empty point set");
    NGPD_REQUIRE(n < (int64_t)2147483647, "ngpd_grid_create: at most 2^31-1 points per grid");
    ngpd_grid* G = new ngpd_grid();
    G->n = n;
    // bounding box
    int* bb_d = nullptr;
    NGPD_CUDA_OK(cudaMalloc(&bb_d, 7 * sizeof(int)));
    int init[7] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000, (int)0x80800000, (int)0x80800000, 0};
    for (int c = 3; c < 6; ++c) { float f = -FLT_MAX; int i; memcpy(&i, &f, 4); init[c] = i >= 0 ? i : i ^ 0x7fffffff; }
    NGPD_CUDA_OK(cudaMemcpyAsync(bb_d, init, sizeof(init), cudaMemcpyHostToDevice, stream));
    int blocks = (int)std::min<int64_t>(cdiv(n, 256), (int64_t)num_sms() * 8);
    bbox_kernel<<<blocks, 256, 0, stream>>>(pos, n, bb_d);
    int bb_h[7];
    NGPD_CUDA_OK(cudaMemcpyAsync(bb_h, bb_d, sizeof(bb_h), cudaMemcpyDeviceToHost, stream));
    NGPD_CUDA_OK(cudaStreamSynchronize(stream));
    cudaFree(bb_d);
    if (bb_h[6]) { delete G; set_error("ngpd_grid_create: non-finite coordinates"); return -1; }
    for (int c = 0; c < 6; ++c) G->bbox[c] = ordered_float(bb_h[c]);
    for (int c = 0; c < 6; ++c)
        if (!std::isfinite(G->bbox[c])) { delete G; set_error("ngpd_grid_create: non-finite coordinates"); return -1; }
    double ex = (double)G->bbox[3] - G->bbox[0], ey = (double)G->bbox[4] - G->bbox[1], ez = (double)G->bbox[5] - G->bbox[2];
    double emax = std::max(ex, std::max(ey, ez));
    if (emax <= 0) emax = 1.0;
    // points per occupied cell: 0.4 k, but no more than for k = 32 -- longer rows of cells would not fit the streaming
    // search's 64-point range slots, and a k = 64 query is answered by the 5x5x5 tier anyway
#ifndef NGPD_GRID_KMAX
#define NGPD_GRID_KMAX 32
#endif
    const double target = std::max(2.0, 0.40 * (double)std::min(k_hint > 0 ? k_hint : 16, NGPD_GRID_KMAX));
    double h;
    bool fixed = cell_size > 0.0f;
    if (fixed) {
        h = cell_size;
    } else {
        // first guess: points spread over the bounding box's surface
        double area = 2.0 * (ex * ey + ey * ez + ez * ex);
        if (area <= 0) area = emax * emax;
        h = std::sqrt(target * area / (double)n);
    }
    const double hmin = emax / 2.0e6;   // 21 bits of cell coordinate per axis
    int rc = 0;
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (h < hmin) h = hmin;
        // keep the dense brick table bounded (<= 2^31 entries, in practice far smaller)
        for (;;) {
            double tb = (std::floor(ex / h) / 8 + 1) * (std::floor(ey / h) / 8 + 1) * (std::floor(ez / h) / 8 + 1);
            if (tb < 1.0e9) break;
            h *= 1.26;
        }
        rc = build_once(G, pos, n, h, stream);
        if (rc) break;
        G->rebuilds = attempt;
        if (fixed) break;
        double occ = (double)n / (double)std::max<int64_t>(1, G->occupied_cells);
        if (occ > target / 1.5 && occ < target * 1.5) break;
        if (G->occupied_cells <= 1 && n > 1 && attempt == 3) break;
        // surfaces: occupancy ~ h^2 ; volumes ~ h^3.  Use the exponent 2.5 as a compromise; the loop converges either way.
        double s = std::pow(target / occ, 1.0 / 2.5);
        if (s > 8) s = 8;
        if (s < 0.125) s = 0.125;
        if (h * s <= hmin && h <= hmin) break;
        h *= s;
    }
    if (rc) { destroy_buffers(G); delete G; return rc; }
    *out = G;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_grid_destroy(ngpd_grid_t* g) {
    if (!g) return 0;
    destroy_buffers(g);
    delete g;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_grid_info(const ngpd_grid_t* g, ngpd_grid_info_t* o) {
    NGPD_REQUIRE(g && o, "ngpd_grid_info: NULL argument");
    o->n = g->n;
    o->cell_size = g->v.h;
    o->dims[0] = g->v.nx; o->dims[1] = g->v.ny; o->dims[2] = g->v.nz;
    o->bricks = g->nbricks;
    o->occupied_cells = g->occupied_cells;
    o->bytes = g->bytes;
    for (int c = 0; c < 6; ++c) o->bbox[c] = g->bbox[c];
    o->rebuilds = g->rebuilds;
    return 0;
}

__global__ void __launch_bounds__(256) perm_kernel(const float4* __restrict__ pts, int64_t n, int32_t* __restrict__ perm) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) perm[s] = __float_as_int(pts[s].w);
}

extern "C" __attribute__((visibility("default"))) int ngpd_grid_order(const ngpd_grid_t* g, int32_t* perm_out, void* stream_) {
    NGPD_REQUIRE(g && perm_out, "ngpd_grid_order: NULL argument");
    perm_kernel<<<(unsigned)cdiv(g->n, 256), 256, 0, (cudaStream_t)stream_>>>(g->pts, g->n, perm_out);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}
