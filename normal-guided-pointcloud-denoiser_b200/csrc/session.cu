// Fused denoising session: the body of Processor.denoise / denoiseUntilMinimumError
// (Processor.py:124-139, 158-176) on device-resident state kept in TREE (Morton) order as float4 arrays,
// so that a point's neighbours are close in memory and every gather is one 16-byte load that hits L1/L2.
//   step = knn(k_f) -> [NVT -> eigh -> smooth] -> [NVT -> eigh -> label, edge vector]
//          -> per class in order 0,1,2: (flat only: cloud-wide centre/delta) -> update into the ping-pong buffer
// Semantics kept from the reference: the index is frozen on the construction-time positions while queries
// are the current positions; the k_u-neighbourhood is the prefix of the k_f one; classes are updated
// one after the other, each from a snapshot that already contains the earlier classes' moves.
// `owned` (optional) marks the rows this rank computes; the remaining rows are halo copies of points owned by
// another GPU, refreshed between phases by ngpd_session_{export,import}_rows.
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda/ptx>
#include "knn_stream.cuh"
#include "point_math.cuh"
#include "../../include/ngpd.h"

#ifndef NGPD_CLASSIFY_BLOCKS
#define NGPD_CLASSIFY_BLOCKS 9   // resident blocks per SM of the stage-2 kernel at k = 16 (56 registers); 8 = 64 registers measured in profiles/
#endif
#define NGPD_PROF_CATEGORIES 6   // 0 knn, 1 nvt+smooth, 2 nvt+classify, 3 flat scalars, 4 update, 5 halo exchange + cross-rank scalars (multi-GPU)

namespace ngpd {
// What the cross-rank kernels need, by value (world <= NGPD_SLAB_MAX_WORLD).  Every rank owns one block of symmetric
// memory with the same layout, mapped into all peers (NVLink / NVSwitch):
//   halo[2][cap] float4   receive buffers of the halo rows, used alternately (parity of the round)
//   scal[2][world][4] f64 one slot per source rank for the cloud-wide scalars of flat_step
//   flag[world] u64       round number last signalled by each source rank
constexpr int SLAB_MAX_WORLD = 16;
struct SlabDev {
    int world, rank;
    long long cap;
    long long seg[SLAB_MAX_WORLD + 1];            // rows [seg[q], seg[q+1]) of the send list go to rank q
    long long first_row[SLAB_MAX_WORLD];          // ... and land at this row of rank q's receive buffer
    unsigned long long base[SLAB_MAX_WORLD];      // address of rank q's block in this process
    __host__ __device__ float4* halo(int q, int parity) const { return reinterpret_cast<float4*>(base[q]) + (long long)parity * cap; }
    __host__ __device__ double* scal(int q, int parity) const {
        return reinterpret_cast<double*>(base[q] + 2ull * (unsigned long long)cap * 16ull) + (long long)parity * world * 4;
    }
    __host__ __device__ unsigned long long* flag(int q) const {
        return reinterpret_cast<unsigned long long*>(base[q] + 2ull * (unsigned long long)cap * 16ull + 2ull * world * 32ull);
    }
};
}

struct ngpd_session {
    ngpd_grid* grid = nullptr;
    int64_t n = 0;
    float4* pos[2] = {nullptr, nullptr};
    int cur = 0;
    float4* nrm = nullptr;    // current normals
    float4* fn = nullptr;     // smoothed normals of this step
    float4* edge = nullptr;   // crease direction (eigenvector of the smallest stage-2 eigenvalue)
    uint8_t* label = nullptr;
    uint8_t* owned = nullptr; // nullable
    int32_t* idx = nullptr;
    int idx_k = 0;
    int32_t* fix = nullptr;   // hand-over lists of the kNN tiers: 3 x n rows + 3 counters (tier 0 -> 1 -> 2 -> exact)
    bool exact_only = false;
    // temporal coherence (knn_stream.cuh, tier 0): candidates + anchors left by the last searches with row template cand_k
    int32_t* cand = nullptr;  // [n * 2 * cand_k]
    float4* anchor = nullptr; // [n]
    int cand_k = 0;           // 0 = nothing stored
    int cand_cap_k = 0;       // row template the candidate rows are allocated (and zero-filled) for
    bool use_rerank = true;
    double* acc = nullptr;    // 4 x 8 bytes: {sum x, sum y, sum z (fixed point, int64), count (int64)} of flat_step's centre; doubles for the edge lengths
    // flat_step's neighbour sums of class 0, produced by the stage-2 kernel itself when class 0 moves first with the flat
    // strategy: one {sum x, sum y, sum z, count} per block, reduced in a fixed order (no atomics: reproducible)
    long long* part = nullptr;  // [cdiv(n,128) * 4], fixed point (sum_scale)
    float sum_scale = 1.0f;     // power of two: a row's neighbour-position sum times this is < 2^32 in magnitude
    double* red2 = nullptr;     // scratch of the two-level reduction (slice sums + ticket counter)
    bool sums_ready = false;  // part[] and blockfar[] describe the current positions and labels
    // flat_step's delta without a second pass over the cloud: the stage-2 kernel also leaves, per block, the largest distance of
    // the class' neighbours from a REFERENCE centre (the previous iteration's); once the true centre is known only the blocks that
    // can still hold the farthest point are looked at again (session_class_max_pruned_kernel)
    float* blockfar = nullptr;  // [cdiv(n,128)] squared distances, + [1] their maximum (uint bits, atomicMax)
    float* cref = nullptr;      // reference centre xyz, [3] = distance between it and the centre now in cd
    // rows of classes 1 and 2 (the minorities: creases and corners), listed by the stage-2 kernel so that their updates
    // touch only their own rows instead of streaming the whole cloud through once more: [2 * n rows][2 counters]
    int32_t* cls = nullptr;
    int32_t* inv = nullptr;   // original index -> tree position (built on the first read-back)
    bool lists_ready = false; // cls[] matches label[]
    float* cd = nullptr;      // centre xyz, delta
    float4* orig = nullptr;   // positions a clamped run measures its displacement from (ngpd_session_set_original), tree order
    int launches = 0;
    int knn_launches = 0;     // kernels of the last kNN pass
    // staging for the host-buffer entry point (packed [n,3] positions, normals, labels), allocated on first use
    float* stage_pos = nullptr;
    float* stage_nrm = nullptr;
    uint8_t* stage_lab = nullptr;
    // The last two search tiers (5x5x5 cells, exact shell search) answer ~0.1 % of the rows but take 0.2 ms of pure
    // latency (a handful of warps walking long candidate lists).  In a step they run on their own stream while the first
    // tensor pass already works on every other row; `late` marks the rows it has to leave for a small second launch.
    cudaStream_t tail = nullptr;
    cudaEvent_t tail_ev[2] = {nullptr, nullptr};
    uint8_t* late = nullptr;              // [n], all zero between steps
    bool overlap_tail = false;            // set by the step's search only
    bool tail_pending = false;            // tiers 2-3 of the current search are still running on `tail`
    cudaStream_t side = nullptr;          // second stream + events of the host-buffer entry point
    cudaEvent_t xfer[3] = {nullptr, nullptr, nullptr};
    // optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers)
    bool profiling = false;
    std::vector<cudaEvent_t> ev;      // pairs
    std::vector<int> ev_cat;
    double prof_ms[NGPD_PROF_CATEGORIES] = {0, 0, 0, 0, 0, 0};
    int prof_n[NGPD_PROF_CATEGORIES] = {0, 0, 0, 0, 0, 0};
    // Morton-slab wiring (ngpd_session_set_slab): halo rows travel by peer stores into symmetric memory, see the slab section
    ngpd::SlabDev* slab = nullptr;     // host copy of what the exchange kernels get by value
    const int32_t* slab_send_rows = nullptr;   // caller-owned device arrays
    const int32_t* slab_recv_rows = nullptr;
    int64_t slab_n_send = 0, slab_n_recv = 0;
    unsigned long long slab_epoch = 0;         // one per cross-rank round, the same sequence on every rank
    unsigned* slab_done = nullptr;             // block counter of the export kernel
    // the last position refresh of a step is only half done when the step returns: the rows are pushed, the wait + scatter is left
    // for the moment somebody needs halo positions -- the next step's search does not, so it overlaps the exchange
    bool slab_pull_pending = false;
    float4* slab_pull_dst = nullptr;
    unsigned long long slab_pull_epoch = 0;
    float* halo_need = nullptr;                // max over the owned rows of (k-th neighbour distance + distance moved from the tree position)
};

namespace ngpd {
struct ProfScope {
    ngpd_session* S; cudaStream_t st; int cat; cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(ngpd_session* S_, cudaStream_t st_, int cat_) : S(S_), st(st_), cat(cat_) {
        if (S->profiling) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (S->profiling) { cudaEventRecord(b, st); S->ev.push_back(a); S->ev.push_back(b); S->ev_cat.push_back(cat); }
    }
};
}

namespace ngpd {

// Slabs only (need != nullptr): running maximum over the owned rows of  d_k + |query - its own tree position|.  Every tree
// point within the halo width of an owned point's TREE position is resident, so a row whose value stays below that width
// has seen every candidate the whole cloud would offer (partition.py checks it after the run).  All lanes must call.
__device__ __forceinline__ void note_halo_need(float* __restrict__ need, bool active, double dk2, float4 q, const float4* __restrict__ tree_pts, int64_t s) {
    float v = 0.0f;
    if (active) {
        const float4 t = __ldg(tree_pts + s);
        const float dx = q.x - t.x, dy = q.y - t.y, dz = q.z - t.z;
        v = sqrtf((float)dk2) * 1.000001f + sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx))) * 1.000001f;   // (rounded up: a bound, not a value)
        if (!(v == v)) v = 3.0e38f;
    }
    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(v));
    // (the running maximum only grows: a stale, L1-cached copy of it is good enough to skip the atomic -- reading it from L2 in every
    // warp was a ~600-cycle stall at the end of 1.6 M warps per pass)
    if ((threadIdx.x & 31) == 0 && m > __ldca(reinterpret_cast<const unsigned*>(need))) atomicMax(reinterpret_cast<unsigned*>(need), m);
}

template <int K>
__global__ void __launch_bounds__(128) session_knn_kernel(GridView g, const float4* __restrict__ pos, const uint8_t* __restrict__ owned,
                                                          int64_t n, int k, int32_t* __restrict__ idx, float* __restrict__ need) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = s < n && (!owned || owned[s]);
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    double dk2 = 0.0;
    if (active) {
        q = __ldg(pos + s);
        TopK<K> top;
        top.init();
        knn_search<K>(top, g, (double)q.x, (double)q.y, (double)q.z, -1);
        int32_t* row = idx + s * k;
#pragma unroll
        for (int a = 0; a < K; ++a)
            if (a < k) { row[a] = top.id[a] >= 0 ? top.id[a] : (int32_t)s; if (a == k - 1 && top.id[a] >= 0) dk2 = top.d[a]; }   // fewer than k tree points: pad with self
    }
    if (need) note_halo_need(need, active, dk2, q, g.pts, active ? s : 0);
}

// What a search leaves behind for the re-ranking tier (knn_stream.cuh, "tier 0"): per row the position it was asked
// from + the certified radius (anchor.w; 0 = nothing stored), and KT = 2K candidate tree positions whose first K are the row.
struct KnnTrack {
    int32_t* cand;      // [n * KT]
    float4* anchor;     // [n]
    float* need;        // nullable: note_halo_need's slot (slabs only)
};

template <int K, int KT>
__device__ __forceinline__ void session_write_row(int64_t s, int k, const int (&id)[KT], int32_t* __restrict__ idx) {
    if (k == K) {
        int4* row = reinterpret_cast<int4*>(idx + s * K);
#pragma unroll
        for (int a = 0; a < K / 4; ++a) row[a] = make_int4(id[4 * a], id[4 * a + 1], id[4 * a + 2], id[4 * a + 3]);
    } else {
#pragma unroll
        for (int a = 0; a < K; ++a)
            if (a < k) idx[s * k + a] = id[a];
    }
}

// tier 0: every row whose stored candidates still answer the query; the others are listed for the search tiers.
// (Measured and dropped: staging the window of 512-768 tree points around the block's rows in shared memory and gathering
// the candidates' coordinates from there when they fall inside it -- 1.00 -> 1.17 ms at 10 M points, k = 16: the extra
// select per candidate costs more than the L1 look-ups it saves.  6 instead of 5 blocks per SM (80 registers): no change.)
#ifndef NGPD_RERANK_BLOCKS
#define NGPD_RERANK_BLOCKS 5
#endif
template <int K, bool LEGACY_STORE>
__global__ void __launch_bounds__(128, K <= 16 ? NGPD_RERANK_BLOCKS : 4) session_knn_rerank_kernel(GridView g, const float4* __restrict__ pos, const uint8_t* __restrict__ owned,
                                                                 int64_t n, int k, int32_t* __restrict__ idx, KnnTrack tr,
                                                                 int32_t* __restrict__ fail_list, int32_t* __restrict__ fail_count) {
    __shared__ KsCandTile<2 * K> tile;
    const int64_t row0 = (int64_t)blockIdx.x * blockDim.x;
    tile.fill(tr.cand, row0, n);
    const int64_t s0 = row0 + threadIdx.x;
    const bool active = s0 < n && (!owned || owned[s0]);
    const int64_t s = active ? s0 : 0;
    const float4 q = __ldg(pos + s), an = __ldg(tr.anchor + s);
    __syncthreads();
    constexpr int KF = ks_kf(K, 2 * K);
    KsTop<KF> top;
    double ex[KF];
    // (lanes whose row is not this rank's, or past the end, re-rank whatever their column holds -- zeros unless searched before --
    // and drop the result; a warp without any row of its own -- halo rows come in runs of tree order -- skips the work)
    bool ok = false;
    if (__any_sync(FULL, active)) ok = ks_rerank<K>(top, g, KsRowShared<2 * K>{tile, (int)threadIdx.x}, an, q.x, q.y, q.z, ex);
    else {
#pragma unroll
        for (int a = 0; a < KF; ++a) { top.id[a] = 0; ex[a] = 0.0; }
    }
    ok = ok && active && an.w > 0.0f;
    if (tr.need) note_halo_need(tr.need, ok, ex[K - 1], q, g.pts, s);
    bool bulk_pending = false;
    if (k == K) {
        if constexpr (K <= 16 && !LEGACY_STORE) {
            // Rows out by TMA: a warp's 32 rows are one contiguous 32*K*4-byte piece of the neighbour table.  Every lane puts its
            // row into a row-major staging area (16-byte stores), one elected lane hands the piece to the bulk-copy engine
            // (cp.async.bulk.global.shared::cta) -- no block barrier, no transposed read-back, no per-lane global stores.
            // Rows that failed the certificate are written too: they are on the hand-over list and every one of them is
            // rewritten by a search tier later in the stream before anything reads the table.
            namespace ptx = cuda::ptx;
            __shared__ alignas(128) int rows_out[128 * K];
            int4* mine = reinterpret_cast<int4*>(rows_out + threadIdx.x * K);
            constexpr int CH = K / 4;                                // 16-byte chunks per row
            static_assert((CH & (CH - 1)) == 0 && CH <= 8, "the chunk rotation wants a power of two");
#if defined(NGPD_ROWS_OUT_PLAIN)
#pragma unroll
            for (int c = 0; c < CH; ++c) mine[c] = make_int4(top.id[4 * c], top.id[4 * c + 1], top.id[4 * c + 2], top.id[4 * c + 3]);
#else
            // A lane's row starts CH chunks after its neighbour's, so the lanes of a quarter warp that store chunk c of their rows hit
            // 8 / CH bank groups CH times over (ncu round 2: 15 M of this kernel's 58 M shared-memory wavefronts were these conflicts).
            // Lane L therefore stores its chunks in the order r, r+1, ... with r = (L / (8 / CH)) mod CH: in every step the eight
            // lanes of a quarter warp cover the eight 16-byte bank groups.  The rotation of the row is log2(CH) rounds of selects.
            int4 ch[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) ch[c] = make_int4(top.id[4 * c], top.id[4 * c + 1], top.id[4 * c + 2], top.id[4 * c + 3]);
            const int rot = CH > 1 ? (((int)threadIdx.x & 31) / (CH > 1 ? 8 / CH : 1)) & (CH - 1) : 0;
#pragma unroll
            for (int bit = 1; bit < CH; bit <<= 1) {
                const bool on = (rot & bit) != 0;
                int4 t[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int4 a = ch[c], b = ch[(c + bit) & (CH - 1)];
                    t[c] = make_int4(on ? b.x : a.x, on ? b.y : a.y, on ? b.z : a.z, on ? b.w : a.w);
                }
#pragma unroll
                for (int c = 0; c < CH; ++c) ch[c] = t[c];
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) mine[(c + rot) & (CH - 1)] = ch[c];          // ch[c] = chunk (c + rot) mod CH of the row
#endif
            ptx::fence_proxy_async(ptx::space_shared);               // generic-proxy stores -> visible to the async proxy
            __syncwarp();
            if ((threadIdx.x & 31) == 0) {
                const int64_t w0 = row0 + (threadIdx.x & ~31);
                const int64_t rows = n - w0 < 32 ? n - w0 : 32;
                if (rows > 0) {
                    ptx::cp_async_bulk(ptx::space_global, ptx::space_shared, idx + w0 * K, rows_out + (threadIdx.x & ~31) * K,
                                       (uint32_t)(rows * K * sizeof(int32_t)));
                    ptx::cp_async_bulk_commit_group();
                    bulk_pending = true;
                }
            }
        } else {
            // rows out through the tile as well: the block's 128 rows are one contiguous 128*K*4-byte piece of the table, written
            // with coalesced 16-byte stores (a lane storing its own 64-byte row touches a cache line per two lanes)
            __shared__ uint8_t row_ok[128];
            __syncthreads();                                   // every lane is done reading candidate ids
#pragma unroll
            for (int a = 0; a < K; ++a) tile.v[a][threadIdx.x] = top.id[a];
            row_ok[threadIdx.x] = ok;
            __syncthreads();
            constexpr int CH = K / 4;
            int4* dst = reinterpret_cast<int4*>(idx + row0 * K);
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int c = (int)threadIdx.x + 128 * i, r = c / CH, a0 = (c % CH) * 4;
                if (row_ok[r]) dst[c] = make_int4(tile.v[a0][r], tile.v[a0 + 1][r], tile.v[a0 + 2][r], tile.v[a0 + 3][r]);
            }
        }
    } else if (ok) {
        session_write_row<K, KF>(s, k, top.id, idx);
    }
    fix_append(active && !ok, (int)s, fail_list, fail_count);
    if (bulk_pending) cuda::ptx::cp_async_bulk_wait_group_read(cuda::ptx::n32_t<0>{});   // shared memory must outlive the copy that reads it
}

// tier 0 with the candidate rows staged by TMA (north star kernel 2: "stages candidates in shared memory via TMA"): see KsCandRows.
// The arithmetic is ks_rerank, as in the kernel above, so the rows are bit-identical; what changes is the data movement:
//   fill   one cp.async.bulk (global -> shared, mbarrier complete_tx) per thread for its own 8 K-byte row instead of coalesced
//          16-byte loads into registers + 2 K scalar shared-memory stores per thread + a block barrier;
//   reads  K / 2 16-byte shared loads per lane instead of 2 K scalar ones;
//   store  the sorted row goes back through the lane's own shared row and one cp.async.bulk (shared -> global) per thread
//          instead of K shared stores, two block barriers and a transposed read-back.
// Every thread touches only its own row of the tile: the only block-wide synchronisation left is the mbarrier itself.
template <int K>
__global__ void __launch_bounds__(128, K <= 16 ? 5 : 4) session_knn_rerank_tma_kernel(GridView g, const float4* __restrict__ pos, const uint8_t* __restrict__ owned,
                                                                     int64_t n, int k, int32_t* __restrict__ idx, KnnTrack tr,
                                                                     int32_t* __restrict__ fail_list, int32_t* __restrict__ fail_count) {
    namespace ptx = cuda::ptx;
    constexpr int KT = 2 * K;
    __shared__ KsCandRows<KT> tile;
    __shared__ alignas(8) uint64_t bar;
    const int tid = threadIdx.x;
    const int64_t s0 = (int64_t)blockIdx.x * blockDim.x + tid;
    if (tid == 0) {
        ptx::mbarrier_init(&bar, 128);
        ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
    }
    __syncthreads();
    int* myrow = tile.v[tid];
    if (s0 < n) {
        ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, &bar, (uint32_t)(KT * sizeof(int32_t)));
        ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, myrow, tr.cand + s0 * KT, (uint32_t)(KT * sizeof(int32_t)), &bar);
    } else {
        // past the end (last block): an all-zero row, valid ids whose result is dropped
#pragma unroll
        for (int c = 0; c < KT / 4; ++c) reinterpret_cast<int4*>(myrow)[c] = make_int4(0, 0, 0, 0);
        ptx::mbarrier_arrive(&bar);
    }
    const bool active = s0 < n && (!owned || owned[s0]);
    const int64_t s = active ? s0 : 0;
    const float4 q = __ldg(pos + s), an = __ldg(tr.anchor + s);     // in flight while the rows arrive
    while (!ptx::mbarrier_try_wait_parity(&bar, 0)) {}
    constexpr int KF = ks_kf(K, 2 * K);
    KsTop<KF> top;
    double ex[KF];
    bool ok = false;
    if (__any_sync(FULL, active)) ok = ks_rerank<K>(top, g, KsRowPadded{myrow}, an, q.x, q.y, q.z, ex);
    else {
#pragma unroll
        for (int a = 0; a < KF; ++a) { top.id[a] = 0; ex[a] = 0.0; }
    }
    ok = ok && active && an.w > 0.0f;
    if (tr.need) note_halo_need(tr.need, ok, ex[K - 1], q, g.pts, s);
    if (k == K) {
        // (the lane is done reading its candidate ids: top.id[] holds what it needs)
#pragma unroll
        for (int c = 0; c < K / 4; ++c)
            reinterpret_cast<int4*>(myrow)[c] = make_int4(top.id[4 * c], top.id[4 * c + 1], top.id[4 * c + 2], top.id[4 * c + 3]);
        ptx::fence_proxy_async(ptx::space_shared);                  // the bulk copy reads shared memory through the async proxy
        if (ok) ptx::cp_async_bulk(ptx::space_global, ptx::space_shared, idx + s * K, myrow, (uint32_t)(K * sizeof(int32_t)));
        ptx::cp_async_bulk_commit_group();
    } else if (ok) {
        session_write_row<K, KF>(s, k, top.id, idx);
    }
    fix_append(active && !ok, (int)s, fail_list, fail_count);
    if (k == K) ptx::cp_async_bulk_wait_group_read(ptx::n32_t<0>{});   // shared memory must outlive the copies that read it
}

// tiers 1 (R = 1) and 2 (R = 2) of the streaming search.  KT == K: plain search.  KT == 2K: the search also stores
// candidates + anchor for tier 0.
template <int K, int KT, int R>
__device__ __forceinline__ void session_knn_body(KsShared<R>& sm, const GridView& g, const float4* __restrict__ pos, int64_t s, bool active,
                                                 int k, int32_t* __restrict__ idx, KnnTrack tr, int32_t* __restrict__ fail_list,
                                                 int32_t* __restrict__ fail_count, uint8_t* __restrict__ late = nullptr) {
    const float4 q = active ? __ldg(pos + s) : make_float4(0.f, 0.f, 0.f, 0.f);
    KsTop<KT> top;
    float rlim = 0.0f;
    const bool ok = knn_stream<K, KT, R, false>(top, sm, g, q.x, q.y, q.z, active, -1, INFINITY, KT > K ? &rlim : nullptr);
    constexpr int KF = ks_kf(K, KT);
    double ex[KF];
    ks_finalize<KF, KT>(top, g.pts, q.x, q.y, q.z, ex);
    if (tr.need) note_halo_need(tr.need, active && ok, ex[K - 1], q, g.pts, s);
    if (active && ok) {
        session_write_row<K, KT>(s, k, top.id, idx);
        if (KT > K) {
            int4* crow = reinterpret_cast<int4*>(tr.cand + s * KT);
#pragma unroll
            for (int a = 0; a < KT / 4; ++a) crow[a] = make_int4(top.id[4 * a], top.id[4 * a + 1], top.id[4 * a + 2], top.id[4 * a + 3]);
            tr.anchor[s] = make_float4(q.x, q.y, q.z, rlim);
        }
    }
    fix_append(active && !ok, (int)s, fail_list, fail_count);
    if (late && active && !ok) late[s] = 1;     // answered later, on the side stream: the first tensor pass skips the row
}

// rows = todo_list[0 .. *todo_count) when a list is given, every (owned) row otherwise
template <int K, int KT>
// resident blocks per SM of the search that keeps 2K = 32 candidates: 5 (96 registers, 76 bytes of spills) beats 4 (128 registers)
// by 5 % and 3 (158, no spill) by 16 % on the first search of a session -- profiles/r2_ab_measurements.md, section 8
#ifndef NGPD_FAST_BLOCKS_KT32
#define NGPD_FAST_BLOCKS_KT32 5
#endif
__global__ void __launch_bounds__(KsCfg<1>::THREADS, KT <= 16 ? 5 : (KT <= 32 ? NGPD_FAST_BLOCKS_KT32 : 1))
session_knn_fast_kernel(GridView g, const float4* __restrict__ pos, const uint8_t* __restrict__ owned, int64_t n, int k,
                        int32_t* __restrict__ idx, KnnTrack tr, const int32_t* __restrict__ todo_list,
                        const int32_t* __restrict__ todo_count, int32_t* __restrict__ fail_list, int32_t* __restrict__ fail_count,
                        uint8_t* __restrict__ late) {
    __shared__ KsShared<1> sm;
    const int64_t cnt = todo_list ? (int64_t)*todo_count : n;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < cnt; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = base + threadIdx.x;
        bool active = i < cnt;
        const int64_t s = active ? (todo_list ? (int64_t)todo_list[i] : i) : 0;
        if (!todo_list) active = active && (!owned || owned[s]);
        session_knn_body<K, KT, 1>(sm, g, pos, s, active, k, idx, tr, fail_list, fail_count, late);
    }
}

template <int K, int KT>
__global__ void __launch_bounds__(KsCfg<2>::THREADS) session_knn_wide_kernel(GridView g, const float4* __restrict__ pos, int k, int32_t* __restrict__ idx,
                                                                             KnnTrack tr, const int32_t* __restrict__ todo_list,
                                                                             const int32_t* __restrict__ todo_count,
                                                                             int32_t* __restrict__ fail_list, int32_t* __restrict__ fail_count) {
    __shared__ KsShared<2> sm;
    const int cnt = *todo_count;
    for (int base = blockIdx.x * blockDim.x; base < cnt; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        const bool active = i < cnt;
        session_knn_body<K, KT, 2>(sm, g, pos, active ? (int64_t)todo_list[i] : 0, active, k, idx, tr, fail_list, fail_count);
    }
}

// last tier: exact shell search (fp64 insertion lists) for what is left -- isolated points, clouds of duplicates
template <int K, int KT>
__global__ void __launch_bounds__(128) session_knn_fix_kernel(GridView g, const float4* __restrict__ pos, int k, int32_t* __restrict__ idx, KnnTrack tr,
                                                              const int32_t* __restrict__ fix_list, const int32_t* __restrict__ fix_count) {
    const int cnt = *fix_count;
    for (int base = blockIdx.x * blockDim.x; base < cnt; base += gridDim.x * blockDim.x) {
        const int i = base + threadIdx.x;
        const bool active = i < cnt;
        const int64_t s = active ? fix_list[i] : 0;
        const float4 q = __ldg(pos + s);
        double dk2 = 0.0;
        // A 64-entry fp64 list lives in local memory (measured: 1.8 ms for a handful of rows) and a 48-entry one at 255
        // registers is not much better (1.3 ms for 100 rows): with K = 32 only the row itself is searched here and nothing is
        // stored for tier 0 (radius 0 = the row goes through the search tiers again next time; ~0.04 % of the rows, the ones
        // in regions dense enough to overflow the 5x5x5 tier's range slots).
        constexpr int KS = (KT > 32 && K <= 32) ? K : KT;
        if (active) {
            TopK<KS> top;
            top.init();
            knn_search<KS>(top, g, (double)q.x, (double)q.y, (double)q.z, -1);
            int32_t* row = idx + s * k;
#pragma unroll
            for (int a = 0; a < K; ++a)
                if (a < k) { row[a] = top.id[a] >= 0 ? top.id[a] : (int32_t)s; if (a == k - 1 && top.id[a] >= 0) dk2 = top.d[a]; }   // fewer than k tree points: pad with self
            if constexpr (KT > K && KS < KT) {
                tr.anchor[s] = make_float4(q.x, q.y, q.z, 0.0f);
            } else if constexpr (KT > K) {
                int4* crow = reinterpret_cast<int4*>(tr.cand + s * KT);
#pragma unroll
                for (int a = 0; a < KT / 4; ++a) crow[a] = make_int4(top.id[4 * a], top.id[4 * a + 1], top.id[4 * a + 2], top.id[4 * a + 3]);
                // a full list: nothing outside it is closer than its last entry; a short one holds the whole tree
                const float rlim = top.id[KT - 1] >= 0 ? __double2float_rd(sqrt(top.d[KT - 1]) * (1.0 - 1e-7)) : 3.0e38f;
                tr.anchor[s] = make_float4(q.x, q.y, q.z, top.id[K - 1] >= 0 ? rlim : 0.0f);
            }
        }
        if (tr.need) note_halo_need(tr.need, active, dk2, q, g.pts, s);
    }
}

// neighbour row of point s in registers: K > 0 = compile-time row length (16-byte vector loads, fully unrolled
// consumers), K = 0 = runtime length read through the pointer
template <int K>
struct RowRegs {
    int32_t v[K > 0 ? K : 1];
    __device__ __forceinline__ int64_t operator()(int a) const { return (int64_t)v[a]; }
    __device__ __forceinline__ void load(const int32_t* __restrict__ row) {
        const int4* r4 = reinterpret_cast<const int4*>(row);
#pragma unroll
        for (int a = 0; a < K / 4; ++a) { int4 q = __ldg(r4 + a); v[4 * a] = q.x; v[4 * a + 1] = q.y; v[4 * a + 2] = q.z; v[4 * a + 3] = q.w; }
    }
};

// stage 1: filtered NVT on the current normals, smoothed normal out
template <int K>
__device__ __forceinline__ void session_nvt_smooth_row(const Quad4& pos, const Quad4& nrm, const int32_t* __restrict__ idx, int64_t s, int k,
                                                       float x_thresh, float tau, float damp, float4* __restrict__ fn) {
    NvtResult o;
    if (K > 0) {
        RowRegs<K> row;
        row.load(idx + s * K);
        nvt_point_row<K>(pos, nrm, s, row, idx + s * K, K, x_thresh, o, nullptr);
    } else {
        nvt_point(pos, nrm, s, idx + s * k, k, x_thresh, o, nullptr);
    }
    V3 f = smooth_normal(o.w, o.V, nrm(s), tau, damp);
    fn[s] = make_float4(f.x, f.y, f.z, 0.0f);
}
// every (owned) row; rows marked in `late` (nullable) are still being searched on the side stream and are left out
template <int K>
__global__ void __launch_bounds__(128) session_nvt_smooth_kernel(Quad4 pos, Quad4 nrm, const uint8_t* __restrict__ owned,
                                                                 const uint8_t* __restrict__ late, const int32_t* __restrict__ idx, int64_t n, int k,
                                                                 float x_thresh, float tau, float damp, float4* __restrict__ fn) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    if (owned && !owned[s]) return;
    if (late && late[s]) return;
    session_nvt_smooth_row<K>(pos, nrm, idx, s, k, x_thresh, tau, damp, fn);
}
// the rows left out above, once their search is done; clears their marks
template <int K>
__global__ void __launch_bounds__(128) session_nvt_smooth_late_kernel(Quad4 pos, Quad4 nrm, const int32_t* __restrict__ list,
                                                                      const int32_t* __restrict__ count, uint8_t* __restrict__ late,
                                                                      const int32_t* __restrict__ idx, int k, float x_thresh, float tau, float damp,
                                                                      float4* __restrict__ fn) {
    const int cnt = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const int64_t s = list[i];
        session_nvt_smooth_row<K>(pos, nrm, idx, s, k, x_thresh, tau, damp, fn);
        late[s] = 0;
    }
}

// stage 2: filtered NVT on the smoothed normals, label + crease direction out.  With part != nullptr the kernel also
// leaves, per block, {sum x, sum y, sum z, count} over the first ku neighbours of the rows it labelled sum_key
// (flat_step's centre, Denoiser.py:106): the positions were just gathered, so the separate pass over the class is saved.
// FAST: labels from closed-form eigenvalues (eig3_fast.cuh); rows whose label is not certain that way, rows that need their crease
// direction (edge_mask: bit l set = rows labelled l are moved by edge_step), and every row when FAST is off go through the
// LAPACK-order solver.
template <int K, bool FAST>
__global__ void __launch_bounds__(128, K == 16 ? NGPD_CLASSIFY_BLOCKS : (K == 32 ? 6 : 1)) session_nvt_classify_kernel(Quad4 pos, Quad4 fn, const uint8_t* __restrict__ owned,
                                                                   const int32_t* __restrict__ idx, int64_t n, int k, float x_thresh,
                                                                   float scale, uint8_t* __restrict__ label, float4* __restrict__ edge, int edge_mask,
                                                                   int sum_key, int ku, long long* __restrict__ part, float sum_scale,
                                                                   const float* __restrict__ cref, float* __restrict__ blockfar,
                                                                   int32_t* __restrict__ cls) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = s < n && (!owned || owned[s]);
    float fx = 0.0f, fy = 0.0f, fz = 0.0f, far2 = 0.0f;
    int lab = -1;
    if (active) {
        // the sum of the first ku neighbour positions (flat_step's centre) is accumulated inside the vote loop, where the
        // positions are in registers anyway: fp32 within the row and the warp (the reference's own mean is fp32), fp64 above
        V3 ps = v3(0.0f, 0.0f, 0.0f);
        float t6[6];
        if (K > 0) {
            RowRegs<K> row;
            row.load(idx + s * K);
            nvt_tensor_row<K>(pos, fn, s, row, idx + s * K, K, x_thresh, t6, part ? &ps : nullptr, ku, cref, part ? &far2 : nullptr);
        } else {
            nvt_tensor_row<0>(pos, fn, s, RowPtr<int32_t>{idx + s * k}, idx + s * k, k, x_thresh, t6, part ? &ps : nullptr, ku, cref,
                              part ? &far2 : nullptr);
        }
        V3 y = v3(0.0f, 0.0f, 0.0f);
        bool full = !FAST;
        if (FAST) {
            // rows that edge_step will move keep the LAPACK-order eigenvector: its solve can be near-singular and
            // then amplifies even a 1e-6 rad change of y -- measured: one fandisk row moved by 3e-4 of the extent with a
            // cross-product eigenvector, so the fused step would no longer equal the operator-by-operator one
            const FastLabel f = classify_fast(t6[0], t6[1], t6[2], t6[3], t6[4], t6[5], scale);
            lab = f.label;
            full = !f.certain || ((edge_mask >> lab) & 1);
        }
        if (full) {
            const LabelVec o = classify_lapack(t6[0], t6[1], t6[2], t6[3], t6[4], t6[5], scale);
            lab = o.label; y = o.y;
        }
        if (lab == sum_key) { fx = ps.x; fy = ps.y; fz = ps.z; }   // (the row's fp32 sum in neighbour order: the same bits on any partition)
        label[s] = (uint8_t)lab;
        if ((edge_mask >> lab) & 1) edge[s] = make_float4(y.x, y.y, y.z, 0.0f);
    }
    if (part) {
        // Above the row everything is integer: the row sums in fixed point (sum_scale = a power of two chosen from the cloud's
        // extent, <= 2^32 per row), so that the total does not depend on which rows share a warp, a block or a GPU -- with fp32
        // warp sums the centre's last bit, and through delta a few positions' last bits, changed with the partition.
        __shared__ long long red[4][4];
        const unsigned members = __ballot_sync(0xffffffffu, lab == sum_key);
        long long qx = __float2ll_rn(fx * sum_scale), qy = __float2ll_rn(fy * sum_scale), qz = __float2ll_rn(fz * sum_scale);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            qx += __shfl_xor_sync(0xffffffffu, qx, o); qy += __shfl_xor_sync(0xffffffffu, qy, o); qz += __shfl_xor_sync(0xffffffffu, qz, o);
        }
        const int w = threadIdx.x >> 5;
        // (squared distances are non-negative floats: their bit patterns order like the values)
        __shared__ unsigned farw[4];
        const unsigned fm = __reduce_max_sync(0xffffffffu, lab == sum_key ? __float_as_uint(far2) : 0u);
        if ((threadIdx.x & 31) == 0) { red[w][0] = qx; red[w][1] = qy; red[w][2] = qz; red[w][3] = (long long)ku * __popc(members); farw[w] = fm; }
        __syncthreads();
        if (threadIdx.x < 4)
            part[(int64_t)blockIdx.x * 4 + threadIdx.x] = ((red[0][threadIdx.x] + red[1][threadIdx.x]) + red[2][threadIdx.x]) + red[3][threadIdx.x];
        if (threadIdx.x == 0 && blockfar) {
            const unsigned m = max(max(farw[0], farw[1]), max(farw[2], farw[3]));
            blockfar[blockIdx.x] = __uint_as_float(m);
            if (m) atomicMax(reinterpret_cast<unsigned*>(blockfar + gridDim.x), m);
        }
    }
    if (cls) {
        // block-aggregated append to the two class lists: one global atomic per block and class
        __shared__ int wcnt[2][4], base[2];
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const unsigned m1 = __ballot_sync(0xffffffffu, lab == 1), m2 = __ballot_sync(0xffffffffu, lab == 2);
        if (lane == 0) { wcnt[0][w] = __popc(m1); wcnt[1][w] = __popc(m2); }
        __syncthreads();
        if (threadIdx.x < 2) {
            const int c = threadIdx.x, tot = wcnt[c][0] + wcnt[c][1] + wcnt[c][2] + wcnt[c][3];
            base[c] = tot ? atomicAdd(cls + 2 * n + c, tot) : 0;
        }
        __syncthreads();
        if (lab == 1 || lab == 2) {
            const int c = lab - 1;
            int off = base[c];
            for (int v = 0; v < w; ++v) off += wcnt[c][v];
            off += __popc((c ? m2 : m1) & ((1u << lane) - 1u));
            cls[(int64_t)c * n + off] = (int32_t)s;
        }
    }
}

// fixed-order reduction of the per-block partial sums into acc (4 doubles): RED_BLOCKS blocks each add up a contiguous
// slice (fixed order inside the slice), the block that finishes last adds the slice sums in slice order.  Which block
// that is varies, the order of the additions does not: the result is reproducible.  scratch = RED_BLOCKS*4 doubles + a counter.
constexpr int RED_BLOCKS = 64;
// T = double (edge lengths) or long long (flat_step's fixed-point sums: exact, any order)
template <class T>
__global__ void __launch_bounds__(256) session_partial_reduce_kernel(const T* __restrict__ part, int64_t blocks, T* __restrict__ scratch,
                                                                     T* __restrict__ acc) {
    __shared__ T red[8][4];
    __shared__ bool last;
    const int64_t per = (blocks + RED_BLOCKS - 1) / RED_BLOCKS, b0 = blockIdx.x * per, b1 = b0 + per < blocks ? b0 + per : blocks;
    T v[4] = {0, 0, 0, 0};
    for (int64_t b = b0 + threadIdx.x; b < b1; b += 256) {
        const T* p4 = part + b * 4;
        v[0] += p4[0]; v[1] += p4[1]; v[2] += p4[2]; v[3] += p4[3];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int c = 0; c < 4; ++c) red[threadIdx.x >> 5][c] = v[c];
    __syncthreads();
    unsigned* counter = reinterpret_cast<unsigned*>(scratch + RED_BLOCKS * 4);
    if (threadIdx.x < 4) {
        T t = 0;
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        scratch[blockIdx.x * 4 + threadIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == RED_BLOCKS - 1;
    __syncthreads();
    if (last && threadIdx.x < 4) {
        __threadfence();
        T t = 0;
        for (int b = 0; b < RED_BLOCKS; ++b) t += *((volatile T*)(scratch + b * 4 + threadIdx.x));
        acc[threadIdx.x] = t;
        if (threadIdx.x == 0) *counter = 0;
    }
}

// flat_step scalars over the neighbour multiset of one class (Denoiser.py:106-107): per row the fp32 sum of its first ku neighbour
// positions in neighbour order (as the stage-2 kernel accumulates it), in fixed point above the row
__global__ void __launch_bounds__(256) session_class_sum_kernel(Quad4 pos, const uint8_t* __restrict__ owned, const uint8_t* __restrict__ label,
                                                                int key, const int32_t* __restrict__ idx, int64_t n, int k, int ku,
                                                                long long* __restrict__ part, float sum_scale) {
    long long sx = 0, sy = 0, sz = 0, cnt = 0;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        if ((owned && !owned[s]) || label[s] != key) continue;
        const int32_t* row = idx + s * k;
        V3 ps = v3(0.0f, 0.0f, 0.0f);
        for (int a = 0; a < ku; ++a) ps = ps + pos((int64_t)row[a]);
        sx += __float2ll_rn(ps.x * sum_scale); sy += __float2ll_rn(ps.y * sum_scale); sz += __float2ll_rn(ps.z * sum_scale);
        cnt += ku;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ long long red[8][4];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[w][0] = sx; red[w][1] = sy; red[w][2] = sz; red[w][3] = cnt; }
    __syncthreads();
    if (threadIdx.x < 4) {
        long long t = 0;
        for (int v = 0; v < 8; ++v) t += red[v][threadIdx.x];
        part[(int64_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}
// acc = {sum x, sum y, sum z in units of 1 / sum_scale, neighbour count} -> centre.  cref (nullable): the reference centre the last
// stage-2 pass measured its per-block maxima from; receives its distance from the new centre ([3], rounded up) and then the new centre
// itself, which is the next iteration's reference.
__global__ void session_center_kernel(const long long* __restrict__ acc, float sum_scale, float* __restrict__ cd, float* __restrict__ cref) {
    const double c = acc[3] > 0 ? (double)acc[3] * (double)sum_scale : 1.0;
    const float x = (float)((double)acc[0] / c), y = (float)((double)acc[1] / c), z = (float)((double)acc[2] / c);
    cd[0] = x; cd[1] = y; cd[2] = z; cd[3] = 0.0f;
    if (cref) {
        const double dx = (double)x - cref[0], dy = (double)y - cref[1], dz = (double)z - cref[2];
        cref[3] = (float)(sqrt(dx * dx + dy * dy + dz * dz) * 1.0001) + 1e-30f;
        cref[0] = x; cref[1] = y; cref[2] = z;
    }
}
// flat_step's delta (Denoiser.py:107) from the per-block maxima of the stage-2 pass.  far2[b] = largest squared distance of block b's
// class neighbours from the REFERENCE centre, far2[blocks] = the largest of them all (d'max^2), eps = |centre - reference|.
// The neighbour farthest from the true centre (distance delta) is at least delta - eps from the reference, and delta >= d'max - eps:
// only blocks with far[b] >= d'max - 2 eps can hold it.  Those are recomputed exactly as the full pass would (same arithmetic, same
// atomicMax), every other block costs one 4-byte load per warp.  One warp per 128-row block.
__global__ void __launch_bounds__(256) session_class_max_pruned_kernel(Quad4 pos, const uint8_t* __restrict__ owned, const uint8_t* __restrict__ label,
                                                                       int key, const int32_t* __restrict__ idx, int64_t n, int k, int ku,
                                                                       const float* __restrict__ far2, int64_t blocks, const float* __restrict__ cref,
                                                                       float* __restrict__ cd) {
    const V3 c = v3(cd[0], cd[1], cd[2]);
    const float eps = cref[3];
    const float thr = sqrtf(far2[blocks]) * 0.99999f - 2.0f * eps;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float mx = 0.0f;
    for (int64_t b = warp; b < blocks; b += warps) {
        const float f = __ldg(far2 + b);
        if (!(sqrtf(f) >= thr) || f == 0.0f) continue;              // (f == 0: no row of the class in the block)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t s = b * 128 + lane + 32 * i;
            if (s >= n || (owned && !owned[s]) || label[s] != key) continue;
            const int32_t* row = idx + s * k;
            for (int a = 0; a < ku; ++a) mx = fmaxf(mx, norm3_fma(pos((int64_t)row[a]) - c));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0 && mx > 0.0f) atomicMax((int*)(cd + 3), __float_as_int(mx));
}
__global__ void __launch_bounds__(256) session_class_max_kernel(Quad4 pos, const uint8_t* __restrict__ owned, const uint8_t* __restrict__ label,
                                                                int key, const int32_t* __restrict__ idx, int64_t n, int k, int ku,
                                                                float* __restrict__ cd) {
    V3 c = v3(cd[0], cd[1], cd[2]);
    const bool vec = (k & 3) == 0 && (ku & 3) == 0;   // 16-byte loads of the row prefix
    float mx = 0.0f;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        if ((owned && !owned[s]) || label[s] != key) continue;
        const int32_t* row = idx + s * k;
        if (vec) {
            const int4* r4 = reinterpret_cast<const int4*>(row);
            for (int a = 0; a < ku; a += 4) {
                const int4 q = __ldg(r4 + (a >> 2));
                mx = fmaxf(fmaxf(mx, norm3_fma(pos((int64_t)q.x) - c)), norm3_fma(pos((int64_t)q.y) - c));
                mx = fmaxf(fmaxf(mx, norm3_fma(pos((int64_t)q.z) - c)), norm3_fma(pos((int64_t)q.w) - c));
            }
        } else {
            for (int a = 0; a < ku; ++a) mx = fmaxf(mx, norm3_fma(pos((int64_t)row[a]) - c));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax((int*)(cd + 3), __float_as_int(mx));
}

// new position of row s under strategy `kind`.  KU > 0: compile-time row length, ids fetched with 16-byte loads into
// registers (a scalar load per neighbour id is one L1 tag look-up per LANE, rows being 64 bytes apart); KU = 0: runtime length.
template <int KU>
__device__ __forceinline__ V3 session_move_point(int kind, const Quad4& pos, const Quad4& fn, const float4* __restrict__ edge, int64_t s,
                                                 const int32_t* __restrict__ row, int ku, const float* __restrict__ cd, float alpha, float dmax) {
    if constexpr (KU > 0) {
        RowRegs<KU> r;
        r.load(row);
        if (kind == NGPD_STEP_FLAT) return flat_point_row<KU>(pos, fn, s, r, KU, cd[3], alpha, dmax);
        if (kind == NGPD_STEP_EDGE) { float4 e = __ldg(edge + s); return edge_point_row<KU>(pos, fn, v3(e.x, e.y, e.z), s, r, KU, alpha, dmax); }
        if (kind == NGPD_STEP_FEATURE) return feature_point_row<KU>(pos, fn, s, r, KU, alpha, dmax);
        return corner_point_row<KU>(pos, fn, s, r, KU, alpha, dmax);
    } else {
        if (kind == NGPD_STEP_FLAT) return flat_point(pos, fn, s, row, ku, cd[3], alpha, dmax);
        if (kind == NGPD_STEP_EDGE) { float4 e = __ldg(edge + s); return edge_point(pos, fn, v3(e.x, e.y, e.z), s, row, ku, alpha, dmax); }
        if (kind == NGPD_STEP_FEATURE) return feature_point(pos, fn, s, row, ku, alpha, dmax);
        return corner_point(pos, fn, s, row, ku, alpha, dmax);
    }
}

// the notebook's displacement clamp (PostProcessing.ipynb#c9: mask = (temp_pos - original_pos).norm(dim=1) < d;
// pos[mask] = temp_pos[mask]): a moved point is kept only while it stays within clamp_r of its original position
__device__ __forceinline__ V3 clamp_to_original(V3 moved, V3 old, const float4* __restrict__ orig, int64_t s, float clamp_r) {
    if (!orig) return moved;
    const float4 o = __ldg(orig + s);
    return norm3_fma(moved - v3(o.x, o.y, o.z)) < clamp_r ? moved : old;
}

// one class: members move, everybody else is copied through to the other buffer (snapshot semantics)
// MINB: resident blocks per SM.  The flat pass is latency-bound and wants every warp slot: at 100 M points 10 blocks per SM
// (48 registers) take 2.98 ms, 12 (40 registers) 2.81 ms, 16 (32 registers; the 580 bytes of spills per thread sit on the other
// strategies' paths) 2.54 ms.  feature_step / edge_step do more arithmetic per row and lose at 16 (0.68 / 0.85 against 0.56 / 0.62 ms
// at 10 M points): they stay at 10 (profiles/r2_ab_measurements.md, section 7).
template <int KU, int MINB>
__global__ void __launch_bounds__(128, MINB) session_update_kernel(int kind, int key, Quad4 pos, Quad4 fn, const float4* __restrict__ edge,
                                                             const uint8_t* __restrict__ label, const uint8_t* __restrict__ owned,
                                                             const int32_t* __restrict__ idx, int64_t n, int k, int ku, float alpha, float dmax,
                                                             const float* __restrict__ cd, const float4* __restrict__ orig, float clamp_r,
                                                             float4* __restrict__ out) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    V3 p = pos(s);
    if (label[s] == key && (!owned || owned[s]) && kind >= NGPD_STEP_FLAT && kind <= NGPD_STEP_CORNER) {
        const V3 q = session_move_point<KU>(kind, pos, fn, edge, s, idx + s * k, ku, cd, alpha, dmax);
        p = clamp_to_original(q, p, orig, s, clamp_r);
    }
    out[s] = make_float4(p.x, p.y, p.z, 0.0f);
}

// one minority class, in place: the listed rows move, nothing else is touched.  Two launches because a row's neighbours
// may be rows of the same class: every new position is computed from the snapshot before any of them is stored.
template <int KU>
__global__ void __launch_bounds__(128) session_update_rows_kernel(int kind, Quad4 pos, Quad4 fn, const float4* __restrict__ edge,
                                                                  const int32_t* __restrict__ list, const int32_t* __restrict__ count,
                                                                  const int32_t* __restrict__ idx, int k, int ku, float alpha, float dmax,
                                                                  const float* __restrict__ cd, const float4* __restrict__ orig, float clamp_r,
                                                                  float4* __restrict__ moved, bool scatter) {
    // scatter == false: moved[i] = new position of the i-th listed row (applied afterwards, in place);
    // scatter == true (every class moves from the same snapshot): `moved` is the other position buffer, written at the row itself
    const int cnt = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        const int64_t s = list[i];
        const V3 q = session_move_point<KU>(kind, pos, fn, edge, s, idx + s * k, ku, cd, alpha, dmax);
        const V3 p = clamp_to_original(q, pos(s), orig, s, clamp_r);
        moved[scatter ? s : (int64_t)i] = make_float4(p.x, p.y, p.z, 0.0f);
    }
}
__global__ void __launch_bounds__(256) session_apply_rows_kernel(const int32_t* __restrict__ list, const int32_t* __restrict__ count,
                                                                 const float4* __restrict__ moved, float4* __restrict__ pos) {
    const int cnt = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) pos[list[i]] = moved[i];
}

__global__ void __launch_bounds__(256) session_scatter_in_kernel(const float4* __restrict__ tree_pts, const float* __restrict__ pos,
                                                                 const float* __restrict__ nrm, int64_t n, float4* __restrict__ pos4,
                                                                 float4* __restrict__ nrm4) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int64_t o = (int64_t)__float_as_int(__ldg(&tree_pts[s].w));
    if (pos) pos4[s] = make_float4(__ldg(pos + 3 * o), __ldg(pos + 3 * o + 1), __ldg(pos + 3 * o + 2), 0.0f);
    if (nrm) nrm4[s] = make_float4(__ldg(nrm + 3 * o), __ldg(nrm + 3 * o + 1), __ldg(nrm + 3 * o + 2), 0.0f);
}
// state back to the caller's order.  One thread per ORIGINAL index: the packed outputs are written fully coalesced and the
// tree-order rows are gathered through the inverse permutation (scattering 12-byte rows from tree order instead costs a
// read-modify-write of a 32-byte sector per row: 1.23 ms against 0.3 ms at 10 M points).
__global__ void __launch_bounds__(256) session_inverse_kernel(const float4* __restrict__ tree_pts, int64_t n, int32_t* __restrict__ inv) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n) inv[__float_as_int(__ldg(&tree_pts[s].w))] = (int32_t)s;
}
__global__ void __launch_bounds__(256) session_scatter_out_kernel(const int32_t* __restrict__ inv, const float4* __restrict__ pos4,
                                                                  const float4* __restrict__ nrm4, const uint8_t* __restrict__ label, int64_t n,
                                                                  float* __restrict__ pos, float* __restrict__ nrm, uint8_t* __restrict__ lab) {
    int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n) return;
    const int64_t s = __ldg(inv + o);
    if (pos) { float4 p = __ldg(pos4 + s); pos[3 * o] = p.x; pos[3 * o + 1] = p.y; pos[3 * o + 2] = p.z; }
    if (nrm) { float4 p = __ldg(nrm4 + s); nrm[3 * o] = p.x; nrm[3 * o + 1] = p.y; nrm[3 * o + 2] = p.z; }
    if (lab) lab[o] = __ldg(label + s);
}

// per block {sum of edge lengths, edge count, 0, 0}; session_partial_reduce_kernel adds the blocks up in a fixed order, so
// the result does not depend on scheduling (two sessions over the same rows report bit-identical sums)
__global__ void __launch_bounds__(256) session_edge_len_kernel(Quad4 pos, const uint8_t* __restrict__ owned, const int32_t* __restrict__ idx,
                                                               int64_t n, int k, double* __restrict__ part) {
    __shared__ double red[8][2];
    double sum = 0, cnt = 0;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        if (owned && !owned[s]) continue;
        V3 c = pos(s);
        for (int a = 0; a < k; ++a) sum += (double)norm3_fma(pos((int64_t)idx[s * k + a]) - c);
        cnt += k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = sum; red[threadIdx.x >> 5][1] = cnt; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0;
        if (threadIdx.x < 2)
            for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        part[(int64_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}

// halo traffic: rows listed by tree position
__global__ void __launch_bounds__(256) session_export_kernel(const float4* __restrict__ src, const int32_t* __restrict__ rows, int64_t m,
                                                             float4* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = src[rows[i]];
}
// the same gather, but every row goes straight into the receive buffer of the rank that needs it (peer memory mapped into
// this process, stores travel over NVLink): rows [seg[p], seg[p+1]) belong to peer p and land at peer_base[p] onwards
__global__ void __launch_bounds__(256) session_export_peers_kernel(const float4* __restrict__ src, const int32_t* __restrict__ rows, int64_t m,
                                                                   const int64_t* __restrict__ seg, const unsigned long long* __restrict__ peer_base,
                                                                   int world) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    int p = 0;
    while (p + 1 < world && i >= seg[p + 1]) ++p;
    float4* dst = reinterpret_cast<float4*>(peer_base[p]);
    dst[i - seg[p]] = src[rows[i]];
}
__global__ void __launch_bounds__(256) session_import_kernel(float4* __restrict__ dst, const int32_t* __restrict__ rows, int64_t m,
                                                             const float4* __restrict__ in) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) dst[rows[i]] = in[i];
}


// ---- Morton slabs: halo rows by peer stores, cloud-wide scalars through the same symmetric block -------------------------
// One cross-rank ROUND = everybody writes into its peers' memory, signals the round number to every peer (release), and waits
// until every peer's number has arrived (acquire) before it reads what the peers wrote.  No NCCL call and no host round trip:
// a halo refresh is two kernels (gather + push, wait + scatter), a scalar all-reduce is one.  Buffers alternate with the
// parity of the round: a rank can only be one round ahead of a peer (it needs that peer's signal to get further), so the
// buffers of round t are not written again before every reader of round t has signalled round t + 1, i.e. is done reading.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// rows of this rank that peers hold as halo copies: gathered and stored straight into the owners' receive buffers; the block
// that finishes last tells every peer that this rank's rows of round `epoch` are in place
__global__ void __launch_bounds__(256) slab_push_kernel(SlabDev d, const float4* __restrict__ src, const int32_t* __restrict__ rows, int64_t m,
                                                        int parity, unsigned long long epoch, unsigned* __restrict__ done) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        int q = 0;
        while (q + 1 < d.world && i >= d.seg[q + 1]) ++q;
        d.halo(q, parity)[d.first_row[q] + (i - d.seg[q])] = src[rows[i]];
    }
    __threadfence_system();
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < d.world && (int)threadIdx.x != d.rank) st_release_sys(d.flag(threadIdx.x) + d.rank, epoch);
        if (threadIdx.x == 0) *done = 0;
    }
}
// the other half: wait for every peer's rows of round `epoch`, then scatter this rank's receive buffer into the halo rows
__global__ void __launch_bounds__(256) slab_pull_kernel(SlabDev d, float4* __restrict__ dst, const int32_t* __restrict__ rows, int64_t m,
                                                        int parity, unsigned long long epoch) {
    if (threadIdx.x < d.world && (int)threadIdx.x != d.rank) {
        const unsigned long long* f = d.flag(d.rank) + threadIdx.x;
        while (ld_acquire_sys(f) < epoch) __nanosleep(64);
    }
    __syncthreads();
    const float4* in = d.halo(d.rank, parity);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
        dst[rows[i]] = __ldcv(in + i);
}
// flat_step's cloud-wide scalars across ranks (Denoiser.py:106-107).  mode 0: acc[0..3] <- sum over ranks of acc[0..3] (added in
// rank order on every rank: identical bits everywhere); mode 1: cd[3] <- max over ranks.  One warp.
__global__ void __launch_bounds__(32) slab_allreduce_kernel(SlabDev d, double* __restrict__ acc, float* __restrict__ cd, int mode, int parity,
                                                            unsigned long long epoch) {
    // (mode 0 moves the four 8-byte words bit for bit and adds them as the integers they are: flat_step's fixed-point sums)
    const int q = threadIdx.x;
    double v[4];
    if (mode == 0) { v[0] = acc[0]; v[1] = acc[1]; v[2] = acc[2]; v[3] = acc[3]; }
    else { v[0] = (double)cd[3]; v[1] = v[2] = v[3] = 0.0; }
    if (q < d.world) {
        double* slot = d.scal(q, parity) + d.rank * 4;
        slot[0] = v[0]; slot[1] = v[1]; slot[2] = v[2]; slot[3] = v[3];
        __threadfence_system();
        if (q != d.rank) {
            st_release_sys(d.flag(q) + d.rank, epoch);
            const unsigned long long* f = d.flag(d.rank) + q;
            while (ld_acquire_sys(f) < epoch) __nanosleep(32);
        }
    }
    __syncwarp();
    const volatile double* mine = d.scal(d.rank, parity);
    if (mode == 0) {
        if (q < 4) {
            const volatile long long* mi = reinterpret_cast<const volatile long long*>(mine);
            long long t = 0;
            for (int r = 0; r < d.world; ++r) t += mi[r * 4 + q];
            reinterpret_cast<long long*>(acc)[q] = t;
        }
    } else if (q == 0) {
        double t = mine[0];
        for (int r = 1; r < d.world; ++r) t = fmax(t, mine[r * 4]);
        cd[3] = (float)t;
    }
}

// ---- order-independent checksum of the owned rows (bench.py: the lines of different GPU counts must agree bit for bit) -------
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {   // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
    return x;
}
constexpr int CHECKSUM_WORDS = 11;
// out: 0 hash of (global id, position bits), 1 hash of (global id, normal bits), 2-4 sum of x, y, z in units of 2^-24 (int64),
// 5 sum of |n| in units of 2^-24, 6-9 label histogram (0, 1, 2, other), 10 rows counted.  All integer sums: any order, any partition.
__global__ void __launch_bounds__(256) session_checksum_kernel(const float4* __restrict__ tree_pts, const float4* __restrict__ pos,
                                                               const float4* __restrict__ nrm, const uint8_t* __restrict__ label,
                                                               const uint8_t* __restrict__ owned, const int64_t* __restrict__ gids, int64_t n,
                                                               unsigned long long* __restrict__ out) {
    unsigned long long a[CHECKSUM_WORDS];
#pragma unroll
    for (int c = 0; c < CHECKSUM_WORDS; ++c) a[c] = 0;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x) {
        if (owned && !owned[s]) continue;
        const int64_t o = (int64_t)__float_as_int(__ldg(&tree_pts[s].w));
        const unsigned long long g = (unsigned long long)(gids ? gids[o] : o);
        const float4 p = pos[s], v = nrm[s];
        a[0] += mix64(g * 0x9e3779b97f4a7c15ull ^ ((unsigned long long)__float_as_uint(p.x) | ((unsigned long long)__float_as_uint(p.y) << 32))) +
                mix64(g * 0xc2b2ae3d27d4eb4full ^ (unsigned long long)__float_as_uint(p.z));
        a[1] += mix64(g * 0x9e3779b97f4a7c15ull ^ ((unsigned long long)__float_as_uint(v.x) | ((unsigned long long)__float_as_uint(v.y) << 32))) +
                mix64(g * 0xc2b2ae3d27d4eb4full ^ (unsigned long long)__float_as_uint(v.z));
        a[2] += (unsigned long long)llrint((double)p.x * 16777216.0);
        a[3] += (unsigned long long)llrint((double)p.y * 16777216.0);
        a[4] += (unsigned long long)llrint((double)p.z * 16777216.0);
        a[5] += (unsigned long long)llrint(sqrt((double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z) * 16777216.0);
        const int l = label[s];
        a[6] += l == 0; a[7] += l == 1; a[8] += l == 2; a[9] += l > 2;
        a[10] += 1;
    }
#pragma unroll
    for (int c = 0; c < CHECKSUM_WORDS; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[c] += __shfl_xor_sync(0xffffffffu, a[c], o);
        if ((threadIdx.x & 31) == 0 && a[c]) atomicAdd(out + c, a[c]);
    }
}

static inline unsigned strided(int64_t n, int threads) {
    int64_t b = cdiv(n, threads), cap = (int64_t)num_sms() * 8;
    return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

static int ensure_idx(ngpd_session* S, int k) {
    if (S->idx && S->idx_k >= k) return 0;
    if (S->idx) cudaFree(S->idx);
    S->idx = nullptr;
    NGPD_CUDA_OK(cudaMalloc(&S->idx, (size_t)S->n * k * sizeof(int32_t)));
    S->idx_k = k;
    return 0;
}

struct KnnLists {
    int32_t *list[3], *cnt[3];
    explicit KnnLists(ngpd_session* S) {
        for (int t = 0; t < 3; ++t) { list[t] = S->fix + (int64_t)t * S->n; cnt[t] = S->fix + 3 * S->n + t; }
    }
};

static inline int stride_blocks(int64_t n, int threads, int per_sm) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(n, threads), (int64_t)num_sms() * per_sm));
}

// the search tiers over `todo` (a list written by tier 0) or over every row (todo == false)
template <int K, int KT>
static void run_knn_tiers(ngpd_session* S, int k, int32_t* idx, cudaStream_t st, bool todo) {
    const GridView& g = S->grid->v;
    const float4* p = S->pos[S->cur];
    KnnLists L(S);
    KnnTrack tr{S->cand, S->anchor, S->owned ? S->halo_need : nullptr};
    uint8_t* late = S->overlap_tail ? S->late : nullptr;
    if (todo)
        session_knn_fast_kernel<K, KT><<<stride_blocks(S->n, KsCfg<1>::THREADS, 16), KsCfg<1>::THREADS, 0, st>>>(
            g, p, S->owned, S->n, k, idx, tr, L.list[0], L.cnt[0], L.list[1], L.cnt[1], late);
    else
        session_knn_fast_kernel<K, KT><<<(unsigned)cdiv(S->n, KsCfg<1>::THREADS), KsCfg<1>::THREADS, 0, st>>>(
            g, p, S->owned, S->n, k, idx, tr, nullptr, nullptr, L.list[1], L.cnt[1], late);
    cudaStream_t ts = st;
    if (late) {
        cudaEventRecord(S->tail_ev[0], st);
        cudaStreamWaitEvent(S->tail, S->tail_ev[0], 0);
        ts = S->tail;
    }
    session_knn_wide_kernel<K, KT><<<stride_blocks(S->n, KsCfg<2>::THREADS, 16), KsCfg<2>::THREADS, 0, ts>>>(g, p, k, idx, tr, L.list[1], L.cnt[1], L.list[2], L.cnt[2]);
    session_knn_fix_kernel<K, KT><<<stride_blocks(S->n, 128, 8), 128, 0, ts>>>(g, p, k, idx, tr, L.list[2], L.cnt[2]);
    if (late) {
        cudaEventRecord(S->tail_ev[1], S->tail);
        S->tail_pending = true;
    }
}

// candidate rows + anchors of the re-ranking tier for row template K, zero-filled (rows never searched -- halo -- stay valid ids,
// anchor radius 0 = nothing stored)
static int reserve_candidates(ngpd_session* S, int K, cudaStream_t st) {
    if (S->cand_cap_k == K && S->cand_k == 0) return 0;       // reserved and untouched
    if (S->cand && S->cand_cap_k != K) { cudaFree(S->cand); S->cand = nullptr; S->cand_cap_k = 0; }
    S->cand_k = 0;
    if (!S->cand) NGPD_CUDA_OK(cudaMalloc(&S->cand, (size_t)S->n * 2 * K * sizeof(int32_t)));
    S->cand_cap_k = K;
    NGPD_CUDA_OK(cudaMemsetAsync(S->cand, 0, (size_t)S->n * 2 * K * sizeof(int32_t), st));
    if (!S->anchor) NGPD_CUDA_OK(cudaMalloc(&S->anchor, (size_t)S->n * sizeof(float4)));
    NGPD_CUDA_OK(cudaMemsetAsync(S->anchor, 0, (size_t)S->n * sizeof(float4), st));
    return 0;
}

// `track`: this is the step's own search -- it may answer from, and refreshes, the stored candidates
template <int K>
static int run_knn_fast(ngpd_session* S, int k, int32_t* idx, cudaStream_t st, bool track) {
    const GridView& g = S->grid->v;
    KnnLists L(S);
    NGPD_CUDA_OK(cudaMemsetAsync(L.cnt[0], 0, 3 * sizeof(int32_t), st));
    constexpr bool CAN_TRACK = K <= 32;                      // 2K keys per lane must stay in registers (K = 32: 64 keys, 128 registers, 4 blocks per SM)
    const bool rerank_ok = S->use_rerank && CAN_TRACK;
    if (track && rerank_ok && S->cand_k != K) {
        // first search with this row length: candidate rows (allocated here unless ngpd_session_reserve did), no anchors yet
        int rc = reserve_candidates(S, K, st);
        if (rc) return rc;
        run_knn_tiers<K, CAN_TRACK ? 2 * K : K>(S, k, idx, st, false);
        S->cand_k = K;
        S->knn_launches = 3;
        return 0;
    }
    if constexpr (CAN_TRACK) {
        if (rerank_ok && S->cand_k == K) {
            // tier 0 first; what it cannot answer is searched (tracked: re-anchored, untracked: just answered)
            // A/B measurements (profiles/): NGPD_RERANK=ldg  round 1's kernel (coalesced loads, transposed tile, rows stored through the tile);
            //                               NGPD_RERANK=rows candidate rows AND result rows by one bulk copy per thread (measured: 15 % slower --
            //                                               128 small bulk copies per block cost the TMA unit more than the stores they replace);
            //                               default        coalesced loads + transposed tile for the candidates, result rows by one bulk copy per warp
            static const char* mode_env = getenv("NGPD_RERANK");
            static const int mode = !mode_env ? 0 : (!strcmp(mode_env, "ldg") ? 1 : (!strcmp(mode_env, "rows") ? 2 : 0));
            const KnnTrack trk{S->cand, S->anchor, S->owned ? S->halo_need : nullptr};
            const unsigned rb = (unsigned)cdiv(S->n, 128);
            if (mode == 1)
                session_knn_rerank_kernel<K, true><<<rb, 128, 0, st>>>(g, S->pos[S->cur], S->owned, S->n, k, idx, trk, L.list[0], L.cnt[0]);
            else if (mode == 2)
                session_knn_rerank_tma_kernel<K><<<rb, 128, 0, st>>>(g, S->pos[S->cur], S->owned, S->n, k, idx, trk, L.list[0], L.cnt[0]);
            else
                session_knn_rerank_kernel<K, false><<<rb, 128, 0, st>>>(g, S->pos[S->cur], S->owned, S->n, k, idx, trk, L.list[0], L.cnt[0]);
            if (track) run_knn_tiers<K, 2 * K>(S, k, idx, st, true);
            else run_knn_tiers<K, K>(S, k, idx, st, true);
            S->knn_launches = 4;
            return 0;
        }
    }
    run_knn_tiers<K, K>(S, k, idx, st, false);
    S->knn_launches = 3;
    return 0;
}

static int run_knn(ngpd_session* S, int k, int32_t* idx, cudaStream_t st, bool track = false) {
    unsigned b = (unsigned)cdiv(S->n, 128);
    const GridView& g = S->grid->v;
    const float4* p = S->pos[S->cur];
    const bool fast = !S->exact_only && k > 4 && k <= 64;
    int rc = 0;
    S->knn_launches = 1;
    if (fast) {
        // a shorter row is a prefix of a longer one (same total order): reuse the stored candidates' template when it fits
        const int kt = (S->use_rerank && S->cand_k >= k && !track) ? S->cand_k : k;
        if (kt <= 8) rc = run_knn_fast<8>(S, k, idx, st, track);
        else if (kt <= 16) rc = run_knn_fast<16>(S, k, idx, st, track);
        else if (kt <= 32) rc = run_knn_fast<32>(S, k, idx, st, track);
        else rc = run_knn_fast<64>(S, k, idx, st, track);
    }
    else if (k <= 4) session_knn_kernel<4><<<b, 128, 0, st>>>(g, p, S->owned, S->n, k, idx, S->owned ? S->halo_need : nullptr);
    else if (k <= 8) session_knn_kernel<8><<<b, 128, 0, st>>>(g, p, S->owned, S->n, k, idx, S->owned ? S->halo_need : nullptr);
    else if (k <= 16) session_knn_kernel<16><<<b, 128, 0, st>>>(g, p, S->owned, S->n, k, idx, S->owned ? S->halo_need : nullptr);
    else if (k <= 32) session_knn_kernel<32><<<b, 128, 0, st>>>(g, p, S->owned, S->n, k, idx, S->owned ? S->halo_need : nullptr);
    else session_knn_kernel<64><<<b, 128, 0, st>>>(g, p, S->owned, S->n, k, idx, S->owned ? S->halo_need : nullptr);
    if (rc) return rc;
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ngpd

using namespace ngpd;

// completes a halo refresh whose second half (wait + scatter) was deferred; defined with the slab code below.  Every entry that
// reads or replaces positions of halo rows, or starts another cross-rank round, calls it first.
static int slab_flush(ngpd_session_t* S, cudaStream_t st);

extern "C" __attribute__((visibility("default"))) int ngpd_session_destroy(ngpd_session_t* S) {
    if (!S) return 0;
    if (S->grid) ngpd_grid_destroy(S->grid);
    void* bufs[] = {S->pos[0], S->pos[1], S->nrm, S->fn, S->edge, S->label, S->owned, S->idx, S->acc, S->cd, S->fix, S->cand, S->anchor, S->part, S->red2, S->cls, S->inv, S->stage_pos, S->stage_nrm, S->stage_lab, S->orig, S->blockfar, S->cref};
    for (void* b : bufs) if (b) cudaFree(b);
    if (S->side) { cudaStreamDestroy(S->side); for (cudaEvent_t e : S->xfer) if (e) cudaEventDestroy(e); }
    if (S->tail) { cudaStreamDestroy(S->tail); for (cudaEvent_t e : S->tail_ev) if (e) cudaEventDestroy(e); }
    if (S->late) cudaFree(S->late);
    if (S->slab_done) cudaFree(S->slab_done);
    if (S->halo_need) cudaFree(S->halo_need);
    delete S->slab;
    delete S;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_create(const float* tree_pos, int64_t n, int k_hint, void* stream_, ngpd_session_t** out) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(out, "ngpd_session_create: out is NULL");
    *out = nullptr;
    ngpd_grid* G = nullptr;
    int rc = ngpd_grid_create(tree_pos, n, 0.0f, k_hint, stream_, &G);
    if (rc) return rc;
    ngpd_session* S = new ngpd_session();
    S->grid = G; S->n = n;
    {
        // fixed-point unit of flat_step's centre sums: a row adds up to 64 neighbour positions, positions may drift to twice the
        // tree's extent -> |row sum| <= 128 max|coordinate|, mapped below 2^32 (2^31 rows then still fit 63 bits)
        float maxabs = 0.0f;
        for (int c = 0; c < 6; ++c) maxabs = std::max(maxabs, std::fabs(G->bbox[c]));
        int e = 0;
        std::frexp(128.0 * (double)std::max(maxabs, 1e-30f), &e);      // 128 maxabs < 2^e
        S->sum_scale = std::ldexp(1.0f, std::max(-100, std::min(100, 32 - e)));
    }
    size_t b4 = (size_t)n * sizeof(float4);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&S->pos[0], b4);
    if (e == cudaSuccess) e = cudaMalloc(&S->pos[1], b4);
    if (e == cudaSuccess) e = cudaMalloc(&S->nrm, b4);
    if (e == cudaSuccess) e = cudaMalloc(&S->fn, b4);
    if (e == cudaSuccess) e = cudaMalloc(&S->edge, b4);
    if (e == cudaSuccess) e = cudaMalloc(&S->label, (size_t)n);
    if (e == cudaSuccess) e = cudaMalloc(&S->acc, 4 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&S->cd, 4 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&S->cref, 4 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&S->red2, (RED_BLOCKS * 4 + 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&S->fix, (3 * (size_t)n + 3) * sizeof(int32_t));
    if (e != cudaSuccess) { set_error("ngpd_session_create: %s", cudaGetErrorString(e)); ngpd_session_destroy(S); return -2; }
    NGPD_CUDA_OK(cudaMemsetAsync(S->label, 0, (size_t)n, st));
    {
        // first reference centre of flat_step's pruned delta pass: the middle of the tree's bounding box (any point is valid)
        const float c0[4] = {0.5f * (G->bbox[0] + G->bbox[3]), 0.5f * (G->bbox[1] + G->bbox[4]), 0.5f * (G->bbox[2] + G->bbox[5]), 0.0f};
        NGPD_CUDA_OK(cudaMemcpyAsync(S->cref, c0, sizeof(c0), cudaMemcpyHostToDevice, st));
        NGPD_CUDA_OK(cudaStreamSynchronize(st));
    }
    NGPD_CUDA_OK(cudaMemsetAsync(S->red2, 0, (RED_BLOCKS * 4 + 1) * sizeof(double), st));
    NGPD_CUDA_OK(cudaMemsetAsync(S->nrm, 0, b4, st));
    // current positions start as the tree positions
    session_scatter_in_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(G->pts, tree_pos, nullptr, n, S->pos[0], nullptr);
    NGPD_CUDA_OK(cudaGetLastError());
    *out = S;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_set_state(ngpd_session_t* S, const float* pos, const float* nrm, void* stream_) {
    NGPD_REQUIRE(S, "ngpd_session_set_state: NULL session");
    { int rc = slab_flush(S, (cudaStream_t)stream_); if (rc) return rc; }
    S->sums_ready = false;
    // (positions replaced from outside keep the stored kNN candidates usable: tier 0 measures the displacement from the anchor)
    session_scatter_in_kernel<<<(unsigned)cdiv(S->n, 256), 256, 0, (cudaStream_t)stream_>>>(S->grid->pts, pos, nrm, S->n, S->pos[S->cur], S->nrm);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_get_state(ngpd_session_t* S, float* pos_out, float* nrm_out, uint8_t* labels_out, void* stream_) {
    NGPD_REQUIRE(S, "ngpd_session_get_state: NULL session");
    { int rc = slab_flush(S, (cudaStream_t)stream_); if (rc) return rc; }
    if (!S->inv) {
        NGPD_CUDA_OK(cudaMalloc(&S->inv, (size_t)S->n * sizeof(int32_t)));
        session_inverse_kernel<<<(unsigned)cdiv(S->n, 256), 256, 0, (cudaStream_t)stream_>>>(S->grid->pts, S->n, S->inv);
    }
    session_scatter_out_kernel<<<(unsigned)cdiv(S->n, 256), 256, 0, (cudaStream_t)stream_>>>(S->inv, S->pos[S->cur], S->nrm, S->label, S->n,
                                                                                             pos_out, nrm_out, labels_out);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_set_owned(ngpd_session_t* S, const uint8_t* owned_tree_order, void* stream_) {
    NGPD_REQUIRE(S, "ngpd_session_set_owned: NULL session");
    S->sums_ready = false;
    S->lists_ready = false;
    if (!owned_tree_order) { if (S->owned) cudaFree(S->owned); S->owned = nullptr; return 0; }
    if (!S->owned) NGPD_CUDA_OK(cudaMalloc(&S->owned, (size_t)S->n));
    if (!S->halo_need) {
        NGPD_CUDA_OK(cudaMalloc(&S->halo_need, sizeof(float)));
        NGPD_CUDA_OK(cudaMemsetAsync(S->halo_need, 0, sizeof(float), (cudaStream_t)stream_));
    }
    NGPD_CUDA_OK(cudaMemcpyAsync(S->owned, owned_tree_order, (size_t)S->n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_order(const ngpd_session_t* S, int32_t* perm_out, void* stream_) {
    NGPD_REQUIRE(S, "ngpd_session_order: NULL session");
    return ngpd_grid_order(S->grid, perm_out, stream_);
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_launch_count(const ngpd_session_t* S) { return S ? S->launches : 0; }

static int ensure_tail(ngpd_session_t* S, cudaStream_t st) {
    if (S->tail) return 0;
    // highest priority: its few blocks must get SM slots as the tensor pass' blocks retire, not after all of them
    int prio_lo = 0, prio_hi = 0;
    NGPD_CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    NGPD_CUDA_OK(cudaStreamCreateWithPriority(&S->tail, cudaStreamNonBlocking, prio_hi));
    for (cudaEvent_t& e : S->tail_ev) NGPD_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    NGPD_CUDA_OK(cudaMalloc(&S->late, (size_t)S->n));
    NGPD_CUDA_OK(cudaMemsetAsync(S->late, 0, (size_t)S->n, st));
    return 0;
}

// Everything a step with neighbourhood size k_feature allocates on first use (neighbour table, candidate rows of the re-ranking
// tier, class lists, partial sums, side stream) now, so that the first iteration of a run costs what its kernels cost.
extern "C" __attribute__((visibility("default"))) int ngpd_session_reserve(ngpd_session_t* S, int k_feature, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(S && k_feature >= 1 && k_feature <= 64, "ngpd_session_reserve: bad argument");
    int rc = ensure_idx(S, k_feature);
    if (rc) return rc;
    if ((rc = ensure_tail(S, st))) return rc;
    if (S->use_rerank && k_feature > 4 && k_feature <= 32 && S->cand_k == 0) {
        const int K = k_feature <= 8 ? 8 : (k_feature <= 16 ? 16 : 32);
        if ((rc = reserve_candidates(S, K, st))) return rc;
    }
    if (!S->part) NGPD_CUDA_OK(cudaMalloc(&S->part, (size_t)cdiv(S->n, 128) * 4 * sizeof(long long)));
    if (!S->blockfar) NGPD_CUDA_OK(cudaMalloc(&S->blockfar, ((size_t)cdiv(S->n, 128) + 1) * sizeof(float)));
    if (!S->cls) NGPD_CUDA_OK(cudaMalloc(&S->cls, (2 * (size_t)S->n + 2) * sizeof(int32_t)));
    return 0;
}

// ---- phases (exposed one by one so that a multi-GPU driver can exchange halos in between) -----------
// part 0 of the feature phase in its two halves (the host-buffer entry point overlaps transfers with each of them)
static int features_knn(ngpd_session_t* S, const ngpd_step_params_t* p, cudaStream_t st) {
    const int kf = p->k_feature;
    int rc = ensure_idx(S, kf);
    if (rc) return rc;
    S->idx_k = kf;
    if ((rc = ensure_tail(S, st))) return rc;
    // Round 1 ran the last two search tiers on a high-priority side stream under the first tensor pass (-0.09 ms per iteration at
    // 10 M points then).  Measured again in round 2 with the wider 5x5x5 tier (36 KB of shared memory per block): equal at 10 M
    // points, but at 100 M points the iteration is 3 ms SLOWER with the overlap (28.1 vs 25.1 ms; ncu: the kernels themselves
    // add up to the shorter figure) -- the search blocks' shared memory is carved out of the L1 that the gather-bound tensor
    // pass lives on (L1 hit 80 %), for as long as the two kernels share the SMs.  Off by default; NGPD_TAIL_OVERLAP=1 turns it on.
    static const bool tail_on = getenv("NGPD_TAIL_OVERLAP") != nullptr;
    S->overlap_tail = tail_on;     // honoured by the tiered search only (5 <= k <= 32, not in exact-only mode)
    { ProfScope ps(S, st, 0); rc = run_knn(S, kf, S->idx, st, true); }
    S->overlap_tail = false;
    if (rc) return rc;
    S->launches += S->knn_launches;
    return 0;
}
static int features_smooth(ngpd_session_t* S, const ngpd_step_params_t* p, cudaStream_t st) {
    const int kf = p->k_feature;
    unsigned b = (unsigned)cdiv(S->n, 128);
    Quad4 pos{S->pos[S->cur]};
    { ProfScope ps(S, st, 1);
      Quad4 nq{S->nrm};
      const uint8_t* late = S->tail_pending ? S->late : nullptr;
      if (kf == 16) session_nvt_smooth_kernel<16><<<b, 128, 0, st>>>(pos, nq, S->owned, late, S->idx, S->n, kf, p->x_thresh, p->tau, p->damp, S->fn);
      else if (kf == 32) session_nvt_smooth_kernel<32><<<b, 128, 0, st>>>(pos, nq, S->owned, late, S->idx, S->n, kf, p->x_thresh, p->tau, p->damp, S->fn);
      else if (kf == 8) session_nvt_smooth_kernel<8><<<b, 128, 0, st>>>(pos, nq, S->owned, late, S->idx, S->n, kf, p->x_thresh, p->tau, p->damp, S->fn);
      else session_nvt_smooth_kernel<0><<<b, 128, 0, st>>>(pos, nq, S->owned, late, S->idx, S->n, kf, p->x_thresh, p->tau, p->damp, S->fn);
      if (S->tail_pending) {
          // the rows the last two search tiers were still working on
          KnnLists L(S);
          NGPD_CUDA_OK(cudaStreamWaitEvent(st, S->tail_ev[1], 0));
          const int lb = stride_blocks(S->n, 128, 8);
          if (kf == 16) session_nvt_smooth_late_kernel<16><<<lb, 128, 0, st>>>(pos, nq, L.list[1], L.cnt[1], S->late, S->idx, kf, p->x_thresh, p->tau, p->damp, S->fn);
          else if (kf == 32) session_nvt_smooth_late_kernel<32><<<lb, 128, 0, st>>>(pos, nq, L.list[1], L.cnt[1], S->late, S->idx, kf, p->x_thresh, p->tau, p->damp, S->fn);
          else if (kf == 8) session_nvt_smooth_late_kernel<8><<<lb, 128, 0, st>>>(pos, nq, L.list[1], L.cnt[1], S->late, S->idx, kf, p->x_thresh, p->tau, p->damp, S->fn);
          else session_nvt_smooth_late_kernel<0><<<lb, 128, 0, st>>>(pos, nq, L.list[1], L.cnt[1], S->late, S->idx, kf, p->x_thresh, p->tau, p->damp, S->fn);
          S->tail_pending = false;
          S->launches += 1;
      } }
    S->launches += 1;
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_phase_features(ngpd_session_t* S, const ngpd_step_params_t* p, int part, void* stream_) {
    // part 0: knn + NVT + smoothing (writes fn);  part 1: NVT on fn + labels + crease direction
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(S && p, "ngpd_session_phase_features: NULL argument");
    const int kf = p->k_feature;
    NGPD_REQUIRE(kf >= 1 && kf <= 64 && p->k_update >= 1 && p->k_update <= kf, "ngpd_session: need 1 <= k_update <= k_feature <= 64");
    unsigned b = (unsigned)cdiv(S->n, 128);
    Quad4 pos{S->pos[S->cur]};
    if (part == 0) {
        // (the search reads the queries' own rows only: a deferred halo refresh of the positions completes behind it)
        int rc = features_knn(S, p, st);
        if (!rc) rc = slab_flush(S, st);
        if (!rc) rc = features_smooth(S, p, st);
        if (rc) return rc;
    } else {
        { int rc = slab_flush(S, st); if (rc) return rc; }
        // class 0 moves first: when it uses the flat strategy its neighbour sums come out of this kernel
        long long* part = nullptr;
        if (p->strategy[0] == NGPD_STEP_FLAT) {
            if (!S->part) NGPD_CUDA_OK(cudaMalloc(&S->part, (size_t)b * 4 * sizeof(long long)));
            part = S->part;
            if (!S->blockfar) NGPD_CUDA_OK(cudaMalloc(&S->blockfar, ((size_t)b + 1) * sizeof(float)));
            NGPD_CUDA_OK(cudaMemsetAsync(S->blockfar + b, 0, sizeof(float), st));
        }
        S->sums_ready = part != nullptr;
        if (!S->cls) NGPD_CUDA_OK(cudaMalloc(&S->cls, (2 * (size_t)S->n + 2) * sizeof(int32_t)));
        int32_t* cls = S->cls;
        NGPD_CUDA_OK(cudaMemsetAsync(cls + 2 * S->n, 0, 2 * sizeof(int32_t), st));
        S->lists_ready = true;
        // which labels' rows need their crease direction (edge_step's y, Processor.py:134): those rows keep the LAPACK-order solver,
        // all others get their label from the closed form; with edge_step on class 0 (the majority) there is nothing to gain
        int edge_mask = 0;
        for (int key = 0; key < 3; ++key) if (p->strategy[key] == NGPD_STEP_EDGE) edge_mask |= 1 << key;
        static const bool fast_off = getenv("NGPD_NO_FAST_LABELS") != nullptr;   // measurements / A-B tests only
        const bool fast = !fast_off && (edge_mask & 1) == 0;
        { ProfScope ps(S, st, 2);
          Quad4 fq{S->fn};
#define NGPD_CLASSIFY(KK, FF) session_nvt_classify_kernel<KK, FF><<<b, 128, 0, st>>>(pos, fq, S->owned, S->idx, S->n, kf, p->x_thresh, p->scale, S->label, S->edge, edge_mask, 0, p->k_update, part, S->sum_scale, S->cref, part ? S->blockfar : nullptr, cls)
          if (fast) {
              if (kf == 16) NGPD_CLASSIFY(16, true); else if (kf == 32) NGPD_CLASSIFY(32, true); else if (kf == 8) NGPD_CLASSIFY(8, true); else NGPD_CLASSIFY(0, true);
          } else {
              if (kf == 16) NGPD_CLASSIFY(16, false); else if (kf == 32) NGPD_CLASSIFY(32, false); else if (kf == 8) NGPD_CLASSIFY(8, false); else NGPD_CLASSIFY(0, false);
          }
#undef NGPD_CLASSIFY
        }
        S->launches += 1;
    }
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

// flat-step scalars of class `key`: part 0 accumulates {sum xyz, count} into acc (4 doubles, device), part 1
// turns acc into the centre and accumulates the max distance.  Between the parts a multi-GPU driver all-reduces
// acc (sum) and afterwards cd[3] (max).
extern "C" __attribute__((visibility("default"))) int ngpd_session_phase_flat_scalars(ngpd_session_t* S, const ngpd_step_params_t* p, int key, int part, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(S && p, "ngpd_session_phase_flat_scalars: NULL argument");
    { int rc = slab_flush(S, st); if (rc) return rc; }
    // (snapshot mode: classes 1 and 2 still read the positions class 0's pass started from)
    Quad4 pos{S->pos[((p->flags & NGPD_STEP_SNAPSHOT_CLASSES) && key > 0) ? S->cur ^ 1 : S->cur]};
    ProfScope ps(S, st, 3);
    if (part == 0 && key == 0 && S->sums_ready) {
        session_partial_reduce_kernel<long long><<<RED_BLOCKS, 256, 0, st>>>(S->part, cdiv(S->n, 128), reinterpret_cast<long long*>(S->red2),
                                                                             reinterpret_cast<long long*>(S->acc));
        S->launches += 1;
    } else if (part == 0) {
        if (!S->part) NGPD_CUDA_OK(cudaMalloc(&S->part, (size_t)cdiv(S->n, 128) * 4 * sizeof(long long)));
        const unsigned blocks = strided(S->n, 256);
        session_class_sum_kernel<<<blocks, 256, 0, st>>>(pos, S->owned, S->label, key, S->idx, S->n, S->idx_k, p->k_update, S->part, S->sum_scale);
        session_partial_reduce_kernel<long long><<<RED_BLOCKS, 256, 0, st>>>(S->part, (int64_t)blocks, reinterpret_cast<long long*>(S->red2),
                                                                             reinterpret_cast<long long*>(S->acc));
        S->sums_ready = false;
        S->launches += 2;
    } else {
        static const bool full_pass = getenv("NGPD_DELTA_FULL_PASS") != nullptr;   // A/B measurements: round 1's pass over every row
        if (key == 0 && S->sums_ready && S->blockfar && !full_pass) {
            // the stage-2 kernel left per-block maxima: only the blocks that can hold the farthest neighbour are looked at again
            session_center_kernel<<<1, 1, 0, st>>>(reinterpret_cast<const long long*>(S->acc), S->sum_scale, S->cd, S->cref);
            session_class_max_pruned_kernel<<<strided(cdiv(S->n, 128) * 32, 256), 256, 0, st>>>(pos, S->owned, S->label, key, S->idx, S->n, S->idx_k, p->k_update,
                                                                                                S->blockfar, cdiv(S->n, 128), S->cref, S->cd);
        } else {
            session_center_kernel<<<1, 1, 0, st>>>(reinterpret_cast<const long long*>(S->acc), S->sum_scale, S->cd, nullptr);
            session_class_max_kernel<<<strided(S->n, 256), 256, 0, st>>>(pos, S->owned, S->label, key, S->idx, S->n, S->idx_k, p->k_update, S->cd);
        }
        S->launches += 2;
    }
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_phase_update(ngpd_session_t* S, const ngpd_step_params_t* p, int key, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(S && p && key >= 0 && key < 3, "ngpd_session_phase_update: bad argument");
    { int rc = slab_flush(S, st); if (rc) return rc; }
    int kind = p->strategy[key];
    const bool snapshot = (p->flags & NGPD_STEP_SNAPSHOT_CLASSES) != 0;
    // snapshot mode: class 0's pass doubles as the copy of the snapshot into the other buffer, so it runs even when class 0 stays put
    if (kind < 0 && !(snapshot && key == 0)) return 0;
    const float4* orig = p->clamp_radius > 0.0f ? S->orig : nullptr;
    NGPD_REQUIRE(!(p->clamp_radius > 0.0f) || S->orig, "ngpd_session_phase_update: clamp_radius needs ngpd_session_set_original first");
    NGPD_REQUIRE(!snapshot || S->lists_ready, "ngpd_session_phase_update: the snapshot mode needs the class lists of phase_features part 1");
    ProfScope ps(S, st, 4);
    const bool fixed8 = p->k_update == 8 && S->idx_k % 4 == 0;   // rows of 8 ids, 16-byte aligned
    if (key > 0 && S->lists_ready) {
        // minority class: its rows only.  Sequential mode: in place (the other position buffer is free and holds the moved rows
        // in between).  Snapshot mode: read the snapshot (the buffer class 0's pass read), write the rows of the current buffer.
        const int32_t *list = S->cls + (int64_t)(key - 1) * S->n, *count = S->cls + 2 * S->n + (key - 1);
        const float4* src = S->pos[snapshot ? S->cur ^ 1 : S->cur];
        float4* dst = S->pos[snapshot ? S->cur : S->cur ^ 1];
        if (fixed8)
            session_update_rows_kernel<8><<<stride_blocks(S->n, 128, 16), 128, 0, st>>>(kind, Quad4{src}, Quad4{S->fn}, S->edge, list, count, S->idx,
                                                                                     S->idx_k, p->k_update, p->alpha[key], p->dmax, S->cd, orig, p->clamp_radius, dst, snapshot);
        else
            session_update_rows_kernel<0><<<stride_blocks(S->n, 128, 16), 128, 0, st>>>(kind, Quad4{src}, Quad4{S->fn}, S->edge, list, count, S->idx,
                                                                                     S->idx_k, p->k_update, p->alpha[key], p->dmax, S->cd, orig, p->clamp_radius, dst, snapshot);
        S->launches += 1;
        if (!snapshot) {
            session_apply_rows_kernel<<<stride_blocks(S->n, 256, 8), 256, 0, st>>>(list, count, S->pos[S->cur ^ 1], S->pos[S->cur]);
            S->launches += 1;
        }
        NGPD_CUDA_OK(cudaGetLastError());
        S->sums_ready = false;
        return 0;
    }
    // registers per thread against resident warps: see session_update_kernel; NGPD_UPDATE_BLOCKS selects for A/B measurements (profiles/)
    static const char* ub_env = getenv("NGPD_UPDATE_BLOCKS");
    static const int ub = ub_env ? atoi(ub_env) : 0;
#define NGPD_UPDATE(KU_, MB_) session_update_kernel<KU_, MB_><<<(unsigned)cdiv(S->n, 128), 128, 0, st>>>(kind, key, Quad4{S->pos[S->cur]}, Quad4{S->fn}, S->edge, S->label, S->owned, \
                                                                            S->idx, S->n, S->idx_k, p->k_update, p->alpha[key], p->dmax, S->cd, orig, p->clamp_radius, S->pos[S->cur ^ 1])
    if (fixed8) {
        // default: 16 blocks per SM for flat_step, 10 for the others; NGPD_UPDATE_BLOCKS = 6 / 8 / 10 / 12 / 16 forces one (A/B runs)
        const int blocks = ub ? ub : (kind == NGPD_STEP_FLAT ? 16 : 10);
        if (blocks == 6) NGPD_UPDATE(8, 6); else if (blocks == 8) NGPD_UPDATE(8, 8); else if (blocks == 12) NGPD_UPDATE(8, 12);
        else if (blocks == 16) NGPD_UPDATE(8, 16); else NGPD_UPDATE(8, 10);
    }
    else NGPD_UPDATE(0, 10);
#undef NGPD_UPDATE
    NGPD_CUDA_OK(cudaGetLastError());
    S->cur ^= 1;
    S->sums_ready = false;
    S->launches += 1;
    return 0;
}

// the positions a clamped run measures its displacement from (original point order; NULL drops them)
extern "C" __attribute__((visibility("default"))) int ngpd_session_set_original(ngpd_session_t* S, const float* pos, void* stream_) {
    NGPD_REQUIRE(S, "ngpd_session_set_original: NULL session");
    if (!pos) { if (S->orig) cudaFree(S->orig); S->orig = nullptr; return 0; }
    if (!S->orig) NGPD_CUDA_OK(cudaMalloc(&S->orig, (size_t)S->n * sizeof(float4)));
    session_scatter_in_kernel<<<(unsigned)cdiv(S->n, 256), 256, 0, (cudaStream_t)stream_>>>(S->grid->pts, pos, nullptr, S->n, S->orig, nullptr);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_phase_commit_normals(ngpd_session_t* S) {
    NGPD_REQUIRE(S, "ngpd_session_phase_commit_normals: NULL session");
    float4* t = S->nrm; S->nrm = S->fn; S->fn = t;   // graph.n = f_n  (Processor.py:139)
    return 0;
}

// everything after the smoothed normals: labels, then the class-sequential position update, then graph.n = f_n
static int step_tail(ngpd_session_t* S, const ngpd_step_params_t* p, void* stream_) {
    int rc;
    if ((rc = ngpd_session_phase_features(S, p, 1, stream_))) return rc;
    for (int key = 0; key < 3; ++key) {
        if (p->strategy[key] == NGPD_STEP_FLAT) {
            if ((rc = ngpd_session_phase_flat_scalars(S, p, key, 0, stream_))) return rc;
            if ((rc = ngpd_session_phase_flat_scalars(S, p, key, 1, stream_))) return rc;
        }
        if ((rc = ngpd_session_phase_update(S, p, key, stream_))) return rc;   // (a class without a strategy is left alone in there)
    }
    return ngpd_session_phase_commit_normals(S);
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_step(ngpd_session_t* S, const ngpd_step_params_t* p, void* stream_) {
    NGPD_REQUIRE(S && p, "ngpd_session_step: NULL argument");
    S->launches = 0;
    int rc;
    if ((rc = ngpd_session_phase_features(S, p, 0, stream_))) return rc;
    return step_tail(S, p, stream_);
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_mean_edge_length(ngpd_session_t* S, int k, double* out_host, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(S && out_host && k >= 1 && k <= 64, "ngpd_session_mean_edge_length: bad argument");
    { int rc = slab_flush(S, st); if (rc) return rc; }
    // stream-ordered scratch (pool-backed after the first call; returned on every exit path)
    StreamBuf<int32_t> idx(st);
    StreamBuf<double> acc(st), part(st);
    const unsigned blocks = strided(S->n, 256);
    NGPD_CUDA_OK(idx.alloc((size_t)S->n * k));
    NGPD_CUDA_OK(acc.alloc(4));
    NGPD_CUDA_OK(part.alloc((size_t)blocks * 4));
    int rc = run_knn(S, k, idx, st);
    if (rc) return rc;
    session_edge_len_kernel<<<blocks, 256, 0, st>>>(Quad4{S->pos[S->cur]}, S->owned, idx, S->n, k, part);
    session_partial_reduce_kernel<double><<<RED_BLOCKS, 256, 0, st>>>(part, (int64_t)blocks, S->red2, acc);
    double h[2];
    NGPD_CUDA_OK(cudaMemcpyAsync(h, acc, sizeof(h), cudaMemcpyDeviceToHost, st));
    NGPD_CUDA_OK(cudaStreamSynchronize(st));
    out_host[0] = h[0]; out_host[1] = h[1];   // {sum of edge lengths, edge count}: callers divide (and all-reduce first on multi-GPU)
    return 0;
}

// knn_mode 0: re-ranking tier + streaming tiers + exact fix-up (default); 1: exact shell search for every query;
// 2: streaming tiers without the re-ranking tier (measurements, tests)
extern "C" __attribute__((visibility("default"))) int ngpd_session_set_knn_mode(ngpd_session_t* S, int mode) {
    NGPD_REQUIRE(S, "ngpd_session_set_knn_mode: NULL session");
    S->exact_only = mode == 1;
    S->use_rerank = mode == 0;
    return 0;
}
// number of queries the last kNN pass handed to the exact search (synchronises the stream)
extern "C" __attribute__((visibility("default"))) int ngpd_session_last_fixups(ngpd_session_t* S, void* stream_) {
    if (!S) return -1;
    int32_t c = 0;
    if (cudaMemcpyAsync(&c, S->fix + 3 * S->n + 2, sizeof(c), cudaMemcpyDeviceToHost, (cudaStream_t)stream_) != cudaSuccess) return -1;
    cudaStreamSynchronize((cudaStream_t)stream_);
    return c;
}
// {rows the re-ranking tier handed to the search, rows the 3x3x3 tier handed on, rows the 5x5x5 tier handed to the exact
// search} of the last kNN pass (synchronises)
extern "C" __attribute__((visibility("default"))) int ngpd_session_knn_stats(ngpd_session_t* S, int32_t* out3_host, void* stream_) {
    NGPD_REQUIRE(S && out3_host, "ngpd_session_knn_stats: NULL argument");
    NGPD_CUDA_OK(cudaMemcpyAsync(out3_host, S->fix + 3 * S->n, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
    NGPD_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream_));
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_set_profiling(ngpd_session_t* S, int on) {
    NGPD_REQUIRE(S, "ngpd_session_set_profiling: NULL session");
    S->profiling = on != 0;
    return 0;
}
// accumulated device time per kernel category since the last call (synchronises the recorded events)
extern "C" __attribute__((visibility("default"))) int ngpd_session_get_profile(ngpd_session_t* S, double* ms_out, int32_t* launches_out) {
    NGPD_REQUIRE(S && ms_out && launches_out, "ngpd_session_get_profile: NULL argument");
    for (size_t i = 0; i < S->ev_cat.size(); ++i) {
        float ms = 0;
        cudaEventSynchronize(S->ev[2 * i + 1]);
        cudaEventElapsedTime(&ms, S->ev[2 * i], S->ev[2 * i + 1]);
        S->prof_ms[S->ev_cat[i]] += ms; S->prof_n[S->ev_cat[i]] += 1;
        cudaEventDestroy(S->ev[2 * i]); cudaEventDestroy(S->ev[2 * i + 1]);
    }
    S->ev.clear(); S->ev_cat.clear();
    for (int c = 0; c < NGPD_PROF_CATEGORIES; ++c) { ms_out[c] = S->prof_ms[c]; launches_out[c] = S->prof_n[c]; S->prof_ms[c] = 0; S->prof_n[c] = 0; }
    return 0;
}

// raw device views for a multi-GPU driver and for tests: which = 0 pos, 1 nrm, 2 fn, 3 acc (double*), 4 cd (float*), 5 label, 6 idx,
// 7 kNN hand-over lists (int32: n rows each of tiers 0, 1, 2, then the three counters)
extern "C" __attribute__((visibility("default"))) void* ngpd_session_buffer(ngpd_session_t* S, int which) {
    if (!S) return nullptr;
    switch (which) {
        case 0: return S->pos[S->cur];
        case 1: return S->nrm;
        case 2: return S->fn;
        case 3: return S->acc;
        case 4: return S->cd;
        case 5: return S->label;
        case 6: return S->idx;
        case 7: return S->fix;
        default: return nullptr;
    }
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_export_rows(ngpd_session_t* S, int which, const int32_t* rows, int64_t m, float* out4, void* stream_) {
    NGPD_REQUIRE(S && (which >= 0 && which <= 2), "ngpd_session_export_rows: bad argument");
    { int rc = slab_flush(S, (cudaStream_t)stream_); if (rc) return rc; }
    if (m <= 0) return 0;
    const float4* src = which == 0 ? S->pos[S->cur] : (which == 1 ? S->nrm : S->fn);
    session_export_kernel<<<(unsigned)cdiv(m, 256), 256, 0, (cudaStream_t)stream_>>>(src, rows, m, (float4*)out4);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" __attribute__((visibility("default"))) int ngpd_session_export_rows_peers(ngpd_session_t* S, int which, const int32_t* rows, int64_t m,
                                                                                     const int64_t* seg, const uint64_t* peer_base, int world, void* stream_) {
    NGPD_REQUIRE(S && (which >= 0 && which <= 2) && world >= 1 && (m <= 0 || (rows && seg && peer_base)), "ngpd_session_export_rows_peers: bad argument");
    if (m <= 0) return 0;
    const float4* src = which == 0 ? S->pos[S->cur] : (which == 1 ? S->nrm : S->fn);
    session_export_peers_kernel<<<(unsigned)cdiv(m, 256), 256, 0, (cudaStream_t)stream_>>>(src, rows, m, seg, reinterpret_cast<const unsigned long long*>(peer_base), world);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}
extern "C" __attribute__((visibility("default"))) int ngpd_session_import_rows(ngpd_session_t* S, int which, const int32_t* rows, int64_t m, const float* in4, void* stream_) {
    NGPD_REQUIRE(S && (which >= 0 && which <= 2), "ngpd_session_import_rows: bad argument");
    { int rc = slab_flush(S, (cudaStream_t)stream_); if (rc) return rc; }
    if (m <= 0) return 0;
    float4* dst = which == 0 ? S->pos[S->cur] : (which == 1 ? S->nrm : S->fn);
    if (which == 0) S->sums_ready = false;
    session_import_kernel<<<(unsigned)cdiv(m, 256), 256, 0, (cudaStream_t)stream_>>>(dst, rows, m, (const float4*)in4);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}


// ---- Morton slabs: wiring, halo refresh, cross-rank scalars and the whole step driven from here -------------------------
extern "C" __attribute__((visibility("default"))) int64_t ngpd_slab_symm_bytes(int world, int64_t cap) {
    if (world < 1 || world > SLAB_MAX_WORLD || cap < 0) return -1;
    return 2 * cap * 16 + 2 * (int64_t)world * 32 + (int64_t)world * 8;
}

extern "C" __attribute__((visibility("default"))) int ngpd_session_set_slab(ngpd_session_t* S, const ngpd_slab_wiring_t* w, void* stream_) {
    NGPD_REQUIRE(S, "ngpd_session_set_slab: NULL session");
    if (!w) { delete S->slab; S->slab = nullptr; S->slab_pull_pending = false; return 0; }
    NGPD_REQUIRE(w->world >= 1 && w->world <= SLAB_MAX_WORLD && w->rank >= 0 && w->rank < w->world, "ngpd_session_set_slab: bad world / rank");
    NGPD_REQUIRE(w->send_seg_host && w->first_row_host && w->symm_base_host, "ngpd_session_set_slab: NULL table");
    NGPD_REQUIRE((w->n_send == 0 || w->send_rows) && (w->n_recv == 0 || w->recv_rows) && w->n_recv <= w->cap, "ngpd_session_set_slab: bad row lists");
    if (!S->slab) S->slab = new SlabDev();
    SlabDev& d = *S->slab;
    d.world = w->world; d.rank = w->rank; d.cap = w->cap;
    for (int q = 0; q <= w->world; ++q) d.seg[q] = w->send_seg_host[q];
    for (int q = 0; q < w->world; ++q) { d.first_row[q] = w->first_row_host[q]; d.base[q] = w->symm_base_host[q]; }
    NGPD_REQUIRE(d.seg[w->world] == w->n_send, "ngpd_session_set_slab: send_seg does not add up to n_send");
    S->slab_send_rows = w->send_rows; S->slab_recv_rows = w->recv_rows;
    S->slab_n_send = w->n_send; S->slab_n_recv = w->n_recv;
    S->slab_epoch = 0;
    S->slab_pull_pending = false;
    if (!S->slab_done) {
        NGPD_CUDA_OK(cudaMalloc(&S->slab_done, sizeof(unsigned)));
        NGPD_CUDA_OK(cudaMemsetAsync(S->slab_done, 0, sizeof(unsigned), (cudaStream_t)stream_));
    }
    return 0;
}

static float4* slab_buffer(ngpd_session_t* S, int which) { return which == 0 ? S->pos[S->cur] : (which == 1 ? S->nrm : S->fn); }

// first half of a refresh: gather + store into the peers' receive buffers + signal
static int slab_push(ngpd_session_t* S, const float4* buf, unsigned long long* epoch_out, cudaStream_t st) {
    const SlabDev& d = *S->slab;
    const unsigned long long epoch = ++S->slab_epoch;
    // (grids small enough to be resident at once: the pull kernel's blocks wait for the peers)
    const int pb = stride_blocks(std::max<int64_t>(S->slab_n_send, 1), 256, 4);
    slab_push_kernel<<<pb, 256, 0, st>>>(d, buf, S->slab_send_rows, S->slab_n_send, (int)(epoch & 1), epoch, S->slab_done);
    NGPD_CUDA_OK(cudaGetLastError());
    *epoch_out = epoch;
    S->launches += 1;
    return 0;
}
// second half: wait for every peer's rows of that round, scatter them into the halo rows
static int slab_pull(ngpd_session_t* S, float4* buf, unsigned long long epoch, cudaStream_t st) {
    const SlabDev& d = *S->slab;
    const int qb = stride_blocks(std::max<int64_t>(S->slab_n_recv, 1), 256, 4);
    slab_pull_kernel<<<qb, 256, 0, st>>>(d, buf, S->slab_recv_rows, S->slab_n_recv, (int)(epoch & 1), epoch);
    NGPD_CUDA_OK(cudaGetLastError());
    S->launches += 1;
    return 0;
}
// complete a refresh whose second half was deferred (must happen before anything reads halo positions, before the next
// cross-rank round -- a peer may only run one round ahead of what this rank has consumed -- and before positions are replaced)
static int slab_flush(ngpd_session_t* S, cudaStream_t st) {
    if (!S->slab_pull_pending) return 0;
    S->slab_pull_pending = false;
    ProfScope ps(S, st, 5);
    return slab_pull(S, S->slab_pull_dst, S->slab_pull_epoch, st);
}

// one halo refresh of buffer `which` (0 positions, 1 normals, 2 smoothed normals): every rank must call it in the same order
extern "C" __attribute__((visibility("default"))) int ngpd_session_slab_refresh(ngpd_session_t* S, int which, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(S && S->slab && which >= 0 && which <= 2, "ngpd_session_slab_refresh: no slab wiring / bad buffer");
    if (S->slab->world == 1) return 0;
    int rc = slab_flush(S, st);
    if (rc) return rc;
    ProfScope ps(S, st, 5);
    float4* buf = slab_buffer(S, which);
    unsigned long long epoch = 0;
    if ((rc = slab_push(S, buf, &epoch, st))) return rc;
    if ((rc = slab_pull(S, buf, epoch, st))) return rc;
    if (which == 0) S->sums_ready = false;
    return 0;
}

// mode 0: buffer 3 (flat-step sums, 4 doubles) <- sum over the ranks; mode 1: buffer 4 [3] (delta) <- max over the ranks
extern "C" __attribute__((visibility("default"))) int ngpd_session_slab_allreduce(ngpd_session_t* S, int mode, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(S && S->slab && (mode == 0 || mode == 1), "ngpd_session_slab_allreduce: no slab wiring / bad mode");
    const SlabDev& d = *S->slab;
    if (d.world == 1) return 0;
    int rc = slab_flush(S, st);
    if (rc) return rc;
    ProfScope ps(S, st, 5);
    const unsigned long long epoch = ++S->slab_epoch;
    slab_allreduce_kernel<<<1, 32, 0, st>>>(d, S->acc, S->cd, mode, (int)(epoch & 1), epoch);
    NGPD_CUDA_OK(cudaGetLastError());
    S->launches += 1;
    return 0;
}

// ngpd_session_step on one Morton slab: the same phases with the halo refreshes and the cross-rank scalars in between
// (smoothed normals after the first tensor pass, positions after every class that moved, Processor.py:127-138)
extern "C" __attribute__((visibility("default"))) int ngpd_session_step_slab(ngpd_session_t* S, const ngpd_step_params_t* p, void* stream_) {
    NGPD_REQUIRE(S && p && S->slab, "ngpd_session_step_slab: NULL argument / no slab wiring");
    S->launches = 0;
    int rc;
    if ((rc = ngpd_session_phase_features(S, p, 0, stream_))) return rc;
    if ((rc = ngpd_session_slab_refresh(S, 2, stream_))) return rc;
    if ((rc = ngpd_session_phase_features(S, p, 1, stream_))) return rc;
    const bool snapshot = (p->flags & NGPD_STEP_SNAPSHOT_CLASSES) != 0;
    for (int key = 0; key < 3; ++key) {
        if (p->strategy[key] == NGPD_STEP_FLAT) {
            if ((rc = ngpd_session_phase_flat_scalars(S, p, key, 0, stream_))) return rc;
            if ((rc = ngpd_session_slab_allreduce(S, 0, stream_))) return rc;
            if ((rc = ngpd_session_phase_flat_scalars(S, p, key, 1, stream_))) return rc;
            if ((rc = ngpd_session_slab_allreduce(S, 1, stream_))) return rc;
        }
        if ((rc = ngpd_session_phase_update(S, p, key, stream_))) return rc;
        // the class' new positions, before the next class reads them (snapshot mode: nobody reads them before the step ends)
        const bool moved = p->strategy[key] >= 0;
        bool later = false;                                  // does a later class still read positions in this step?
        for (int q = key + 1; q < 3; ++q) later = later || p->strategy[q] >= 0;
        if ((moved && !snapshot && later) || (snapshot && key == 2 && false))
            if ((rc = ngpd_session_slab_refresh(S, 0, stream_))) return rc;
    }
    // the step's last position refresh: pushed now, pulled when the next reader of halo positions comes along (the next step's
    // search reads owned rows only, so the exchange and the slowest peer's tail hide behind it)
    if (S->slab->world > 1) {
        cudaStream_t st = (cudaStream_t)stream_;
        ProfScope ps(S, st, 5);
        S->slab_pull_dst = S->pos[S->cur];
        if ((rc = slab_push(S, S->slab_pull_dst, &S->slab_pull_epoch, st))) return rc;
        S->slab_pull_pending = true;
        S->sums_ready = false;
    }
    return ngpd_session_phase_commit_normals(S);
}

// max over the owned rows, since the session was created (or this was last reset), of: distance of the k-th neighbour found
// + distance of the query from its own tree position.  A slab's searches equal the whole cloud's iff this stays below the halo width.
extern "C" __attribute__((visibility("default"))) int ngpd_session_halo_need(ngpd_session_t* S, int reset, float* out_host, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(S && out_host, "ngpd_session_halo_need: NULL argument");
    *out_host = 0.0f;
    if (S->halo_need) {
        NGPD_CUDA_OK(cudaMemcpyAsync(out_host, S->halo_need, sizeof(float), cudaMemcpyDeviceToHost, st));
        if (reset) NGPD_CUDA_OK(cudaMemsetAsync(S->halo_need, 0, sizeof(float), st));
        NGPD_CUDA_OK(cudaStreamSynchronize(st));
    }
    return 0;
}

// global_ids (nullable, device, int64): original index of the session -> id in the whole cloud (slabs); synchronises
extern "C" __attribute__((visibility("default"))) int ngpd_session_checksum(ngpd_session_t* S, const int64_t* global_ids, uint64_t* out11_host, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(S && out11_host, "ngpd_session_checksum: NULL argument");
    { int rc = slab_flush(S, st); if (rc) return rc; }
    unsigned long long* d = nullptr;
    NGPD_CUDA_OK(cudaMallocAsync(&d, CHECKSUM_WORDS * sizeof(unsigned long long), st));
    cudaError_t e = cudaMemsetAsync(d, 0, CHECKSUM_WORDS * sizeof(unsigned long long), st);
    if (e == cudaSuccess) {
        session_checksum_kernel<<<strided(S->n, 256), 256, 0, st>>>(S->grid->pts, S->pos[S->cur], S->nrm, S->label, S->owned, global_ids, S->n, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out11_host, d, CHECKSUM_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(d, st);
    NGPD_CUDA_OK(e);
    return 0;
}

// End-to-end entry point with HOST buffers (the path the e2e measurement times): this step's positions and normals in,
// `iterations` steps, positions / normals / labels out.  Synchronous.  The PCIe transfers (49 bytes per point) cost more
// than the iteration itself, so they are taken off the critical path where the data flow allows it:
//   in : positions first; the k-NN search needs nothing else and runs while the normals are still arriving on a second stream;
//   out: the normals of the result are final as soon as the last smoothing pass is done, so they leave on the second stream
//        while labels and the position updates are still being computed; positions and labels follow.
extern "C" __attribute__((visibility("default"))) int ngpd_session_run_host(ngpd_session_t* S, const ngpd_step_params_t* p, int iterations, const float* pos_host,
                                     const float* nrm_host, float* pos_out_host, float* nrm_out_host, uint8_t* labels_out_host,
                                     void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(S && p && pos_host && nrm_host, "ngpd_session_run_host: NULL argument");
    { int rc = slab_flush(S, st); if (rc) return rc; }
    NGPD_REQUIRE(p->k_feature >= 1 && p->k_feature <= 64 && p->k_update >= 1 && p->k_update <= p->k_feature,
                 "ngpd_session: need 1 <= k_update <= k_feature <= 64");
    size_t b3 = (size_t)S->n * 3 * sizeof(float);
    // device staging lives with the session: a stream-ordered allocation per call would hand its memory back to the
    // driver at every synchronisation and pay for mapping it again (measured: 60-340 ms per call at 10 M points)
    if (!S->stage_pos) NGPD_CUDA_OK(cudaMalloc(&S->stage_pos, b3));
    if (!S->stage_nrm) NGPD_CUDA_OK(cudaMalloc(&S->stage_nrm, b3));
    if (!S->stage_lab) NGPD_CUDA_OK(cudaMalloc(&S->stage_lab, (size_t)S->n));
    if (!S->side) {
        NGPD_CUDA_OK(cudaStreamCreateWithFlags(&S->side, cudaStreamNonBlocking));
        for (cudaEvent_t& e : S->xfer) NGPD_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const unsigned b = (unsigned)cdiv(S->n, 256);
    if (!S->inv) {
        NGPD_CUDA_OK(cudaMalloc(&S->inv, (size_t)S->n * sizeof(int32_t)));
        session_inverse_kernel<<<b, 256, 0, st>>>(S->grid->pts, S->n, S->inv);
    }
    float *dpos = S->stage_pos, *dnrm = S->stage_nrm;
    uint8_t* dlab = S->stage_lab;
    S->sums_ready = false;
    // in: positions on the caller's stream, normals behind them on the side stream
    NGPD_CUDA_OK(cudaMemcpyAsync(dpos, pos_host, b3, cudaMemcpyHostToDevice, st));
    NGPD_CUDA_OK(cudaEventRecord(S->xfer[0], st));
    NGPD_CUDA_OK(cudaStreamWaitEvent(S->side, S->xfer[0], 0));
    NGPD_CUDA_OK(cudaMemcpyAsync(dnrm, nrm_host, b3, cudaMemcpyHostToDevice, S->side));
    session_scatter_in_kernel<<<b, 256, 0, S->side>>>(S->grid->pts, nullptr, dnrm, S->n, nullptr, S->nrm);
    NGPD_CUDA_OK(cudaEventRecord(S->xfer[1], S->side));
    session_scatter_in_kernel<<<b, 256, 0, st>>>(S->grid->pts, dpos, nullptr, S->n, S->pos[S->cur], nullptr);
    NGPD_CUDA_OK(cudaGetLastError());
    bool nrm_left = false;
    int rc = 0;
    for (int i = 0; i < iterations && !rc; ++i) {
        S->launches = 0;
        rc = features_knn(S, p, st);
        if (i == 0) NGPD_CUDA_OK(cudaStreamWaitEvent(st, S->xfer[1], 0));
        if (!rc) rc = features_smooth(S, p, st);
        if (!rc && i == iterations - 1 && nrm_out_host) {
            // f_n is what graph.n will be after this iteration (Processor.py:139): send it home now
            NGPD_CUDA_OK(cudaEventRecord(S->xfer[2], st));
            NGPD_CUDA_OK(cudaStreamWaitEvent(S->side, S->xfer[2], 0));
            session_scatter_out_kernel<<<b, 256, 0, S->side>>>(S->inv, nullptr, S->fn, nullptr, S->n, nullptr, dnrm, nullptr);
            NGPD_CUDA_OK(cudaMemcpyAsync(nrm_out_host, dnrm, b3, cudaMemcpyDeviceToHost, S->side));
            nrm_left = true;
        }
        if (!rc) rc = step_tail(S, p, stream_);
    }
    if (iterations <= 0) NGPD_CUDA_OK(cudaStreamWaitEvent(st, S->xfer[1], 0));
    if (rc) { cudaStreamSynchronize(S->side); return rc; }
    session_scatter_out_kernel<<<b, 256, 0, st>>>(S->inv, S->pos[S->cur], nrm_left ? nullptr : S->nrm, S->label, S->n,
                                                   pos_out_host ? dpos : nullptr, (nrm_out_host && !nrm_left) ? dnrm : nullptr,
                                                   labels_out_host ? dlab : nullptr);
    NGPD_CUDA_OK(cudaGetLastError());
    if (pos_out_host) NGPD_CUDA_OK(cudaMemcpyAsync(pos_out_host, dpos, b3, cudaMemcpyDeviceToHost, st));
    if (nrm_out_host && !nrm_left) NGPD_CUDA_OK(cudaMemcpyAsync(nrm_out_host, dnrm, b3, cudaMemcpyDeviceToHost, st));
    if (labels_out_host) NGPD_CUDA_OK(cudaMemcpyAsync(labels_out_host, dlab, (size_t)S->n, cudaMemcpyDeviceToHost, st));
    NGPD_CUDA_OK(cudaStreamSynchronize(S->side));
    NGPD_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_denoise_host(const float* tree_pos_host, const float* pos_host, const float* nrm_host, int64_t n,
                                 const ngpd_step_params_t* p, int iterations, float* pos_out_host, float* nrm_out_host,
                                 uint8_t* labels_out_host) {
    NGPD_REQUIRE(tree_pos_host && pos_host && nrm_host && p && n > 0, "ngpd_denoise_host: bad argument");
    float* dtree = nullptr;
    size_t b3 = (size_t)n * 3 * sizeof(float);
    NGPD_CUDA_OK(cudaMalloc(&dtree, b3));
    NGPD_CUDA_OK(cudaMemcpy(dtree, tree_pos_host, b3, cudaMemcpyHostToDevice));
    ngpd_session_t* S = nullptr;
    int rc = ngpd_session_create(dtree, n, p->k_feature, nullptr, &S);
    cudaFree(dtree);
    if (rc) return rc;
    rc = ngpd_session_run_host(S, p, iterations, pos_host, nrm_host, pos_out_host, nrm_out_host, labels_out_host, nullptr);
    ngpd_session_destroy(S);
    return rc;
}
