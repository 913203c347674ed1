// Fast path of the k-NN search (kernel group 2): warp-lockstep candidate streaming.
//
// The 32 queries of a warp are neighbours in tree order, so their 3x3x3 cell blocks overlap heavily.  Instead of
// letting every lane walk its own cells (different trip counts, different insert times: ncu showed 9.7 of 32 lanes
// active), the warp walks the UNION box of its lanes' blocks row by row; every lane sees the same candidate at the
// same time (one broadcast 16-byte load per candidate), computes its own fp32 distance and appends survivors to a
// small per-lane queue in shared memory.  Queues are drained together into per-lane sorted lists, so the expensive
// insertion chain runs with most lanes busy.
//
// Exactness is kept without fp64 in the inner loop:
//   * the k survivors are re-evaluated in fp64 ((dx^2+dy^2)+dz^2, SciPy's order) and sorted by (distance, original
//     index) before they are written;
//   * a lane remembers the smallest fp32 distance it ever rejected or evicted; if that is not clearly (2e-6 relative,
//     > 5x the fp32 evaluation error) above its k-th kept distance, or if the k-th distance is not inside the lane's
//     own 3x3x3 block, the query is put on a fix-up list and redone by the exact shell search (knn.cuh).
// A query answered here therefore has exactly the result the exact kernel would give.
#pragma once
#include "knn.cuh"

namespace ngpd {

constexpr int KF_THREADS = 128;
constexpr int KF_BUF = 12;        // queue slots per lane
constexpr int KF_GROUP = 4;       // candidates per inner step (4 loads in flight)
constexpr unsigned FULL = 0xffffffffu;

template <int K>
struct Near {
    float d[K];
    int id[K];
    // `bound`: no neighbour of the final answer is farther than this (INFINITY when nothing is known)
    __device__ __forceinline__ void init(float bound = INFINITY) {
#pragma unroll
        for (int a = 0; a < K; ++a) { d[a] = bound; id[a] = -1; }
    }
};

struct KfShared {
    float d[KF_BUF][KF_THREADS];
    int j[KF_BUF][KF_THREADS];
};

template <int K>
__device__ __forceinline__ void near_insert(Near<K>& t, float d2, int j, float& rej) {
    rej = fminf(rej, t.d[K - 1]);          // the evicted tail
    bool prev = true;
#pragma unroll
    for (int a = K - 1; a > 0; --a) {
        bool sh = d2 < t.d[a - 1];
        float nd = sh ? t.d[a - 1] : (prev ? d2 : t.d[a]);
        int ni = sh ? t.id[a - 1] : (prev ? j : t.id[a]);
        t.d[a] = nd; t.id[a] = ni;
        prev = sh;
    }
    if (prev) { t.d[0] = d2; t.id[0] = j; }
}

template <int K>
__device__ __forceinline__ void near_flush(Near<K>& t, KfShared& sm, int& cnt, float& rej) {
    const int tid = threadIdx.x;
#pragma unroll 1
    for (int b = 0; b < KF_BUF; ++b) {
        bool has = b < cnt;
        if (!__any_sync(FULL, has)) break;
        if (has) {
            float d2 = sm.d[b][tid];
            if (d2 < t.d[K - 1]) near_insert<K>(t, d2, sm.j[b][tid], rej);
            else rej = fminf(rej, d2);
        }
    }
    cnt = 0;
}

template <int K, bool SELF, bool TAIL>
__device__ __forceinline__ void near_group(Near<K>& t, KfShared& sm, int& cnt, float& rej, const float4* __restrict__ pts,
                                           int j0, int e, float qx, float qy, float qz, bool active, int self_orig) {
    const int tid = threadIdx.x;
    float4 p[KF_GROUP];
#pragma unroll
    for (int u = 0; u < KF_GROUP; ++u) p[u] = __ldg(pts + (TAIL ? min(j0 + u, e - 1) : j0 + u));
    const float worst = t.d[K - 1];
#pragma unroll
    for (int u = 0; u < KF_GROUP; ++u) {
        float dx = p[u].x - qx, dy = p[u].y - qy, dz = p[u].z - qz;
        float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        bool valid = active;
        if (TAIL) valid = valid && (j0 + u < e);
        if (SELF) valid = valid && (__float_as_int(p[u].w) != self_orig);
        bool take = valid && d2 < worst;
        if (take) { sm.d[cnt][tid] = d2; sm.j[cnt][tid] = j0 + u; ++cnt; }
        if (valid && !take) rej = fminf(rej, d2);
    }
}

template <int K, bool SELF>
__device__ __forceinline__ void near_stream(Near<K>& t, KfShared& sm, int& cnt, float& rej, const float4* __restrict__ pts,
                                            int s, int e, float qx, float qy, float qz, bool active, int self_orig) {
    int j0 = s;
    for (; j0 + KF_GROUP <= e; j0 += KF_GROUP) {
        if (__any_sync(FULL, cnt > KF_BUF - KF_GROUP)) near_flush<K>(t, sm, cnt, rej);
        near_group<K, SELF, false>(t, sm, cnt, rej, pts, j0, e, qx, qy, qz, active, self_orig);
    }
    if (j0 < e) {
        if (__any_sync(FULL, cnt > KF_BUF - KF_GROUP)) near_flush<K>(t, sm, cnt, rej);
        near_group<K, SELF, true>(t, sm, cnt, rej, pts, j0, e, qx, qy, qz, active, self_orig);
    }
}

// Lockstep search of the lanes' 3x3x3 blocks.  Returns true when this lane's list is final up to the exact re-sort.
// Rows are visited in two passes: first the rows that hold the lanes' own cells (their candidates are the nearest, so
// the lists fill with good values and the accept threshold is tight before the surrounding shell streams by).
template <int K, bool SELF>
__device__ __forceinline__ bool knn_lockstep(Near<K>& t, KfShared& sm, const GridView& g, float qx, float qy, float qz,
                                             bool active, int self_orig) {
    const double rx = (double)qx - g.ox, ry = (double)qy - g.oy, rz = (double)qz - g.oz;
    const int cx = min(max((int)floor(rx * g.inv_h), 0), g.nx - 1);
    const int cy = min(max((int)floor(ry * g.inv_h), 0), g.ny - 1);
    const int cz = min(max((int)floor(rz * g.inv_h), 0), g.nz - 1);
    const int BIG = 1 << 29;
    const int my0 = __reduce_min_sync(FULL, active ? cy : BIG), my1 = __reduce_max_sync(FULL, active ? cy : -BIG);
    const int mz0 = __reduce_min_sync(FULL, active ? cz : BIG), mz1 = __reduce_max_sync(FULL, active ? cz : -BIG);
    if (my1 < my0) return false;                                      // no active lane in this warp
    const int ly = max(my0 - 1, 0), hy = min(my1 + 1, g.ny - 1), lz = max(mz0 - 1, 0), hz = min(mz1 + 1, g.nz - 1);
    {
        int mx0 = __reduce_min_sync(FULL, active ? cx : BIG), mx1 = __reduce_max_sync(FULL, active ? cx : -BIG);
        if ((hy - ly + 1) * (hz - lz + 1) > 576 || (mx1 - mx0 + 3) > 48) return false;   // lanes far apart (curve jump): exact path
    }
    int cnt = 0;
    float rej = INFINITY;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        for (int z = lz; z <= hz; ++z) {
            for (int y = ly; y <= hy; ++y) {
                const bool core = (y >= my0 && y <= my1 && z >= mz0 && z <= mz1);
                if (core != (pass == 0)) continue;
                bool rel = active && abs(z - cz) <= 1 && abs(y - cy) <= 1;
                if (!__any_sync(FULL, rel)) continue;
                int x0 = max(__reduce_min_sync(FULL, rel ? cx - 1 : BIG), 0);
                int x1 = min(__reduce_max_sync(FULL, rel ? cx + 1 : -BIG), g.nx - 1);
                const int64_t trow = ((int64_t)(z >> 3) * g.tby + (y >> 3)) * g.tbx;
                const int lrow = ((z & 7) << 6) | ((y & 7) << 3);
                while (x0 <= x1) {
                    int xe = min(x1, (x0 | 7));
                    int b = __ldg(g.top + trow + (x0 >> 3));
                    if (b >= 0) {
                        const int* f = g.fine + (int64_t)b * 513 + lrow;
                        int s = __ldg(f + (x0 & 7)), e = __ldg(f + (xe & 7) + 1);
                        if (e > s) near_stream<K, SELF>(t, sm, cnt, rej, g.pts, s, e, qx, qy, qz, active, self_orig);
                    }
                    x0 = xe + 1;
                }
            }
        }
    }
    near_flush<K>(t, sm, cnt, rej);
    if (!active) return false;
    // the lane saw (at least) every point of its own 3x3x3 block: final iff the list is full, its k-th distance lies
    // inside that block, and no rejected candidate is within the fp32 evaluation error of it
    const float worst = t.d[K - 1];
    double reach = DBL_MAX;
    if (cx - 1 > 0) reach = fmin(reach, rx - (double)(cx - 1) * g.h);
    if (cx + 1 < g.nx - 1) reach = fmin(reach, (double)(cx + 2) * g.h - rx);
    if (cy - 1 > 0) reach = fmin(reach, ry - (double)(cy - 1) * g.h);
    if (cy + 1 < g.ny - 1) reach = fmin(reach, (double)(cy + 2) * g.h - ry);
    if (cz - 1 > 0) reach = fmin(reach, rz - (double)(cz - 1) * g.h);
    if (cz + 1 < g.nz - 1) reach = fmin(reach, (double)(cz + 2) * g.h - rz);
    const double wmax = (double)worst * (1.0 + 2e-6);
    const double rr = reach - g.h * 1e-9;
    bool inside = (reach == DBL_MAX) ? true : (rr > 0.0 && wmax < rr * rr);
    bool clear = (double)rej > wmax;
    return t.id[K - 1] >= 0 && inside && clear;
}

// exact re-evaluation and ordering of the k kept candidates: (fp64 distance, original index)
template <int K>
__device__ __forceinline__ void near_finalize(Near<K>& t, const float4* __restrict__ pts, float qx, float qy, float qz, double ex[K]) {
    int orig[K];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        int j = t.id[a];
        float4 p = __ldg(pts + max(j, 0));
        double dx = (double)qx - (double)p.x, dy = (double)qy - (double)p.y, dz = (double)qz - (double)p.z;
        ex[a] = j >= 0 ? __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) : DBL_MAX;
        orig[a] = j >= 0 ? __float_as_int(p.w) : 0x7fffffff;
    }
    // nearly sorted already (fp32 order): odd-even transposition passes until nothing moves anywhere in the warp
    bool moved = true;
    while (__any_sync(FULL, moved)) {
        moved = false;
#pragma unroll
        for (int par = 0; par < 2; ++par) {
#pragma unroll
            for (int a = par; a + 1 < K; a += 2) {
                bool sw = ex[a + 1] < ex[a] || (ex[a + 1] == ex[a] && orig[a + 1] < orig[a]);
                if (sw) {
                    double td = ex[a]; ex[a] = ex[a + 1]; ex[a + 1] = td;
                    int ti = orig[a]; orig[a] = orig[a + 1]; orig[a + 1] = ti;
                    int tj = t.id[a]; t.id[a] = t.id[a + 1]; t.id[a + 1] = tj;
                    moved = true;
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < K; ++a) t.d[a] = __int_as_float(orig[a]);   // hand the original indices back in d[]
}

// warp-aggregated append of the lanes with `flag` to a global list
__device__ __forceinline__ void fix_append(bool flag, int value, int32_t* __restrict__ list, int32_t* __restrict__ count) {
    unsigned m = __ballot_sync(FULL, flag);
    if (!m) return;
    int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(FULL, base, __ffs(m) - 1);
    if (flag) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

}  // namespace ngpd
