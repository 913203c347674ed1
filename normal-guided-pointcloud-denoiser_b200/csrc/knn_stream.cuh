// Fast path of the k-NN search (kernel group 2): per-lane candidate streams + sorting-network selection.
//
// One query per thread.  A query only ever needs the points of the 3x3x3 block of cells around its own cell (checked
// afterwards; the rare query whose k-th neighbour lies outside goes to the exact shell search of knn.cuh), and in the
// sorted point array that block is at most 18 contiguous ranges (9 cell rows along x, each split at most once by a
// brick boundary).  The kernel therefore runs in three phases:
//   1. ranges   every lane looks its ranges up in the two-level grid and parks them in shared memory (core row first);
//   2. stream   a flat, warp-uniform loop: each lane walks ITS OWN ranges four candidates at a time.  Lanes are
//               neighbours in tree order, so the lanes of one cell walk the same addresses (one L1 wavefront per
//               distinct cell, 5-6 per warp) while no lane evaluates candidates of somebody else's block
//               (the previous warp-union walk evaluated ~300 candidates per lane for ~80 useful ones).
//               A candidate is turned into ONE 32-bit key: 21 bits of squared distance in fixed point (units of
//               h^2 / 65536) above 11 bits of (range slot, offset in range).  Keys below the lane's current threshold
//               are appended to a 16-entry batch in shared memory; nothing is inserted one by one.
//   3. select   whenever some lane's batch is nearly full, every lane sorts its batch with Batcher's odd-even merge
//               network (63 compare-exchanges, 2 instructions each since a key carries its own id) and merges it
//               into its sorted top-k with a bitonic half-cleaner + merge.  All lanes do the same thing at the same
//               time: no divergence, no per-candidate insertion chains (ncu on the previous insertion kernel:
//               10 of 32 lanes active in 34 % of its instructions).
// Exactness is kept without fp64 in the inner loop:
//   * keys order candidates by distance to 1.5e-5 h^2; a lane remembers the smallest key it ever dropped, and unless
//     that is clearly (3 units: key truncation + fp32 evaluation error) above its k-th kept key and the k-th distance
//     lies inside the lane's own 3x3x3 block, the query is put on a fix-up list and redone by the next tier;
//   * the k survivors are re-evaluated in fp64 ((dx^2+dy^2)+dz^2, SciPy's order) and sorted by (distance, original
//     index) before they are written.
// A query answered here therefore has exactly the rows the exact kernel would give.
//
// The search is a template on the block radius R.  R = 1 (3x3x3 cells, 128 threads) answers ~97.5 % of the queries of
// a surface cloud; its failures are compacted into a list and retried with R = 2 (5x5x5 cells, 64 threads per block
// because of the 50 range slots per lane), and only what fails there too (isolated points, k-th neighbour farther than
// two cells) reaches the one-query-per-thread exact shell search.  ncu on the first version, whose failures all
// went to the exact search: that kernel took half as long as the fast kernel for 2.4 % of the queries.
#pragma once
#include "knn.cuh"
#include "sortnet.cuh"

namespace ngpd {

constexpr int KS_BATCH = 16;          // keys per lane between two selection rounds
#ifndef NGPD_KS_GROUP_R1
#define NGPD_KS_GROUP_R1 4
#endif
#ifndef NGPD_KS_GROUP_R2
#define NGPD_KS_GROUP_R2 8
#endif
constexpr unsigned KS_NONE = 0xffffffffu;
constexpr unsigned FULL = 0xffffffffu;

// Key of a candidate: [ squared distance in units of h^2 * 2^-16 (R = 1) | range slot | offset in the slot ].
// The distance field is the mantissa of  y = d2 / h^2 + 32  (y in [32,64): fixed exponent, so the mantissa IS the
// fixed-point value; every candidate of a (2R+1)^3 block has d2 < 27 h^2 < 32 h^2), which costs one FFMA and two
// logic instructions.  Keys of different candidates are different, unsigned compare orders them by distance to
// 1.5e-5 h^2, and min/max on them moves distance and id together.
template <int R>
struct KsCfg {
    static constexpr int ROWS = (2 * R + 1) * (2 * R + 1);
    // A row of 2R+1 <= 8 cells crosses at most one brick boundary: 2 * ROWS pieces, one range slot per piece unless the piece is
    // longer than a slot.  R = 1: 64-point slots, 18 of them.  R = 2: 128-point slots and 64 of them instead of 50 -- where two
    // sheets of a surface meet (ncu round 2: the rows around the torus / cube-face intersection of the bench cloud) or with
    // k = 64 on cells sized for k <= 32, rows of 5 cells hold more than 64 points and 50 slots overflowed, which sent those
    // queries to the one-thread-per-query exact search (0.8 ms of latency per iteration for a few dozen rows).
    // R = 2 also PRUNES: every slot carries a lower bound of its candidates' keys (distance from the query to the slot's cells), the
    // rows of the inner 3x3 ring come first, and a slot whose bound is not below the lane's current threshold is skipped --
    // a 5x5x5 block holds 750-900 candidates of which the outer shell (98 of 125 cells) almost never contributes, and the
    // kernel's latency (one wave of blocks, 222 us whatever the cloud size) is the per-lane candidate count.
    static constexpr bool PRUNE = R == 2;
    static constexpr int OFFBITS = R == 1 ? 6 : 7;
    static constexpr int SLOTS = R == 1 ? 2 * ROWS : 56;
    static constexpr int SLOTBITS = R == 1 ? 5 : 6;
    static constexpr int IDBITS = SLOTBITS + OFFBITS;
    static constexpr unsigned IDMASK = (1u << IDBITS) - 1u;
    static constexpr int THREADS = R == 1 ? 128 : 64;
    // candidates per inner step = independent loads in flight per lane.  The 5^3 tier runs one wave of blocks at 3 % occupancy: its
    // duration is (steps per lane) x (latency of a divergent 16-byte gather), so eight loads per step instead of four nearly halve it.
    static constexpr int GROUP = R == 1 ? NGPD_KS_GROUP_R1 : NGPD_KS_GROUP_R2;
    static_assert(GROUP <= KS_PAD && 2 * GROUP <= KS_BATCH, "overrun past the padding / a step must fit the batch");
    static constexpr double UNIT = 1.0 / (double)(1 << (18 - (IDBITS - 9)));   // of the distance field, in h^2
};

template <int R>
struct KsShared {
    int2 rng[KsCfg<R>::SLOTS][KsCfg<R>::THREADS];   // (first point, one past the last) of each slot
    unsigned batch[KS_BATCH][KsCfg<R>::THREADS];
    unsigned mink[KsCfg<R>::PRUNE ? KsCfg<R>::SLOTS : 1][KsCfg<R>::THREADS];   // PRUNE: lower bound of the keys in the slot
};

// visiting order of the (dy, dz) rows of the block: the lane's own row first; R = 2: then the inner ring, then the outer ring
template <int R>
__device__ __forceinline__ int ks_row_order(int tt) {
    constexpr int W = 2 * R + 1, MID = (W * W - 1) / 2;
    if (R == 2) {
        constexpr int order[25] = {12, 6, 7, 8, 11, 13, 16, 17, 18, 0, 1, 2, 3, 4, 5, 9, 10, 14, 15, 19, 20, 21, 22, 23, 24};
        return order[tt];
    }
    return tt == 0 ? MID : (tt <= MID ? tt - 1 : tt);
}

// Near-ties: when the K-th and (K+1)-th candidates are closer together than the keys can tell, the first KF = K + 4
// keys are re-evaluated exactly and the row is taken from those; only a list without spare entries (KT == K) has to
// hand such a query on.
__host__ __device__ constexpr int ks_kf(int K, int KT) { return KT >= K + 4 ? K + 4 : K; }

template <int K>
struct KsTop {
    unsigned key[K];  // ascending
    int id[K];        // filled by the decode step
};

template <int R>
__device__ __forceinline__ unsigned ks_dist_field(float d2, float inv_h2) {
    float y = fminf(fmaf(d2, inv_h2, 32.0f), 63.99999f);      // NaN and far-away candidates saturate
    return (__float_as_uint(y) << 9) & ~KsCfg<R>::IDMASK;
}

// one selection round: the lane's batch (cnt keys in shared memory) is merged into its top-k
template <int K, int R>
__device__ __forceinline__ void ks_round(KsTop<K>& t, KsShared<R>& sm, int cnt, unsigned& rej) {
    const int tid = threadIdx.x;
    unsigned b[KS_BATCH];
#pragma unroll
    for (int i = 0; i < KS_BATCH; ++i) { unsigned v = sm.batch[i][tid]; b[i] = i < cnt ? v : KS_NONE; }
    ks_sort<KS_BATCH>(b);
    constexpr int M = K < KS_BATCH ? K : KS_BATCH;
    // top (ascending) against the batch (descending): the minima are the k smallest of both and form a bitonic sequence
#pragma unroll
    for (int i = 0; i < M; ++i) {
        unsigned lo = min(t.key[K - 1 - i], b[i]), hi = max(t.key[K - 1 - i], b[i]);
        t.key[K - 1 - i] = lo;
        rej = min(rej, hi);
    }
    ks_bitonic_merge<K>(t.key);
    if (K < KS_BATCH) rej = min(rej, b[K < KS_BATCH ? K : 0]);
}

// Search of the (2R+1)^3 block of cells around the lane's own cell.  Returns true when t.id[] holds the lane's final
// neighbours (tree positions, still in key order); false = hand the query to the next tier.
// KT >= K keys are kept.  With KT > K the extra KT-K entries are spare candidates for the re-ranking tier (ks_rerank):
// *rlim_out receives a radius (rounded down) such that EVERY tree point closer than that to the query is among the KT
// kept ones -- the smaller of the block's reach and the distance of the nearest candidate that was not kept.
template <int K, int KT, int R, bool SELF>
__device__ __forceinline__ bool knn_stream(KsTop<KT>& t, KsShared<R>& sm, const GridView& g, float qx, float qy, float qz,
                                           bool active, int self_orig, float bound, float* rlim_out = nullptr) {
    static_assert(KT >= K, "the kept list cannot be shorter than the row");
    using C = KsCfg<R>;
    const int tid = threadIdx.x;
    const double rx = (double)qx - g.ox, ry = (double)qy - g.oy, rz = (double)qz - g.oz;
    const int cx = min(max((int)floor(rx * g.inv_h), 0), g.nx - 1);
    const int cy = min(max((int)floor(ry * g.inv_h), 0), g.ny - 1);
    const int cz = min(max((int)floor(rz * g.inv_h), 0), g.nz - 1);
    const float inv_h2 = (float)(g.inv_h * g.inv_h);

    // ---- phase 1: the block's point ranges, the lane's own row first
    int nr = 0;
    bool over = false;
    if (active) {
        const int xa = max(cx - R, 0), xb = min(cx + R, g.nx - 1);
        constexpr int W = 2 * R + 1, MID = (C::ROWS - 1) / 2;
#pragma unroll(R == 1 ? 9 : 5)
        for (int tt = 0; tt < C::ROWS; ++tt) {
            const int o = ks_row_order<R>(tt);                              // (dy 0, dz 0) first
            const int y = cy + (o % W) - R, z = cz + (o / W) - R;
            if ((unsigned)y >= (unsigned)g.ny || (unsigned)z >= (unsigned)g.nz) continue;
            double gyz2 = 0.0;                                              // PRUNE: squared distance from the query to the row of cells
            if (C::PRUNE) {
                // (fp64 like the binning itself: in fp32 the offsets would carry 2e-5 h of error, more than the slack below)
                const double fy = ry - (double)y * g.h, fz = rz - (double)z * g.h;                  // offsets from the cell's low faces
                const double gy = fy < 0.0 ? -fy : (fy > g.h ? fy - g.h : 0.0), gz = fz < 0.0 ? -fz : (fz > g.h ? fz - g.h : 0.0);
                gyz2 = gy * gy + gz * gz;
            }
            const int64_t trow = ((int64_t)(z >> 3) * g.tby + (y >> 3)) * g.tbx;
            const int lrow = ((z & 7) << 6) | ((y & 7) << 3);
            int x0 = xa;
#pragma unroll
            for (int piece = 0; piece < 2; ++piece) {
                if (x0 > xb) break;
                const int xe = min(xb, x0 | 7);
                const int b = __ldg(g.top + trow + (x0 >> 3));
                if (b >= 0) {
                    const int* f = g.fine + (int64_t)b * 513 + lrow;
                    int s = __ldg(f + (x0 & 7));
                    const int e = __ldg(f + (xe & 7) + 1);
                    unsigned lowkey = 0;
                    if (C::PRUNE) {
                        // every point of cells x0 .. xe of this row is at least this far from the query; 1e-5 of relative slack covers the
                        // fp32 evaluation of the candidates' own distances (a few 1e-7), the absolute term the binning (1e-13 h);
                        // the key of a candidate is monotone in its distance, so no candidate of the slot has a smaller key
                        const double fx0 = rx - (double)x0 * g.h, fx1 = rx - (double)(xe + 1) * g.h;
                        const double gx = fx0 < 0.0 ? -fx0 : (fx1 > 0.0 ? fx1 : 0.0);
                        const double lb = (gx * gx + gyz2) * (1.0 - 1e-5) - 1e-8 * g.h * g.h;
                        lowkey = ks_dist_field<R>(lb > 0.0 ? (float)lb * 0.999999f : 0.0f, inv_h2);
                    }
                    while (s < e) {                                    // one slot per 2^OFFBITS points
                        const int ee = min(e, s + (1 << C::OFFBITS));
                        if (nr < C::SLOTS) { sm.rng[nr][tid] = make_int2(s, ee); if (C::PRUNE) sm.mink[nr][tid] = lowkey; ++nr; } else over = true;
                        s = ee;
                    }
                }
                x0 = xe + 1;
            }
        }
    }

    // ---- phase 2 + 3: stream the ranges four candidates at a time, select in rounds
#pragma unroll
    for (int a = 0; a < KT; ++a) t.key[a] = KS_NONE;
    // the temporal bound is a bound on the K-th distance: it cannot prune a longer list
    const unsigned bkey = (KT == K && bound < INFINITY) ? (ks_dist_field<R>(bound, inv_h2) | C::IDMASK) : KS_NONE;
    unsigned tau = bkey, rej = KS_NONE;
    const float4* pp = g.pts;
    int rem = 0, r = 0;
    unsigned idv = 0;
    unsigned* const b0 = &sm.batch[0][tid];
    unsigned* bp = b0;
    if (nr > 0) { int2 q = sm.rng[0][tid]; pp = g.pts + q.x; rem = q.y - q.x; r = 1; }
    while (__any_sync(FULL, rem > 0)) {
        float4 p[C::GROUP];
#pragma unroll
        for (int u = 0; u < C::GROUP; ++u) p[u] = __ldg(pp + u);       // may run past the range: the array is padded, `rem` masks
#pragma unroll
        for (int u = 0; u < C::GROUP; ++u) {
            float dx = p[u].x - qx, dy = p[u].y - qy, dz = p[u].z - qz;
            float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            bool valid = u < rem;
            if (SELF) valid = valid && (__float_as_int(p[u].w) != self_orig);
            const unsigned key = valid ? (ks_dist_field<R>(d2, inv_h2) | (idv + u)) : KS_NONE;
            const bool take = key < tau;
            *bp = key;                                                 // always stored, kept only if the pointer moves on
            bp += take ? C::THREADS : 0;
            rej = min(rej, take ? KS_NONE : key);
        }
        rem -= C::GROUP;
        if (rem > 0) { pp += C::GROUP; idv += C::GROUP; }
        else {
            if (C::PRUNE) {
                // slots none of whose candidates could be taken: they count as dropped at their lower bound
                while (r < nr && sm.mink[r][tid] >= tau) { rej = min(rej, sm.mink[r][tid]); ++r; }
            }
            if (r < nr) { int2 q = sm.rng[r][tid]; pp = g.pts + q.x; rem = q.y - q.x; idv = (unsigned)r << C::OFFBITS; ++r; }
            else { pp = g.pts; rem = 0; }                               // done: keep the speculative loads in bounds
        }
        if (__any_sync(FULL, bp > b0 + (KS_BATCH - C::GROUP) * C::THREADS)) {
            ks_round<KT, R>(t, sm, (int)(bp - b0) / C::THREADS, rej);
            bp = b0;
            tau = min(bkey, t.key[KT - 1]);
        }
    }
    if (__any_sync(FULL, bp > b0)) ks_round<KT, R>(t, sm, (int)(bp - b0) / C::THREADS, rej);

    // ---- decode the survivors' ids
#pragma unroll
    for (int a = 0; a < KT; ++a) {
        const unsigned bits = t.key[a];
        int slot = min((int)((bits & C::IDMASK) >> C::OFFBITS), C::SLOTS - 1);
        t.id[a] = bits != KS_NONE ? sm.rng[slot][tid].x + (int)(bits & ((1u << C::OFFBITS) - 1u)) : -1;
    }
    if (rlim_out) *rlim_out = 0.0f;
    if (!active || over) return false;
    // the lane saw every point of its block: final iff the list is full, its k-th distance lies inside the block, and
    // no dropped candidate is within the key resolution (+ the fp32 evaluation error) of it
    const unsigned worst = t.key[K - 1] >> C::IDBITS;
    constexpr int KF = ks_kf(K, KT);
    const unsigned rejv = (KF < KT ? t.key[KF < KT ? KF : 0] : rej) >> C::IDBITS;   // nearest candidate that is not re-evaluated
    double reach = DBL_MAX;
    if (cx - R > 0) reach = fmin(reach, rx - (double)(cx - R) * g.h);
    if (cx + R < g.nx - 1) reach = fmin(reach, (double)(cx + R + 1) * g.h - rx);
    if (cy - R > 0) reach = fmin(reach, ry - (double)(cy - R) * g.h);
    if (cy + R < g.ny - 1) reach = fmin(reach, (double)(cy + R + 1) * g.h - ry);
    if (cz - R > 0) reach = fmin(reach, rz - (double)(cz - R) * g.h);
    if (cz + R < g.nz - 1) reach = fmin(reach, (double)(cz + R + 1) * g.h - rz);
    const double wmax = (double)(worst + 3u) * C::UNIT * (g.h * g.h);
    const double rr = reach - g.h * 1e-9;
    const bool inside = (reach == DBL_MAX) ? true : (rr > 0.0 && wmax < rr * rr);
    const bool clear = rejv > worst + 3u;
    if (rlim_out) {
        // candidates of the block that were not kept have a key >= rej, i.e. lie at least sqrt((rej - 3) units) away;
        // points outside the block lie at least `reach` away
        double lim = reach == DBL_MAX ? 3.0e38 : fmax(rr, 0.0);
        if (rej != KS_NONE) {
            const unsigned rj = rej >> C::IDBITS;
            lim = fmin(lim, sqrt((double)(rj > 3u ? rj - 3u : 0u) * C::UNIT) * g.h);
        }
        *rlim_out = __double2float_rd(lim * (1.0 - 1e-7));
    }
    return t.key[K - 1] != KS_NONE && inside && clear;
}

// exact re-evaluation and ordering of the k kept candidates: (fp64 distance, original index).  On return t.id[] are
// tree positions in final order and ex[] the fp64 squared distances.  (Original indices only matter between candidates
// at exactly the same distance; they are fetched on demand.)
template <int K, int KT>
__device__ __forceinline__ void ks_finalize(KsTop<KT>& t, const float4* __restrict__ pts, float qx, float qy, float qz, double (&ex)[K]) {
    const double dqx = (double)qx, dqy = (double)qy, dqz = (double)qz;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        int j = t.id[a];
        float4 p = __ldg(pts + max(j, 0));
        double dx = dqx - (double)p.x, dy = dqy - (double)p.y, dz = dqz - (double)p.z;
        ex[a] = j >= 0 ? __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) : DBL_MAX;
    }
    // nearly sorted already (key order): odd-even transposition passes until nothing moves anywhere in the warp
    bool moved = true;
    while (__any_sync(FULL, moved)) {
        moved = false;
#pragma unroll
        for (int par = 0; par < 2; ++par) {
#pragma unroll
            for (int a = par; a + 1 < K; a += 2) {
                bool sw = ex[a + 1] < ex[a];
                if (ex[a + 1] == ex[a] && t.id[a] >= 0 && t.id[a + 1] >= 0)
                    sw = __float_as_int(__ldg(&pts[t.id[a + 1]].w)) < __float_as_int(__ldg(&pts[t.id[a]].w));
                if (sw) {
                    double td = ex[a]; ex[a] = ex[a + 1]; ex[a + 1] = td;
                    int tj = t.id[a]; t.id[a] = t.id[a + 1]; t.id[a + 1] = tj;
                    moved = true;
                }
            }
        }
    }
}

// ---- tier 0: re-ranking ---------------------------------------------------------------------------------------------
// The index is frozen, only the queries move.  A search (any tier) leaves behind, per query: the position it was asked
// from (anchor), KT = 2K candidates and a radius rlim such that every tree point closer than rlim to the anchor is among
// the candidates.  When the query has moved by delta since, a tree point that is not a candidate is at least
// rlim - delta away from it; so if the K-th nearest CANDIDATE is closer than that, the K nearest candidates ARE the K
// nearest tree points and no grid walk is needed: 2K gathers, two sorting networks and a merge.  Same keys, same
// exact fp64 ordering of the survivors and same "clear of the first rejected key" rule as the streaming search, so an
// answer given here is bit-identical to the full search's.  Returns false = ask the full search (and re-anchor).
// The lane's candidate row is read through `ids(slot)`: shared memory in the session kernel (KsCandTile below), which also
// serves the slot -> id look-ups after the sort; a plain row in global memory costs one L1 tag look-up per LANE for
// every such access (each lane's 128-byte row is a cache line of its own).
struct KsRowGlobal {
    const int32_t* __restrict__ row;
    __device__ __forceinline__ int operator()(int a) const { return __ldg(row + a); }
    __device__ __forceinline__ int4 chunk(int c) const { return __ldg(reinterpret_cast<const int4*>(row) + c); }
};

// Candidate rows of the 128 consecutive queries of a block, staged through shared memory: the block's rows are one
// contiguous piece of the candidate array, so the loads are fully coalesced 16-byte vectors (4 cache lines per warp
// instruction instead of 32), and the transposed tile [slot][query] (+ padding chosen per row length) is free of bank
// conflicts both when it is filled and when a lane reads its own column.
template <int KT>
struct KsCandTile {
    static constexpr int THREADS = 128;
    static constexpr int STRIDE = KT == 16 ? 130 : 129;
    int v[KT][STRIDE];
    __device__ __forceinline__ void fill(const int32_t* __restrict__ cand, int64_t row0, int64_t n) {
        constexpr int CH = KT / 4;                               // 16-byte chunks per row
        const int4* src = reinterpret_cast<const int4*>(cand + row0 * KT);
        const int64_t avail = (n - row0) * CH;                   // chunks that exist (tail block)
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int c = (int)threadIdx.x + THREADS * i;
            const int4 q = c < avail ? __ldg(src + c) : make_int4(0, 0, 0, 0);
            const int r = c / CH, a0 = (c % CH) * 4;
            v[a0][r] = q.x; v[a0 + 1][r] = q.y; v[a0 + 2][r] = q.z; v[a0 + 3][r] = q.w;
        }
    }
};
template <int KT>
struct KsRowShared {
    const KsCandTile<KT>& tile;
    int col;
    __device__ __forceinline__ int operator()(int a) const { return tile.v[a][col]; }
    __device__ __forceinline__ int4 chunk(int c) const {
        return make_int4(tile.v[4 * c][col], tile.v[4 * c + 1][col], tile.v[4 * c + 2][col], tile.v[4 * c + 3][col]);
    }
};

// The same rows staged by the TMA unit (cp.async.bulk, completion on an mbarrier): every query's candidate row is one contiguous,
// 16-byte-aligned piece of the candidate array, copied by ONE bulk-copy instruction of its own thread into a row of shared
// memory -- no register round trip, no per-element shared-memory stores, no block barrier after the fill.  Rows are padded to an
// odd number of 16-byte units, so the 16-byte vector reads of a lane's own row are free of bank conflicts (8 lanes x 16 bytes
// cover the 32 banks), and the sorted row leaves the same way: written into the lane's row, stored to the neighbour table by
// one bulk copy per thread (cp.async.bulk.global.shared::cta).
template <int KT>
struct KsCandRows {
    static constexpr int STRIDE = KT + 4;                        // ints: KT / 4 + 1 units of 16 bytes, odd for KT = 16, 32, 64
    alignas(128) int v[128][STRIDE];
};
struct KsRowPadded {
    const int* row;
    __device__ __forceinline__ int operator()(int a) const { return row[a]; }
    __device__ __forceinline__ int4 chunk(int c) const { return *reinterpret_cast<const int4*>(row + 4 * c); }
};

// Keys of the re-ranking tier: a candidate is one of at most 64 slots of the row, so only 6 id bits are needed and the
// distance field keeps all 23 mantissa bits of y = d2/h^2 + 32: its unit is 2^-18 h^2 instead of the 2^-16 of the streaming
// tiers.  Error budget per key: 0.5 unit (rounding of y) + 0.15 (fp32 evaluation of d2) -> two candidates whose fields differ
// by more than KS_RR_MARGIN units are in their exact order.
constexpr int KS_RR_SHIFT = 9;
constexpr unsigned KS_RR_IDMASK = (1u << KS_RR_SHIFT) - 1u;
constexpr unsigned KS_RR_MARGIN = 3;

template <int K, class Ids>
__device__ __forceinline__ void ks_rerank_keys(const Ids& ids, int first, const float4* __restrict__ pts,
                                               float qx, float qy, float qz, float inv_h2, unsigned (&out)[K]) {
#pragma unroll
    for (int c = 0; c < K / 4; ++c) {
        const int4 q4 = ids.chunk(first / 4 + c);                        // 16 bytes of ids per access
        const int jj[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int a = 4 * c + e, j = jj[e];
            const float4 p = __ldg(pts + max(j, 0));
            const float dx = p.x - qx, dy = p.y - qy, dz = p.z - qz;
            const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float y = fminf(fmaf(d2, inv_h2, 32.0f), 63.99999f);  // NaN and far-away candidates saturate
            out[a] = j >= 0 ? ((__float_as_uint(y) << KS_RR_SHIFT) | (unsigned)(first + a)) : KS_NONE;
        }
    }
}

template <int K, class Ids>
__device__ __forceinline__ bool ks_rerank(KsTop<ks_kf(K, 2 * K)>& t, const GridView& g, const Ids& ids, float4 anchor,
                                          float qx, float qy, float qz, double (&ex)[ks_kf(K, 2 * K)]) {
    constexpr int KF = ks_kf(K, 2 * K);
    static_assert(K % 4 == 0 && 2 * K <= (1 << KS_RR_SHIFT) && KF < 2 * K, "candidate slots must fit the key's id field");
    const float inv_h2 = (float)(g.inv_h * g.inv_h);
    unsigned lo[K], hi[K];
    ks_rerank_keys<K>(ids, 0, g.pts, qx, qy, qz, inv_h2, lo);
    ks_rerank_keys<K>(ids, K, g.pts, qx, qy, qz, inv_h2, hi);
    ks_sort<K>(lo);
    ks_sort<K>(hi);
#pragma unroll
    for (int i = 0; i < K / 2; ++i) { unsigned x = hi[i]; hi[i] = hi[K - 1 - i]; hi[K - 1 - i] = x; }   // descending (register renaming)
#pragma unroll
    for (int i = 0; i < K; ++i) ks_ce(lo[i], hi[i]);              // ascending against descending: two bitonic halves
    ks_bitonic_merge<K>(lo);
    ks_bitonic_merge<K>(hi);                                        // lo, hi = the 2K keys in ascending order
    // Are the first K places decided by the keys alone?  Only if every neighbouring pair up to (K, K+1) differs by more than
    // the margin; otherwise (1-2 % of the lanes at this resolution; exact duplicates always) the first KF candidates are
    // re-evaluated in fp64 and ordered by (distance, original index) as in the search tiers.
    bool tight = false;
#pragma unroll
    for (int a = 0; a < K; ++a) {
        const unsigned x = lo[a] >> KS_RR_SHIFT, y = (a + 1 < K ? lo[a + 1 < K ? a + 1 : 0] : hi[0]) >> KS_RR_SHIFT;
        tight = tight || y <= x + KS_RR_MARGIN;
    }
#pragma unroll
    for (int a = 0; a < KF; ++a) {
        const unsigned key = a < K ? lo[a < K ? a : 0] : hi[a >= K ? a - K : 0];
        t.key[a] = key;
        t.id[a] = key != KS_NONE ? ids((int)(key & KS_RR_IDMASK)) : -1;
    }
    const unsigned worst = lo[K - 1], next = hi[KF - K];
    if (__any_sync(FULL, tight)) {
        ks_finalize<KF, KF>(t, g.pts, qx, qy, qz, ex);            // (lanes whose order was already decided pass through unchanged)
    } else {
        // only the K-th distance is needed, for the certificate below
        const int j = t.id[K - 1];
        const float4 p = __ldg(g.pts + max(j, 0));
        const double dx = (double)qx - (double)p.x, dy = (double)qy - (double)p.y, dz = (double)qz - (double)p.z;
        ex[K - 1] = j >= 0 ? __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) : DBL_MAX;
    }
    // distance travelled since the anchor, rounded up
    const float ax = qx - anchor.x, ay = qy - anchor.y, az = qz - anchor.z;
    const float delta = sqrtf(fmaf(az, az, fmaf(ay, ay, ax * ax))) * 1.000001f;
    const bool full = worst != KS_NONE;
    const bool clear = (next >> KS_RR_SHIFT) > (worst >> KS_RR_SHIFT) + KS_RR_MARGIN;
    const bool inside = sqrt(ex[K - 1]) * (1.0 + 1e-12) + (double)delta < (double)anchor.w;
    return full && clear && inside;
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
