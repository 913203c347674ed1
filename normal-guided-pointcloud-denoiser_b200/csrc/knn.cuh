// Kernel group 2: exact k-nearest-neighbour search over the two-level grid.
// One query per thread; the k best are a sorted list in registers, keyed on the fp64 squared distance
// of the fp32 coordinates ((dx^2+dy^2)+dz^2, the order SciPy's KD-tree sums in) with ties broken by the
// original tree index, so the result is a pure function of the inputs (SciPy leaves tie order open).
// Search: the query's own cell, then cubic shells of cells around it; a shell is walked as rows along x,
// each row being one contiguous range of the sorted point array per brick.  The search stops when the
// k-th distance is inside the block of cells already visited.
#pragma once
#include <cfloat>
#include "grid.cuh"

namespace ngpd {

template <int K>
struct TopK {
    double d[K];
    int id[K];   // sorted position in the tree array, -1 = empty
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int a = 0; a < K; ++a) { d[a] = DBL_MAX; id[a] = -1; }
    }
    __device__ __forceinline__ double worst() const { return d[K - 1]; }
};

__device__ __forceinline__ bool key_less(double d2, int orig, double D, int J, const float4* __restrict__ pts) {
    if (d2 < D) return true;
    if (d2 > D) return false;
    if (J < 0) return true;
    return orig < __float_as_int(__ldg(&pts[J].w));
}

template <int K>
__device__ __forceinline__ void topk_insert(TopK<K>& t, double d2, int j, int orig, const float4* __restrict__ pts) {
    // precondition: (d2, orig) < last element.  One pass from the tail: shift while the new key is smaller.
    bool prev = true;
#pragma unroll
    for (int a = K - 1; a > 0; --a) {
        bool sh = key_less(d2, orig, t.d[a - 1], t.id[a - 1], pts);
        double nd = sh ? t.d[a - 1] : (prev ? d2 : t.d[a]);
        int ni = sh ? t.id[a - 1] : (prev ? j : t.id[a]);
        t.d[a] = nd; t.id[a] = ni;
        prev = sh;
    }
    if (prev) { t.d[0] = d2; t.id[0] = j; }
}

template <int K>
__device__ __forceinline__ void scan_range(TopK<K>& t, const float4* __restrict__ pts, int s, int e,
                                           double qx, double qy, double qz, int self_orig) {
    for (int j = s; j < e; ++j) {
        float4 p = __ldg(pts + j);
        double dx = qx - (double)p.x, dy = qy - (double)p.y, dz = qz - (double)p.z;
        double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (d2 > t.worst()) continue;
        int orig = __float_as_int(p.w);
        if (orig == self_orig) continue;
        if (key_less(d2, orig, t.d[K - 1], t.id[K - 1], pts)) topk_insert<K>(t, d2, j, orig, pts);
    }
}

// cells [x0,x1] of row (y,z): split at brick boundaries, each piece is one contiguous point range
template <int K>
__device__ __forceinline__ void scan_row(TopK<K>& t, const GridView& g, int x0, int x1, int y, int z,
                                         double qx, double qy, double qz, int self_orig) {
    if ((unsigned)y >= (unsigned)g.ny || (unsigned)z >= (unsigned)g.nz) return;
    x0 = max(x0, 0); x1 = min(x1, g.nx - 1);
    const int64_t trow = ((int64_t)(z >> 3) * g.tby + (y >> 3)) * g.tbx;
    const int lrow = ((z & 7) << 6) | ((y & 7) << 3);
    while (x0 <= x1) {
        int xe = min(x1, (x0 | 7));
        int b = __ldg(g.top + trow + (x0 >> 3));
        if (b >= 0) {
            const int* f = g.fine + (int64_t)b * 513 + lrow;
            int s = __ldg(f + (x0 & 7)), e = __ldg(f + (xe & 7) + 1);
            scan_range<K>(t, g.pts, s, e, qx, qy, qz, self_orig);
        }
        x0 = xe + 1;
    }
}

// squared distance from coordinate q to the slab [lo, lo+h] of one cell index along an axis
__device__ __forceinline__ double axis_gap(double qrel, int c, double h) {
    double lo = (double)c * h, hi = lo + h;
    double g = qrel < lo ? lo - qrel : (qrel > hi ? qrel - hi : 0.0);
    return g;
}

template <int K>
__device__ __forceinline__ void knn_search(TopK<K>& t, const GridView& g, double qx, double qy, double qz, int self_orig) {
    const double rx = qx - g.ox, ry = qy - g.oy, rz = qz - g.oz;
    const int cx = min(max((int)floor(rx * g.inv_h), 0), g.nx - 1);
    const int cy = min(max((int)floor(ry * g.inv_h), 0), g.ny - 1);
    const int cz = min(max((int)floor(rz * g.inv_h), 0), g.nz - 1);
    const double slack = g.h * 1e-9;
    const int rmax = max(g.nx, max(g.ny, g.nz));
    for (int r = 0; r <= rmax; ++r) {
        if (r == 0) {
            scan_row<K>(t, g, cx, cx, cy, cz, qx, qy, qz, self_orig);
        } else {
            for (int dz = -r; dz <= r; ++dz) {
                const int z = cz + dz;
                if ((unsigned)z >= (unsigned)g.nz) continue;
                const double gz = axis_gap(rz, z, g.h);
                for (int dy = -r; dy <= r; ++dy) {
                    const int y = cy + dy;
                    if ((unsigned)y >= (unsigned)g.ny) continue;
                    const double gy = axis_gap(ry, y, g.h);
                    if (gz * gz + gy * gy > t.worst()) continue;   // the whole row is farther than the k-th
                    if (max(abs(dy), abs(dz)) == r) {
                        scan_row<K>(t, g, cx - r, cx + r, y, z, qx, qy, qz, self_orig);
                    } else {
                        scan_row<K>(t, g, cx - r, cx - r, y, z, qx, qy, qz, self_orig);
                        scan_row<K>(t, g, cx + r, cx + r, y, z, qx, qy, qz, self_orig);
                    }
                }
            }
        }
        // every point outside the visited block [c-r, c+r]^3 is at least `reach` away; sides where the
        // block already touches the grid boundary have nothing beyond them
        double reach = DBL_MAX;
        if (cx - r > 0) reach = fmin(reach, rx - (double)(cx - r) * g.h);
        if (cx + r < g.nx - 1) reach = fmin(reach, (double)(cx + r + 1) * g.h - rx);
        if (cy - r > 0) reach = fmin(reach, ry - (double)(cy - r) * g.h);
        if (cy + r < g.ny - 1) reach = fmin(reach, (double)(cy + r + 1) * g.h - ry);
        if (cz - r > 0) reach = fmin(reach, rz - (double)(cz - r) * g.h);
        if (cz + r < g.nz - 1) reach = fmin(reach, (double)(cz + r + 1) * g.h - rz);
        if (reach == DBL_MAX) break;                       // whole grid visited
        reach -= slack;
        if (reach > 0.0 && t.worst() < reach * reach) break;
    }
}

}  // namespace ngpd
