// Radius ("ball") selection over the frozen grid: replaces scipy.spatial.KDTree.query_ball_point(pos, radii)
// (Selector.getPointsInRangeSelectionVectorized, Selector.py:214-229) for the Yadav-2018 baseline path.
// Per query: every tree point whose fp64 squared distance ((dx^2+dy^2)+dz^2 of the fp32 coordinates, SciPy's order of
// summation) is <= radius^2 (radius upcast to fp64, squared in fp64), listed in ascending ORIGINAL index -- the order SciPy
// gives for multi-point queries.  Rows have different lengths, so the call is made twice: a counting pass, then (after an
// exclusive scan of the counts by the caller) a filling pass that also sorts each row in place.
// One query per thread; queries are visited in a spatially coherent order when the caller says they are the tree points.
#include "knn.cuh"
#include "../../include/ngpd.h"

namespace ngpd {

template <class Visit>
__device__ __forceinline__ void ball_visit(const GridView& g, double qx, double qy, double qz, double r, Visit&& visit) {
    const double r2 = r * r;
    const double pad = r + g.h * 1e-9;                       // binning error of the build (grid.cuh: cell_of)
    const double rx = qx - g.ox, ry = qy - g.oy, rz = qz - g.oz;
    const int x0 = max((int)floor((rx - pad) * g.inv_h), 0), x1 = min((int)floor((rx + pad) * g.inv_h), g.nx - 1);
    const int y0 = max((int)floor((ry - pad) * g.inv_h), 0), y1 = min((int)floor((ry + pad) * g.inv_h), g.ny - 1);
    const int z0 = max((int)floor((rz - pad) * g.inv_h), 0), z1 = min((int)floor((rz + pad) * g.inv_h), g.nz - 1);
    for (int z = z0; z <= z1; ++z) {
        const double gz = axis_gap(rz, z, g.h);
        for (int y = y0; y <= y1; ++y) {
            const double gy = axis_gap(ry, y, g.h);
            if (gz * gz + gy * gy > r2 + pad * 1e-6) continue;            // the whole row of cells is outside the ball
            const int64_t trow = ((int64_t)(z >> 3) * g.tby + (y >> 3)) * g.tbx;
            const int lrow = ((z & 7) << 6) | ((y & 7) << 3);
            int xa = x0;
            while (xa <= x1) {
                const int xe = min(x1, xa | 7);
                const int b = __ldg(g.top + trow + (xa >> 3));
                if (b >= 0) {
                    const int* f = g.fine + (int64_t)b * 513 + lrow;
                    const int s = __ldg(f + (xa & 7)), e = __ldg(f + (xe & 7) + 1);
                    for (int j = s; j < e; ++j) {
                        const float4 p = __ldg(g.pts + j);
                        const double dx = qx - (double)p.x, dy = qy - (double)p.y, dz = qz - (double)p.z;
                        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        if (d2 <= r2) visit(__float_as_int(p.w));
                    }
                }
                xa = xe + 1;
            }
        }
    }
}

__global__ void __launch_bounds__(128) ball_kernel(GridView g, const float* __restrict__ query, const int32_t* __restrict__ qorder, int64_t m,
                                                   const float* __restrict__ radii, int32_t* __restrict__ counts,
                                                   const int32_t* __restrict__ offsets, int32_t* __restrict__ idx) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m) return;
    const int64_t q = qorder ? (int64_t)qorder[t] : t;
    const double qx = (double)__ldg(query + 3 * q), qy = (double)__ldg(query + 3 * q + 1), qz = (double)__ldg(query + 3 * q + 2);
    const float rf = __ldg(radii + q);
    const double r = (double)rf;
    if (!(rf >= 0.0f)) {                                       // negative or NaN radius: empty row
        if (!idx) counts[q] = 0;
        return;
    }
    if (!idx) {
        int c = 0;
        ball_visit(g, qx, qy, qz, r, [&](int) { ++c; });
        counts[q] = c;
        return;
    }
    int32_t* row = idx + offsets[q];
    const int cap = offsets[q + 1] - offsets[q];
    int c = 0;
    // insertion into the sorted row as the points come (rows are short: a ball of two mean edge lengths holds ~12 points)
    ball_visit(g, qx, qy, qz, r, [&](int orig) {
        if (c >= cap) return;                                  // cannot happen when offsets come from the counting pass
        int a = c++;
        while (a > 0 && row[a - 1] > orig) { row[a] = row[a - 1]; --a; }
        row[a] = orig;
    });
}

__global__ void __launch_bounds__(256) ball_order_kernel(const float4* __restrict__ pts, int64_t m, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = __float_as_int(pts[i].w);
}

}  // namespace ngpd

using namespace ngpd;

extern "C" __attribute__((visibility("default"))) int ngpd_ball_query(const ngpd_grid_t* grid, const float* query, int64_t m, const float* radii, int flags,
                                                                  int32_t* counts_out, const int32_t* offsets, int32_t* idx_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(grid && query && radii, "ngpd_ball_query: NULL argument");
    NGPD_REQUIRE((counts_out && !idx_out) || (idx_out && offsets), "ngpd_ball_query: pass counts_out (counting pass) or offsets + idx_out (filling pass)");
    NGPD_REQUIRE(!(flags & NGPD_KNN_QUERY_IS_TREE) || m == grid->n, "ngpd_ball_query: NGPD_KNN_QUERY_IS_TREE needs one query per tree point");
    if (m <= 0) return 0;
    int32_t* order = nullptr;
    if (flags & NGPD_KNN_QUERY_IS_TREE) {
        // query i is (the possibly moved) tree point i: walk them in tree order so that a warp's balls overlap
        NGPD_CUDA_OK(cudaMallocAsync(&order, (size_t)m * sizeof(int32_t), st));
        ball_order_kernel<<<(unsigned)cdiv(m, 256), 256, 0, st>>>(grid->pts, m, order);
    }
    ball_kernel<<<(unsigned)cdiv(m, 128), 128, 0, st>>>(grid->v, query, order, m, radii, counts_out, offsets, idx_out);
    NGPD_CUDA_OK(cudaGetLastError());
    if (order) NGPD_CUDA_OK(cudaFreeAsync(order, st));
    return 0;
}
