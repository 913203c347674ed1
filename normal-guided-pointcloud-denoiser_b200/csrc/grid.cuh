// Frozen spatial index ("tree") over a point set: a two-level sparse uniform grid.
//   level 0: dense table over 8x8x8-cell bricks  -> slot of the occupied brick, or -1
//   level 1: per occupied brick, 513 start offsets into the sorted point array (cells x-fastest
//            inside the brick, so a run of cells along x is one contiguous range of points)
// Points are sorted by (Morton code of the brick, linear cell id inside the brick); bricks that are
// close in space are close in memory, which keeps neighbour gathers inside L2 and makes contiguous
// key ranges ("Morton slabs") compact pieces of the surface for multi-GPU partitioning.
// Replaces scipy.spatial.KDTree (Selector.py:141) / torch_cluster's nanoflann trees.
#pragma once
#include "common.cuh"

namespace ngpd {

constexpr int KS_PAD = 8;   // padding entries behind GridView::pts: the streaming k-NN steps load up to KS_PAD candidates and mask the overrun

struct GridView {
    const float4* pts;   // sorted points: x, y, z, original index (int bits)
    const int* top;      // [tbx*tby*tbz]
    const int* fine;     // [nbricks*513]
    double ox, oy, oz, inv_h, h;
    int nx, ny, nz;      // cells per axis
    int tbx, tby, tbz;   // bricks per axis
    int n;
};

}  // namespace ngpd

struct ngpd_grid {
    ngpd::GridView v;
    float4* pts = nullptr;
    int* top = nullptr;
    int* fine = nullptr;
    int64_t n = 0;
    int nbricks = 0;
    int64_t occupied_cells = 0;
    int64_t bytes = 0;
    int rebuilds = 0;
    float bbox[6];
};

namespace ngpd {

#if defined(__CUDACC__)
__host__ __device__ __forceinline__ uint64_t spread3(uint64_t v) {  // 21 bits -> every third bit
    v &= 0x1fffffULL;
    v = (v | (v << 32)) & 0x1f00000000ffffULL;
    v = (v | (v << 16)) & 0x1f0000ff0000ffULL;
    v = (v | (v << 8)) & 0x100f00f00f00f00fULL;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ULL;
    v = (v | (v << 2)) & 0x1249249249249249ULL;
    return v;
}
__host__ __device__ __forceinline__ uint32_t compact3(uint64_t v) {
    v &= 0x1249249249249249ULL;
    v = (v ^ (v >> 2)) & 0x10c30c30c30c30c3ULL;
    v = (v ^ (v >> 4)) & 0x100f00f00f00f00fULL;
    v = (v ^ (v >> 8)) & 0x1f0000ff0000ffULL;
    v = (v ^ (v >> 16)) & 0x1f00000000ffffULL;
    v = (v ^ (v >> 32)) & 0x1fffffULL;
    return (uint32_t)v;
}
__host__ __device__ __forceinline__ uint64_t cell_key(int cx, int cy, int cz) {
    uint64_t brick = spread3((uint64_t)(cx >> 3)) | (spread3((uint64_t)(cy >> 3)) << 1) | (spread3((uint64_t)(cz >> 3)) << 2);
    return (brick << 9) | (uint64_t)(((cz & 7) << 6) | ((cy & 7) << 3) | (cx & 7));
}
// fp64 binning: a point can only be mis-binned by ~1e-13 of a cell, far below the slack the
// search keeps on its guaranteed radius
__device__ __forceinline__ int cell_of(double p, double o, double inv_h, int n) {
    int c = (int)floor((p - o) * inv_h);
    return min(max(c, 0), n - 1);
}
#endif

// device-wide primitives implemented in grid.cu and reused by other translation units
int exclusive_scan_i32(int* data, int64_t n, int* total_out_host, cudaStream_t stream);
int radix_sort_pairs(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp, int64_t n,
                     int significant_bits, cudaStream_t stream, bool* result_in_tmp);
// cell keys of arbitrary points w.r.t. a grid (used to put foreign queries in a coherent order)
int point_keys(const GridView& g, const float* pos_packed, int64_t n, uint64_t* keys, uint32_t* vals, cudaStream_t stream);
int key_bits(const GridView& g);

}  // namespace ngpd
