// Kernel groups 3 and 4 behind the public ABI: per-row neighbourhood tensors, eigen-decomposition,
// smoothing, labels and the class-wise position updates, on packed [n,3] arrays in ORIGINAL point order.
// One row per thread; the arithmetic lives in point_math.cuh.  (The fused tree-order variants used by
// Processor.denoise are in session.cu.)
#include "point_math.cuh"
#include "../../include/ngpd.h"

namespace ngpd {

struct RowSpan { const int32_t* nbr; int cnt; int64_t centre; };

__device__ __forceinline__ RowSpan row_span(int64_t r, const int32_t* idx, const int32_t* offsets, const int32_t* rows, int k) {
    RowSpan s;
    if (offsets) { int a = offsets[r]; s.nbr = idx + a; s.cnt = offsets[r + 1] - a; }
    else { s.nbr = idx + r * k; s.cnt = k; }
    s.centre = rows ? (int64_t)rows[r] : r;
    return s;
}

__global__ void __launch_bounds__(128) pca_kernel(Packed3 pos, const int32_t* __restrict__ idx, const int32_t* __restrict__ offsets,
                                                  int64_t m, int k, float* __restrict__ normals, float* __restrict__ eigval,
                                                  float* __restrict__ eigvec) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    RowSpan s = row_span(r, idx, offsets, nullptr, k);
    float w[3], V[9];
    pca_point(pos, s.nbr, s.cnt, w, V);
    if (normals) { normals[3 * r] = V[0]; normals[3 * r + 1] = V[3]; normals[3 * r + 2] = V[6]; }
    if (eigval) { eigval[3 * r] = w[0]; eigval[3 * r + 1] = w[1]; eigval[3 * r + 2] = w[2]; }
    if (eigvec) {
#pragma unroll
        for (int c = 0; c < 9; ++c) eigvec[9 * r + c] = V[c];
    }
}

__global__ void __launch_bounds__(128) nvt_kernel(Packed3 pos, Packed3 nrm, const int32_t* __restrict__ idx, const int32_t* __restrict__ offsets,
                                                  const int32_t* __restrict__ rows, int64_t m, int k, float x_thresh,
                                                  float* __restrict__ eigval, float* __restrict__ eigvec, float* __restrict__ tensor,
                                                  int32_t* __restrict__ sumw) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    RowSpan s = row_span(r, idx, offsets, rows, k);
    NvtResult o;
    float t6[6];
    nvt_point(pos, nrm, s.centre, s.nbr, s.cnt, x_thresh, o, t6);
#pragma unroll
    for (int c = 0; c < 3; ++c) eigval[3 * r + c] = o.w[c];
#pragma unroll
    for (int c = 0; c < 9; ++c) eigvec[9 * r + c] = o.V[c];
    if (tensor) {
        float* T = tensor + 9 * r;
        T[0] = t6[0]; T[1] = t6[1]; T[2] = t6[2]; T[3] = t6[1]; T[4] = t6[3]; T[5] = t6[4]; T[6] = t6[2]; T[7] = t6[4]; T[8] = t6[5];
    }
    if (sumw) sumw[r] = o.sumw;
}

// Yadav-2018 baseline tensors: neighbours filtered by the angle between the normals.  mode 0 = getNormalFilteredNVT,
// mode 1 = getNormalFilteredPVT (Decompositionor.py:260-276, 172-211)
__global__ void __launch_bounds__(128) normal_filtered_kernel(int mode, Packed3 pos, Packed3 nrm, const int32_t* __restrict__ idx,
                                                              const int32_t* __restrict__ offsets, const int32_t* __restrict__ rows, int64_t m, int k,
                                                              float x_le, float* __restrict__ eigval, float* __restrict__ eigvec,
                                                              float* __restrict__ tensor, int32_t* __restrict__ sumw) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    RowSpan s = row_span(r, idx, offsets, rows, k);
    NvtResult o;
    float t6[6];
    if (mode == 0) nvt_normal_point(nrm, s.centre, s.nbr, s.cnt, x_le, o, t6);
    else pvt_normal_point(pos, nrm, s.centre, s.nbr, s.cnt, x_le, o, t6);
#pragma unroll
    for (int c = 0; c < 3; ++c) eigval[3 * r + c] = o.w[c];
#pragma unroll
    for (int c = 0; c < 9; ++c) eigvec[9 * r + c] = o.V[c];
    if (tensor) {
        float* T = tensor + 9 * r;
        T[0] = t6[0]; T[1] = t6[1]; T[2] = t6[2]; T[3] = t6[1]; T[4] = t6[3]; T[5] = t6[4]; T[6] = t6[2]; T[7] = t6[4]; T[8] = t6[5];
    }
    if (sumw) sumw[r] = o.sumw;
}

__global__ void __launch_bounds__(128) eigh3_kernel(const float* __restrict__ T, int64_t m, float* __restrict__ eigval, float* __restrict__ eigvec) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const float* a = T + 9 * r;
    float w[3], V[9];
    eigh3_lapack(a[0], a[3], a[6], a[4], a[7], a[8], w, V);
#pragma unroll
    for (int c = 0; c < 3; ++c) eigval[3 * r + c] = w[c];
#pragma unroll
    for (int c = 0; c < 9; ++c) eigvec[9 * r + c] = V[c];
}

__global__ void __launch_bounds__(128) smooth_kernel(const float* __restrict__ eigval, const float* __restrict__ eigvec, Packed3 nrm,
                                                     int64_t m, float tau, float damp, float* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    float w[3], V[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) w[c] = eigval[3 * r + c];
#pragma unroll
    for (int c = 0; c < 9; ++c) V[c] = eigvec[9 * r + c];
    V3 f = smooth_normal(w, V, nrm(r), tau, damp);
    out[3 * r] = f.x; out[3 * r + 1] = f.y; out[3 * r + 2] = f.z;
}

__global__ void __launch_bounds__(256) classify_kernel(const float* __restrict__ eigval, int64_t m, float scale,
                                                       uint8_t* __restrict__ labels, float* __restrict__ features) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    float w[3] = {eigval[3 * r], eigval[3 * r + 1], eigval[3 * r + 2]};
    labels[r] = (uint8_t)classify(w, scale);
    if (features) {
        float l1 = w[2], l2 = w[1], l3 = w[0];
        features[3 * r] = (l1 - l2) / l1; features[3 * r + 1] = (l2 - l3) / l1; features[3 * r + 2] = l3 / l1;
    }
}

// ---- flat_step's global scalars: two passes (sum, then max distance from the mean) -----------------
__global__ void __launch_bounds__(256) nbr_sum_kernel(Packed3 pos, const int32_t* __restrict__ idx, int64_t total, double* __restrict__ acc3) {
    double sx = 0, sy = 0, sz = 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        V3 p = pos((int64_t)idx[e]);
        sx += p.x; sy += p.y; sz += p.z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(acc3, sx); atomicAdd(acc3 + 1, sy); atomicAdd(acc3 + 2, sz); }
}
__global__ void center_finalize_kernel(const double* __restrict__ acc3, int64_t total, float* __restrict__ out4) {
    out4[0] = (float)(acc3[0] / (double)total); out4[1] = (float)(acc3[1] / (double)total); out4[2] = (float)(acc3[2] / (double)total);
    out4[3] = 0.0f;
}
__global__ void __launch_bounds__(256) nbr_maxdist_kernel(Packed3 pos, const int32_t* __restrict__ idx, int64_t total, float* __restrict__ out4) {
    V3 c = v3(out4[0], out4[1], out4[2]);
    float mx = 0.0f;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
        mx = fmaxf(mx, norm3_fma(pos((int64_t)idx[e]) - c));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax((int*)(out4 + 3), __float_as_int(mx));   // non-negative floats order as ints
}

__global__ void __launch_bounds__(128) update_kernel(int kind, Packed3 pos, Packed3 nrm, Packed3 edge, const int32_t* __restrict__ idx,
                                                     const int32_t* __restrict__ offsets, const int32_t* __restrict__ rows, int64_t m, int k,
                                                     float alpha, float dmax, const float* __restrict__ center_delta, float* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    RowSpan s = row_span(r, idx, offsets, rows, k);
    V3 p;
    if (kind == NGPD_STEP_FLAT) p = flat_point(pos, nrm, s.centre, s.nbr, s.cnt, center_delta[3], alpha, dmax);
    else if (kind == NGPD_STEP_EDGE) p = edge_point(pos, nrm, edge(s.centre), s.centre, s.nbr, s.cnt, alpha, dmax);
    else if (kind == NGPD_STEP_FEATURE) p = feature_point(pos, nrm, s.centre, s.nbr, s.cnt, alpha, dmax);
    else p = corner_point(pos, nrm, s.centre, s.nbr, s.cnt, alpha, dmax);
    out[3 * r] = p.x; out[3 * r + 1] = p.y; out[3 * r + 2] = p.z;
}

__global__ void __launch_bounds__(256) edge_length_kernel(Packed3 pos, const int32_t* __restrict__ idx, const int32_t* __restrict__ rows,
                                                          int64_t m, int k, double* __restrict__ out2) {
    double s = 0;
    int64_t total = m * k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e / k;
        int64_t c = rows ? (int64_t)rows[r] : r;
        s += (double)norm3_fma(pos((int64_t)idx[e]) - pos(c));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out2, s);
    if (blockIdx.x == 0 && threadIdx.x == 0) out2[1] = (double)total;
}

static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)cdiv(n, threads); }
static inline unsigned strided_grid(int64_t n, int threads) {
    int64_t b = cdiv(n, threads), cap = (int64_t)num_sms() * 8;
    return (unsigned)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace ngpd

using namespace ngpd;

extern "C" __attribute__((visibility("default"))) int ngpd_pca_normals(const float* pos, const int32_t* idx, const int32_t* offsets, int64_t m, int k,
                                float* normals_out, float* eigval_out, float* eigvec_out, void* stream) {
    NGPD_REQUIRE(pos && idx && (normals_out || eigvec_out), "ngpd_pca_normals: NULL argument");
    if (m <= 0) return 0;
    pca_kernel<<<grid_for(m, 128), 128, 0, (cudaStream_t)stream>>>(Packed3{pos}, idx, offsets, m, k, normals_out, eigval_out, eigvec_out);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_nvt(const float* pos, const float* nrm, const int32_t* idx, const int32_t* offsets, const int32_t* rows,
                        int64_t m, int k, float x_thresh, float* eigval_out, float* eigvec_out, float* tensor_out,
                        int32_t* sumw_out, void* stream) {
    NGPD_REQUIRE(pos && nrm && idx && eigval_out && eigvec_out, "ngpd_nvt: NULL argument");
    if (m <= 0) return 0;
    nvt_kernel<<<grid_for(m, 128), 128, 0, (cudaStream_t)stream>>>(Packed3{pos}, Packed3{nrm}, idx, offsets, rows, m, k, x_thresh,
                                                                  eigval_out, eigvec_out, tensor_out, sumw_out);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_nvt_normal(const float* nrm, const int32_t* idx, const int32_t* offsets, const int32_t* rows, int64_t m, int k,
                               float x_le, float* eigval_out, float* eigvec_out, float* tensor_out, int32_t* sumw_out, void* stream) {
    NGPD_REQUIRE(nrm && idx && eigval_out && eigvec_out, "ngpd_nvt_normal: NULL argument");
    if (m <= 0) return 0;
    normal_filtered_kernel<<<grid_for(m, 128), 128, 0, (cudaStream_t)stream>>>(0, Packed3{nrm}, Packed3{nrm}, idx, offsets, rows, m, k, x_le, eigval_out,
                                                                              eigvec_out, tensor_out, sumw_out);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_pvt_normal(const float* pos, const float* nrm, const int32_t* idx, const int32_t* offsets, const int32_t* rows,
                               int64_t m, int k, float x_le, float* eigval_out, float* eigvec_out, float* tensor_out, int32_t* sumw_out,
                               void* stream) {
    NGPD_REQUIRE(pos && nrm && idx && eigval_out && eigvec_out, "ngpd_pvt_normal: NULL argument");
    if (m <= 0) return 0;
    normal_filtered_kernel<<<grid_for(m, 128), 128, 0, (cudaStream_t)stream>>>(1, Packed3{pos}, Packed3{nrm}, idx, offsets, rows, m, k, x_le, eigval_out,
                                                                              eigvec_out, tensor_out, sumw_out);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_eigh3(const float* tensors, int64_t m, float* eigval_out, float* eigvec_out, void* stream) {
    NGPD_REQUIRE(tensors && eigval_out && eigvec_out, "ngpd_eigh3: NULL argument");
    if (m <= 0) return 0;
    eigh3_kernel<<<grid_for(m, 128), 128, 0, (cudaStream_t)stream>>>(tensors, m, eigval_out, eigvec_out);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_smooth_normals(const float* eigval, const float* eigvec, const float* nrm, int64_t m, float tau, float damp,
                                   float* out, void* stream) {
    NGPD_REQUIRE(eigval && eigvec && nrm && out, "ngpd_smooth_normals: NULL argument");
    if (m <= 0) return 0;
    smooth_kernel<<<grid_for(m, 128), 128, 0, (cudaStream_t)stream>>>(eigval, eigvec, Packed3{nrm}, m, tau, damp, out);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_classify(const float* eigval, int64_t m, float scale, uint8_t* labels_out, float* features_out, void* stream) {
    NGPD_REQUIRE(eigval && labels_out, "ngpd_classify: NULL argument");
    if (m <= 0) return 0;
    classify_kernel<<<grid_for(m, 256), 256, 0, (cudaStream_t)stream>>>(eigval, m, scale, labels_out, features_out);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_center_delta(const float* pos, const int32_t* idx, int64_t total, float* out4, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    NGPD_REQUIRE(pos && idx && out4 && total > 0, "ngpd_center_delta: bad argument");
    double* acc = nullptr;
    NGPD_CUDA_OK(cudaMallocAsync(&acc, 3 * sizeof(double), stream));
    NGPD_CUDA_OK(cudaMemsetAsync(acc, 0, 3 * sizeof(double), stream));
    nbr_sum_kernel<<<strided_grid(total, 256), 256, 0, stream>>>(Packed3{pos}, idx, total, acc);
    center_finalize_kernel<<<1, 1, 0, stream>>>(acc, total, out4);
    nbr_maxdist_kernel<<<strided_grid(total, 256), 256, 0, stream>>>(Packed3{pos}, idx, total, out4);
    NGPD_CUDA_OK(cudaGetLastError());
    NGPD_CUDA_OK(cudaFreeAsync(acc, stream));
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_update(int kind, const float* pos, const float* nrm, const float* edge_vec, const int32_t* idx,
                           const int32_t* offsets, const int32_t* rows, int64_t m, int k, float alpha, float dmax,
                           const float* center_delta, float* pos_out, void* stream) {
    NGPD_REQUIRE(pos && nrm && idx && pos_out, "ngpd_update: NULL argument");
    NGPD_REQUIRE(kind >= 0 && kind <= 3, "ngpd_update: unknown step kind");
    NGPD_REQUIRE(kind != NGPD_STEP_EDGE || edge_vec, "ngpd_update: edge step needs edge_vec");
    NGPD_REQUIRE(kind != NGPD_STEP_FLAT || center_delta, "ngpd_update: flat step needs center_delta");
    if (m <= 0) return 0;
    update_kernel<<<grid_for(m, 128), 128, 0, (cudaStream_t)stream>>>(kind, Packed3{pos}, Packed3{nrm}, Packed3{edge_vec}, idx, offsets, rows,
                                                                     m, k, alpha, dmax, center_delta, pos_out);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ngpd_edge_length_sum(const float* pos, const int32_t* idx, const int32_t* rows, int64_t m, int k, double* out2, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    NGPD_REQUIRE(pos && idx && out2, "ngpd_edge_length_sum: NULL argument");
    NGPD_CUDA_OK(cudaMemsetAsync(out2, 0, 2 * sizeof(double), stream));
    if (m <= 0) return 0;
    edge_length_kernel<<<strided_grid(m * k, 256), 256, 0, stream>>>(Packed3{pos}, idx, rows, m, k, out2);
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}
