// TEST-ONLY: compiles the product's per-point math headers (csrc/point_math.cuh, csrc/eig3.cuh) for the host so
// that the arithmetic can be checked against the golden vectors on a machine without a GPU.  Nothing in the
// product package loads this library; the CUDA kernels include the very same headers.
#include <cstdint>
#include <cmath>
using std::signbit;
using std::isfinite;
#include "../../normal-guided-pointcloud-denoiser_b200/csrc/point_math.cuh"
#include "../../normal-guided-pointcloud-denoiser_b200/csrc/eig3_fast.cuh"
#include "eig3_generic.h"
#include "../../normal-guided-pointcloud-denoiser_b200/csrc/sortnet.cuh"

using namespace ngpd;

struct HostPacked3 {
    const float* p;
    V3 operator()(int64_t i) const { return v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
};

extern "C" {

void hm_eigh3(const float* T, int64_t m, float* w, float* V) {
    for (int64_t r = 0; r < m; ++r) {
        const float* a = T + 9 * r;
        eigh3_lapack(a[0], a[3], a[6], a[4], a[7], a[8], w + 3 * r, V + 9 * r);
    }
}

// the array-indexed transcription of the LAPACK loops (cross-check for the n = 3 specialisation)
void hm_eigh3_generic(const float* T, int64_t m, float* w, float* V) {
    for (int64_t r = 0; r < m; ++r) {
        const float* a = T + 9 * r;
        ngpd_generic::eigh3_lapack(a[0], a[3], a[6], a[4], a[7], a[8], w + 3 * r, V + 9 * r);
    }
}

void hm_nvt(const float* pos, const float* nrm, const int32_t* idx, const int32_t* rows, int64_t m, int k, float x_thresh,
            float* w, float* V, float* T, int32_t* sumw) {
    HostPacked3 P{pos}, N{nrm};
    for (int64_t r = 0; r < m; ++r) {
        NvtResult o;
        float t6[6];
        nvt_point(P, N, rows ? (int64_t)rows[r] : r, idx + r * k, k, x_thresh, o, t6);
        for (int c = 0; c < 3; ++c) w[3 * r + c] = o.w[c];
        for (int c = 0; c < 9; ++c) V[9 * r + c] = o.V[c];
        float* t = T + 9 * r;
        t[0] = t6[0]; t[1] = t6[1]; t[2] = t6[2]; t[3] = t6[1]; t[4] = t6[3]; t[5] = t6[4]; t[6] = t6[2]; t[7] = t6[4]; t[8] = t6[5];
        sumw[r] = o.sumw;
    }
}

// neighbour-filter decision: the reference's exact sequence and the division-free shortcut, for the same edges
void hm_weight(const float* vi, const float* vj, const float* nj, int64_t m, float x_thresh, uint8_t* exact, uint8_t* quick) {
    HostPacked3 A{vi}, B{vj}, N{nj};
    NvtThreshold th(x_thresh);
    for (int64_t r = 0; r < m; ++r) {
        exact[r] = nvt_weight(A(r), B(r), N(r), x_thresh);
        float slack = th.quick ? 1.0f : -1.0f;
        bool q = nvt_weight_quick(A(r), B(r), N(r), th, slack);
        quick[r] = slack > NGPD_NVT_SLACK_FLOOR ? q : nvt_weight(A(r), B(r), N(r), x_thresh);
    }
}

void hm_smooth(const float* w, const float* V, const float* nrm, int64_t m, float tau, float damp, float* out) {
    HostPacked3 N{nrm};
    for (int64_t r = 0; r < m; ++r) {
        V3 f = smooth_normal(w + 3 * r, V + 9 * r, N(r), tau, damp);
        out[3 * r] = f.x; out[3 * r + 1] = f.y; out[3 * r + 2] = f.z;
    }
}

void hm_classify(const float* w, int64_t m, float scale, uint8_t* lab) {
    for (int64_t r = 0; r < m; ++r) lab[r] = (uint8_t)classify(w + 3 * r, scale);
}

// stage-2 shortcut (csrc/eig3_fast.cuh): label, certainty flag, smallest eigenvalue per tensor
void hm_classify_fast(const float* T, int64_t m, float scale, uint8_t* lab, uint8_t* certain, float* l3) {
    for (int64_t r = 0; r < m; ++r) {
        const float* a = T + 9 * r;
        FastLabel f = classify_fast(a[0], a[3], a[6], a[4], a[7], a[8], scale);
        lab[r] = (uint8_t)f.label; certain[r] = f.certain; l3[r] = f.l3;
    }
}

void hm_pca(const float* pos, const int32_t* idx, int64_t m, int k, float* normals) {
    HostPacked3 P{pos};
    for (int64_t r = 0; r < m; ++r) {
        float w[3], V[9];
        pca_point(P, idx + r * k, k, w, V);
        normals[3 * r] = V[0]; normals[3 * r + 1] = V[3]; normals[3 * r + 2] = V[6];
    }
}

void hm_update(int kind, const float* pos, const float* nrm, const float* edge, const int32_t* idx, const int32_t* rows, int64_t m, int k,
               float alpha, float dmax, float delta, float* out) {
    HostPacked3 P{pos}, N{nrm}, E{edge};
    for (int64_t r = 0; r < m; ++r) {
        int64_t c = rows ? (int64_t)rows[r] : r;
        V3 p;
        if (kind == 0) p = flat_point(P, N, c, idx + r * k, k, delta, alpha, dmax);
        else if (kind == 1) p = edge_point(P, N, E(c), c, idx + r * k, k, alpha, dmax);
        else if (kind == 2) p = feature_point(P, N, c, idx + r * k, k, alpha, dmax);
        else p = corner_point(P, N, c, idx + r * k, k, alpha, dmax);
        out[3 * r] = p.x; out[3 * r + 1] = p.y; out[3 * r + 2] = p.z;
    }
}

// Yadav-2018 baseline tensors on CSR rows: mode 0 = normal-filtered NVT, 1 = normal-filtered PVT
void hm_normal_filtered(int mode, const float* pos, const float* nrm, const int32_t* idx, const int32_t* offsets, int64_t m, float x_le,
                        float* w, float* V, float* T, int32_t* sumw) {
    HostPacked3 P{pos}, N{nrm};
    for (int64_t r = 0; r < m; ++r) {
        NvtResult o;
        float t6[6];
        const int32_t* row = idx + offsets[r];
        const int cnt = offsets[r + 1] - offsets[r];
        if (mode == 0) nvt_normal_point(N, r, row, cnt, x_le, o, t6);
        else pvt_normal_point(P, N, r, row, cnt, x_le, o, t6);
        for (int c = 0; c < 3; ++c) w[3 * r + c] = o.w[c];
        for (int c = 0; c < 9; ++c) V[9 * r + c] = o.V[c];
        float* t = T + 9 * r;
        t[0] = t6[0]; t[1] = t6[1]; t[2] = t6[2]; t[3] = t6[1]; t[4] = t6[3]; t[5] = t6[4]; t[6] = t6[2]; t[7] = t6[4]; t[8] = t6[5];
        sumw[r] = o.sumw;
    }
}

// sorting networks of the k-NN selection (csrc/sortnet.cuh): sort / bitonic-merge one array of n = 8, 16 or 32 keys
int hm_sortnet(unsigned* keys, int n, int bitonic) {
#define NGPD_SN(N) if (n == N) { unsigned a[N]; for (int i = 0; i < N; ++i) a[i] = keys[i]; \
        if (bitonic) ks_bitonic_merge<N>(a); else ks_sort<N>(a); for (int i = 0; i < N; ++i) keys[i] = a[i]; return 0; }
    NGPD_SN(8) NGPD_SN(16) NGPD_SN(32)
#undef NGPD_SN
    return -1;
}

}  // extern "C"
