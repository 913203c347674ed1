// Shared helpers for the ngpd CUDA library (sm_100a).  No torch types anywhere in csrc/.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cmath>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define NGPD_HD __host__ __device__ __forceinline__
#define NGPD_HD_COLD static __host__ __device__ __noinline__     // rare paths: kept out of line so the hot code stays small
#else
#define NGPD_HD inline
#define NGPD_HD_COLD static inline
#endif

namespace ngpd {

struct V3 {
    float x, y, z;
};

NGPD_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
NGPD_HD V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
NGPD_HD V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
NGPD_HD V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
// three-term dot with the reference's rounding: products rounded, then (p0+p1)+p2.
// (the library is compiled with --fmad=false so none of this contracts)
NGPD_HD float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// torch-CPU vector_norm over 3 elements: fma chain  fma(z,z,fma(y,y,x*x))  (probed)
NGPD_HD float norm3_fma(V3 a) { return sqrtf(fmaf(a.z, a.z, fmaf(a.y, a.y, a.x * a.x))); }

}  // namespace ngpd

#if defined(__CUDACC__)
namespace ngpd {

void set_error(const char* fmt, ...);

#define NGPD_CUDA_OK(expr)                                                                    \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ngpd::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -2;                                                                        \
        }                                                                                     \
    } while (0)

#define NGPD_REQUIRE(cond, msg)                                        \
    do {                                                               \
        if (!(cond)) {                                                 \
            ngpd::set_error("%s:%d: %s", __FILE__, __LINE__, msg);     \
            return -1;                                                 \
        }                                                              \
    } while (0)

// positions / normals as the public ABI passes them: packed [n,3] fp32
struct Packed3 {
    const float* p;
    __device__ __forceinline__ V3 operator()(int64_t i) const {
        const float* q = p + 3 * i;
        return v3(__ldg(q), __ldg(q + 1), __ldg(q + 2));
    }
};
// internal (tree-order) arrays: float4 per point, one 16-byte load per gather
struct Quad4 {
    const float4* p;
    __device__ __forceinline__ V3 operator()(int64_t i) const {
        float4 q = __ldg(p + i);
        return v3(q.x, q.y, q.z);
    }
};

// stream-ordered scratch that is returned on every exit path (early error returns included)
template <class T>
struct StreamBuf {
    T* p = nullptr;
    cudaStream_t st;
    explicit StreamBuf(cudaStream_t s) : st(s) {}
    StreamBuf(const StreamBuf&) = delete;
    StreamBuf& operator=(const StreamBuf&) = delete;
    cudaError_t alloc(size_t count) { return cudaMallocAsync(&p, count * sizeof(T), st); }
    T* release() { T* q = p; p = nullptr; return q; }
    ~StreamBuf() { if (p) cudaFreeAsync(p, st); }
    operator T*() const { return p; }
};

static inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace ngpd
#endif
