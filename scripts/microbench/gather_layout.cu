// Micro-benchmark behind DESIGN 10.2: what would the neighbourhood gathers of the tensor / update passes cost if positions and
// (smoothed) normals were interleaved in one 32-byte record and fetched with ONE 256-bit load (LDG.E.256 exists on sm_100a)
// instead of two 16-byte loads from two arrays?  Rows in tree order, 16 neighbours per row, 88 % of them within +-64 positions
// (the locality measured on the bench cloud), the rest within +-5000.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/gather_layout scripts/microbench/gather_layout.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

struct alignas(32) Rec { float4 p, n; };

__device__ __forceinline__ void ld256(const Rec* r, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(r));
}

template <int MODE, int MINB>   // 0: two arrays, two 16-byte loads; 1: records, two 16-byte loads; 2: records, one 32-byte load
__global__ void __launch_bounds__(128, MINB) gather(const float4* __restrict__ pos, const float4* __restrict__ nrm, const Rec* __restrict__ rec,
                                                   const int* __restrict__ idx, float4* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4* row = reinterpret_cast<const int4*>(idx + (size_t)i * 16);
    const float4 pi = MODE == 0 ? __ldg(pos + i) : __ldg(&rec[i].p);
    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int4 q = __ldg(row + c);
        const int jj[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float4 p, m;
            if (MODE == 0) { p = __ldg(pos + jj[e]); m = __ldg(nrm + jj[e]); }
            else if (MODE == 1) { p = __ldg(&rec[jj[e]].p); m = __ldg(&rec[jj[e]].n); }
            else ld256(rec + jj[e], p, m);
            const float d = (p.x - pi.x) * m.x + (p.y - pi.y) * m.y + (p.z - pi.z) * m.z;
            ax += d * m.x; ay += d * m.y; az += d * m.z;
        }
    }
    out[i] = make_float4(pi.x + ax, pi.y + ay, pi.z + az, 0.f);
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 10000000;
    std::vector<int> idx((size_t)n * 16);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 16; ++a) {
            const int span = (rnd() % 100) < 88 ? 64 : 5000;
            long j = (long)i + (long)(rnd() % (2 * span + 1)) - span;
            if (j < 0) j = 0;
            if (j >= n) j = n - 1;
            idx[(size_t)i * 16 + a] = (int)j;
        }
    float4 *pos, *nrm, *out; Rec* rec; int* didx;
    cudaMalloc(&pos, (size_t)n * 16); cudaMalloc(&nrm, (size_t)n * 16); cudaMalloc(&out, (size_t)n * 16);
    cudaMalloc(&rec, (size_t)n * 32); cudaMalloc(&didx, (size_t)n * 64);
    cudaMemset(pos, 0, (size_t)n * 16); cudaMemset(nrm, 0, (size_t)n * 16); cudaMemset(rec, 0, (size_t)n * 32);
    cudaMemcpy(didx, idx.data(), (size_t)n * 64, cudaMemcpyHostToDevice);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = (n + 127) / 128;
    const char* names[3] = {"two arrays, 2 x LDG.128 per neighbour", "32-byte records, 2 x LDG.128", "32-byte records, 1 x LDG.256"};
    for (int occ = 0; occ < 2; ++occ)
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(a);
            if (occ == 0) {
                if (mode == 0) gather<0, 10><<<blocks, 128>>>(pos, nrm, rec, didx, out, n);
                else if (mode == 1) gather<1, 10><<<blocks, 128>>>(pos, nrm, rec, didx, out, n);
                else gather<2, 10><<<blocks, 128>>>(pos, nrm, rec, didx, out, n);
            } else {
                if (mode == 0) gather<0, 16><<<blocks, 128>>>(pos, nrm, rec, didx, out, n);
                else if (mode == 1) gather<1, 16><<<blocks, 128>>>(pos, nrm, rec, didx, out, n);
                else gather<2, 16><<<blocks, 128>>>(pos, nrm, rec, didx, out, n);
            }
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep && ms < best) best = ms;
        }
        printf("%2d blocks/SM  %-40s %8.3f ms for %d rows x 16 neighbours (%s)\n", occ ? 16 : 10, names[mode], best, n, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
