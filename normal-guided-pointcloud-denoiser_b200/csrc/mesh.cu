// Mesh vertex update of the reference's Vertex_updating notebook / PatchGeneration.Modules.Mesh.updateVertices
// (Mesh.py:377-418): one Jacobi sweep
//     v_i <- v_i + 1 / (3 deg_i) * sum_{f incident to i} sum_{k in f} n_f (n_f . (v_k - v_i))
// over the vertex-triangle adjacency in CSR form (igl.vertex_triangle_adjacency's VF / NI pair).  fp64 like the
// reference's numpy arrays.  One vertex per thread; a sweep reads v (24 B) + 3 vertex ids, 3 positions and a normal per
// incident face through L2 and writes 24 B: bandwidth-bound, no tensor cores.  All vertices are updated from the same
// snapshot (the library method is the vectorised variant of the notebook, Vertex_updating.ipynb#c11).
#include "common.cuh"
#include "../../include/ngpd.h"

namespace ngpd {

__global__ void __launch_bounds__(128) mesh_vertex_update_kernel(const double* __restrict__ v, const int32_t* __restrict__ faces,
                                                                 const double* __restrict__ fn, const int32_t* __restrict__ vf,
                                                                 const int32_t* __restrict__ ni, int64_t nv, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const double vx = v[3 * i], vy = v[3 * i + 1], vz = v[3 * i + 2];
    const int a = ni[i], b = ni[i + 1];
    // the reference sums over the faces first (per corner and axis), then over the three corners (Mesh.py:410)
    double s[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int e = a; e < b; ++e) {
        const int64_t f = vf[e];
        const double nx = fn[3 * f], ny = fn[3 * f + 1], nz = fn[3 * f + 2];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t k = faces[3 * f + c];
            const double dx = v[3 * k] - vx, dy = v[3 * k + 1] - vy, dz = v[3 * k + 2] - vz;
            const double dot = __dadd_rn(__dadd_rn(__dmul_rn(nx, dx), __dmul_rn(ny, dy)), __dmul_rn(nz, dz));
            s[c][0] = __dadd_rn(s[c][0], __dmul_rn(dot, nx));
            s[c][1] = __dadd_rn(s[c][1], __dmul_rn(dot, ny));
            s[c][2] = __dadd_rn(s[c][2], __dmul_rn(dot, nz));
        }
    }
    const double den = 3.0 * (double)(b - a);          // an isolated vertex divides 0 by 0 in the reference as well
    out[3 * i] = vx + ((s[0][0] + s[1][0]) + s[2][0]) / den;
    out[3 * i + 1] = vy + ((s[0][1] + s[1][1]) + s[2][1]) / den;
    out[3 * i + 2] = vz + ((s[0][2] + s[1][2]) + s[2][2]) / den;
}

}  // namespace ngpd

using namespace ngpd;

extern "C" __attribute__((visibility("default"))) int ngpd_mesh_vertex_update(const double* v, int64_t nv, const int32_t* faces, const double* face_normals,
                                                                          const int32_t* vta_faces, const int32_t* vta_offsets, int iterations,
                                                                          double* scratch, double* v_out, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    NGPD_REQUIRE(v && faces && face_normals && vta_faces && vta_offsets && v_out, "ngpd_mesh_vertex_update: NULL argument");
    NGPD_REQUIRE(iterations >= 0 && (iterations <= 1 || scratch), "ngpd_mesh_vertex_update: more than one sweep needs a scratch buffer of nv*3 doubles");
    if (nv <= 0) return 0;
    if (iterations == 0) {
        if (v_out != v) NGPD_CUDA_OK(cudaMemcpyAsync(v_out, v, (size_t)nv * 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    // ping-pong so that the last sweep lands in v_out
    const double* src = v;
    for (int it = 0; it < iterations; ++it) {
        double* dst = ((iterations - 1 - it) % 2 == 0) ? v_out : scratch;
        mesh_vertex_update_kernel<<<(unsigned)cdiv(nv, 128), 128, 0, st>>>(src, faces, face_normals, vta_faces, vta_offsets, nv, dst);
        src = dst;
    }
    NGPD_CUDA_OK(cudaGetLastError());
    return 0;
}
