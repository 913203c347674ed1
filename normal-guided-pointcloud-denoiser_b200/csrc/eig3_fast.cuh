// Labels of the SECOND tensor pass without the LAPACK-order eigensolver.
//
// What stage 2 consumes (Decompositionor.py:57-69, Processor.py:134): the argmax of three ratios of the eigenvalues, and, for
// the rows that edge_step will move, the eigenvector of the smallest eigenvalue.  The label does not depend on LAPACK's
// eigenvector conventions -- unlike the smoothing of stage 1, which is why eig3.cuh exists -- so for every row that does not
// need the vector the QL sweeps (~900 of the kernel's ~1 900 warp-instructions per 32 rows, profiles/r1m_instruction_mix.md)
// are replaced by:
//   * eigenvalues in closed form (trigonometric solution of the characteristic cubic of K = T - tr/3 I).  The closed form is
//     unstable in fp32 (cancellation in p^3 - q^2 near a double root), so the invariants are accumulated in fp64 (about 25 fused
//     multiply-adds); square roots and the angle functions act on well-conditioned quantities and stay fp32.
//     Absolute error of the eigenvalues <= ~1e-6 for the tensors of this path (unit trace).
//   * the label from g0 = scale (l1 - l2), g1 = l2 - l3, g2 = l3 (the reference's ratios times l1 > 0).  The reference decides
//     on fp32 ratios of LAPACK's fp32 eigenvalues, themselves a few 1e-7 off: a decision is CERTAIN here only when the winner
//     leads by more than NGPD_FAST_LABEL_MARGIN; every other row (about 1 in 10^4, and all degenerate input) is handed to the
//     LAPACK-order path by the caller, so the labels stay bit-identical to the reference's
//     (tests/test_hostmath.py::test_fast_labels_agree_with_lapack_order, test_gpu_parity.py teacher-forced label tests).
// Rows moved by edge_step keep the LAPACK-order path: their 3x3 solve can be near-singular and amplifies a 1e-6 rad change of
// the crease direction (measured with a cross-product eigenvector: one fandisk row off by 3e-4 of the extent).
#pragma once
#include "common.cuh"

namespace ngpd {

#define NGPD_FAST_LABEL_MARGIN 2.5e-5f

struct FastLabel {
    int label;       // 0 flat, 1 edge, 2 corner
    bool certain;    // false: decide with eigh3_lapack + classify()
    float l3;        // smallest eigenvalue
};

NGPD_HD FastLabel classify_fast(float xx, float xy, float xz, float yy, float yz, float zz, float scale) {
    FastLabel o;
    o.label = 0; o.certain = false; o.l3 = 0.0f;
    const double dxx = xx, dxy = xy, dxz = xz, dyy = yy, dyz = yz, dzz = zz;
    const double m = (dxx + dyy + dzz) * (1.0 / 3.0);
    const double a = dxx - m, b = dyy - m, c = dzz - m;
    const double off = fma(dxy, dxy, fma(dxz, dxz, dyz * dyz));
    const double p2 = fma(a, a, fma(b, b, fma(c, c, 2.0 * off))) * (1.0 / 6.0);            // p^2 = |K|_F^2 / 6
    // q = det(K) / 2
    const double q = 0.5 * fma(a, fma(b, c, -dyz * dyz), fma(-dxy, fma(dxy, c, -dyz * dxz), dxz * fma(dxy, dyz, -b * dxz)));
    const double p6 = p2 * p2 * p2;
    const double disc = p6 - q * q;                                                           // >= 0 up to rounding
    const float p2f = (float)p2;
    if (!(p2f > 1e-12f) || !(p2f < 1e12f)) return o;                                          // (nearly) isotropic, empty or non-finite: not here
    const float s = sqrtf(fmaxf((float)disc, 0.0f));
    const float ang = atan2f(s, (float)q) * (1.0f / 3.0f);                                    // in [0, pi/3]
    float sn, cs;
#if defined(__CUDA_ARCH__)
    sincosf(ang, &sn, &cs);
#else
    sn = sinf(ang); cs = cosf(ang);
#endif
    const float p = sqrtf(p2f), mf = (float)m;
    const float r3 = 1.7320508075688772f * sn;
    const float l1 = fmaf(2.0f * p, cs, mf);
    const float l3 = fmaf(-p, cs + r3, mf);
    const float l2 = fmaf(-p, cs - r3, mf);
    if (!(l1 > 0.0f)) return o;
    const float g0 = scale * (l1 - l2), g1 = l2 - l3, g2 = l3;
    // argmax with the lead over the runner-up
    int lab = 0;
    float best = g0, second = fmaxf(g1, g2);
    if (g1 > best) { lab = 1; best = g1; second = fmaxf(g0, g2); }
    if (g2 > best) { lab = 2; best = g2; second = fmaxf(g0, g1); }
    o.label = lab;
    o.l3 = l3;
    o.certain = (best - second) > NGPD_FAST_LABEL_MARGIN * fmaxf(l1, 1.0f);
    return o;
}

}  // namespace ngpd
