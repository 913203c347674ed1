// TEST-ONLY cross-check implementation (array-indexed transcription of the LAPACK routines for general loop
// bounds).  The product uses the statically-indexed n = 3 specialisation in csrc/eig3.cuh; tests/test_hostmath.py
// checks that both give bit-identical results.
// Symmetric 3x3 eigen-decomposition that follows LAPACK's SSYEVD code path for n = 3
// (what torch.linalg.eigh runs on CPU: Decompositionor.py:300, GraphBuilder.py:110):
//   SSYTD2('L') Householder tridiagonalisation -> SSTEQR('I') implicit QL/QR with Wilkinson
//   shifts -> back-multiplication by the reflector (SORMTR) -> ascending selection sort.
// The plane-rotation generator uses the classic (LAPACK <= 3.9) SLARTG sign convention; that is
// the convention MKL's eigenvectors follow (probed: 99.7 % of fandisk NVT tensors agree in all
// three column signs with torch+MKL, the rest are rank-deficient tensors).  The reference's
// normal smoothing (Decompositionor.py:92-106) depends on these signs, so a closed-form or
// Jacobi solver with an arbitrary sign rule would not reproduce the reference.
// Written from the published algorithm; fp32 throughout, no FMA contraction (--fmad=false).
#pragma once
#include "../../normal-guided-pointcloud-denoiser_b200/csrc/common.cuh"

namespace ngpd_generic {
using ngpd::V3;

struct Rot { float c, s, r; };

NGPD_HD float sign_of(float mag, float sgn) {  // Fortran SIGN(a,b)
    float a = fabsf(mag);
    return signbit(sgn) ? -a : a;
}

NGPD_HD float lapy2(float x, float y) {
    float xa = fabsf(x), ya = fabsf(y);
    float w = fmaxf(xa, ya), z = fminf(xa, ya);
    if (z == 0.0f) return w;
    float q = z / w;
    return w * sqrtf(1.0f + q * q);
}

NGPD_HD Rot lartg(float f, float g) {
    Rot o;
    if (g == 0.0f) { o.c = 1.0f; o.s = 0.0f; o.r = f; return o; }
    if (f == 0.0f) { o.c = 0.0f; o.s = 1.0f; o.r = g; return o; }
    // classic SLARTG safe scaling: safmn2 = 2^-51 (= base^int(log(safmin/eps)/log(base)/2) in fp32), safmx2 = 2^51
    const float safmn2 = 4.44089209850062616e-16f, safmx2 = 2251799813685248.0f;
    float f1 = f, g1 = g;
    float scale = fmaxf(fabsf(f1), fabsf(g1));
    float r;
    if (scale >= safmx2) {
        int count = 0;
        do { ++count; f1 *= safmn2; g1 *= safmn2; scale = fmaxf(fabsf(f1), fabsf(g1)); } while (scale >= safmx2 && count < 20);
        r = sqrtf(f1 * f1 + g1 * g1);
        o.c = f1 / r; o.s = g1 / r;
        for (int i = 0; i < count; ++i) r *= safmx2;
    } else if (scale <= safmn2) {
        int count = 0;
        do { ++count; f1 *= safmx2; g1 *= safmx2; scale = fmaxf(fabsf(f1), fabsf(g1)); } while (scale <= safmn2 && count < 20);
        r = sqrtf(f1 * f1 + g1 * g1);
        o.c = f1 / r; o.s = g1 / r;
        for (int i = 0; i < count; ++i) r *= safmn2;
    } else {
        r = sqrtf(f1 * f1 + g1 * g1);
        o.c = f1 / r; o.s = g1 / r;
    }
    o.r = r;
    if (fabsf(f) > fabsf(g) && o.c < 0.0f) { o.c = -o.c; o.s = -o.s; o.r = -o.r; }
    return o;
}

// eigen-system of [[a,b],[b,c]]: rt1 = eigenvalue of larger magnitude, (cs,sn) its unit eigenvector
NGPD_HD void laev2(float a, float b, float c, float& rt1, float& rt2, float& cs1, float& sn1) {
    float sm = a + c, df = a - c, adf = fabsf(df), tb = b + b, ab = fabsf(tb);
    float acmx, acmn;
    if (fabsf(a) > fabsf(c)) { acmx = a; acmn = c; } else { acmx = c; acmn = a; }
    float rt;
    if (adf > ab)      { float q = ab / adf; rt = adf * sqrtf(1.0f + q * q); }
    else if (adf < ab) { float q = adf / ab; rt = ab * sqrtf(1.0f + q * q); }
    else               rt = ab * sqrtf(2.0f);
    int sgn1, sgn2;
    if (sm < 0.0f)      { rt1 = 0.5f * (sm - rt); sgn1 = -1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
    else if (sm > 0.0f) { rt1 = 0.5f * (sm + rt); sgn1 = 1;  rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
    else                { rt1 = 0.5f * rt; rt2 = -0.5f * rt; sgn1 = 1; }
    float cs;
    if (df >= 0.0f) { cs = df + rt; sgn2 = 1; } else { cs = df - rt; sgn2 = -1; }
    if (fabsf(cs) > ab) {
        float ct = -tb / cs;
        sn1 = 1.0f / sqrtf(1.0f + ct * ct);
        cs1 = ct * sn1;
    } else if (ab == 0.0f) {
        cs1 = 1.0f; sn1 = 0.0f;
    } else {
        float tn = -cs / tb;
        cs1 = 1.0f / sqrtf(1.0f + tn * tn);
        sn1 = tn * cs1;
    }
    if (sgn1 == sgn2) { float tn = cs1; cs1 = -sn1; sn1 = tn; }
}

struct Tri3 {
    float d1, d2, d3, e1, e2;
    float z[9];  // row-major, columns are the accumulated eigenvectors
    NGPD_HD float d(int i) const { return i == 1 ? d1 : (i == 2 ? d2 : d3); }
    NGPD_HD float e(int i) const { return i == 1 ? e1 : e2; }
    NGPD_HD void setd(int i, float v) { if (i == 1) d1 = v; else if (i == 2) d2 = v; else d3 = v; }
    NGPD_HD void sete(int i, float v) { if (i == 1) e1 = v; else e2 = v; }
    // SLASR('R','V'): rotate columns j, j+1 (1-based j)
    NGPD_HD void rot(int j, float c, float s) {
        if (c == 1.0f && s == 0.0f) return;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float lo = (j == 1) ? z[r * 3 + 0] : z[r * 3 + 1];
            float hi = (j == 1) ? z[r * 3 + 1] : z[r * 3 + 2];
            float nhi = c * hi - s * lo;
            float nlo = s * hi + c * lo;
            if (j == 1) { z[r * 3 + 0] = nlo; z[r * 3 + 1] = nhi; }
            else        { z[r * 3 + 1] = nlo; z[r * 3 + 2] = nhi; }
        }
    }
};

// SSTEQR(COMPZ='I') for n = 3 (scaling branch omitted: callers pass tensors whose largest entry
// is O(1e-30 .. 1e30), far inside [ssfmin, ssfmax]).
NGPD_HD void steqr3(Tri3& t) {
    const float eps = 5.9604644775390625e-08f;   // SLAMCH('E') = 2^-24
    const float eps2 = eps * eps;
    const float safmin = 1.17549435e-38f;
    const int n = 3, nmaxit = 90;
    int jtot = 0, l1 = 1;
    float wc[4], ws[4];
    while (l1 <= n) {
        if (l1 > 1) t.sete(l1 - 1, 0.0f);
        int m = n;
        for (int mm = l1; mm <= n - 1; ++mm) {
            float tst = fabsf(t.e(mm));
            if (tst == 0.0f) { m = mm; break; }
            if (tst <= (sqrtf(fabsf(t.d(mm))) * sqrtf(fabsf(t.d(mm + 1)))) * eps) {
                t.sete(mm, 0.0f); m = mm; break;
            }
        }
        int l = l1, lend = m;
        const int lsv = l, lendsv = lend;
        l1 = m + 1;
        if (lend == l) continue;
        if (fabsf(t.d(lend)) < fabsf(t.d(l))) { lend = lsv; l = lendsv; }
        if (lend > l) {
            // QL iteration
            for (;;) {
                m = lend;
                if (l != lend) {
                    for (int mm = l; mm <= lend - 1; ++mm) {
                        float a = fabsf(t.e(mm));
                        float tst = a * a;
                        if (tst <= (eps2 * fabsf(t.d(mm))) * fabsf(t.d(mm + 1)) + safmin) { m = mm; break; }
                    }
                }
                if (m < lend) t.sete(m, 0.0f);
                float p = t.d(l);
                if (m == l) {
                    ++l;
                    if (l <= lend) continue;
                    break;
                }
                if (m == l + 1) {
                    float rt1, rt2, c, s;
                    laev2(t.d(l), t.e(l), t.d(l + 1), rt1, rt2, c, s);
                    t.rot(l, c, s);
                    t.setd(l, rt1); t.setd(l + 1, rt2); t.sete(l, 0.0f);
                    l += 2;
                    if (l <= lend) continue;
                    break;
                }
                if (jtot == nmaxit) break;
                ++jtot;
                float g = (t.d(l + 1) - p) / (2.0f * t.e(l));
                float r = lapy2(g, 1.0f);
                g = t.d(m) - p + (t.e(l) / (g + sign_of(r, g)));
                float s = 1.0f, c = 1.0f;
                p = 0.0f;
                for (int i = m - 1; i >= l; --i) {
                    float f = s * t.e(i), b = c * t.e(i);
                    Rot q = lartg(g, f);
                    c = q.c; s = q.s; r = q.r;
                    if (i != m - 1) t.sete(i + 1, r);
                    g = t.d(i + 1) - p;
                    r = (t.d(i) - g) * s + 2.0f * c * b;
                    p = s * r;
                    t.setd(i + 1, g + p);
                    g = c * r - b;
                    wc[i] = c; ws[i] = -s;
                }
                for (int j = m - 1; j >= l; --j) t.rot(j, wc[j], ws[j]);
                t.setd(l, t.d(l) - p);
                t.sete(l, g);
            }
        } else {
            // QR iteration
            for (;;) {
                m = lend;
                if (l != lend) {
                    for (int mm = l; mm >= lend + 1; --mm) {
                        float a = fabsf(t.e(mm - 1));
                        float tst = a * a;
                        if (tst <= (eps2 * fabsf(t.d(mm))) * fabsf(t.d(mm - 1)) + safmin) { m = mm; break; }
                    }
                }
                if (m > lend) t.sete(m - 1, 0.0f);
                float p = t.d(l);
                if (m == l) {
                    --l;
                    if (l >= lend) continue;
                    break;
                }
                if (m == l - 1) {
                    float rt1, rt2, c, s;
                    laev2(t.d(l - 1), t.e(l - 1), t.d(l), rt1, rt2, c, s);
                    t.rot(l - 1, c, s);
                    t.setd(l - 1, rt1); t.setd(l, rt2); t.sete(l - 1, 0.0f);
                    l -= 2;
                    if (l >= lend) continue;
                    break;
                }
                if (jtot == nmaxit) break;
                ++jtot;
                float g = (t.d(l - 1) - p) / (2.0f * t.e(l - 1));
                float r = lapy2(g, 1.0f);
                g = t.d(m) - p + (t.e(l - 1) / (g + sign_of(r, g)));
                float s = 1.0f, c = 1.0f;
                p = 0.0f;
                for (int i = m; i <= l - 1; ++i) {
                    float f = s * t.e(i), b = c * t.e(i);
                    Rot q = lartg(g, f);
                    c = q.c; s = q.s; r = q.r;
                    if (i != m) t.sete(i - 1, r);
                    g = t.d(i) - p;
                    r = (t.d(i + 1) - g) * s + 2.0f * c * b;
                    p = s * r;
                    t.setd(i, g + p);
                    g = c * r - b;
                    wc[i] = c; ws[i] = s;
                }
                for (int j = m; j <= l - 1; ++j) t.rot(j, wc[j], ws[j]);
                t.setd(l, t.d(l) - p);
                t.sete(l - 1, g);
            }
        }
        if (jtot >= nmaxit) break;  // LAPACK would report INFO > 0; keep the partial result
    }
}

// A given by its lower triangle.  w ascending; V row-major 3x3, V[r*3+c] = component r of eigenvector c.
NGPD_HD void eigh3_lapack(float a11, float a21, float a31, float a22, float a32, float a33,
                          float w[3], float V[9]) {
    Tri3 t;
    float tau = 0.0f, v2 = 0.0f;
    // SLARFG(2, a21, a31): reflector that annihilates a31
    float xnorm = fabsf(a31);
    if (xnorm == 0.0f) {
        t.e1 = a21;
    } else {
        float beta = -sign_of(lapy2(a21, xnorm), a21);
        tau = (beta - a21) / beta;
        v2 = a31 * (1.0f / (a21 - beta));
        t.e1 = beta;
        // two-sided update of the trailing 2x2 block (SSYMV, SAXPY, SSYR2)
        float w1 = tau * (a22 + a32 * v2);
        float w2 = tau * (a32 + a33 * v2);
        float alpha = -0.5f * tau * (w1 + w2 * v2);
        w1 = w1 + alpha;
        w2 = w2 + alpha * v2;
        a22 = a22 - w1 - w1;
        a32 = a32 - v2 * w1 - w2;
        a33 = a33 - v2 * w2 - w2 * v2;
    }
    t.d1 = a11; t.d2 = a22; t.d3 = a33; t.e2 = a32;
#pragma unroll
    for (int i = 0; i < 9; ++i) t.z[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    steqr3(t);
    // ascending selection sort with column swaps
    float d[3] = {t.d1, t.d2, t.d3};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int k = i;
        float p = d[i];
#pragma unroll
        for (int j = i + 1; j < 3; ++j)
            if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i]; d[i] = p;
#pragma unroll
            for (int r = 0; r < 3; ++r) { float tmp = t.z[r * 3 + i]; t.z[r * 3 + i] = t.z[r * 3 + k]; t.z[r * 3 + k] = tmp; }
        }
    }
    // Q * Z with Q = I - tau v v^T acting on rows 2..3, v = (1, v2)
    if (tau != 0.0f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float s = t.z[3 + c] + v2 * t.z[6 + c];
            t.z[3 + c] = t.z[3 + c] - tau * s;
            t.z[6 + c] = t.z[6 + c] - (tau * v2) * s;
        }
    }
    w[0] = d[0]; w[1] = d[1]; w[2] = d[2];
#pragma unroll
    for (int i = 0; i < 9; ++i) V[i] = t.z[i];
}

}  // namespace ngpd_generic
