// Sorting networks on small register arrays of 32-bit keys.  Every index is a template parameter, so an array element
// is always one fixed register: loops with `#pragma unroll` are not enough -- when ptxas leaves one level of Batcher's
// loop nest rolled, the "array" is indexed dynamically and each access turns into a chain of predicated moves
// (ncu on the re-ranking kernel: 1 700 of its 3 000 instructions per warp were spent in 206 compare-exchanges).
// Host-compilable (tests/hostmath checks the networks with the 0-1 principle).
#pragma once
#include "common.cuh"

namespace ngpd {

NGPD_HD void ks_ce(unsigned& a, unsigned& b) {
#if defined(__CUDA_ARCH__)
    const unsigned lo = min(a, b), hi = max(a, b);
#else
    const unsigned lo = a < b ? a : b, hi = a < b ? b : a;
#endif
    a = lo; b = hi;
}

namespace sortnet {
// Batcher's odd-even merge sort on a[LO..HI] (inclusive), HI - LO + 1 a power of two
template <int I, int END, int R, int STEP, int N>
NGPD_HD void oem_row(unsigned (&a)[N]) {
    if constexpr (I < END) { ks_ce(a[I], a[I + R]); oem_row<I + STEP, END, R, STEP, N>(a); }
}
template <int LO, int HI, int R, int N>
NGPD_HD void oem_merge(unsigned (&a)[N]) {
    constexpr int STEP = 2 * R;
    if constexpr (STEP < HI - LO) {
        oem_merge<LO, HI, STEP, N>(a);
        oem_merge<LO + R, HI, STEP, N>(a);
        oem_row<LO + R, HI - R, R, STEP, N>(a);
    } else {
        ks_ce(a[LO], a[LO + R]);
    }
}
template <int LO, int HI, int N>
NGPD_HD void oem_sort(unsigned (&a)[N]) {
    if constexpr (HI > LO) {
        constexpr int MID = LO + (HI - LO) / 2;
        oem_sort<LO, MID, N>(a);
        oem_sort<MID + 1, HI, N>(a);
        oem_merge<LO, HI, 1, N>(a);
    }
}
// ascending sort of a bitonic sequence a[LO .. LO+LEN)
template <int I, int END, int H, int N>
NGPD_HD void bitonic_row(unsigned (&a)[N]) {
    if constexpr (I < END) { ks_ce(a[I], a[I + H]); bitonic_row<I + 1, END, H, N>(a); }
}
template <int LO, int LEN, int N>
NGPD_HD void bitonic_merge(unsigned (&a)[N]) {
    if constexpr (LEN > 1) {
        constexpr int H = LEN / 2;
        bitonic_row<LO, LO + H, H, N>(a);
        bitonic_merge<LO, H, N>(a);
        bitonic_merge<LO + H, H, N>(a);
    }
}
}  // namespace sortnet

// ascending sort of N = 2^m keys (Batcher's odd-even merge sort: 63 compare-exchanges for N = 16, 19 for N = 8)
template <int N>
NGPD_HD void ks_sort(unsigned (&a)[N]) {
    static_assert((N & (N - 1)) == 0, "power of two");
    sortnet::oem_sort<0, N - 1, N>(a);
}
// ascending sort of a bitonic sequence of N = 2^m keys
template <int N>
NGPD_HD void ks_bitonic_merge(unsigned (&a)[N]) {
    static_assert((N & (N - 1)) == 0, "power of two");
    sortnet::bitonic_merge<0, N, N>(a);
}

}  // namespace ngpd
