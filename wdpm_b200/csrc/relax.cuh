// The 3x3 relax step: one wet centre pushes water to its lower neighbours.
//
// Semantics follow the reference's OpenCL kernels bit for bit on finite inputs:
//   ADD       src/runoff.cl:24-55   (runoffadd)
//   SUBTRACT  src/runoff.cl:57-88   (runoffsubtract)
//   DRAIN     src/runoff.cl:90-134  (runoffdrain)
// with the centre guard of src/runoff.cl:145 / :160 / :177-179.
//
// The functions are written against three water rows and three elevation rows
// (any address space: global memory for the colour kernel, the shared-memory row
// ring for the fused kernel) and a centre column index. They are __host__
// __device__ so the schedule emulator in tests/ can run the very same arithmetic
// on the CPU; the product never calls them on the host.
//
// Arithmetic notes (each keeps the reference's result bit-identical):
//  * x/8.0 is computed as x*0.125: scaling by a power of two is exact, and in the
//    subnormal range both round the same exact quotient.
//  * runoffadd's maxi(flow,0) and maxi(w-flow,0) are dropped: there flow is
//    w/8 or h/8 with w>0, h>0, so 0 <= flow, and mini(flow,w) <= w makes
//    w-flow >= +0 exactly. runoffdrain's four-term flow can be <= 0, so its
//    maxi(flow,0) stays (as a compare+select, which is what maxi is); its second
//    clamp is a no-op for the same reason as above.
//  * mini(a,b) = (a<=b)?a:b equals fmin(a,b) unless an operand is NaN or the
//    operands are zeros of opposite sign; neither arises from finite inputs
//    (see DESIGN.md "bit-exactness").
//  * the file is compiled with -fmad=false; there is no multiply-add to contract
//    anyway (only + - *0.125 min max compare).
#pragma once

#ifdef __CUDACC__
#define WDPM_HD __host__ __device__ __forceinline__
#else
#define WDPM_HD inline
#include <cmath>
#endif

namespace wdpm {

enum : int { kAdd = 0, kSubtract = 1, kDrain = 2 };

template <typename T>
WDPM_HD T min_finite(T a, T b) {
#ifdef __CUDA_ARCH__
    return fmin(a, b);
#else
    return (a <= b) ? a : b;
#endif
}

// One neighbour step. wc is the centre's running water, (dn, wn) the neighbour.
template <typename T, int MODULE>
WDPM_HD void push(T dc, T& wc, T dn, T& wn, T nodata) {
    if (dn > nodata) {
        const T sn = dn + wn;
        const T sc = dc + wc;
        const T h = sc - sn;
        if (h > T(0)) {
            T flow;
            if (MODULE == kAdd) {
                flow = (dc > sn) ? wc * T(0.125) : h * T(0.125);
                flow = min_finite(flow, wc);
            } else {
                flow = (dc > sn) ? wc * T(0.125) : ((dc - dn) + (wc - wn)) * T(0.125);
                if (MODULE == kDrain) flow = (flow <= T(0)) ? T(0) : flow;
                flow = min_finite(flow, wc);
            }
            wc = wc - flow;
            wn = wn + flow;
        }
    }
}

// Relax the tile centred at column j of rows (w0,w1,w2)/(d0,d1,d2).
// Returns true if the centre was wet and valid (work was done).
template <typename T, int MODULE>
WDPM_HD bool relax_tile(T* w0, T* w1, T* w2, const T* d0, const T* d1, const T* d2, int j, T nodata) {
    T wc = w1[j];
    if (!(wc > T(0))) return false;
    const T dc = d1[j];
    if (!(dc > nodata)) return false;

    T wn0 = w0[j - 1], wn1 = w0[j], wn2 = w0[j + 1];
    T wn3 = w1[j - 1], wn4 = w1[j + 1];
    T wn5 = w2[j - 1], wn6 = w2[j], wn7 = w2[j + 1];
    const T dn0 = d0[j - 1], dn1 = d0[j], dn2 = d0[j + 1];
    const T dn3 = d1[j - 1], dn4 = d1[j + 1];
    const T dn5 = d2[j - 1], dn6 = d2[j], dn7 = d2[j + 1];

    // neighbour order: row offset outer, column offset inner (src/runoff.cl:28-30)
    push<T, MODULE>(dc, wc, dn0, wn0, nodata);
    push<T, MODULE>(dc, wc, dn1, wn1, nodata);
    push<T, MODULE>(dc, wc, dn2, wn2, nodata);
    push<T, MODULE>(dc, wc, dn3, wn3, nodata);
    push<T, MODULE>(dc, wc, dn4, wn4, nodata);
    push<T, MODULE>(dc, wc, dn5, wn5, nodata);
    push<T, MODULE>(dc, wc, dn6, wn6, nodata);
    push<T, MODULE>(dc, wc, dn7, wn7, nodata);

    w0[j - 1] = wn0; w0[j] = wn1; w0[j + 1] = wn2;
    w1[j - 1] = wn3; w1[j] = wc;  w1[j + 1] = wn4;
    w2[j - 1] = wn5; w2[j] = wn6; w2[j + 1] = wn7;
    return true;
}

// Drain, centre adjacent to the outlet (rare: at most 8 centres per iteration).
// (orow, ocol) is the outlet's position relative to the centre, each in {-1,0,1},
// not both 0. The outlet test precedes the height test (src/runoff.cl:104-111):
//   totaldrain = totaldrain + w[outlet] + w[centre]; both set to 0; the walk goes on.
// The two addends are returned so the caller can fold them into totaldrain in
// sub-pass order: *ev_outlet = w[outlet], *ev_centre = w[centre] at that moment.
template <typename T>
WDPM_HD bool relax_tile_at_outlet(T* w0, T* w1, T* w2, const T* d0, const T* d1, const T* d2, int j,
                                  T nodata, int orow, int ocol, T* ev_outlet, T* ev_centre, bool* drained) {
    *drained = false;
    T wc = w1[j];
    if (!(wc > T(0))) return false;
    const T dc = d1[j];
    if (!(dc > nodata)) return false;
    T* wr[3] = {w0, w1, w2};
    const T* dr[3] = {d0, d1, d2};
    for (int a = -1; a <= 1; a++) {
        for (int b = -1; b <= 1; b++) {
            if (a == 0 && b == 0) continue;
            const T dn = dr[a + 1][j + b];
            if (!(dn > nodata)) continue;
            T wn = wr[a + 1][j + b];
            if (a == orow && b == ocol) {
                *ev_outlet = wn;
                *ev_centre = wc;
                *drained = true;
                wn = T(0);
                wc = T(0);
            } else {
                push<T, kDrain>(dc, wc, dn, wn, nodata);
            }
            wr[a + 1][j + b] = wn;
        }
    }
    w1[j] = wc;
    return true;
}

}  // namespace wdpm
