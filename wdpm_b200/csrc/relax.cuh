// The 3x3 relax step: one wet centre pushes water to its lower neighbours.
//
// Semantics follow the reference's OpenCL kernels bit for bit on finite inputs:
//   ADD       src/runoff.cl:24-55   (runoffadd)
//   SUBTRACT  src/runoff.cl:57-88   (runoffsubtract)
//   DRAIN     src/runoff.cl:90-134  (runoffdrain)
// with the centre guard of src/runoff.cl:145 / :160 / :177-179.
//
// MASKED ELEVATIONS. The solver stores the DEM with every invalid cell
// (dem <= nodata, i.e. NODATA cells, the halo ring and the device margins; also
// NaN, infinite or absurdly large inputs) replaced by a sentinel S = the largest
// power of two of the format. A neighbour at S has surface sn = S, so the height
// difference h = sc - sn is -S (finite, hugely negative) and the `h > 0` test fails
// exactly where the reference's `bigdem[n] > missingvalue` guard (runoff.cl:33) would
// have skipped the neighbour: the validity compare and its branch disappear
// from the inner loop. A centre is valid iff its elevation is below S.
//
// The step is written branch-free (selects / predicated adds) against
// register-resident values: the eight neighbour steps of a tile form one
// dependent chain, and anything that lengthens it costs directly.
//
// The functions are __host__ __device__ so the schedule emulator in tests/ can
// run the very same arithmetic on the CPU; the product never calls them there.
//
// Arithmetic notes (each keeps the reference's result bit-identical):
//  * x/8.0 is computed as x*0.125: scaling by a power of two is exact, and in the
//    subnormal range both round the same exact quotient. Selecting the operand
//    first ((dc>sn ? wc : h) * 0.125) is the same operation on the same value.
//  * runoffadd's maxi(flow,0) and maxi(w-flow,0) are dropped: there flow is
//    w/8 or h/8 with w>0, h>0, so 0 <= flow, and mini(flow,w) <= w makes
//    w-flow >= +0 exactly. runoffdrain's four-term flow can be <= 0, so its
//    maxi(flow,0) stays; its second clamp is a no-op for the same reason.
//  * mini(a,b) is written as the reference writes it, (a<=b)?a:b.
//  * compiled with -fmad=false; there is no multiply-add to contract anyway.
#pragma once

#ifdef __CUDACC__
#define WDPM_HD __host__ __device__ __forceinline__
#else
#define WDPM_HD inline
#include <cmath>
#endif

#include <limits>

namespace wdpm {

enum : int { kAdd = 0, kSubtract = 1, kDrain = 2 };

// What the solver stores for an invalid cell: the largest power of two of the format (2^1023, 2^127).
// Finite on purpose: surfaces and height differences against it stay finite (-2^1023 etc.), so a
// closed gate can be expressed as a multiplication by zero (push_add_fast) without producing NaN.
template <typename T>
WDPM_HD T invalid_elevation() {
#ifdef __CUDA_ARCH__
    return sizeof(T) == 8 ? (T)__longlong_as_double(0x7fe0000000000000LL) : (T)__int_as_float(0x7f000000);
#else
    return sizeof(T) == 8 ? (T)8.98846567431158e307 : (T)1.7014118346046923e38f;
#endif
}
template <typename T>
WDPM_HD bool is_valid_elevation(T d) { return d < invalid_elevation<T>(); }
// what the upload path stores for a raw elevation
template <typename T>
WDPM_HD T mask_elevation(T d, T nodata) { return (d > nodata && d < invalid_elevation<T>()) ? d : invalid_elevation<T>(); }

// A 3x3 tile held in registers. Neighbour order: row offset outer, column
// offset inner (src/runoff.cl:28-30): 0 1 2 / 3 c 4 / 5 6 7.
template <typename T>
struct Tile {
    T wc, dc;
    T wn[8], dn[8];
    bool active;  // centre wet and valid: the reference would have called runoff*()
};

template <typename T>
WDPM_HD void tile_load(Tile<T>& t, const T* w0, const T* w1, const T* w2, const T* d0, const T* d1, const T* d2, int j) {
    t.wc = w1[j];
    t.dc = d1[j];
    t.active = (t.wc > T(0)) && is_valid_elevation(t.dc);
    t.wn[0] = w0[j - 1]; t.wn[1] = w0[j]; t.wn[2] = w0[j + 1];
    t.wn[3] = w1[j - 1]; t.wn[4] = w1[j + 1];
    t.wn[5] = w2[j - 1]; t.wn[6] = w2[j]; t.wn[7] = w2[j + 1];
    t.dn[0] = d0[j - 1]; t.dn[1] = d0[j]; t.dn[2] = d0[j + 1];
    t.dn[3] = d1[j - 1]; t.dn[4] = d1[j + 1];
    t.dn[5] = d2[j - 1]; t.dn[6] = d2[j]; t.dn[7] = d2[j + 1];
}

template <typename T>
WDPM_HD void tile_store(const Tile<T>& t, T* w0, T* w1, T* w2, int j) {
    w0[j - 1] = t.wn[0]; w0[j] = t.wn[1]; w0[j + 1] = t.wn[2];
    w1[j - 1] = t.wn[3]; w1[j] = t.wc;    w1[j + 1] = t.wn[4];
    w2[j - 1] = t.wn[5]; w2[j] = t.wn[6]; w2[j + 1] = t.wn[7];
}

// mini(a, b) = (a <= b) ? a : b (runoff.cl:14-22). For fp32 on the device the single-instruction
// fminf is used instead: it differs only for NaN operands and for zeros of opposite sign, neither
// of which can reach this call (b = the centre's water is never -0; a = -0 would need a four-term
// flow of (-0)+(-0), and x-x is +0).
template <typename T>
WDPM_HD T mini(T a, T b) {
#ifdef __CUDA_ARCH__
    if (sizeof(T) == 4) return (T)fminf((float)a, (float)b);
#endif
    return (a <= b) ? a : b;
}

// if (take) { wc -= flow; wn += flow; }
// fp32, device: two predicated round-to-nearest adds in PTX (left to itself the compiler turns the
// `if` into add + select per operand). fp64: ptxas turns predicated DADDs back into selects, so the
// cheapest form is ONE select of the amount moved - when nothing moves the amount is -0.0, the
// additive identity that preserves even a negative zero in wn (x + -0.0 == x bit for bit) and
// leaves wc unchanged (wc is never negative nor -0 for an active centre).
template <typename T>
WDPM_HD void move_if(bool take, T& wc, T& wn, T flow) {
#ifdef __CUDA_ARCH__
    if (sizeof(T) == 4) {
        float c = (float)wc, n = (float)wn;
        asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p sub.rn.f32 %0, %0, %3;\n\t@p add.rn.f32 %1, %1, %3;\n\t}"
            : "+f"(c), "+f"(n) : "r"((int)take), "f"((float)flow));
        wc = (T)c; wn = (T)n;
        return;
    }
#endif
    const T give = take ? flow : T(-0.0);
    wc = wc - give;
    wn = wn + give;
}

// One neighbour step, branch-free. wc is the centre's running water.
//
// The reference does  flow = mini(flow, wc); wc -= flow; wn += flow  under `if (h > 0)`; here
// the flow is computed unconditionally and only the two updates depend on h > 0 (move_if).
template <typename T, int MODULE>
WDPM_HD void push(T dc, T& wc, T dn, T& wn) {
    const T sn = dn + wn;
    const T sc = dc + wc;
    const T h = sc - sn;
    const bool pos = h > T(0);
    T x;
    if (MODULE == kAdd) x = (dc > sn) ? wc : h;
    else x = (dc > sn) ? wc : ((dc - dn) + (wc - wn));
    T flow = x * T(0.125);
    if (MODULE == kDrain) flow = (flow <= T(0)) ? T(0) : flow;
    flow = mini(flow, wc);  // runoff.cl:46
    move_if(pos, wc, wn, flow);
}

// ---------------------------------------------------------------------------------------------
// Add module: the same step with fewer instructions on the FP64 pipe (the pipe that bounds the
// fused kernel while the tile chains run). Two rewrites, both value-preserving:
//
//  (1) NO CAP. In runoffadd, mini(flow, wc) (runoff.cl:46) never changes flow. For dc > sn,
//      flow = wc/8 <= wc. For dc <= sn: sc = fl(dc + wc) is the floating-point number nearest to
//      dc + wc, and dc itself is a floating-point number at distance wc from it, so
//      |sc - (dc + wc)| <= wc, i.e. sc <= dc + 2 wc. With sn >= dc, sc - sn <= 2 wc, and rounding
//      is monotonic, so h = fl(sc - sn) <= 2 wc and flow = fl(h/8) <= fl(wc/4) <= wc. (Holds in
//      any binary format with round-to-nearest, for every wc >= 0; tests/test_relax_rewrites.py
//      hammers it with adversarial chains.) The four-term flow of runoffsubtract / runoffdrain has
//      no such bound (it can exceed a tiny wc next to a deep neighbour), so those keep the cap.
//
//  (2) SIGN GATE. With x = (dc > sn) ? wc : h the reference's `if (h > 0)` is equivalent to "x is
//      not negative": dc > sn implies sc = fl(dc + wc) >= dc > sn (wc >= 0), hence h > 0; and when
//      dc <= sn, x = h. So the gate is read off x's sign bit with an integer compare (ALU pipe)
//      instead of a DSETP. The one value the two tests disagree on is h == +0: the sign gate lets
//      flow = +0 through, which changes nothing (wc - 0 = wc, wn + 0 = wn) unless wn is -0.0,
//      which then becomes +0.0. Water is never -0.0 unless a water FILE holds a negative zero and
//      the zero threshold is 0 (and Add rewrites every valid cell first, WDPMCL.c:778-792), so
//      the Add module cannot observe it.
template <typename T>
WDPM_HD bool sign_clear(T x) {
#ifdef __CUDA_ARCH__
    if (sizeof(T) == 8) return __double2hiint((double)x) >= 0;
    return __float_as_int((float)x) >= 0;
#else
    return !std::signbit(x);
#endif
}

template <typename T>
WDPM_HD void push_add_fast(T dc, T& wc, T dn, T& wn) {
    const T sn = dn + wn;
    const T sc = dc + wc;
    const T h = sc - sn;
    const bool q = dc > sn;
    const T x = q ? wc : h;
#ifdef __CUDA_ARCH__
    if (sizeof(T) == 8) {
        // Gate folded into the scaling: x * (open ? 0.125 : +0.0). A closed gate has x = h < 0 and
        // FINITE (invalid elevations are a huge finite number, not infinity), so the product is
        // exactly -0.0, the additive identity. open = "h is not negative" (dc > sn implies h > 0,
        // see above), read off h's sign bit; only the high word of the factor is selected.
        int mhi;
        asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %1, 0;\n\tselp.b32 %0, 0x3fc00000, 0, p;\n\t}" : "=r"(mhi) : "r"(__double2hiint((double)h)));
        const T give = (T)((double)x * __hiloint2double(mhi, 0));
        wc = wc - give;
        wn = wn + give;
        return;
    }
#endif
    move_if(sign_clear(h), wc, wn, x * T(0.125));  // dc > sn implies h > 0: the gate can be read off h
}

// fp32 Add in the warp-autonomous kernel: rewrite (1) only - no cap - with the reference's own gate `h > 0` and the
// predicated adds of move_if (in fp32 the sign-gated forms measured slower, profiles/micro_chain_throughput.txt):
// 9 instructions per neighbour instead of 10.
template <typename T>
WDPM_HD void push_add_nocap(T dc, T& wc, T dn, T& wn) {
    const T sn = dn + wn;
    const T sc = dc + wc;
    const T h = sc - sn;
    const T x = (dc > sn) ? wc : h;
    move_if(h > T(0), wc, wn, x * T(0.125));
}

// Drain, fp64: runoffdrain's step (runoff.cl:112-127) with its gate `if (h > 0)` and its clamp
// maxi(flow, 0) folded into the scaling factor, as in push_add_fast: give = x * (h > 0 && x >= +0 ?
// 0.125 : +0.0). A closed gate or a clamped flow then moves -0.0 or +0.0, where the reference moves
// nothing or +0.0 - the same bits unless the neighbour's water is -0.0. The solver selects this form
// only when its zero threshold is > 0: the block prologue (WDPMCL.c:1055-1065) then turns any -0.0
// from a water file into +0.0 before the first iteration, and no step can produce one (x - x = +0,
// +0 + -0 = +0). The cap mini(flow, wc) stays: the four-term flow can exceed a tiny wc.
template <typename T>
WDPM_HD void push_drain_fast(T dc, T& wc, T dn, T& wn) {
    if (sizeof(T) != 8) {  // fp32 keeps the reference form (predicated adds)
        push<T, kDrain>(dc, wc, dn, wn);
        return;
    }
    const double sn = (double)dn + (double)wn;
    const double sc = (double)dc + (double)wc;
    const double h = sc - sn;
    const bool pos = h > 0.0;
    const double x = ((double)dc > sn) ? (double)wc : (((double)dc - (double)dn) + ((double)wc - (double)wn));
#ifdef __CUDA_ARCH__
    int mhi;  // factor = (pos && x's sign bit clear) ? 0.125 : +0.0; only its high word is selected
    asm("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %2, 0;\n\tsetp.ge.and.s32 p, %1, 0, q;\n\tselp.b32 %0, 0x3fc00000, 0, p;\n\t}"
        : "=r"(mhi) : "r"(__double2hiint(x)), "r"((int)pos));
    const double m = __hiloint2double(mhi, 0);
#else
    const double m = (pos && sign_clear(x)) ? 0.125 : 0.0;
#endif
    double flow = x * m;
    flow = (flow <= (double)wc) ? flow : (double)wc;  // mini, as the reference writes it
    wc = (T)((double)wc - flow);
    wn = (T)((double)wn + flow);
}

// The eight steps on a register window that "slides" one column per colour sub-pass: w and d hold the
// tile's water / elevations over ALL three positions (3x5); sub-pass COFS works on columns
// COFS..COFS+2. Indices are compile-time constants, so the slide is register renaming - no moves.
template <typename T, int MODULE, int COFS, bool FAST>
WDPM_HD void relax_window5(T (&w)[3][5], const T (&d)[3][5]) {
    const T dc = d[1][COFS + 1];
    T wc = w[1][COFS + 1];
#define WDPM_PUSH5(r, c)                                                                     \
    do {                                                                                     \
        if (FAST && MODULE == kAdd && sizeof(T) == 8) push_add_fast<T>(dc, wc, d[r][COFS + c], w[r][COFS + c]); \
        else if (FAST && MODULE == kDrain && sizeof(T) == 8) push_drain_fast<T>(dc, wc, d[r][COFS + c], w[r][COFS + c]); \
        else push<T, MODULE>(dc, wc, d[r][COFS + c], w[r][COFS + c]);                        \
    } while (0)
    WDPM_PUSH5(0, 0); WDPM_PUSH5(0, 1); WDPM_PUSH5(0, 2);
    WDPM_PUSH5(1, 0); WDPM_PUSH5(1, 2);
    WDPM_PUSH5(2, 0); WDPM_PUSH5(2, 1); WDPM_PUSH5(2, 2);
#undef WDPM_PUSH5
    w[1][COFS + 1] = wc;
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous kernel (kernels.cuh, k_fused_wa): a lane owns TWO column-adjacent tiles and keeps a
// 3 x 8 register window (water and elevations) that slides one column per colour sub-pass; in sub-pass
// C the tiles are columns C..C+2 and C+3..C+5. The two tiles of a sub-pass are disjoint, so their
// eight-step chains are independent and are written interleaved (instruction-level parallelism: the
// FP64 latency of one chain hides behind the other).
//
// No branch around an inactive centre (it would serialise the two chains): a centre the reference
// would not relax (dry, or not a valid cell: runoff.cl:145) runs the same eight steps on the
// substitutes dc := -S/2, wc := +0. Then sc = -S/2, h = sc - sn < 0 for every neighbour (surfaces are
// > -S/2 and finite), every gate is closed, every amount moved is -0.0 (or nothing, for the predicated
// forms): an exact no-op on all nine cells; the centre's own value is put back afterwards.
template <typename T>
WDPM_HD T inactive_elevation() { return T(-0.5) * invalid_elevation<T>(); }

template <typename T, int MODULE, bool FAST>
WDPM_HD void push_sel(T dc, T& wc, T dn, T& wn) {
    if (FAST && MODULE == kAdd && sizeof(T) == 8) push_add_fast<T>(dc, wc, dn, wn);
    else if (FAST && MODULE == kAdd) push_add_nocap<T>(dc, wc, dn, wn);
    else if (FAST && MODULE == kDrain && sizeof(T) == 8) push_drain_fast<T>(dc, wc, dn, wn);
    else push<T, MODULE>(dc, wc, dn, wn);
}

// GUARD = false: the caller guarantees that every centre may run unguarded, i.e. water is +0 wherever the
// reference would skip the centre (no water on invalid cells, no negative water, no -0.0). A skipped centre then
// has wc = +0, and with wc = +0 every step moves +0 or nothing: an open gate h > 0 means dc = sc > sn, so x = wc = 0;
// an invalid centre (dc = S) likewise has x = wc = 0. Adding +0 to water that is never -0.0 changes nothing.
// Add only, with the cap-free steps (solver.cu decides when).
// PART: 0 = all eight neighbour steps; 1 = the first four, 2 = the last four (the staggered schedule cuts the
// middle sub-pass of a step in two, kernels.cuh). Between the halves the centre's running water lives in the
// window like any other cell, and the second half decides again whether the centre is active: a centre that ran
// dry in the first half has wc = +0, with which the remaining steps move nothing in the reference either.
template <typename T, int MODULE, int C, bool FAST, bool GUARD, int PART = 0>
WDPM_HD void wa_relax_pair(T (&w)[3][8], const T (&d)[3][8]) {
    T dcA = d[1][C + 1], dcB = d[1][C + 4];
    T wcA = w[1][C + 1], wcB = w[1][C + 4];
    const T keepA = wcA, keepB = wcB;
    bool actA = true, actB = true;
    if (GUARD) {
        actA = (wcA > T(0)) && is_valid_elevation(dcA);
        actB = (wcB > T(0)) && is_valid_elevation(dcB);
        dcA = actA ? dcA : inactive_elevation<T>();
        dcB = actB ? dcB : inactive_elevation<T>();
        wcA = actA ? wcA : T(0);
        wcB = actB ? wcB : T(0);
    }
#define WDPM_PUSH2(r, c)                                                  \
    do {                                                                  \
        push_sel<T, MODULE, FAST>(dcA, wcA, d[r][C + c], w[r][C + c]);          \
        push_sel<T, MODULE, FAST>(dcB, wcB, d[r][C + 3 + c], w[r][C + 3 + c]);  \
    } while (0)
    if (PART != 2) {
        WDPM_PUSH2(0, 0); WDPM_PUSH2(0, 1); WDPM_PUSH2(0, 2);
        WDPM_PUSH2(1, 0);
    }
    if (PART != 1) {
        WDPM_PUSH2(1, 2);
        WDPM_PUSH2(2, 0); WDPM_PUSH2(2, 1); WDPM_PUSH2(2, 2);
    }
#undef WDPM_PUSH2
    w[1][C + 1] = (!GUARD || actA) ? wcA : keepA;
    w[1][C + 4] = (!GUARD || actB) ? wcB : keepB;
}

// The eight neighbour steps of one tile (only meaningful when t.active). FAST selects the fp64 Add
// rewrite (push_add_fast; in fp32 the predicated form of move_if measured faster); the plain form is what the colour kernel - the
// second, independent CUDA path - and the CPU checks run.
template <typename T, int MODULE, bool FAST = false>
WDPM_HD void tile_relax(Tile<T>& t) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int n = 0; n < 8; n++) {
        if (FAST && MODULE == kAdd && sizeof(T) == 8) push_add_fast<T>(t.dc, t.wc, t.dn[n], t.wn[n]);
        else push<T, MODULE>(t.dc, t.wc, t.dn[n], t.wn[n]);
    }
}

// Convenience: load, relax if active, store. Returns whether work was done.
template <typename T, int MODULE, bool FAST = false>
WDPM_HD bool relax_tile(T* w0, T* w1, T* w2, const T* d0, const T* d1, const T* d2, int j) {
    Tile<T> t;
    tile_load(t, w0, w1, w2, d0, d1, d2, j);
    if (!t.active) return false;
    tile_relax<T, MODULE, FAST>(t);
    tile_store(t, w0, w1, w2, j);
    return true;
}

// Drain outlets. The reference has exactly one outlet (WDPMCL.c:1005-1017); this solver takes a SET
// of outlet cells (BASELINE configs[4]; one outlet reproduces the reference). The solver marks an
// outlet cell in its elevation grid with -S (S = invalid_elevation): runoffdrain never reads an
// outlet's elevation - the outlet test precedes the height test (runoff.cl:104-111) and an outlet
// is never a centre (runoff.cl:179) - so the mark costs no information, keeps the cell "valid" for
// the reductions, and lets every kernel recognise an outlet from the data it already holds.
template <typename T>
WDPM_HD T outlet_mark() { return -invalid_elevation<T>(); }
template <typename T>
WDPM_HD bool is_outlet(T d) { return d == outlet_mark<T>(); }

// Which cells of the 3x3 around (w1, d1)[j] are outlets: bit (a+1)*3 + (b+1) for row offset a,
// column offset b (bit 4 = the centre).
template <typename T>
WDPM_HD int outlet_mask_3x3(const T* d0, const T* d1, const T* d2, int j) {
    int m = 0;
    const T* dr[3] = {d0, d1, d2};
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++)
            if (is_outlet(dr[a][j + b - 1])) m |= 1 << (a * 3 + b);
    return m;
}

// Drain tile with outlets among its cells (rare). `mask` = outlet_mask_3x3. The outlet test precedes
// the height test (src/runoff.cl:104-111):
//   totaldrain = totaldrain + w[outlet] + w[centre]; both set to 0; the walk goes on.
// Each contact's two addends are returned so the caller can fold them into that outlet's total in
// sub-pass order: ev_outlet[i] = w[outlet], ev_centre[i] = w[centre] at that moment, ev_pos[i] = the
// outlet's bit position in the 3x3. Returns the number of contacts (<= 8).
// Kept out of line on the device: inlining its pointer tables into the hot loop costs every thread
// registers (and spills).
#ifdef __CUDACC__
#define WDPM_RARE __host__ __device__ __noinline__
#else
#define WDPM_RARE inline
#endif
template <typename T>
WDPM_RARE int relax_tile_near_outlets(T* w0, T* w1, T* w2, const T* d0, const T* d1, const T* d2, int j,
                                      int mask, T* ev_outlet, T* ev_centre, int* ev_pos) {
    if (mask & 16) return 0;  // the centre is an outlet: never a centre (src/runoff.cl:179)
    T wc = w1[j];
    if (!(wc > T(0))) return 0;
    const T dc = d1[j];
    if (!is_valid_elevation(dc)) return 0;
    T* wr[3] = {w0, w1, w2};
    const T* dr[3] = {d0, d1, d2};
    int n = 0;
    for (int a = 0; a < 3; a++) {
        for (int b = 0; b < 3; b++) {
            if (a == 1 && b == 1) continue;
            const T dn = dr[a][j + b - 1];
            if (!is_valid_elevation(dn)) continue;
            T wn = wr[a][j + b - 1];
            if (mask & (1 << (a * 3 + b))) {
                ev_outlet[n] = wn;
                ev_centre[n] = wc;
                ev_pos[n] = a * 3 + b;
                n++;
                wn = T(0);
                wc = T(0);
            } else {
                push<T, kDrain>(dc, wc, dn, wn);
            }
            wr[a][j + b - 1] = wn;
        }
    }
    w1[j] = wc;
    return n;
}

// Warp-autonomous kernel, Drain, a lane whose 3 x 8 elevation window holds outlet marks (rare): the two tiles of
// sub-pass C relaxed one after the other with the outlet rule of runoffdrain (src/runoff.cl:104-111: the outlet
// test precedes the height test; both cells are emptied into the outlet's total and the walk goes on) and the
// centre guard of src/runoff.cl:177-179 (an outlet is never a centre). Every contact is handed to
// on_contact(tile, row offset, column offset, w_outlet, w_centre) in walk order. Fully unrolled: the window must stay
// in registers, so every index is a compile-time constant.
template <typename T, int C, bool FAST, typename F>
WDPM_HD void wa_relax_pair_outlets(T (&w)[3][8], const T (&d)[3][8], F&& on_contact) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int tile = 0; tile < 2; tile++) {
        const int c0 = C + 3 * tile;  // first column of the tile; the centre is (1, c0 + 1)
        const T dc = d[1][c0 + 1];
        T wc = w[1][c0 + 1];
        const bool act = !is_outlet(dc) && (wc > T(0)) && is_valid_elevation(dc);
        if (act) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int a = 0; a < 3; a++) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
                for (int b = 0; b < 3; b++) {
                    if (a == 1 && b == 1) continue;
                    const T dn = d[a][c0 + b];
                    if (is_outlet(dn)) {
                        on_contact(tile, a - 1, b - 1, w[a][c0 + b], wc);
                        w[a][c0 + b] = T(0);
                        wc = T(0);
                    } else if (is_valid_elevation(dn)) {
                        push_sel<T, kDrain, FAST>(dc, wc, dn, w[a][c0 + b]);
                    }
                }
            }
            w[1][c0 + 1] = wc;
        }
    }
}

}  // namespace wdpm
