// CPU check of the fp64 Add rewrites in wdpm_b200/csrc/relax.cuh (test infrastructure).
//
// Runs random and adversarial eight-neighbour chains through
//   (a) push<T, kAdd>     - the reference form (src/runoff.cl:24-55), cap included
//   (b) push_add_fast<T>  - no cap, sign gate
// and counts chains whose results differ in any bit, plus the steps on which the reference's cap
// mini(flow, wc) changed the flow (relax.cuh proves there are none). The sign of a zero in wn is
// compared too, so the generator never produces -0.0 water (the documented exception).
#include <cstdint>
#include <cstring>
#include <random>

#include "../../wdpm_b200/csrc/relax.cuh"

using namespace wdpm;

template <typename T>
static bool same_bits(T a, T b) { return std::memcmp(&a, &b, sizeof(T)) == 0; }

template <typename T>
static long long run(long long n, unsigned long long seed, long long* n_cap_bites) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    long long bad = 0;
    *n_cap_bites = 0;
    const T inf = std::numeric_limits<T>::infinity();
    const double emin = sizeof(T) == 8 ? -330.0 : -46.0;  // decimal exponent range of the water values
    for (long long it = 0; it < n; it++) {
        // elevation scale and water scale drawn over wide ranges, including flats and tiny water
        const int mode = (int)(rng() % 8);
        double base = 0.0, relief = 1.0;
        switch (mode) {
            case 0: base = 500.0; relief = 3.0; break;
            case 1: base = 500.0; relief = 1e-9; break;     // almost flat
            case 2: base = 0.0; relief = 1e-3; break;       // near zero, signs mixed
            case 3: base = -50.0; relief = 10.0; break;     // negative elevations
            case 4: base = 8000.0; relief = 0.0; break;     // exactly flat
            case 5: base = 1e-30; relief = 1e-30; break;    // absurdly small
            case 6: base = 500.0; relief = sizeof(T) == 8 ? 1e-13 : 1e-4; break;  // differences of a few ulps
            default: base = -512.0; relief = 1e-12; break;  // binade boundary
        }
        const T dc = (T)(base + relief * (U(rng) - 0.5));
        T dn[8], wn[8];
        T wc = (T)(std::pow(10.0, emin * U(rng) * (rng() % 3 == 0 ? 1.0 : 0.05)) * U(rng));
        if (rng() % 16 == 0) wc = (T)std::ldexp(U(rng), sizeof(T) == 8 ? -1060 : -140);  // subnormal
        if (rng() % 5 == 0) {  // water of the order of one ulp of the elevation: where the surface sum rounds
            const T a = dc < 0 ? -dc : dc;
            wc = (T)((std::nextafter(a, inf) - a) * (0.25 + 1.5 * U(rng)));
        }
        for (int k = 0; k < 8; k++) {
            dn[k] = (T)(base + relief * (U(rng) - 0.5));
            if (rng() % 9 == 0) dn[k] = inf;  // masked neighbour
            if (rng() % 7 == 0) dn[k] = dc;   // tie
            if (rng() % 7 == 0) dn[k] = std::nextafter(dc, rng() % 2 ? inf : -inf);
            wn[k] = (rng() % 4 == 0) ? (T)0 : (T)(std::pow(10.0, emin * U(rng) * (rng() % 3 == 0 ? 1.0 : 0.05)) * U(rng));
            if (rng() % 11 == 0) wn[k] = wc;
        }
        if (!(wc > (T)0)) continue;  // inactive centres never run the chain
        T a_c = wc, b_c = wc, a_n[8], b_n[8];
        for (int k = 0; k < 8; k++) {
            a_n[k] = b_n[k] = wn[k];
            {   // does the cap change anything on this step of the reference chain?
                const T sn = dn[k] + a_n[k], sc = dc + a_c, h = sc - sn;
                if (h > (T)0) { const T f = ((dc > sn) ? a_c : h) * (T)0.125; if (!(f <= a_c)) (*n_cap_bites)++; }
            }
            push<T, kAdd>(dc, a_c, dn[k], a_n[k]);
            push_add_fast<T>(dc, b_c, dn[k], b_n[k]);
        }
        bool ok = same_bits(a_c, b_c);
        for (int k = 0; k < 8; k++) ok = ok && same_bits(a_n[k], b_n[k]);
        if (!ok) bad++;
    }
    return bad;
}

// Drain: push<T, kDrain> (reference form) against push_drain_fast<T> (gate and max(flow,0) folded into the
// factor) on chains without -0.0 water - the condition under which the solver selects the fast form.
template <typename T>
static long long run_drain(long long n, unsigned long long seed) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    long long bad = 0;
    const T S = invalid_elevation<T>();
    const double emin = sizeof(T) == 8 ? -330.0 : -46.0;
    for (long long it = 0; it < n; it++) {
        const int mode = (int)(rng() % 6);
        double base = 500.0, relief = 3.0;
        if (mode == 1) relief = 1e-9;
        if (mode == 2) { base = 0.0; relief = 1e-3; }
        if (mode == 3) { base = -50.0; relief = 10.0; }
        if (mode == 4) relief = 0.0;
        if (mode == 5) relief = sizeof(T) == 8 ? 1e-13 : 1e-4;
        const T dc = (T)(base + relief * (U(rng) - 0.5));
        T wc = (T)(std::pow(10.0, emin * U(rng) * (rng() % 3 == 0 ? 1.0 : 0.05)) * U(rng));
        if (rng() % 5 == 0) { const T a = dc < 0 ? -dc : dc; wc = (T)((std::nextafter(a, S) - a) * (0.25 + 1.5 * U(rng))); }
        if (!(wc > (T)0)) continue;
        T a_c = wc, b_c = wc;
        bool ok = true;
        for (int k = 0; k < 8; k++) {
            T dn = (T)(base + relief * (U(rng) - 0.5));
            if (rng() % 9 == 0) dn = S;   // masked neighbour (the sentinel the solver stores)
            if (rng() % 7 == 0) dn = dc;
            if (rng() % 7 == 0) dn = std::nextafter(dc, rng() % 2 ? S : -S);
            T wn = (rng() % 4 == 0) ? (T)0 : (T)(std::pow(10.0, emin * U(rng) * (rng() % 3 == 0 ? 1.0 : 0.05)) * U(rng) * (rng() % 4 == 0 ? 1000.0 : 1.0));
            T a_n = wn, b_n = wn;
            push<T, kDrain>(dc, a_c, dn, a_n);
            push_drain_fast<T>(dc, b_c, dn, b_n);
            ok = ok && same_bits(a_n, b_n);
        }
        ok = ok && same_bits(a_c, b_c);
        if (!ok) bad++;
    }
    return bad;
}
extern "C" long long relax_equiv_drain_f64(long long n, unsigned long long seed) { return run_drain<double>(n, seed); }

extern "C" long long relax_equiv_run_f64(long long n, unsigned long long seed, long long* n_cap_bites) { return run<double>(n, seed, n_cap_bites); }
extern "C" long long relax_equiv_run_f32(long long n, unsigned long long seed, long long* n_cap_bites) { return run<float>(n, seed, n_cap_bites); }
