/*
 * What a "%f" text round trip does to a value.
 *
 * The reference hands grids from one module to the next through ESRI ASCII files: written with
 * fprintf("%f ") (src/WDPMCL.c:1546), read back with fscanf("%lf") (:1576). Both conversions are
 * correctly rounded in glibc, so the round trip is the function
 *
 *     q(w) = nearest double to ( w rounded to 6 decimals, ties to even on the EXACT value of w ).
 *
 * wdpm_quantize6 computes q without going through text, so that chained modules can keep their
 * grids in memory (host) or in HBM (device, k_quantize_water) and still start from exactly the
 * values a file would have given them:
 *   p + e = w * 10^6 exactly (p the rounded product, e its error, from one fused multiply-add);
 *   n = p rounded to an integer, ties to even - and when p sits exactly on a tie the sign of e says
 *   on which side of it the true value lies; n / 10^6 with IEEE division is the double nearest to
 *   the exact quotient, which is what strtod returns for the printed digits.
 * Valid for finite |w| < 4.5e9 (w * 10^6 below 2^52); larger magnitudes are returned unchanged.
 * The sign of a zero result follows w, as "-0.000000" does. Checked bit for bit against
 * snprintf/strtod in tests/test_quantize.py.
 *
 * Plain C99 / CUDA: included by the command-line host and by the CUDA library.
 */
#ifndef WDPM_QUANTIZE_H
#define WDPM_QUANTIZE_H

#include <math.h>

#ifdef __CUDACC__
#define WDPM_Q_HD __host__ __device__ inline
#else
#define WDPM_Q_HD static inline
#endif

WDPM_Q_HD double wdpm_quantize6(double w)
{
    if (!(fabs(w) < 4.5e9)) return w; /* also NaN and infinities */
    const double p = w * 1.0e6;
    const double e = fma(w, 1.0e6, -p);
    double n = rint(p); /* ties to even */
    const double d = p - n; /* exact */
    if (d == 0.5 || d == -0.5) { /* p is a tie; the true value p + e is one only if e == 0 */
        const double lo = floor(p);
        if (e > 0.0) n = lo + 1.0;
        else if (e < 0.0) n = lo;
    }
    const double q = n / 1.0e6;
    return (q == 0.0) ? copysign(0.0, w) : q;
}

#endif
