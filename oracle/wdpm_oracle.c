/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE. See wdpm_oracle_impl.h.
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (oracle/Makefile).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "wdpm_oracle.h"

#define REAL double
#define SFX _f64
#include "wdpm_oracle_impl.h"
#undef REAL
#undef SFX

#define REAL float
#define SFX _f32
#include "wdpm_oracle_impl.h"
#undef REAL
#undef SFX

int wdpm_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
