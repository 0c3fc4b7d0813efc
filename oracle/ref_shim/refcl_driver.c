/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * ctypes-facing driver around the verbatim reference kernels (runoffcl_tu.c):
 * launches one colour sub-pass exactly as the reference host does - NDRange of
 * WDPMCL.c:1192-1198, kernel arguments of WDPMCL.c:1155-1190, COLUMN-MAJOR
 * flattened grids (index row + (numrows+2)*col, WDPMCL.c:1129-1134) - and loops
 * iterations in the order of WDPMCL.c:1184-1186. Work-items of one sub-pass own
 * disjoint 3x3 tiles, so running the NDRange under OpenMP yields the same bits
 * as any OpenCL device would.
 */
#include <stddef.h>

__thread int refcl_gid[2];

void refcl_add_f64(double *, double *, const double, const int, const int, const int, const int, const int);
void refcl_subtract_f64(double *, double *, const double, const int, const int, const int, const int, const int);
void refcl_ddrain_f64(double *, double *, double, const int, const int, const int, const int, const int, double *, const int, const int);
void refcl_add_f32(float *, float *, const float, const int, const int, const int, const int, const int);
void refcl_subtract_f32(float *, float *, const float, const int, const int, const int, const int, const int);
void refcl_ddrain_f32(float *, float *, float, const int, const int, const int, const int, const int, float *, const int, const int);

#define DRIVER(REAL, SFX)                                                                        \
    void refcl_subpass##SFX(int which, REAL *w, REAL *d, REAL nodata, int numrows, int numcols,  \
                            int oi, int oj, REAL *totaldrain, int drainrow, int draincol)       \
    {                                                                                            \
        const int offset = 4;                                                                    \
        const long g0 = (((numrows + 2) / (offset - 1)) / 32 + 1) * 32;                          \
        const long g1 = (((numcols + 2) / (offset - 1)) / 32 + 1) * 32;                          \
        if (which == 2) { /* single writer of totaldrain per sub-pass, see oracle impl */      \
            _Pragma("omp parallel for schedule(static)")                                         \
            for (long b = 0; b < g1; b++)                                                        \
                for (long a = 0; a < g0; a++) {                                                  \
                    refcl_gid[0] = (int)a;                                                       \
                    refcl_gid[1] = (int)b;                                                       \
                    refcl_ddrain##SFX(w, d, nodata, numrows, numcols, offset, oi, oj,            \
                                      totaldrain, drainrow, draincol);                           \
                }                                                                                \
            return;                                                                              \
        }                                                                                        \
        _Pragma("omp parallel for schedule(static)")                                             \
        for (long b = 0; b < g1; b++)                                                            \
            for (long a = 0; a < g0; a++) {                                                      \
                refcl_gid[0] = (int)a;                                                           \
                refcl_gid[1] = (int)b;                                                           \
                if (which == 0)                                                                  \
                    refcl_add##SFX(w, d, nodata, numrows, numcols, offset, oi, oj);              \
                else                                                                             \
                    refcl_subtract##SFX(w, d, nodata, numrows, numcols, offset, oi, oj);         \
            }                                                                                    \
    }                                                                                            \
    void refcl_iterate##SFX(int which, REAL *w, REAL *d, REAL nodata, int numrows, int numcols,  \
                            int n_iters, REAL *totaldrain, int drainrow, int draincol)          \
    {                                                                                            \
        for (int it = 0; it < n_iters; it++)                                                     \
            for (int oi = 1; oi < 4; oi++)                                                       \
                for (int oj = 1; oj < 4; oj++)                                                   \
                    refcl_subpass##SFX(which, w, d, nodata, numrows, numcols, oi, oj,            \
                                       totaldrain, drainrow, draincol);                          \
    }

DRIVER(double, _f64)
DRIVER(float, _f32)
