/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE. See wdpm_oracle_impl.h.
 *
 * C API of the CPU oracle (restatement of /root/reference/src/runoff.cl and the
 * solver loop of /root/reference/src/WDPMCL.c:1054-1268). Loaded through ctypes
 * by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
 */
#ifndef WDPM_ORACLE_H
#define WDPM_ORACLE_H

#define WDPM_ORACLE_ADD 0
#define WDPM_ORACLE_SUBTRACT 1
#define WDPM_ORACLE_DRAIN 2

#define WDPM_ORACLE_SCHED_OPENCL 0 /* runoff.cl arithmetic, WDPMCL.c:1184-1216 */
#define WDPM_ORACLE_SCHED_SERIAL 1 /* runoffs/runoffd + drain(), WDPMCL.c:1074-1125 */

#ifdef __cplusplus
extern "C" {
#endif

#define WDPM_ORACLE_DECL(REAL, SFX)                                                              \
    void wdpm_oracle_subpass##SFX(REAL *w, const REAL *d, int R, int C, REAL nodata, int module, \
                                  int schedule, int oi, int oj, int drainrow, int draincol,      \
                                  REAL *totaldrain);                                             \
    void wdpm_oracle_iterate##SFX(REAL *w, const REAL *d, int R, int C, REAL nodata, int module, \
                                  int schedule, int n_iters, int drainrow, int draincol,         \
                                  REAL *totaldrain);                                             \
    void wdpm_oracle_block##SFX(REAL *w, REAL *oldw, const REAL *d, int R, int C, REAL nodata,   \
                                int module, int schedule, REAL thres, int n_iters, int drainrow, \
                                int draincol, REAL *totaldrain, double *max_diff,                \
                                double *masked_sum);                                             \
    int wdpm_oracle_find_outlet##SFX(const REAL *d, int R, int C, int *drainrow, int *draincol);     \
    int wdpm_oracle_iterate_outlets##SFX(REAL *w, const REAL *d, int R, int C, REAL nodata,          \
                                         int n_iters, int n_outlets, const int *rows,               \
                                         const int *cols, REAL *totals);

WDPM_ORACLE_DECL(double, _f64)
WDPM_ORACLE_DECL(float, _f32)

int wdpm_oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
