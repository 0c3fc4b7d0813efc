/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * CPU restatement of WDPM's water-redistribution path, instantiated once per
 * precision by wdpm_oracle.c (REAL = double / float, SFX = _f64 / _f32).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may call into this; the shipped library never does.
 *
 * Layout: padded row-major grids of (R+2) x (C+2) cells, index i*(C+2)+j with
 * i in [0,R+1], j in [0,C+1] (the reference's bigdem/bigwater, WDPMCL.c:795-807).
 * The reference's OpenCL path flattens column-major (runoff.cl:33); the layout
 * does not affect any value, only addresses.
 *
 * Parity status: pinned. tests/test_oracle.py checks these
 * functions bit-for-bit against the verbatim runoff.cl compiled through
 * oracle/ref_shim (oracle/_ref/librunoffcl_ref.so) and against outputs of the
 * unmodified WDPMCL.c serial backend (oracle/_ref/WDPMCL_ref), and the
 * validation/ awk goldens (validate_WDPM.sh:48-70) are checked on both schedules.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SFX)

/* runoff.cl:3-22 - tie semantics: maxi returns b when a<=b, mini returns a. */
static inline REAL FN(cl_maxi)(REAL a, REAL b) { return (a <= b) ? b : a; }
static inline REAL FN(cl_mini)(REAL a, REAL b) { return (a <= b) ? a : b; }
/* WDPMCL.c:19-20 - the host's macros have the opposite tie choice. */
#define HOST_MAX(a, b) (((a) > (b)) ? (a) : (b))
#define HOST_MIN(a, b) (((a) < (b)) ? (a) : (b))

#define IDX(i, j) ((size_t)(i) * (size_t)pitch + (size_t)(j))

/* ---- OpenCL-branch arithmetic ------------------------------------------ */

/* runoff.cl:24-55 (runoffadd). */
static void FN(relax_add_cl)(REAL *w, const REAL *d, int pitch, int ci, int cj, REAL nodata)
{
    for (int i = ci - 1; i <= ci + 1; i++) {
        for (int j = cj - 1; j <= cj + 1; j++) {
            if ((i != ci || j != cj) && d[IDX(i, j)] > nodata) {
                REAL cell_elev = d[IDX(i, j)] + w[IDX(i, j)];
                REAL centre_elev = d[IDX(ci, cj)] + w[IDX(ci, cj)];
                REAL h = centre_elev - cell_elev;
                if (h > 0) {
                    REAL flow;
                    if (d[IDX(ci, cj)] > cell_elev)
                        flow = w[IDX(ci, cj)] / (REAL)8.0;
                    else
                        flow = h / (REAL)8.0;
                    flow = FN(cl_mini)(FN(cl_maxi)(flow, (REAL)0.0), w[IDX(ci, cj)]);
                    w[IDX(ci, cj)] = FN(cl_maxi)(w[IDX(ci, cj)] - flow, (REAL)0.0);
                    w[IDX(i, j)] = w[IDX(i, j)] + flow;
                }
            }
        }
    }
}

/* runoff.cl:57-88 (runoffsubtract): four-term else branch, no max clamps. */
static void FN(relax_sub_cl)(REAL *w, const REAL *d, int pitch, int ci, int cj, REAL nodata)
{
    for (int i = ci - 1; i <= ci + 1; i++) {
        for (int j = cj - 1; j <= cj + 1; j++) {
            if ((i != ci || j != cj) && d[IDX(i, j)] > nodata) {
                REAL h = (d[IDX(ci, cj)] + w[IDX(ci, cj)]) - (d[IDX(i, j)] + w[IDX(i, j)]);
                if (h > 0) {
                    REAL flow;
                    if (d[IDX(ci, cj)] > (d[IDX(i, j)] + w[IDX(i, j)]))
                        flow = w[IDX(ci, cj)] / (REAL)8.0;
                    else
                        flow = ((d[IDX(ci, cj)] - d[IDX(i, j)]) +
                                (w[IDX(ci, cj)] - w[IDX(i, j)])) / (REAL)8.0;
                    flow = FN(cl_mini)(flow, w[IDX(ci, cj)]);
                    w[IDX(ci, cj)] = w[IDX(ci, cj)] - flow;
                    w[IDX(i, j)] = w[IDX(i, j)] + flow;
                }
            }
        }
    }
}

/* runoff.cl:90-134 (runoffdrain): outlet test precedes the height test. */
static void FN(relax_drain_cl)(REAL *w, const REAL *d, int pitch, int ci, int cj, REAL nodata,
                               int oi, int oj, REAL *totaldrain)
{
    for (int i = ci - 1; i <= ci + 1; i++) {
        for (int j = cj - 1; j <= cj + 1; j++) {
            if ((i != ci || j != cj) && d[IDX(i, j)] > nodata) {
                REAL centre_elev = d[IDX(ci, cj)] + w[IDX(ci, cj)];
                REAL cell_elev = d[IDX(i, j)] + w[IDX(i, j)];
                if (j == oj && i == oi) {
                    totaldrain[0] = totaldrain[0] + w[IDX(oi, oj)] + w[IDX(ci, cj)];
                    w[IDX(oi, oj)] = (REAL)0.0;
                    w[IDX(ci, cj)] = (REAL)0.0;
                } else {
                    REAL h = centre_elev - cell_elev;
                    if (h > 0) {
                        REAL flow;
                        if (d[IDX(ci, cj)] > cell_elev)
                            flow = w[IDX(ci, cj)] / (REAL)8.0;
                        else
                            flow = ((d[IDX(ci, cj)] - d[IDX(i, j)]) +
                                    (w[IDX(ci, cj)] - w[IDX(i, j)])) / (REAL)8.0;
                        flow = FN(cl_mini)(FN(cl_maxi)(flow, (REAL)0.0), w[IDX(ci, cj)]);
                        w[IDX(ci, cj)] = FN(cl_maxi)(w[IDX(ci, cj)] - flow, (REAL)0.0);
                        w[IDX(i, j)] = w[IDX(i, j)] + flow;
                    }
                }
            }
        }
    }
}

/* ---- serial-branch arithmetic ------------------------------------------ */

/* WDPMCL.c:1934-1964 (runoffs), used for Add AND Subtract (:1100, :1116). The
 * four-term value assigned at :1953 is dead; :1955 overwrites it with h/8. */
static void FN(relax_serial_s)(REAL *w, const REAL *d, int pitch, int ci, int cj, REAL nodata)
{
    for (int i = ci - 1; i <= ci + 1; i++) {
        for (int j = cj - 1; j <= cj + 1; j++) {
            if ((i != ci || j != cj) && d[IDX(i, j)] > nodata) {
                REAL h = (d[IDX(ci, cj)] + w[IDX(ci, cj)]) - (d[IDX(i, j)] + w[IDX(i, j)]);
                if (h > 0) {
                    REAL flow;
                    if (d[IDX(ci, cj)] > (d[IDX(i, j)] + w[IDX(i, j)]))
                        flow = w[IDX(ci, cj)] / (REAL)8.0;
                    else
                        flow = h / (REAL)8.0;
                    flow = HOST_MIN(flow, w[IDX(ci, cj)]);
                    w[IDX(ci, cj)] = w[IDX(ci, cj)] - flow;
                    w[IDX(i, j)] = w[IDX(i, j)] + flow;
                }
            }
        }
    }
}

/* WDPMCL.c:1967-2006 (runoffd). */
static void FN(relax_serial_d)(REAL *w, const REAL *d, int pitch, int ci, int cj, REAL nodata,
                               int oi, int oj, REAL *totaldrain)
{
    for (int i = ci - 1; i <= ci + 1; i++) {
        for (int j = cj - 1; j <= cj + 1; j++) {
            if ((i != ci || j != cj) && d[IDX(i, j)] > nodata) {
                REAL centre_elev = d[IDX(ci, cj)] + w[IDX(ci, cj)];
                REAL cell_elev = d[IDX(i, j)] + w[IDX(i, j)];
                if (j == oj && i == oi) {
                    totaldrain[0] = totaldrain[0] + w[IDX(oi, oj)] + w[IDX(ci, cj)];
                    w[IDX(oi, oj)] = (REAL)0.0;
                    w[IDX(ci, cj)] = (REAL)0.0;
                } else {
                    REAL h = centre_elev - cell_elev;
                    if (h > 0) {
                        REAL flow;
                        if (d[IDX(ci, cj)] > cell_elev)
                            flow = w[IDX(ci, cj)] / (REAL)8.0;
                        else
                            flow = ((d[IDX(ci, cj)] - d[IDX(i, j)]) +
                                    (w[IDX(ci, cj)] - w[IDX(i, j)])) / (REAL)8.0;
                        flow = HOST_MIN(HOST_MAX(flow, (REAL)0.0), w[IDX(ci, cj)]);
                        w[IDX(ci, cj)] = HOST_MAX(w[IDX(ci, cj)] - flow, (REAL)0.0);
                        w[IDX(i, j)] = w[IDX(i, j)] + flow;
                    }
                }
            }
        }
    }
}

/* WDPMCL.c:1859-1897 (drain): sum wet valid cells of the outlet's 3x3, then
 * zero all nine cells (valid or not). Serial backend only has an effect. */
static REAL FN(outlet_wipe)(REAL *w, const REAL *d, int pitch, REAL nodata, int oi, int oj)
{
    REAL got = 0;
    for (int i = oi - 1; i <= oi + 1; i++)
        for (int j = oj - 1; j <= oj + 1; j++)
            if (d[IDX(i, j)] > nodata && w[IDX(i, j)] > 0)
                got += w[IDX(i, j)];
    for (int i = oi - 1; i <= oi + 1; i++)
        for (int j = oj - 1; j <= oj + 1; j++)
            w[IDX(i, j)] = (REAL)0.0;
    return got;
}

/* One colour sub-pass. Centres: row = oi+3k, col = oj+3m, 1-based padded
 * coordinates, inside [1,R]x[1,C], wet and valid (runoff.cl:142-145, :157-160,
 * :174-179; serial twin WDPMCL.c:1079-1084, :1097-1101). Centres of one colour
 * own disjoint 3x3 tiles, so the order inside a sub-pass cannot change a bit
 * and the loop may run under OpenMP. That includes Drain: only a centre adjacent
 * to the outlet touches totaldrain, and the outlet's 3x3 neighbourhood holds at
 * most one centre of any colour (centres are 3 apart), so there is a single
 * writer per sub-pass. */
void FN(wdpm_oracle_subpass)(REAL *w, const REAL *d, int R, int C, REAL nodata,
                             int module, int schedule, int oi, int oj,
                             int drainrow, int draincol, REAL *totaldrain)
{
    const int pitch = C + 2;
    if (module == WDPM_ORACLE_DRAIN) {
#pragma omp parallel for schedule(static)
        for (int i = oi; i <= R; i += 3)
            for (int j = oj; j <= C; j += 3)
                if (w[IDX(i, j)] > (REAL)0.0 && d[IDX(i, j)] > nodata &&
                    (i != drainrow || j != draincol)) {
                    if (schedule == WDPM_ORACLE_SCHED_OPENCL)
                        FN(relax_drain_cl)(w, d, pitch, i, j, nodata, drainrow, draincol, totaldrain);
                    else
                        FN(relax_serial_d)(w, d, pitch, i, j, nodata, drainrow, draincol, totaldrain);
                }
        return;
    }
#pragma omp parallel for schedule(static)
    for (int i = oi; i <= R; i += 3)
        for (int j = oj; j <= C; j += 3)
            if (w[IDX(i, j)] > (REAL)0.0 && d[IDX(i, j)] > nodata) {
                if (schedule == WDPM_ORACLE_SCHED_SERIAL)
                    FN(relax_serial_s)(w, d, pitch, i, j, nodata);
                else if (module == WDPM_ORACLE_ADD)
                    FN(relax_add_cl)(w, d, pitch, i, j, nodata);
                else
                    FN(relax_sub_cl)(w, d, pitch, i, j, nodata);
            }
}

/* n_iters full iterations: oi outer, oj inner (WDPMCL.c:1184-1186, :1077-1078).
 * Serial Drain additionally wipes the outlet's 3x3 after every iteration
 * (WDPMCL.c:1089); in the OpenCL branch that call acts on a stale host copy and
 * is discarded by the readback (:1214 vs :1217-1233), so it is not applied. */
void FN(wdpm_oracle_iterate)(REAL *w, const REAL *d, int R, int C, REAL nodata,
                             int module, int schedule, int n_iters,
                             int drainrow, int draincol, REAL *totaldrain)
{
    const int pitch = C + 2;
    for (int it = 0; it < n_iters; it++) {
        for (int oi = 1; oi <= 3; oi++)
            for (int oj = 1; oj <= 3; oj++)
                FN(wdpm_oracle_subpass)(w, d, R, C, nodata, module, schedule, oi, oj,
                                        drainrow, draincol, totaldrain);
        if (module == WDPM_ORACLE_DRAIN && schedule == WDPM_ORACLE_SCHED_SERIAL)
            totaldrain[0] = totaldrain[0] + FN(outlet_wipe)(w, d, pitch, nodata, drainrow, draincol);
    }
}

/* EXTENSION (not in the reference, which has exactly one outlet): Drain with a SET of outlets,
 * the semantics wdpm_set_outlets() of the product defines and BASELINE configs[4] asks for. Every
 * outlet behaves as runoff.cl:104-111 prescribes for the one outlet - it is never a centre
 * (runoff.cl:179), and a centre next to it adds w[outlet] + w[centre] to THAT outlet's total and
 * zeroes both, before any height test. `oid` maps a padded cell to its outlet index or -1. With one
 * outlet this is relax_drain_cl. OpenCL-branch arithmetic only. */
static void FN(relax_drain_set)(REAL *w, const REAL *d, int pitch, int ci, int cj, REAL nodata,
                                const int *oid, REAL *totals)
{
    for (int i = ci - 1; i <= ci + 1; i++) {
        for (int j = cj - 1; j <= cj + 1; j++) {
            if ((i != ci || j != cj) && d[IDX(i, j)] > nodata) {
                REAL centre_elev = d[IDX(ci, cj)] + w[IDX(ci, cj)];
                REAL cell_elev = d[IDX(i, j)] + w[IDX(i, j)];
                const int k = oid[IDX(i, j)];
                if (k >= 0) {
                    totals[k] = totals[k] + w[IDX(i, j)] + w[IDX(ci, cj)];
                    w[IDX(i, j)] = (REAL)0.0;
                    w[IDX(ci, cj)] = (REAL)0.0;
                } else {
                    REAL h = centre_elev - cell_elev;
                    if (h > 0) {
                        REAL flow;
                        if (d[IDX(ci, cj)] > cell_elev)
                            flow = w[IDX(ci, cj)] / (REAL)8.0;
                        else
                            flow = ((d[IDX(ci, cj)] - d[IDX(i, j)]) +
                                    (w[IDX(ci, cj)] - w[IDX(i, j)])) / (REAL)8.0;
                        flow = FN(cl_mini)(FN(cl_maxi)(flow, (REAL)0.0), w[IDX(ci, cj)]);
                        w[IDX(ci, cj)] = FN(cl_maxi)(w[IDX(ci, cj)] - flow, (REAL)0.0);
                        w[IDX(i, j)] = w[IDX(i, j)] + flow;
                    }
                }
            }
        }
    }
}

/* n_iters Drain iterations with an outlet set; totals[k] accumulates outlet k's contacts. Serial
 * inside a sub-pass (an outlet has at most one neighbouring centre per colour, so the order is
 * immaterial; kept serial for simplicity). Returns -1 on allocation failure. */
int FN(wdpm_oracle_iterate_outlets)(REAL *w, const REAL *d, int R, int C, REAL nodata, int n_iters,
                                    int n_outlets, const int *rows, const int *cols, REAL *totals)
{
    const int pitch = C + 2;
    const size_t n = (size_t)(R + 2) * (size_t)(C + 2);
    int *oid = (int *)malloc(n * sizeof *oid);
    if (!oid) return -1;
    for (size_t k = 0; k < n; k++) oid[k] = -1;
    for (int k = 0; k < n_outlets; k++) oid[IDX(rows[k], cols[k])] = k;
    for (int it = 0; it < n_iters; it++)
        for (int oi = 1; oi <= 3; oi++)
            for (int oj = 1; oj <= 3; oj++)
                for (int i = oi; i <= R; i += 3)
                    for (int j = oj; j <= C; j += 3)
                        if (w[IDX(i, j)] > (REAL)0.0 && d[IDX(i, j)] > nodata && oid[IDX(i, j)] < 0)
                            FN(relax_drain_set)(w, d, pitch, i, j, nodata, oid, totals);
    free(oid);
    return 0;
}

/* One convergence block (WDPMCL.c:1054-1268): zero-threshold over the whole
 * padded grid (:1055-1065), snapshot (:1069-1073), n_iters iterations, then
 * max |w-old| over valid cells seeded with cell [0][0] (:1239-1254) and the
 * masked sum of water in row-major order (:1259-1266, accumulated in double).
 * olddrain is the caller's business (it owns totaldrain). */
void FN(wdpm_oracle_block)(REAL *w, REAL *oldw, const REAL *d, int R, int C, REAL nodata,
                           int module, int schedule, REAL thres, int n_iters,
                           int drainrow, int draincol, REAL *totaldrain,
                           double *max_diff, double *masked_sum)
{
    const int pitch = C + 2;
    const size_t n = (size_t)(R + 2) * (size_t)(C + 2);
    for (size_t k = 0; k < n; k++) {
        if (w[k] < thres) w[k] = 0;
        oldw[k] = w[k];
    }
    FN(wdpm_oracle_iterate)(w, d, R, C, nodata, module, schedule, n_iters,
                            drainrow, draincol, totaldrain);
    REAL md = (REAL)fabs((double)(w[0] - oldw[0]));
    double sum = 0.0;
    for (int i = 0; i < R + 2; i++)
        for (int j = 0; j < C + 2; j++)
            if (d[IDX(i, j)] > nodata) {
                REAL df = w[IDX(i, j)] - oldw[IDX(i, j)];
                if (df < 0) df = -df;
                if (df > md) md = df;
                sum += (double)w[IDX(i, j)];
            }
    *max_diff = (double)md;
    *masked_sum = sum;
}

/* Outlet = lowest padded cell with dem > 0 (not > nodata), strict '<', rows then
 * columns, so the first in row-major order wins ties (WDPMCL.c:1005-1017).
 * Returns 0 and leaves the outputs untouched if no cell qualifies. */
int FN(wdpm_oracle_find_outlet)(const REAL *d, int R, int C, int *drainrow, int *draincol)
{
    const int pitch = C + 2;
    REAL lowest = (REAL)100000000;
    int found = 0;
    for (int i = 0; i < R + 2; i++)
        for (int j = 0; j < C + 2; j++)
            if (d[IDX(i, j)] > 0 && d[IDX(i, j)] < lowest) {
                lowest = d[IDX(i, j)];
                *drainrow = i;
                *draincol = j;
                found = 1;
            }
    return found;
}

#undef IDX
#undef HOST_MAX
#undef HOST_MIN
#undef FN
#undef CAT
#undef CAT_
