/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * "minicl": a single-device CPU stand-in for the OpenCL runtime, just large
 * enough to carry the UNMODIFIED /root/reference/src/WDPMCL.c through its
 * OpenCL branch (cpu=1) in an image with no ICD. Programs are not compiled from
 * source: clCreateKernel binds the three kernel names to the verbatim runoff.cl
 * functions that runoffcl_tu.c compiled as C (fp64 - the reference's only
 * precision). Buffers alias their host pointer (the reference always passes
 * CL_MEM_USE_HOST_PTR, WDPMCL.c:1138-1140, :1168), every enqueue runs to
 * completion before returning, and an NDRange is executed under OpenMP (tiles
 * of one colour are disjoint, so any execution order gives identical bits).
 */
#include <stdlib.h>
#include <string.h>
#include "CL/cl.h"

extern __thread int refcl_gid[2];
void refcl_add_f64(double *, double *, const double, const int, const int, const int, const int, const int);
void refcl_subtract_f64(double *, double *, const double, const int, const int, const int, const int, const int);
void refcl_ddrain_f64(double *, double *, double, const int, const int, const int, const int, const int, double *, const int, const int);

struct minicl_platform { int unused; };
struct minicl_device { int unused; };
struct minicl_context { int unused; };
struct minicl_queue { int unused; };
struct minicl_program { int unused; };
struct minicl_event { int unused; };
struct minicl_mem { void *ptr; size_t size; };
struct minicl_kernel {
    int which; /* 0 add, 1 subtract, 2 ddrain */
    union { void *p; double f; int i; } arg[11];
};

static struct minicl_platform the_platform;
static struct minicl_device the_device;
static struct minicl_context the_context;
static struct minicl_queue the_queue;
static struct minicl_program the_program;
static struct minicl_event the_event;

cl_int clGetPlatformIDs(cl_uint n, cl_platform_id *out, cl_uint *count)
{
    if (count) *count = 1;
    if (out && n > 0) out[0] = &the_platform;
    return CL_SUCCESS;
}

cl_int clGetDeviceIDs(cl_platform_id p, cl_device_type t, cl_uint n, cl_device_id *out, cl_uint *count)
{
    (void)t;
    if (p != &the_platform) return CL_INVALID_PLATFORM;
    if (count) *count = 1;
    if (out && n > 0) out[0] = &the_device;
    return CL_SUCCESS;
}

cl_context clCreateContext(const cl_context_properties *props, cl_uint n, const cl_device_id *devs,
                           void (*cb)(const char *, const void *, size_t, void *), void *ud, cl_int *err)
{
    (void)props; (void)n; (void)devs; (void)cb; (void)ud;
    if (err) *err = CL_SUCCESS;
    return &the_context;
}

cl_program clCreateProgramWithSource(cl_context c, cl_uint n, const char **src, const size_t *len, cl_int *err)
{
    (void)c; (void)n; (void)src; (void)len;
    if (err) *err = CL_SUCCESS;
    return &the_program;
}

cl_int clBuildProgram(cl_program p, cl_uint n, const cl_device_id *d, const char *opt,
                      void (*cb)(cl_program, void *), void *ud)
{
    (void)p; (void)n; (void)d; (void)opt; (void)cb; (void)ud;
    return CL_SUCCESS;
}

cl_int clGetProgramBuildInfo(cl_program p, cl_device_id d, cl_program_build_info what, size_t sz,
                             void *out, size_t *ret)
{
    (void)p; (void)d; (void)what;
    if (out && sz > 0) ((char *)out)[0] = 0;
    if (ret) *ret = 1;
    return CL_SUCCESS;
}

cl_command_queue clCreateCommandQueue(cl_context c, cl_device_id d, cl_command_queue_properties pr, cl_int *err)
{
    (void)c; (void)d; (void)pr;
    if (err) *err = CL_SUCCESS;
    return &the_queue;
}

cl_kernel clCreateKernel(cl_program p, const char *name, cl_int *err)
{
    (void)p;
    int which = -1;
    if (strcmp(name, "add") == 0) which = 0;
    else if (strcmp(name, "subtract") == 0) which = 1;
    else if (strcmp(name, "ddrain") == 0) which = 2;
    if (which < 0) {
        if (err) *err = CL_INVALID_KERNEL_NAME;
        return NULL;
    }
    struct minicl_kernel *k = calloc(1, sizeof *k);
    k->which = which;
    if (err) *err = CL_SUCCESS;
    return k;
}

cl_mem clCreateBuffer(cl_context c, cl_mem_flags flags, size_t size, void *host, cl_int *err)
{
    (void)c;
    struct minicl_mem *m = calloc(1, sizeof *m);
    m->size = size;
    m->ptr = (flags & CL_MEM_USE_HOST_PTR) ? host : malloc(size);
    if (err) *err = CL_SUCCESS;
    return m;
}

cl_int clEnqueueWriteBuffer(cl_command_queue q, cl_mem m, cl_bool blocking, size_t off, size_t size,
                            const void *src, cl_uint nw, const cl_event *wl, cl_event *ev)
{
    (void)q; (void)blocking; (void)nw; (void)wl;
    if ((char *)m->ptr + off != (const char *)src) memmove((char *)m->ptr + off, src, size);
    if (ev) *ev = &the_event;
    return CL_SUCCESS;
}

cl_int clEnqueueReadBuffer(cl_command_queue q, cl_mem m, cl_bool blocking, size_t off, size_t size,
                           void *dst, cl_uint nw, const cl_event *wl, cl_event *ev)
{
    (void)q; (void)blocking; (void)nw; (void)wl;
    if ((char *)m->ptr + off != (char *)dst) memmove(dst, (char *)m->ptr + off, size);
    if (ev) *ev = &the_event;
    return CL_SUCCESS;
}

/* Argument kinds by index (runoff.cl:137-138, :168-170): 0,1,8 buffers; 2 a
 * double; the rest ints. */
cl_int clSetKernelArg(cl_kernel k, cl_uint idx, size_t size, const void *value)
{
    if (idx > 10) return CL_INVALID_ARG_INDEX;
    if (idx == 0 || idx == 1 || idx == 8) {
        if (size != sizeof(cl_mem)) return CL_INVALID_ARG_SIZE;
        k->arg[idx].p = (*(const cl_mem *)value)->ptr;
    } else if (idx == 2) {
        if (size != sizeof(double)) return CL_INVALID_ARG_SIZE;
        k->arg[idx].f = *(const double *)value;
    } else {
        if (size != sizeof(int)) return CL_INVALID_ARG_SIZE;
        k->arg[idx].i = *(const int *)value;
    }
    return CL_SUCCESS;
}

cl_int clEnqueueNDRangeKernel(cl_command_queue q, cl_kernel k, cl_uint dim, const size_t *goff,
                              const size_t *gsz, const size_t *lsz, cl_uint nw, const cl_event *wl,
                              cl_event *ev)
{
    (void)q; (void)goff; (void)lsz; (void)nw; (void)wl;
    if (dim != 2) return CL_INVALID_WORK_DIMENSION;
    const long g0 = (long)gsz[0], g1 = (long)gsz[1];
    double *w = k->arg[0].p, *d = k->arg[1].p;
    const double nodata = k->arg[2].f;
    const int nr = k->arg[3].i, nc = k->arg[4].i, off = k->arg[5].i, oi = k->arg[6].i, oj = k->arg[7].i;
    if (k->which == 2) {
        double *td = k->arg[8].p;
        const int dr = k->arg[9].i, dc = k->arg[10].i;
        for (long a = 0; a < g0; a++)
            for (long b = 0; b < g1; b++) {
                refcl_gid[0] = (int)a;
                refcl_gid[1] = (int)b;
                refcl_ddrain_f64(w, d, nodata, nr, nc, off, oi, oj, td, dr, dc);
            }
    } else {
#pragma omp parallel for schedule(static)
        for (long b = 0; b < g1; b++)
            for (long a = 0; a < g0; a++) {
                refcl_gid[0] = (int)a;
                refcl_gid[1] = (int)b;
                if (k->which == 0) refcl_add_f64(w, d, nodata, nr, nc, off, oi, oj);
                else refcl_subtract_f64(w, d, nodata, nr, nc, off, oi, oj);
            }
    }
    if (ev) *ev = &the_event;
    return CL_SUCCESS;
}

cl_int clWaitForEvents(cl_uint n, const cl_event *ev) { (void)n; (void)ev; return CL_SUCCESS; }
cl_int clReleaseEvent(cl_event e) { (void)e; return CL_SUCCESS; }
cl_int clReleaseMemObject(cl_mem m) { free(m); return CL_SUCCESS; }
cl_int clReleaseKernel(cl_kernel k) { free(k); return CL_SUCCESS; }
cl_int clReleaseProgram(cl_program p) { (void)p; return CL_SUCCESS; }
cl_int clReleaseCommandQueue(cl_command_queue q) { (void)q; return CL_SUCCESS; }
cl_int clReleaseContext(cl_context c) { (void)c; return CL_SUCCESS; }
