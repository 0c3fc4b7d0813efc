/*
 * TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * Compiles the VERBATIM reference kernel file as C. RUNOFF_CL is the quoted
 * path of /root/reference/src/runoff.cl, passed by oracle/Makefile; the file is
 * included where it lies and never copied into this repository. OpenCL C's
 * address-space and kernel qualifiers are defined away and get_global_id reads
 * a thread-local set by the NDRange loop (minicl.c / refcl_driver.c).
 *
 * With -DREF_FP32 every `double` in the kernel file becomes `float` (the
 * literals 8.0 / 0.0 stay double: x/8.0 is evaluated in double and rounded
 * back, which is exact for a power of two) - the fp32 twin of the reference.
 */
#ifndef RUNOFF_CL
#error "pass -DRUNOFF_CL='\"/root/reference/src/runoff.cl\"'"
#endif

extern __thread int refcl_gid[2];
static inline int get_global_id(int dim) { return refcl_gid[dim]; }

#define __global
#define __kernel

#ifdef REF_FP32
#define maxi refcl_maxi_f32
#define mini refcl_mini_f32
#define runoffadd refcl_runoffadd_f32
#define runoffsubtract refcl_runoffsubtract_f32
#define runoffdrain refcl_runoffdrain_f32
#define add refcl_add_f32
#define subtract refcl_subtract_f32
#define ddrain refcl_ddrain_f32
#define double float
#else
#define maxi refcl_maxi_f64
#define mini refcl_mini_f64
#define runoffadd refcl_runoffadd_f64
#define runoffsubtract refcl_runoffsubtract_f64
#define runoffdrain refcl_runoffdrain_f64
#define add refcl_add_f64
#define subtract refcl_subtract_f64
#define ddrain refcl_ddrain_f64
#endif

#include RUNOFF_CL
