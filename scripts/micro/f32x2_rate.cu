// Does a packed add.f32x2 cost one issue cycle or two? (If two, packing the fp32 relax step two
// tiles at a time cannot shorten it: the fused kernel is bound by issue cycles, not instruction count.)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
template <int MODE>
__global__ void k(float* out, long long* cyc, float seed, int reps) {
    float a[16];
    unsigned long long p[8];
    for (int i = 0; i < 16; i++) a[i] = seed + i + threadIdx.x;
    for (int i = 0; i < 8; i++) p[i] = ((unsigned long long)__float_as_uint(a[2 * i]) << 32) | __float_as_uint(a[2 * i + 1]);
    const unsigned long long inc = ((unsigned long long)__float_as_uint(seed) << 32) | __float_as_uint(seed);
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; r++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) a[i] = a[i] + seed;      // 16 independent FADD
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = add2(p[i], inc);   // 8 independent add.f32x2 (same 16 additions)
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 16; i++) s += a[i];
    for (int i = 0; i < 8; i++) s += __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    float* o; long long* c;
    cudaMalloc(&o, 1024 * 4); cudaMalloc(&c, 8);
    const int reps = 4000;
    for (int mode = 0; mode < 2; mode++)
        for (int wps = 1; wps <= 8; wps *= 2) {
            if (mode == 0) k<0><<<1, 128 * wps>>>(o, c, 1.0f, reps); else k<1><<<1, 128 * wps>>>(o, c, 1.0f, reps);
            long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
            printf("%s, %d warps/scheduler: %.2f cycles per 16 additions per warp\n", mode ? "8 x add.f32x2" : "16 x FADD     ", wps, (double)cy / reps / wps);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
