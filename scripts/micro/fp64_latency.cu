// Microbenchmark: dependent-issue latency of the instructions on the relax chain (one warp, B200).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b) {
    double x = a, y = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) x = x + y;   // dependent DADD
    }
    long long t1 = clock64();
    double z = a;
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) z = z * 1.0000001;   // dependent DMUL
    }
    long long t2 = clock64();
    double s = a;
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) s = (s > y) ? s - y : s + b;   // DSETP + DADDs + select
    }
    long long t3 = clock64();
    float f = (float)a, g = (float)b;
#pragma unroll 1
    for (int i = 0; i < 256; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) f = f + g;
    }
    long long t4 = clock64();
    out[threadIdx.x] = x + z + s + f;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 4 * 8);
    for (int w = 1; w <= 8; w *= 2) {
        k<<<1, 32 * w>>>(o, c, 1.5, 1e-9);
        long long h[4]; cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
        printf("warps %d: DADD %.2f  DMUL %.2f  DSETP+DADD+SEL %.2f  FADD %.2f cycles per dependent op\n", w, h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / 4096.0);
    }
    return 0;
}
