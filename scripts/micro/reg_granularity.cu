// Does a CTA of 25 warps fit at 80 registers per thread (64000 of 65536), or does the register file
// hand out registers to groups of 4 warps (28 x 32 x 80 = 71680 > 65536)? Asks the occupancy calculator
// and then really launches.
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void __maxnreg__(N) k(double* out, const double* in, int n) {
    double acc[36];
#pragma unroll
    for (int i = 0; i < 36; i++) acc[i] = in[(threadIdx.x + i * 37) % n];
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 36; i++) acc[i] = acc[i] * acc[(i + 1) % 36] + in[(r + i) % n];
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 36; i++) s += acc[i];
    out[threadIdx.x] = s;
}
template <int N>
void probe(double* o, double* in) {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k<N>);
    for (int threads : {768, 800, 832, 896}) {
        int blocks = -1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k<N>, threads, 0);
        k<N><<<1, threads>>>(o, in, 64);
        cudaError_t e = cudaDeviceSynchronize();
        cudaError_t e2 = cudaGetLastError();
        printf("maxnreg %d: kernel uses %d regs; %d threads -> occupancy %d blocks/SM, launch: %s / %s\n", N, a.numRegs, threads, blocks,
               cudaGetErrorString(e), cudaGetErrorString(e2));
    }
}
int main() {
    double *o, *in;
    cudaMalloc(&o, 1024 * 8); cudaMalloc(&in, 64 * 8); cudaMemset(in, 0, 64 * 8);
    probe<72>(o, in);
    probe<80>(o, in);
    probe<88>(o, in);
    return 0;
}
