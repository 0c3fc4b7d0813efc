// Microbenchmark (B200): throughput of the eight-neighbour relax chain per SM sub-partition as a
// function of resident warps, for the reference-form step and the Add fast step, plus the raw
// issue rate of independent DADDs. One CTA, W warps per scheduler; cycles per chain per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../wdpm_b200/csrc/relax.cuh"
using namespace wdpm;

// candidate formulations of the Add neighbour step (all value-identical on non-negative finite inputs)
template <int V>
__device__ __forceinline__ void push_cand(double dc, double& wc, double dn, double& wn) {
    const double sn = dn + wn;
    const double sc = dc + wc;
    const double h = sc - sn;
    bool q;
    if (V & 2) q = __double_as_longlong(dc) > __double_as_longlong(sn);  // integer compare (non-negative operands)
    else q = dc > sn;
    const double x = q ? wc : h;
    if (V & 1) {  // gate folded into the scaling factor
        int mhi;
        asm("{\n\t.reg .pred p, pq;\n\tsetp.ne.s32 pq, %2, 0;\n\tsetp.ge.or.s32 p, %1, 0, pq;\n\tselp.b32 %0, 0x3fc00000, 0, p;\n\t}"
            : "=r"(mhi) : "r"(__double2hiint(h)), "r"((int)q));
        const double give = x * __hiloint2double(mhi, 0);
        wc = wc - give;
        wn = wn + give;
    } else {
        const double flow = x * 0.125;
        const double give = (__double2hiint(x) >= 0) ? flow : -0.0;
        wc = wc - give;
        wn = wn + give;
    }
}
template <int V>
__device__ __forceinline__ void relax_cand(double (&w)[3][5], const double (&d)[3][5]) {
    const double dc = d[1][1];
    double wc = w[1][1];
    push_cand<V>(dc, wc, d[0][0], w[0][0]);
    push_cand<V>(dc, wc, d[0][1], w[0][1]);
    push_cand<V>(dc, wc, d[0][2], w[0][2]);
    push_cand<V>(dc, wc, d[1][0], w[1][0]);
    push_cand<V>(dc, wc, d[1][2], w[1][2]);
    push_cand<V>(dc, wc, d[2][0], w[2][0]);
    push_cand<V>(dc, wc, d[2][1], w[2][1]);
    push_cand<V>(dc, wc, d[2][2], w[2][2]);
    w[1][1] = wc;
}

template <int MODE>
__global__ void k(double* out, long long* cyc, const double* in, int reps) {
    double w[3][5], d[3][5];
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 5; c++) w[r][c] = in[(threadIdx.x * 31 + r * 3 + c) % 256];
        for (int c = 0; c < 5; c++) d[r][c] = 500.0 + in[(threadIdx.x * 17 + r * 5 + c) % 256];
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < reps; i++) {
        if (MODE == 0) relax_window5<double, kAdd, 0, false>(w, d);
        if (MODE == 1) relax_window5<double, kAdd, 0, true>(w, d);
        if (MODE == 2) {  // 9 independent DADDs
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) w[r][c] = w[r][c] + d[r][c];
        }
        if (MODE == 3) {  // 9 independent DSETP + select pairs
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) w[r][c] = (w[r][c] > d[r][c + 1]) ? d[r][c] : w[r][c];
        }
        if (MODE == 4) relax_cand<0>(w, d);
        if (MODE == 5) relax_cand<1>(w, d);
        if (MODE == 6) relax_cand<2>(w, d);
        if (MODE == 7) relax_cand<3>(w, d);
        w[1][1] += 0.3;  // keep the centre wet
    }
    long long t1 = clock64();
    __syncthreads();
    double s = 0;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) s += w[r][c];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE>
__global__ void kf(float* out, long long* cyc, const double* in, int reps) {
    float w[3][5], d[3][5];
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 5; c++) {
            w[r][c] = (float)in[(threadIdx.x * 31 + r * 3 + c) % 256];
            d[r][c] = 5.0f + (float)in[(threadIdx.x * 17 + r * 5 + c) % 256];
        }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < reps; i++) {
        if (MODE == 0) relax_window5<float, kAdd, 0, false>(w, d);
        if (MODE == 1) {  // the fp64 fast formulation transplanted to fp32: no cap, gate read off h
            const float dc = d[1][1];
            float wc = w[1][1];
#define P32(r, c) push_add_fast<float>(dc, wc, d[r][c], w[r][c])
            P32(0, 0); P32(0, 1); P32(0, 2); P32(1, 0); P32(1, 2); P32(2, 0); P32(2, 1); P32(2, 2);
#undef P32
            w[1][1] = wc;
        }
        w[1][1] += 0.3f;
    }
    long long t1 = clock64();
    __syncthreads();
    float s = 0;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) s += w[r][c];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    double *o, *in; long long* c;
    cudaMalloc(&o, 1024 * 8); cudaMalloc(&c, 8); cudaMalloc(&in, 256 * 8);
    double h[256];
    for (int i = 0; i < 256; i++) h[i] = 0.05 + 0.3 * ((i * 7919) % 97) / 97.0;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const int reps = 2000;
    const char* names[8] = {"reference step x8 (15 instr/nbr)", "relax.cuh add fast step x8", "9 independent DADD", "9 independent DSETP+FSELx2",
                            "cand 0: sign gate, select", "cand 1: gate in the factor", "cand 2: sign gate + integer q", "cand 3: factor gate + integer q"};
    for (int mode = 0; mode < 8; mode++) {
        for (int wps = 1; wps <= 8; wps += (wps < 4 ? 1 : 2)) {
            const int threads = 128 * wps;
            if (mode == 0) k<0><<<1, threads>>>(o, c, in, reps);
            if (mode == 1) k<1><<<1, threads>>>(o, c, in, reps);
            if (mode == 2) k<2><<<1, threads>>>(o, c, in, reps);
            if (mode == 3) k<3><<<1, threads>>>(o, c, in, reps);
            if (mode == 4) k<4><<<1, threads>>>(o, c, in, reps);
            if (mode == 5) k<5><<<1, threads>>>(o, c, in, reps);
            if (mode == 6) k<6><<<1, threads>>>(o, c, in, reps);
            if (mode == 7) k<7><<<1, threads>>>(o, c, in, reps);
            long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
            printf("%-36s warps/scheduler %d: %8.1f cycles per rep per warp-slot, %7.2f cycles per rep per scheduler-warp\n", names[mode], wps,
                   (double)cy / reps, (double)cy / reps / wps);
        }
    }
    const char* fnames[2] = {"fp32 reference step x8 (fminf, predicated adds)", "fp32 no cap, sign gate"};
    float* of; cudaMalloc(&of, 1024 * 4);
    for (int mode = 0; mode < 2; mode++) {
        for (int wps = 1; wps <= 8; wps += (wps < 4 ? 1 : 2)) {
            const int threads = 128 * wps;
            if (mode == 0) kf<0><<<1, threads>>>(of, c, in, reps);
            if (mode == 1) kf<1><<<1, threads>>>(of, c, in, reps);
            long long cy; cudaMemcpy(&cy, c, 8, cudaMemcpyDeviceToHost);
            printf("%-48s warps/scheduler %d: %8.1f cycles per rep per warp-slot, %7.2f cycles per rep per scheduler-warp\n", fnames[mode], wps,
                   (double)cy / reps, (double)cy / reps / wps);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
