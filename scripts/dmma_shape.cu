// dmma_shape.cu — DMMA throughput in the SHAPE the 20-state walker uses it: per step 30 (child pass: 2 children x 5 k-tiles x 3
// state tiles, distinct A per k-tile, 15 distinct B fragments) + 15 (back pass) DMMAs, with the product / shuffle in between.
// Variants: 0 = child+back passes chained as in the kernel; 1 = same but no product/shuffle dependency (independent steps).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/dmma_shape.bin scripts/dmma_shape.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int VAR>
__global__ void __launch_bounds__(128) k(const double *in, double *out, int iters) {
    const int lane = threadIdx.x & 31, q = lane & 3;
    double Be[5][3], Bi[5][3];
    for (int kt = 0; kt < 5; kt++) for (int nt = 0; nt < 3; nt++) { Be[kt][nt] = in[(kt * 3 + nt) * 32 + lane]; Bi[kt][nt] = in[480 + (kt * 3 + nt) * 32 + lane]; }
    double fA[5], fB[5];
    for (int kt = 0; kt < 5; kt++) { fA[kt] = in[960 + kt * 32 + lane]; fB[kt] = in[1120 + kt * 32 + lane]; }
    double acc = 0.0;
    for (int it = 0; it < iters; it++) {
        double a[3][2] = {}, b[3][2] = {};
#pragma unroll
        for (int kt = 0; kt < 5; kt++) { fA[kt] *= 1.0000001; fB[kt] *= 0.9999999; }   // the exp scaling of the real kernel: 10 DMULs per step
#pragma unroll
        for (int kt = 0; kt < 5; kt++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) { mma(a[nt][0], a[nt][1], fA[kt], Be[kt][nt]); mma(b[nt][0], b[nt][1], fB[kt], Be[kt][nt]); }
        double ft[5];
        if (VAR == 0) {
            ft[0] = a[0][0] * b[0][0]; ft[1] = a[0][1] * b[0][1]; ft[2] = a[1][0] * b[1][0]; ft[3] = a[1][1] * b[1][1];
            const double t0 = a[2][0] * b[2][0], t1 = a[2][1] * b[2][1];
            const double r = __shfl_sync(0xffffffffu, t1, (lane & 28) | (q & 1));
            ft[4] = q < 2 ? t0 : r;
        } else {
            for (int kt = 0; kt < 5; kt++) ft[kt] = fA[kt];
            acc += a[0][0] + a[1][0] + a[2][0] + b[0][1] + b[1][1] + b[2][1];
        }
        double o[3][2] = {};
#pragma unroll
        for (int kt = 0; kt < 5; kt++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) mma(o[nt][0], o[nt][1], ft[kt], Bi[kt][nt]);
        if (VAR == 0) {
            // feed the result back so that steps depend on each other only weakly (as the next step's data comes from memory)
            acc += o[0][0] + o[1][1] + o[2][0];
        } else acc += o[0][0] + o[1][1] + o[2][0];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int VAR>
void run(const double *in, double *out, int sms, int warps_per_smsp) {
    int blocks = sms * warps_per_smsp, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<VAR><<<blocks, 128>>>(in, out, 16);
    float best = 1e30f, ms;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k<VAR><<<blocks, 128>>>(in, out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double tf = 512.0 * 45 * iters * (double)blocks * 4 / (best * 1e-3) / 1e12;
    printf(" var%d_w%d=%.1f", VAR, warps_per_smsp, tf);
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    double *in, *out; cudaMalloc(&in, 2048 * 8); cudaMalloc(&out, 8 * 148 * 1024 * 8);
    double h[2048]; for (int i = 0; i < 2048; i++) h[i] = 1.0 + 1e-3 * (i % 97);
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int w : {1, 2, 3, 4}) { run<0>(in, out, prop.multiProcessorCount, w); run<1>(in, out, prop.multiProcessorCount, w); printf("\n"); }
    return 0;
}
