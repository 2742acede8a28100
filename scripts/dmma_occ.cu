// dmma_occ.cu — DMMA (mma.sync m8n8k4.f64) throughput against warps per SM sub-partition and independent chains per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_occ scripts/dmma_occ.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define MMA(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
template <int NCH>
__global__ void k(double *out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    double c[NCH][2] = {};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k2 = 0; k2 < NCH; k2++) MMA(c[k2])
    }
    double s = 0;
    for (int k2 = 0; k2 < NCH; k2++) s += c[k2][0] + c[k2][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NCH>
void run(double *out, int sms, int warps_per_sm) {
    int threads = 128, blocks = sms * warps_per_sm / 4;
    int iters = 1 << 13;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NCH><<<blocks, threads>>>(out, 64);
    float best = 1e30f, ms;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k<NCH><<<blocks, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double tf = 512.0 * NCH * iters * (double)blocks * 4 / (best * 1e-3) / 1e12;
    printf(" w%d_ch%d=%.1f", warps_per_sm / 4, NCH, tf);
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    double *out; cudaMalloc(&out, 8 * 148 * 2048 * 8);
    for (int w : {4, 8, 12, 16, 24, 32}) { run<1>(out, prop.multiProcessorCount, w); run<3>(out, prop.multiProcessorCount, w); run<6>(out, prop.multiProcessorCount, w); run<12>(out, prop.multiProcessorCount, w); printf("\n"); }
    return 0;
}
