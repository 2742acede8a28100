// fp64_peaks.cu — measures the FP64 roofline denominators MEASURED_PEAKS.json does not carry:
//   (1) DFMA (vector FP64 FMA) throughput, (2) DMMA m8n8k4 (FP64 tensor) throughput, (3) a read+write copy.
// Build & run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peaks scripts/fp64_peaks.cu && /tmp/fp64_peaks
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void k_dfma(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void k_dmma(double *out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0}, c4[2] = {0, 0}, c5[2] = {0, 0}, c6[2] = {0, 0}, c7[2] = {0, 0};
    for (int i = 0; i < iters; i++) {
#define MMA(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
        MMA(c0) MMA(c1) MMA(c2) MMA(c3) MMA(c4) MMA(c5) MMA(c6) MMA(c7)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[1] + c2[0] + c3[1] + c4[0] + c5[1] + c6[0] + c7[1];
}

__global__ void k_copy(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    double *out;
    CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int threads : {256, 512, 1024}) {
        int blocks = sms * (2048 / threads);
        int iters = 1 << 16;
        k_dfma<<<blocks, threads>>>(out, 1024);
        double best = 0;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            k_dfma<<<blocks, threads>>>(out, iters);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
            double tf = 2.0 * 8 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
            if (tf > best) best = tf;
        }
        printf(", \"dfma_tflops_t%d\": %.2f", threads, best);
    }
    for (int threads : {128, 256, 512}) {
        int blocks = sms * (1024 / threads);
        int iters = 1 << 14;
        k_dmma<<<blocks, threads>>>(out, 64);
        double best = 0;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            k_dmma<<<blocks, threads>>>(out, iters);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
            // per warp-level mma: 8*8*4 FMAs = 512 flops
            double tf = 512.0 * 8 * (double)iters * blocks * (threads / 32) / (ms * 1e-3) / 1e12;
            if (tf > best) best = tf;
        }
        printf(", \"dmma_tflops_t%d\": %.2f", threads, best);
    }
    {
        size_t n = (size_t)1 << 28;   // 256 Mi double2 = 4 GiB each way
        double2 *a, *b;
        CK(cudaMalloc(&a, n * sizeof(double2)));
        CK(cudaMalloc(&b, n * sizeof(double2)));
        CK(cudaMemset(a, 1, n * sizeof(double2)));
        double best = 0;
        for (int rep = 0; rep < 6; rep++) {
            cudaEventRecord(e0);
            k_copy<<<sms * 16, 512>>>(a, b, n);
            cudaEventRecord(e1);
            CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1);
            double gbs = 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9;
            if (gbs > best) best = gbs;
        }
        printf(", \"copy_gbs\": %.1f", best);
        // write-only and read-only streams
        cudaEventRecord(e0);
        CK(cudaMemsetAsync(b, 0, n * sizeof(double2)));
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        printf(", \"memset_gbs\": %.1f", n * sizeof(double2) / (ms * 1e-3) / 1e9);
    }
    printf("}\n");
    return 0;
}
