// store_probe.cu — write-only bandwidth in the access pattern of the 4-state walker: each 128-thread CTA owns a 128-pattern tile
// and, op after op, writes that tile's 16 KB of a different CLV slot (slots 128 MB apart).  Variants: 8-byte stores (k*256 B
// apart per thread, as the warp-interleaved layout gives) vs 16-byte stores; walker order vs slot-major streaming order.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/store_probe.bin scripts/store_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int VEC>
__global__ void __launch_bounds__(128, 4) k(double *base, long long slot_stride, int nslots, long long ntiles, unsigned long long *counter, double v) {
    __shared__ long long s_item;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = (long long)atomicAdd(counter, 1ULL);
        __syncthreads();
        const long long tile = s_item;
        if (tile >= ntiles) break;
        const long long p = tile * 128 + threadIdx.x;
        const long long pb = (p >> 5) * 16 * 32 + (p & 31);
        for (int o = 0; o < nslots; o++) {
            double *O = base + (long long)o * slot_stride;
            if (VEC == 1) {
#pragma unroll
                for (int kk = 0; kk < 16; kk++) O[pb + kk * 32] = v + kk;
            } else {   // 16-byte stores: thread t of the warp writes doubles [2t, 2t+1] of each 64-double row pair
                double2 *O2 = reinterpret_cast<double2 *>(O + (p >> 5) * 16 * 32) + (p & 31);
#pragma unroll
                for (int kk = 0; kk < 8; kk++) O2[kk * 32] = make_double2(v + kk, v - kk);
            }
        }
    }
}
int main() {
    const long long P = 1 << 20, slot_stride = P * 16;      // 128 MB per slot
    const int nslots = 400;                                  // 51 GB
    double *buf; cudaMalloc(&buf, sizeof(double) * slot_stride * nslots);
    unsigned long long *ctr; cudaMalloc(&ctr, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int vec = 1; vec <= 2; vec++) {
        float best = 1e30f, ms;
        for (int r = 0; r < 3; r++) {
            cudaMemset(ctr, 0, 8);
            cudaEventRecord(e0);
            if (vec == 1) k<1><<<148 * 4, 128>>>(buf, slot_stride, nslots, P / 128, ctr, 1.0);
            else k<2><<<148 * 4, 128>>>(buf, slot_stride, nslots, P / 128, ctr, 1.0);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("walker-order stores, %d-byte: %.1f GB/s (%.2f ms)\n", vec * 8, 8.0 * slot_stride * nslots / best / 1e6, best);
    }
    float ms; cudaEventRecord(e0); cudaMemsetAsync(buf, 0, sizeof(double) * slot_stride * nslots); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemset same bytes: %.1f GB/s\n", 8.0 * slot_stride * nslots / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
