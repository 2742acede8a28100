// fp64_mix.cu — design probes for the 20-state kernel (run on the GPU box):
//   (1) DFMA and DMMA issued from the SAME warps, interleaved: do the two share one FP64 pipe (sum ~ 36 TFLOP/s) or not?
//   (2) DFMA with a uniform-register operand refreshed from the constant bank every instruction (the k_walk_n20 v1 pattern)
//   (3) DFMA with a shared-memory broadcast operand, register-blocked over NP patterns per thread (LDS.128 feeds 2*NP DFMAs)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_mix scripts/fp64_mix.cu && /tmp/fp64_mix
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
#define MMA(c) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));

template <int NF, int NM>   // NF DFMAs and NM DMMAs per inner iteration
__global__ void k_mix(double *out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    double c[8][2] = {};
    double f[8];
    for (int k = 0; k < 8; k++) f[k] = a + k;
    const double m = 1.0000001, d = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k < NM) MMA(c[k])
            if (k < NF) { f[k] = fma(f[k], m, d); }
        }
    }
    double s = 0;
    for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1] + f[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Mat { double m[400]; };
__global__ void k_dfma_const(double *out, const __grid_constant__ Mat M, int iters) {
    double v[20], acc[20];
    for (int k = 0; k < 20; k++) { v[k] = threadIdx.x * 1e-3 + k; acc[k] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int x = 0; x < 20; x++)
#pragma unroll
            for (int i = 0; i < 20; i++) acc[x] = fma(M.m[x * 20 + i], v[i], acc[x]);
#pragma unroll
        for (int k = 0; k < 20; k++) v[k] = acc[k] * 1e-3;
    }
    double s = 0;
    for (int k = 0; k < 20; k++) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NP>
__global__ void k_dfma_smem(double *out, const double *Mg, int iters) {
    __shared__ __align__(16) double sm[400];
    for (int k = threadIdx.x; k < 400; k += blockDim.x) sm[k] = Mg[k];
    __syncthreads();
    double v[NP][20], acc[NP][20];
    for (int p = 0; p < NP; p++)
        for (int k = 0; k < 20; k++) { v[p][k] = threadIdx.x * 1e-3 + k + p; acc[p][k] = 0; }
    for (int it = 0; it < iters; it++) {
        asm volatile("" ::: "memory");   // the matrix is re-read every pass, as in the real kernel (it cannot live in registers)
#pragma unroll
        for (int x = 0; x < 20; x++)
#pragma unroll
            for (int i = 0; i < 20; i += 2) {
                const double2 m = *reinterpret_cast<const double2 *>(sm + x * 20 + i);
#pragma unroll
                for (int p = 0; p < NP; p++) { acc[p][x] = fma(m.x, v[p][i], acc[p][x]); acc[p][x] = fma(m.y, v[p][i + 1], acc[p][x]); }
            }
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int k = 0; k < 20; k++) v[p][k] = acc[p][k] * 1e-3;
    }
    double s = 0;
    for (int p = 0; p < NP; p++) for (int k = 0; k < 20; k++) s += acc[p][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_best(F launch, int reps = 4) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    double best = 1e30; float ms;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    double *Mg; CK(cudaMalloc(&Mg, 400 * 8)); CK(cudaMemset(Mg, 0, 400 * 8));
    Mat M; for (int k = 0; k < 400; k++) M.m[k] = 1e-3 * k;
    const int threads = 256, blocks = sms * 4, iters = 1 << 14;
    const double nthr = (double)blocks * threads, nwarp = nthr / 32;
    printf("{\"gpu\": \"%s\"", prop.name);
#define RUN(NF, NM) { double s = time_best([&] { k_mix<NF, NM><<<blocks, threads>>>(out, iters); }); \
        double tf_f = 2.0 * NF * iters * nthr / s / 1e12, tf_m = 512.0 * NM * iters * nwarp / s / 1e12; \
        printf(", \"mix_f%d_m%d\": {\"dfma_tf\": %.2f, \"dmma_tf\": %.2f, \"sum_tf\": %.2f}", NF, NM, tf_f, tf_m, tf_f + tf_m); }
    RUN(8, 0) RUN(0, 8) RUN(8, 8) RUN(8, 4) RUN(4, 8) RUN(8, 2) RUN(8, 1)
    {
        const int it2 = 2048;
        for (int thr : {128, 256}) {
            int bl = sms * (768 / thr);
            double s = time_best([&] { k_dfma_const<<<bl, thr>>>(out, M, it2); });
            printf(", \"dfma_const_t%d_tf\": %.2f", thr, 2.0 * 400 * it2 * bl * thr / s / 1e12);
        }
        for (int thr : {128, 256}) {
            int bl = sms * (768 / thr);
            double s = time_best([&] { k_dfma_smem<1><<<bl, thr>>>(out, Mg, it2); });
            printf(", \"dfma_smem_np1_t%d_tf\": %.2f", thr, 2.0 * 400 * it2 * bl * thr / s / 1e12);
            bl = sms * (384 / thr);
            s = time_best([&] { k_dfma_smem<2><<<bl, thr>>>(out, Mg, it2); });
            printf(", \"dfma_smem_np2_t%d_tf\": %.2f", thr, 2.0 * 800 * it2 * bl * thr / s / 1e12);
        }
    }
    printf("}\n");
    return 0;
}
