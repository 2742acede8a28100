// b200lk_mixlen.cuh — derivative kernels of trees with MIXED BRANCH LENGTHS (heterotachy, GHOST `+H` / `*H`; PhyloTree::isMixlen(),
// tree/phylotreemixlen.h:89): every branch carries nmixlen = ncat lengths (PhyloNeighborMixlen::getLength(c)).
//   * the per-class length itself needs no kernel of its own: every place the reference reads child->getLength(cat_id[c]) /
//     dad_branch->getLength(mycat) (tree/phylokernelnew.h:1017, 1069, 2325, 2348, 2724, 2738, 3286, 3301) goes through
//     resolve_len_cat() — the info kernels of all walkers, k_vals / stage_vals — so partial likelihoods, the branch log-likelihood and
//     the from-buffer form run on the tuned kernels unchanged;
//   * k_derv_mixlen: computeLikelihoodDerv's mixlen branch (:2391-2460, 2556-2568) — the GRADIENT over the K lengths of the branch,
//     the K x K HESSIAN and the log-likelihood, from the resident theta buffer;
//   * k_derv_mixlen_cur: computeLikelihoodDervMixlen (:3431-3616) — (df, ddf) of the current class's length alone (the EM algorithm of
//     PhyloTreeMixlen::optimizeOneBranch; ptn_freq then holds the class's posterior weights).
#pragma once
#include "b200lk_kernels.cuh"

namespace b200lk {

constexpr int kMaxMixlen = 16;

// dotProductTriple<ADD = true> over one category's n values (:666-725 with N = nstates: two lane accumulators, odd tail)
__device__ __forceinline__ void dot_triple_cat(const double *v0, const double *v1, const double *v2, const double *theta, int n, double &x, double &y, double &z) {
    double ad[2] = {v0[0] * theta[0], n > 1 ? v0[1] * theta[kWarp] : 0.0};
    double bd[2] = {v1[0] * theta[0], n > 1 ? v1[1] * theta[kWarp] : 0.0};
    double cd[2] = {v2[0] * theta[0], n > 1 ? v2[1] * theta[kWarp] : 0.0};
    const int stop = (n & 1) ? n - 1 : n;
    for (int k = 2; k < stop; k += 2) {
        const double t0 = theta[(long long)k * kWarp], t1 = theta[(long long)(k + 1) * kWarp];
        ad[0] = fma(v0[k], t0, ad[0]); ad[1] = fma(v0[k + 1], t1, ad[1]);
        bd[0] = fma(v1[k], t0, bd[0]); bd[1] = fma(v1[k + 1], t1, bd[1]);
        cd[0] = fma(v2[k], t0, cd[0]); cd[1] = fma(v2[k + 1], t1, cd[1]);
    }
    if (n & 1) {
        if (n == 1) { x = ad[0]; y = bd[0]; z = cd[0]; return; }
        const double t = theta[(long long)(n - 1) * kWarp];
        x = fma(v0[n - 1], t, ad[0] + ad[1]); y = fma(v1[n - 1], t, bd[0] + bd[1]); z = fma(v2[n - 1], t, cd[0] + cd[1]);
    } else { x = ad[0] + ad[1]; y = bd[0] + bd[1]; z = cd[0] + cd[1]; }
}

// sum of v over the CTA's kTile threads -> thread 0 (fixed order: lanes by shuffle tree, then warps in order)
__device__ __forceinline__ double cta_sum(double v, double *s_w) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kTile / 32; w++) t += s_w[w];
    return t;
}

// red[block][K + K*K + 1]: gradient, Hessian (row-major), log-likelihood; k_mixlen_reduce adds the blocks in order
__global__ void __launch_bounds__(kTile) k_derv_mixlen(Dev D, double length, int fuse_vals, double *red) {
    extern __shared__ double s_vals_dyn[];
    __shared__ double s_w[kTile / 32];
    const double *vals = fuse_vals ? stage_vals(D, length, 1, s_vals_dyn) : D.val;
    const int n = D.n, C = D.C, B = D.B, K = D.mixlen;
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    const double *theta = D.theta + pbase(p, B);
    double df[kMaxMixlen], ddf[kMaxMixlen];
#pragma unroll
    for (int i = 0; i < kMaxMixlen; i++) df[i] = ddf[i] = 0.0;
    double lh = 0.0;
    for (int c = 0; c < C; c++) {
        double x, y, z;
        dot_triple_cat(vals + c * n, vals + B + c * n, vals + 2 * B + c * n, theta + (long long)c * n * kWarp, n, x, y, z);
        lh += x;
        const int i = c % D.ncat;   // cat_id[c]
#pragma unroll
        for (int k = 0; k < kMaxMixlen; k++)
            if (k == i) { df[k] += y; ddf[k] += z; }
    }
    lh = fabs(lh) + D.invar[p];
    const bool live = p < D.P;
    const double freq = live ? D.freq[p] : 0.0, inv = 1.0 / lh;
#pragma unroll
    for (int k = 0; k < kMaxMixlen; k++) { df[k] *= inv; ddf[k] *= inv; }
    double *out = red + (long long)blockIdx.x * (K + K * K + 1);
    for (int i = 0; i < K; i++) {
        double gi = 0.0;
#pragma unroll
        for (int k = 0; k < kMaxMixlen; k++) if (k == i) gi = df[k];
        const double s = cta_sum(live ? gi * freq : 0.0, s_w);
        if (threadIdx.x == 0) out[i] = s;
        for (int j = 0; j < K; j++) {
            double gj = 0.0, hj = 0.0;
#pragma unroll
            for (int k = 0; k < kMaxMixlen; k++) if (k == j) { gj = df[k]; hj = ddf[k]; }
            const double h = i == j ? fma(-gi, gi, hj) * freq : -(gi * gj * freq);   // :2437-2444
            const double t = cta_sum(live ? h : 0.0, s_w);
            if (threadIdx.x == 0) out[K + i * K + j] = t;
        }
    }
    const double l = cta_sum(live ? (log(lh) + D.buffer_scale[p]) * freq : 0.0, s_w);
    if (threadIdx.x == 0) out[K + K * K] = l;
}

__global__ void __launch_bounds__(256) k_mixlen_reduce(int nblocks, int nv, const double *red, double *out) {
    __shared__ double s[256];
    for (int j = 0; j < nv; j++) {
        double a = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += 256) a += red[(long long)b * nv + j];
        s[threadIdx.x] = a;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[j] = s[0];
        __syncthreads();
    }
}

// computeLikelihoodDervMixlen (:3431-3616): class cur's slice of theta ((cur * nmix + c) * n, as the reference indexes it), values
// exp(eval * len_cur) * weight without rate or category proportion
__global__ void __launch_bounds__(kTile) k_derv_mixlen_cur(Dev D, double length, int cur, double *red, unsigned *ticket, double *result) {
    extern __shared__ double s_v[];   // val0 | val1 | val2, nmix * n each
    const int n = D.n, B = D.B;
    const int fused = D.denom == 1 && D.nmix > 1;
    const int nmix = D.denom == 1 ? 1 : D.nmix;
    const int W = nmix * n;
    const double len = resolve_len_cat(D, length, cur);   // :3508
    for (int k = threadIdx.x; k < W; k += blockDim.x) {
        const int c = k / n, i = k - c * n;
        const int cur_mix = fused ? cur : c;
        const double cof = D.eval[cur_mix * n + i], val = exp(cof * len) * D.mix_weight[cur_mix], v1 = cof * val;
        s_v[k] = val; s_v[W + k] = v1; s_v[2 * W + k] = cof * v1;
    }
    __syncthreads();
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    const double *theta = D.theta + pbase(p, B) + (long long)cur * n * nmix * kWarp;   // :3548
    double lh = 0.0, df = 0.0, ddf = 0.0;
    for (int c = 0; c < nmix; c++) {
        double x, y, z;
        dot_triple_cat(s_v + c * n, s_v + W + c * n, s_v + 2 * W + c * n, theta + (long long)c * n * kWarp, n, x, y, z);
        lh += x; df += y; ddf += z;
    }
    lh = fabs(lh) + D.invar[p];
    double acc[2] = {0.0, 0.0};
    if (p < D.P) {
        const double inv = 1.0 / lh, freq = D.freq[p];
        df *= inv; ddf *= inv;
        acc[0] = df * freq;
        acc[1] = fma(-df, df, ddf) * freq;
    }
    block_reduce_finish<2>(acc, red, ticket, result);
}

}  // namespace b200lk
