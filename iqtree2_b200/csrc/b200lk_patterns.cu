// b200lk_patterns.cu — site-pattern compression on the device: the distinct alignment columns in first-occurrence order, their
// frequencies and the site -> pattern map, i.e. what Alignment::addPattern builds one site at a time through a hash map
// (alignment/alignment.cpp:1092-1133; callers buildPattern :1725 and createBootstrapAlignment, which compresses a fresh
// alignment for every standard-bootstrap replicate).  Columns are hashed (one thread per site, coalesced over the
// sequence-major states), radix-sorted by hash with the site index as stable tie-break (cub::DeviceRadixSort — a library
// sort on a set-up path, not a likelihood kernel), runs of identical columns become groups, and the groups are ranked by their
// first site.  Equal hashes are always verified by comparing the columns; two different columns with the same 64-bit hash
// would show up as adjacent groups with equal hash and make the call retry with another hash seed.
#include "b200lk.h"

#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

namespace {

__global__ void k_col_hash(const uint8_t *states, int ntaxa, long long nsite, unsigned long long seed, unsigned long long *h, unsigned *idx) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsite) return;
    unsigned long long a = 0x9E3779B97F4A7C15ULL ^ seed;
    for (int t = 0; t < ntaxa; t++) {
        a ^= states[(long long)t * nsite + s];
        a *= 0x100000001B3ULL;            // FNV-1a step
        a ^= a >> 29;
    }
    a *= 0xBF58476D1CE4E5B9ULL; a ^= a >> 32;
    h[s] = a;
    idx[s] = (unsigned)s;
}

__global__ void k_flag_runs(const uint8_t *states, int ntaxa, long long nsite, const unsigned long long *h, const unsigned *idx,
                            unsigned *flag, int *collision) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsite) return;
    unsigned f = 1;
    if (i > 0 && h[i] == h[i - 1]) {
        const unsigned a = idx[i - 1], b = idx[i];
        f = 0;
        for (int t = 0; t < ntaxa; t++)
            if (states[(long long)t * nsite + a] != states[(long long)t * nsite + b]) { f = 1; break; }
        if (f) *collision = 1;   // same hash, different columns: groups may interleave -> retry with another seed
    }
    flag[i] = f;
}

__global__ void k_group_first(long long nsite, const unsigned *flag, const unsigned *gid, const unsigned *idx, unsigned *first, unsigned *gnum) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsite || !flag[i]) return;
    first[gid[i] - 1] = idx[i];      // stable sort: the first member of a run is its smallest site
    gnum[gid[i] - 1] = gid[i] - 1;
}

__global__ void k_rank(long long ngroups, const unsigned *sorted_group, unsigned *pid_of_group) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < ngroups) pid_of_group[sorted_group[r]] = (unsigned)r;
}

__global__ void k_assign(long long nsite, const unsigned *gid, const unsigned *idx, const unsigned *pid_of_group, int *site_pattern, double *freq) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsite) return;
    const unsigned pid = pid_of_group[gid[i] - 1];
    site_pattern[idx[i]] = (int)pid;
    atomicAdd(freq + pid, 1.0);      // integer counts: exact in any order
}

__global__ void k_gather(const uint8_t *states, int ntaxa, long long nsite, long long nptn, const unsigned *first_sorted, uint8_t *patterns) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nptn * ntaxa) return;
    const long long t = k / nptn, p = k - t * nptn;
    patterns[k] = states[t * nsite + first_sorted[p]];
}

struct DevBuf {
    void *p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace

extern "C" int b200lk_compress_patterns(int32_t device, const uint8_t *states, int32_t ntaxa, int64_t nsite, uint8_t *patterns,
                                        double *ptn_freq, int32_t *site_pattern, int64_t *nptn_out) {
    if (!states || !patterns || !ptn_freq || !site_pattern || !nptn_out || ntaxa < 1 || nsite < 1 || nsite > 0x7fffffffLL) return B200LK_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return B200LK_ERR_CUDA;
    const long long n = nsite;
    DevBuf d_states, d_h, d_h2, d_idx, d_idx2, d_flag, d_gid, d_first, d_first2, d_gnum, d_gnum2, d_pid, d_sp, d_freq, d_pat, d_tmp, d_coll;
#define A(buf, bytes) if (buf.alloc(bytes) != cudaSuccess) return B200LK_ERR_NOMEM
    A(d_states, (size_t)ntaxa * n); A(d_h, 8 * n); A(d_h2, 8 * n); A(d_idx, 4 * n); A(d_idx2, 4 * n); A(d_flag, 4 * n); A(d_gid, 4 * n);
    A(d_first, 4 * n); A(d_first2, 4 * n); A(d_gnum, 4 * n); A(d_gnum2, 4 * n); A(d_pid, 4 * n); A(d_sp, 4 * n); A(d_freq, 8 * n);
    A(d_pat, (size_t)ntaxa * n); A(d_coll, 4);
    size_t tmp1 = 0, tmp2 = 0, tmp3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp1, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (unsigned *)nullptr, (unsigned *)nullptr, (int)n);
    cub::DeviceScan::InclusiveSum(nullptr, tmp2, (unsigned *)nullptr, (unsigned *)nullptr, (int)n);
    cub::DeviceRadixSort::SortPairs(nullptr, tmp3, (unsigned *)nullptr, (unsigned *)nullptr, (unsigned *)nullptr, (unsigned *)nullptr, (int)n);
    A(d_tmp, std::max(tmp1, std::max(tmp2, tmp3)));
#undef A
    size_t tmp_bytes = std::max(tmp1, std::max(tmp2, tmp3));
    if (cudaMemcpy(d_states.p, states, (size_t)ntaxa * n, cudaMemcpyHostToDevice) != cudaSuccess) return B200LK_ERR_CUDA;
    const int T = 256;
    const unsigned G = (unsigned)((n + T - 1) / T);
    unsigned ngroups = 0;
    for (int attempt = 0; attempt < 4; attempt++) {
        k_col_hash<<<G, T>>>((const uint8_t *)d_states.p, ntaxa, n, 0x51ED27ULL * (unsigned long long)(attempt + 1), (unsigned long long *)d_h.p, (unsigned *)d_idx.p);
        cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, (const unsigned long long *)d_h.p, (unsigned long long *)d_h2.p, (const unsigned *)d_idx.p,
                                        (unsigned *)d_idx2.p, (int)n);
        cudaMemset(d_coll.p, 0, 4);
        k_flag_runs<<<G, T>>>((const uint8_t *)d_states.p, ntaxa, n, (const unsigned long long *)d_h2.p, (const unsigned *)d_idx2.p, (unsigned *)d_flag.p, (int *)d_coll.p);
        int coll = 0;
        if (cudaMemcpy(&coll, d_coll.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return B200LK_ERR_CUDA;
        if (coll) { if (attempt == 3) return B200LK_ERR_STATE; continue; }
        cub::DeviceScan::InclusiveSum(d_tmp.p, tmp_bytes, (const unsigned *)d_flag.p, (unsigned *)d_gid.p, (int)n);
        if (cudaMemcpy(&ngroups, (unsigned *)d_gid.p + (n - 1), 4, cudaMemcpyDeviceToHost) != cudaSuccess) return B200LK_ERR_CUDA;
        break;
    }
    k_group_first<<<G, T>>>(n, (const unsigned *)d_flag.p, (const unsigned *)d_gid.p, (const unsigned *)d_idx2.p, (unsigned *)d_first.p, (unsigned *)d_gnum.p);
    cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, (const unsigned *)d_first.p, (unsigned *)d_first2.p, (const unsigned *)d_gnum.p, (unsigned *)d_gnum2.p,
                                    (int)ngroups);
    const unsigned GG = (ngroups + T - 1) / T;
    k_rank<<<GG, T>>>(ngroups, (const unsigned *)d_gnum2.p, (unsigned *)d_pid.p);
    cudaMemset(d_freq.p, 0, 8 * (size_t)ngroups);
    k_assign<<<G, T>>>(n, (const unsigned *)d_gid.p, (const unsigned *)d_idx2.p, (const unsigned *)d_pid.p, (int *)d_sp.p, (double *)d_freq.p);
    const long long np = (long long)ngroups * ntaxa;
    k_gather<<<(unsigned)((np + T - 1) / T), T>>>((const uint8_t *)d_states.p, ntaxa, n, ngroups, (const unsigned *)d_first2.p, (uint8_t *)d_pat.p);
    if (cudaGetLastError() != cudaSuccess) return B200LK_ERR_CUDA;
    if (cudaMemcpy(patterns, d_pat.p, (size_t)np, cudaMemcpyDeviceToHost) != cudaSuccess) return B200LK_ERR_CUDA;
    if (cudaMemcpy(ptn_freq, d_freq.p, 8 * (size_t)ngroups, cudaMemcpyDeviceToHost) != cudaSuccess) return B200LK_ERR_CUDA;
    if (cudaMemcpy(site_pattern, d_sp.p, 4 * (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) return B200LK_ERR_CUDA;
    *nptn_out = ngroups;
    return 0;
}
