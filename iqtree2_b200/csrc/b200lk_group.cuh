// b200lk_group.cuh — GROUPED launches: the kernels of one evaluation for MANY engines (the partition trees of an edge-linked
// partition model, tree/phylosupertreeplen.cpp:282-290, 322-341, 380-409) in O(1) launches.
//
// The reference evaluates its partition trees one after the other from an OpenMP loop, each over its slice of ONE pooled arena
// with per-partition offsets (tree/phylosupertreeplen.cpp:1624-1720).  Round 1 mirrored the loop — one engine, one stream and
// ~5 launches per partition: 2 500 launches per evaluation of BASELINE config 5, 8.5 of 13.3 ms spent queueing on the host.
// Here the member engines of a group (same device, same kernel family: 4 states x C categories, or 20 states) are described by a
// table of GroupEntry records in device memory — the arena's descriptor table: every entry carries its engine's buffer pointers —
// and ONE launch per kernel type serves them all:
//   * walkers: a work item is (member, pattern tile); items are handed out by one atomic counter, most expensive members first;
//   * root / theta / derivative / from-buffer: the grid is the concatenation of the members' 128-pattern blocks, every member
//     reduces into ITS OWN partial-sum array, ticket and (mapped host) result slot, in the same fixed order as the single-engine
//     kernels (bit-identical results);
//   * the per-op staging (exp, tip tables, transition matrices) and the state-space -> eigen-space conversion likewise.
// The bodies are the single-engine kernels' own device functions (root_body, walk4_tile, ...): no second implementation.
#pragma once
#include "b200lk_kernels.cuh"
#include "b200lk_walk20ss.cuh"
#include "b200lk_walk20v2.cuh"
#include "b200lk_walk4.cuh"

namespace b200lk {

struct GroupEntry {
    Dev D;                    // the member's buffers; D.opinfo = its staging blocks
    Model4 M;                 // 4-state members: evec | inv_evec (register operands of the walker)
    long long job_off;        // first WalkJob of this member in the group's job array
    int nops;                 // ops of this call (0: nothing to update)
    int item0;                // first walker work item of this member
    int nitems;               // its walker work items (tiles)
    int blk0;                 // first 128-pattern block of this member in the streaming kernels' grids
    int op0;                  // first CTA of this member in the info kernel's grid
    int node_slot, dad_kind, dad_id, save_log;   // the evaluation branch (eigen-space slots for state-space members)
    int do_theta;                                 // derivative calls: build theta_all first
    int conv_src[2], conv_dst[2];                 // state-space -> eigen-space conversions of this call (-1: none)
    int pad_;
    double length;
    double *red;
    unsigned *ticket;
    double *result;           // mapped host memory: [0] logL, [1] df, [2] ddf
};

// the member whose range [first, first + count) contains `item`; `key(e)` = first item of entry e, entries sorted by it
template <typename Key>
__device__ __forceinline__ int group_find(const GroupEntry *E, int n, int item, Key key) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (key(E[mid]) <= item) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// a CTA's private copy of its member's Dev (every thread helps; words)
__device__ __forceinline__ void group_load_dev(Dev *sD, const Dev *src) {
    const unsigned *s = reinterpret_cast<const unsigned *>(src);
    unsigned *d = reinterpret_cast<unsigned *>(sD);
    for (int k = threadIdx.x; k < (int)(sizeof(Dev) / 4); k += blockDim.x) d[k] = s[k];
}

// ---- per-op staging --------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_group_info_n4(const GroupEntry *E, int n, const WalkJob4 *jobs, unsigned long long *counter) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0ULL;
    extern __shared__ double s_exp4[];
    __shared__ Dev sD;
    const int e = group_find(E, n, blockIdx.x, [](const GroupEntry &g) { return g.op0; });
    group_load_dev(&sD, &E[e].D);
    __syncthreads();
    const int op = blockIdx.x - E[e].op0;
    walk_info4_body(sD, jobs[E[e].job_off + op], sD.opinfo + (size_t)op * stage4_doubles(sD.C), s_exp4);
}

__global__ void __launch_bounds__(256) k_group_info_n20ss(const GroupEntry *E, int n, const WalkJob20 *jobs, unsigned long long *counter) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0ULL;
    extern __shared__ double s_exp20[];
    __shared__ Dev sD;
    const int e = group_find(E, n, blockIdx.x, [](const GroupEntry &g) { return g.op0; });
    group_load_dev(&sD, &E[e].D);
    __syncthreads();
    const int nblk = sD.nmix, cpb = sD.C / nblk;
    const int vo = blockIdx.x - E[e].op0, op = vo / nblk, m = vo - op * nblk;
    walk_info20ss_body(sD, jobs[E[e].job_off + op], sD.opinfo + (size_t)vo * ss_stage_doubles(cpb, sD.su), m, s_exp20);
}

// ---- walkers ------------------------------------------------------------------------------------------------------------------
template <int C, bool SAFE>
__global__ void __launch_bounds__(kTile, (C <= 4 ? 3 : 2)) k_group_walk_n4(const GroupEntry *E, int n, int total_items, unsigned long long *counter) {
    __shared__ __align__(16) double s_stage[3][stage4_doubles(C)];
    __shared__ double s_tip[kMaxLeafRows4 * 4];
    __shared__ __align__(16) unsigned char s_ts[2][2][kTile];
    __shared__ Dev sD;
    __shared__ int s_item, s_ent;
    const int tid = threadIdx.x;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const int item = (int)atomicAdd(counter, 1ULL);
            s_item = item;
            s_ent = item < total_items ? group_find(E, n, item, [](const GroupEntry &g) { return g.item0; }) : -1;
        }
        __syncthreads();
        if (s_ent < 0) break;
        const GroupEntry &ge = E[s_ent];
        group_load_dev(&sD, &ge.D);
        const Model4 M = ge.M;   // registers
        __syncthreads();
        for (int k = tid; k < (sD.su + 1) * 4; k += kTile) s_tip[k] = sD.tip_lh[k];
        __syncthreads();
        if (ge.nops > 0) walk4_tile<C, SAFE>(sD, M, sD.opinfo, ge.nops, (long long)(s_item - ge.item0), s_stage, s_tip, s_ts);
    }
}

// 20-state members: the TMA-ring walker's tile function over (member, tile) items; tiles of a member are balanced over its
// 32-pattern blocks, at most NW per tile
template <bool SAFE, bool MIX, int NW, int SD>
__global__ void __launch_bounds__(NW * 32, 1) k_group_walk_n20v2(const GroupEntry *E, int n, int total_items, unsigned long long *counter) {
    extern __shared__ __align__(16) double s_dyn[];
    __shared__ __align__(8) V2Shared<NW, MIX, SD> sh;
    __shared__ Dev sD;
    __shared__ int s_item, s_ent;
    if (threadIdx.x == 0)
        for (int k = 0; k < SD; k++) { mbar_init(&sh.full[k], 1); sh.release[k] = 0; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    unsigned seq = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const int item = (int)atomicAdd(counter, 1ULL);
            s_item = item;
            s_ent = item < total_items ? group_find(E, n, item, [](const GroupEntry &g) { return g.item0; }) : -1;
        }
        __syncthreads();
        if (s_ent < 0) break;
        const GroupEntry &ge = E[s_ent];
        group_load_dev(&sD, &ge.D);
        __syncthreads();
        const long long nblocks = sD.Ppad >> 5, t = s_item - ge.item0, nt = ge.nitems;
        if (ge.nops > 0) walk20v2_tile<SAFE, MIX, NW, 1, SD>(sD, sD.opinfo, ge.nops, t * nblocks / nt, (t + 1) * nblocks / nt, s_dyn, sh, seq);
    }
}

// ---- state-space -> eigen-space (root / theta operands of 20-state members) ---------------------------------------------------
__global__ void __launch_bounds__(128) k_group_clv20_to_eigen(const GroupEntry *E, int n, int total_blocks) {
    __shared__ Dev sD;
    __shared__ int s_ent;
    for (int blk = blockIdx.x; blk < total_blocks; blk += gridDim.x) {   // one 128-pattern block (4 warps x 32 patterns) per iteration
        __syncthreads();
        if (threadIdx.x == 0) s_ent = group_find(E, n, blk, [](const GroupEntry &g) { return g.blk0; });
        __syncthreads();
        const GroupEntry &ge = E[s_ent];
        group_load_dev(&sD, &ge.D);
        __syncthreads();
        const long long wblk = (long long)(blk - ge.blk0) * 4 + (threadIdx.x >> 5);   // this warp's 32-pattern block of the member
        for (int k = 0; k < 2; k++)
            if (ge.conv_src[k] >= 0) clv20_to_eigen_block(sD, ge.conv_src[k], ge.conv_dst[k], wblk);
    }
}

// ---- streaming kernels ----------------------------------------------------------------------------------------------------------
template <bool SAFE>
__global__ void __launch_bounds__(kTile) k_group_root(const GroupEntry *E, int n, int fuse_vals) {
    extern __shared__ double s_vals_dyn[];
    __shared__ Dev sD;
    const int e = group_find(E, n, blockIdx.x, [](const GroupEntry &g) { return g.blk0; });
    const GroupEntry &ge = E[e];
    group_load_dev(&sD, &ge.D);
    __syncthreads();
    root_body<SAFE, false>(sD, blockIdx.x - ge.blk0, (int)(sD.Ppad / kTile), ge.node_slot, ge.dad_kind, ge.dad_id, ge.save_log, ge.red, ge.ticket, ge.result,
                           ge.length, fuse_vals, s_vals_dyn);
}

template <bool SAFE>
__global__ void __launch_bounds__(kTile) k_group_theta(const GroupEntry *E, int n) {
    __shared__ Dev sD;
    const int e = group_find(E, n, blockIdx.x, [](const GroupEntry &g) { return g.blk0; });
    const GroupEntry &ge = E[e];
    if (!ge.do_theta) return;
    group_load_dev(&sD, &ge.D);
    __syncthreads();
    theta_body<SAFE>(sD, blockIdx.x - ge.blk0, ge.node_slot, ge.dad_kind, ge.dad_id);
}

template <int BT>
__global__ void __launch_bounds__(kTile) k_group_derv(const GroupEntry *E, int n, int fuse_vals) {
    extern __shared__ double s_vals_dyn[];
    __shared__ Dev sD;
    const int e = group_find(E, n, blockIdx.x, [](const GroupEntry &g) { return g.blk0; });
    const GroupEntry &ge = E[e];
    group_load_dev(&sD, &ge.D);
    __syncthreads();
    derv_body<BT, false>(sD, blockIdx.x - ge.blk0, (int)(sD.Ppad / kTile), ge.red, ge.ticket, ge.result + 1, ge.length, fuse_vals, s_vals_dyn);
}

template <int BT>
__global__ void __launch_bounds__(kTile) k_group_from_buffer(const GroupEntry *E, int n, int fuse_vals) {
    extern __shared__ double s_vals_dyn[];
    __shared__ Dev sD;
    const int e = group_find(E, n, blockIdx.x, [](const GroupEntry &g) { return g.blk0; });
    const GroupEntry &ge = E[e];
    group_load_dev(&sD, &ge.D);
    __syncthreads();
    from_buffer_body<BT, false>(sD, blockIdx.x - ge.blk0, (int)(sD.Ppad / kTile), ge.red, ge.ticket, ge.result, ge.length, fuse_vals, s_vals_dyn);
}

}  // namespace b200lk
