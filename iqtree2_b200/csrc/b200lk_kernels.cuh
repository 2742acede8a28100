// b200lk_kernels.cuh — hand-written sm_100a kernels of the B200 likelihood engine.
//
// Device data layout: every per-pattern vector is WARP-INTERLEAVED, i.e. the reference's
// SIMD layout (tree/phylokernelnew.h:1914-1916) with V = 32:
//     partial(slot, p, k) = partial[slot*Ppad*B + (p/32)*B*32 + k*32 + p%32]      k = c*n + i
// so that a warp working on 32 consecutive patterns reads/writes 256 contiguous bytes per element index k,
// transition matrices are warp-uniform operands, and per-pattern reductions (scaling max, lh, df, ddf) never
// leave the thread.  Patterns are padded to a multiple of kTile with all-unknown states and zero frequency.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200lk {

constexpr int kWarp = 32;
constexpr int kTile = 128;                                   // patterns per CTA tile == threads per CTA
constexpr double kScalingThreshold = 8.636168555094444625386e-78;   // tree/phylotree.h:77  (2^-256)
constexpr double kLogScalingThreshold = -177.4456782233459932741;   // tree/phylotree.h:80
constexpr double kTwo256 = 0x1p256;                          // ldexp(x, SCALING_THRESHOLD_EXP)

// one TraversalInfo after host-side normalisation (a single leaf child is always `left`,
// tree/phylokernelnew.h:1395-1405)
struct MultiChild {   // one child of a multifurcating node (all of its children, in neighbour order)
    int32_t kind, id;     // 0 = tip (taxon row), 1 = CLV slot
    int64_t e_off;        // echild[c][x][i] block inside opinfo
    int64_t leaf_off;     // partial_lh_leaves table inside opinfo, -1 for an internal child
};

struct DevOp {
    int32_t dst;     // destination CLV slot
    int32_t ncase;   // 0 tip-tip, 1 tip-inner, 2 inner-inner, 3 multifurcating (children in Dev::mchild[mc_off .. mc_off+nchild))
    int32_t lid;     // taxon row (tip) or slot
    int32_t rid;
    int64_t eL, eR;        // offsets (in doubles) of echild[c][x][i] blocks inside opinfo
    int64_t leafL, leafR;  // offsets of partial_lh_leaves[state][c][x] tables, -1 if the child is internal
    int32_t nchild, pad_;
    int64_t mc_off;
};

struct ChildJob {
    double length;
    int64_t e_off;
    int64_t leaf_off;  // -1: internal child, no table
};

struct Dev {
    int n, C, ncat, nmix, denom, B, S, su;
    int quirks;              // B200LK_QUIRK_* (include/b200lk.h)
    int asc;                 // +ASC: 1 = Lewis' correction, 2 = Holder's (+ASC_MIS); patterns [Pobs, P) are the unobserved constant patterns
    long long Pobs, asc_nsite;   // Pobs = P without +ASC; asc_nsite = aln->getNSite()
    long long asc_step;      // Holder: the unobserved block is n groups of asc_step patterns, group s = pattern p with state s
    double *asc_tmp;         // Holder: (lh, df, ddf) of every unobserved pattern, three arrays of P - Pobs doubles
    int clv_f20;             // CLV slots use the 20-state fragment-major layout (clv20_off) instead of the warp-interleaved one
    int clv_f8, n8;          // general state counts: fragment-major layout with n padded to n8 = 8*ceil(n/8) (b200lk_walkgen.cuh)
    int clv_state;           // CLV slots hold the STATE-SPACE vector t (partial_lh = inv_evec . t), see b200lk_walk20ss.cuh
    unsigned ss_lo, ss_hi;   // exponent words: max|t| < ss_lo => max|partial_lh| < 2^-256 for sure, >= ss_hi => certainly not
    const double *inv_frag20;   // inv_evec as DMMA B fragments, [nmix][15][32] (k_inv_frag20)
    long long P, Ppad;
    long long slot_stride;   // Ppad * B   doubles per CLV slot
    long long scale_stride;  // Ppad * S   uint16 per slot
    double *partial;
    uint16_t *scale;
    const uint8_t *tips;     // [ntaxa][Ppad]
    const double *freq, *invar;
    const double *eval, *evec, *inv_evec, *tip_lh, *cat_rate, *cat_prop, *mix_weight;
    double *opinfo;
    double *theta, *buffer_scale, *pattern_lh, *pattern_scaling, *pattern_lh_cat;
    double *val;             // val0 | val1 | val2, B doubles each
    const MultiChild *mchild;   // children of the multifurcating ops of the current call
    const double *dev_len;      // device-resident branch lengths (asynchronous branch sweep): a child length < 0 in an op means
                                // "slot -(length) - 1 of this table", written by an earlier k_newton_coop of the same sweep
    const double *mixlen_tab;   // mixed branch lengths (PhyloTree::isMixlen(), b200lk_set_mixlen): [rows][ncat]; a length is then the
    int mixlen;                 // ROW of the branch, and rate category r reads mixlen_tab[row * ncat + r] (child->getLength(cat_id[c]))
};

__device__ __forceinline__ double resolve_len(const Dev &D, double len) { return len >= 0.0 ? len : D.dev_len[(int)(-len) - 1]; }
// child->getLength(cat_id[c]) / dad_branch->getLength(mycat) (tree/phylokernelnew.h:1017, 1069, 2325, 2348, 2724, 2738, 3286, 3301)
__device__ __forceinline__ double resolve_len_cat(const Dev &D, double len, int r) {
    return D.mixlen ? D.mixlen_tab[(long long)(int)len * D.ncat + r] : resolve_len(D, len);
}

__device__ __forceinline__ long long pbase(long long p, int width) {
    return (p >> 5) * (long long)width * kWarp + (p & 31);
}

// 20-state engines keep their CLV slots FRAGMENT-MAJOR (the order the FP64-MMA walkers produce and consume them in, see
// b200lk_walk20ss.cuh): a 32-pattern block is [M-tile mt = 4][category c = C][160 doubles], and the 160 doubles of 8 patterns
// (g = pattern % 8) x 20 values are three state tiles in C-fragment order:
//     i < 16:  (i/8)*64 + g*8 + i%8        lane 4g+q of a warp owns values 2q, 2q+1 of a tile = the 16-byte chunk number `lane`
//     i >= 16: 128 + g*4 + (i-16)
// so that a lane stores its two accumulators of a state tile with one 16-byte store (the warp: 512 contiguous bytes), loads them
// back as MMA A fragments the same way, and the same 16-byte chunks are read without bank conflicts when a step's 1280 bytes sit in
// shared memory.  Same size as the warp-interleaved layout (B*32 doubles per 32 patterns); theta, _pattern_lh_cat and scale_num
// keep pbase().
constexpr int kClv20Step = 160;
__host__ __device__ __forceinline__ int clv20_off(int i, int g) {
    return i < 16 ? (i >> 3) * 64 + g * 8 + (i & 7) : 128 + g * 4 + (i - 16);
}
// element (category c, index i) of pattern p inside a CLV slot, for the thread-per-pattern kernels
struct ClvIdx {
    long long base;
    int g, n, f20, f8, cstep;
    __device__ __forceinline__ ClvIdx(const Dev &D, long long p) {
        f20 = D.clv_f20; f8 = D.clv_f8; n = D.n; g = (int)(p & 7);
        cstep = f8 ? D.n8 * 8 : kClv20Step;
        if (f20) base = (p >> 5) * (long long)D.B * kWarp + ((p >> 3) & 3) * (long long)D.C * kClv20Step;
        else if (f8) base = (p >> 5) * (long long)D.C * D.n8 * kWarp + ((p >> 3) & 3) * (long long)D.C * cstep;
        else base = pbase(p, D.B);
    }
    __device__ __forceinline__ long long operator()(int c, int i) const {
        if (f20) return base + c * kClv20Step + clv20_off(i, g);
        if (f8) return base + (long long)c * cstep + (i >> 3) * 64 + g * 8 + (i & 7);
        return base + (c * n + i) * kWarp;
    }
};

// ------------------------------------------------------------------------------------------------------------
// computePartialInfo (tree/phylokernelnew.h:900-1110): echild = evec * exp(eval * rate_c * len) and, for leaf
// children, the per-state lookup table.  One CTA per (op, child).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_partial_info(Dev D, const ChildJob *jobs) {
    extern __shared__ double s_exp[];  // [C][n]
    const ChildJob job = jobs[blockIdx.x];
    const int n = D.n, C = D.C, B = D.B, nn = D.n * D.n;
    for (int k = threadIdx.x; k < B; k += blockDim.x) {
        int c = k / n, i = k - c * n;
        int m = c / D.denom, r = c % D.ncat;
        double len_child = D.cat_rate[r] * resolve_len_cat(D, job.length, r);           // :1017
        s_exp[k] = exp(D.eval[m * n + i] * len_child);           // :1022
    }
    __syncthreads();
    double *E = D.opinfo + job.e_off;
    for (int k = threadIdx.x; k < C * nn; k += blockDim.x) {
        int c = k / nn, rem = k - c * nn;
        int i = rem % n;
        int m = c / D.denom;
        E[k] = D.evec[m * nn + rem] * s_exp[c * n + i];          // :1027
    }
    if (job.leaf_off < 0) {
        if (D.clv_state) {   // the child's CLV is the state-space vector t: fold its transform in, P_c = echild_c . inv_evec
            __syncthreads();
            for (int row = threadIdx.x; row < C * n; row += blockDim.x) {
                double *er = E + (long long)row * n;
                const double *iv = D.inv_evec + (long long)((row / n) / D.denom) * nn;
                double erow[64], prow[64];
                for (int i = 0; i < n; i++) erow[i] = er[i];
                for (int y = 0; y < n; y++) {
                    double a0 = 0.0, a1 = 0.0;
                    int i = 0;
                    for (; i + 1 < n; i += 2) { a0 = fma(erow[i], iv[i * n + y], a0); a1 = fma(erow[i + 1], iv[(i + 1) * n + y], a1); }
                    if (i < n) a0 = fma(erow[i], iv[i * n + y], a0);
                    prow[y] = a0 + a1;
                }
                for (int y = 0; y < n; y++) er[y] = prow[y];
            }
        }
        return;
    }
    __syncthreads();
    double *leaf = D.opinfo + job.leaf_off;
    const int tip_block = n * D.nmix;
    const int total = (D.su + 1) * B;
    for (int k = threadIdx.x; k < total; k += blockDim.x) {
        int state = k / B, cx = k - state * B;
        if (state == D.su) { leaf[k] = 1.0; continue; }          // :1052-1055
        int c = cx / n;
        int m = c / D.denom;
        const double *ep = E + (long long)cx * n;
        const double *tip = D.tip_lh + state * tip_block + m * n;
        double v;
        if ((n & 3) == 0) {   // Vec4d lane accumulators + horizontal_add (:1036-1051)
            double a0 = ep[0] * tip[0], a1 = ep[1] * tip[1], a2 = ep[2] * tip[2], a3 = ep[3] * tip[3];
            for (int i = 4; i < n; i += 4) {
                a0 = fma(ep[i], tip[i], a0); a1 = fma(ep[i + 1], tip[i + 1], a1);
                a2 = fma(ep[i + 2], tip[i + 2], a2); a3 = fma(ep[i + 3], tip[i + 3], a3);
            }
            v = (a0 + a1) + (a2 + a3);
        } else {              // scalar branch (:1085-1099)
            v = ep[0] * tip[0];
            for (int i = 1; i < n; i++) v += ep[i] * tip[i];
        }
        leaf[k] = v;
    }
}

// tip_partial_lh[state][m][i] = sum_j inv_evec_m[i][j] * ind_state[j]  (tree/phylotreesse.cpp:367-375,
// model/modelsubst.cpp:182-207, modeldna.cpp:462-475, modelprotein.cpp:1149-1166)
__global__ void k_tip_partial_lh(Dev D, double *tip_lh, int seq_type) {
    const int n = D.n;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int total = (D.su + 1) * D.nmix * n;
    if (idx >= total) return;
    int state = idx / (D.nmix * n);
    int rem = idx - state * D.nmix * n;
    int m = rem / n, i = rem - m * n;
    const double *iv = D.inv_evec + (long long)m * n * n + i * n;
    double acc = 0.0;
    for (int j = 0; j < n; j++) {
        double ind;
        if (state < n) ind = (j == state) ? 1.0 : 0.0;
        else if (seq_type == 1 && state < 18) ind = ((state - n + 1) >> j) & 1 ? 1.0 : 0.0;
        else if (seq_type == 2 && state < n + 3) {
            const int a0 = (state - n == 0) ? 2 : (state - n == 1) ? 5 : 9;
            ind = (j == a0 || j == a0 + 1) ? 1.0 : 0.0;
        } else ind = 1.0;
        acc += iv[j] * ind;
    }
    tip_lh[idx] = acc;
}

// val0/val1/val2 (tree/phylokernelnew.h:2322-2358, 2714-2745)
__global__ void k_vals(Dev D, double length, int with_derv) {
    const int n = D.n;
    for (int k = threadIdx.x; k < D.B; k += blockDim.x) {
        int c = k / n, i = k - c * n;
        int m = c / D.denom, r = c % D.ncat;
        double prop = D.cat_prop[r] * D.mix_weight[m];
        if (with_derv) {
            double cof = D.eval[m * n + i] * D.cat_rate[r];
            double val = exp(cof * resolve_len_cat(D, length, r)) * prop;
            double v1 = cof * val;
            D.val[k] = val;
            D.val[D.B + k] = v1;
            D.val[2 * D.B + k] = cof * v1;
        } else {
            double len = D.cat_rate[r] * resolve_len_cat(D, length, r);
            D.val[k] = exp(D.eval[m * n + i] * len) * prop;
        }
    }
}

// The same values computed by every CTA of a consumer kernel into its own shared memory (small blocks only: 16 or 80 exps per
// CTA cost nothing, a separate k_vals launch costs a launch — a third of a Newton-Raphson step on a small partition).
__device__ __forceinline__ const double *stage_vals(const Dev &D, double length, int with_derv, double *s_val) {
    const int n = D.n;
    for (int k = threadIdx.x; k < D.B; k += blockDim.x) {
        int c = k / n, i = k - c * n;
        int m = c / D.denom, r = c % D.ncat;
        double prop = D.cat_prop[r] * D.mix_weight[m];
        if (with_derv) {
            double cof = D.eval[m * n + i] * D.cat_rate[r];
            double val = exp(cof * resolve_len_cat(D, length, r)) * prop;
            double v1 = cof * val;
            s_val[k] = val;
            s_val[D.B + k] = v1;
            s_val[2 * D.B + k] = cof * v1;
        } else {
            double len = D.cat_rate[r] * resolve_len_cat(D, length, r);
            s_val[k] = exp(D.eval[m * n + i] * len) * prop;
        }
    }
    __syncthreads();
    return s_val;
}
constexpr int kFuseValsMaxB = 512;   // consumer kernels compute val0/1/2 themselves up to this block size

// ------------------------------------------------------------------------------------------------------------
// work distribution shared by the partial-likelihood kernels.
//   walk = 1 : "pattern-tile walker" — a work item is one tile of kTile patterns; the CTA walks the whole
//              post-order op list for it (each thread keeps its pattern, so a child CLV written at op k is re-read
//              by the same thread at its parent and is still L2-resident); this is the packet loop of the
//              reference (tree/phylokernelnew.h:1270-1277) with the packet shrunk to a tile.
//   walk = 0 : "levelled" — ops[op_begin, op_end) are mutually independent (one tree depth); a work item is
//              (tile, op).
// ------------------------------------------------------------------------------------------------------------
struct WorkItem { long long tile; int o0, o1; };

__device__ __forceinline__ bool next_item(unsigned long long *counter, long long *s_item, long long ntiles, int op_begin,
                                          int op_end, int walk, WorkItem &w) {
    __syncthreads();
    if (threadIdx.x == 0) *s_item = (long long)atomicAdd(counter, 1ULL);
    __syncthreads();
    long long item = *s_item;
    int nops = op_end - op_begin;
    long long nitems = walk ? ntiles : ntiles * nops;
    if (item >= nitems) return false;
    if (walk) { w.tile = item; w.o0 = op_begin; w.o1 = op_end; }
    else { w.tile = item / nops; w.o0 = op_begin + (int)(item - w.tile * nops); w.o1 = w.o0 + 1; }
    return true;
}

// ------------------------------------------------------------------------------------------------------------
// Generic partial-likelihood kernel (any nstates, any ncat_mix, NORM_LH or SAFE_LH): thread per pattern, matrices
// read as warp-uniform loads from the per-op info block.  computePartialLikelihood{,Generic}SIMD, binary nodes
// (tree/phylokernelnew.h:1632-2032).
// ------------------------------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ double dot4acc(const double *__restrict__ m, const double *v, int n) {
    // 4 lane accumulators as the reference's N%4==0 paths; tail handled sequentially
    const int nn = NT ? NT : n;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int i = 0;
#pragma unroll
    for (; i + 3 < nn; i += 4) {
        a0 = fma(__ldg(m + i), v[i], a0);
        a1 = fma(__ldg(m + i + 1), v[i + 1], a1);
        a2 = fma(__ldg(m + i + 2), v[i + 2], a2);
        a3 = fma(__ldg(m + i + 3), v[i + 3], a3);
    }
    double s = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (; i < nn; i++) s = fma(__ldg(m + i), v[i], s);
    return s;
}

template <int NT, bool SAFE>
__device__ __forceinline__ void process_op_generic(const Dev &D, const DevOp &op, long long p) {
    constexpr int NMAX = NT ? NT : 64;
    const int n = NT ? NT : D.n;
    const int C = D.C, B = D.B;
    const ClvIdx at(D, p);
    const long long sb = SAFE ? pbase(p, C) : p;
    double *out = D.partial + (long long)op.dst * D.slot_stride;
    uint16_t *osc = D.scale + (long long)op.dst * D.scale_stride + sb;
    const double *Lp = (op.ncase == 2) ? D.partial + (long long)op.lid * D.slot_stride : nullptr;
    const double *Rp = (op.ncase >= 1) ? D.partial + (long long)op.rid * D.slot_stride : nullptr;
    const uint16_t *lsc = (op.ncase == 2) ? D.scale + (long long)op.lid * D.scale_stride + sb : nullptr;
    const uint16_t *rsc = (op.ncase >= 1) ? D.scale + (long long)op.rid * D.scale_stride + sb : nullptr;
    const double *leafL = nullptr, *leafR = nullptr;
    if (op.ncase <= 1) leafL = D.opinfo + op.leafL + (long long)D.tips[(long long)op.lid * D.Ppad + p] * B;
    if (op.ncase == 0) leafR = D.opinfo + op.leafR + (long long)D.tips[(long long)op.rid * D.Ppad + p] * B;
    const double *eL = D.opinfo + op.eL, *eR = D.opinfo + op.eR;
    const bool check = (op.ncase != 0) && (D.invar[p] == 0.0);   // tip-tip never rescales (:1640)
    double lh_max = 0.0;
    unsigned norm_scale = 0;
    if (!SAFE) norm_scale = (lsc ? lsc[0] : 0u) + (rsc ? rsc[0] : 0u);

    double lv[NMAX], rv[NMAX], tmp[NMAX];
    for (int c = 0; c < C; c++) {
        const int m = c / D.denom;
        const double *inv = D.inv_evec + (long long)m * n * n;
        if (Lp) {
#pragma unroll
            for (int i = 0; i < n; i++) lv[i] = Lp[at(c, i)];
        }
        if (Rp) {
#pragma unroll
            for (int i = 0; i < n; i++) rv[i] = Rp[at(c, i)];
        }
#pragma unroll
        for (int x = 0; x < n; x++) {
            double a = Lp ? dot4acc<NT>(eL + (long long)(c * n + x) * n, lv, n) : leafL[c * n + x];
            double b = Rp ? dot4acc<NT>(eR + (long long)(c * n + x) * n, rv, n) : leafR[c * n + x];
            tmp[x] = a * b;
        }
        double cmax = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) {
            double v = dot4acc<NT>(inv + i * n, tmp, n);
            lv[i] = v;   // reuse as output staging
            cmax = fmax(cmax, fabs(v));
        }
        if (D.clv_state) {   // the slot keeps t; the underflow test above is the reference's, on partial_lh = inv_evec . t
#pragma unroll
            for (int i = 0; i < n; i++) lv[i] = tmp[i];
        }
        if (SAFE) {
            unsigned sc = (lsc ? lsc[c * kWarp] : 0u) + (rsc ? rsc[c * kWarp] : 0u);
            if (check && cmax < kScalingThreshold) {
#pragma unroll
                for (int i = 0; i < n; i++) lv[i] *= kTwo256;
                sc += 1;
            }
            osc[c * kWarp] = (uint16_t)sc;
        } else {
            lh_max = fmax(lh_max, cmax);
        }
#pragma unroll
        for (int i = 0; i < n; i++) out[at(c, i)] = lv[i];
    }
    if (!SAFE) {
        if (check && lh_max < kScalingThreshold) {
            for (int c = 0; c < C; c++)
                for (int i = 0; i < n; i++) out[at(c, i)] *= kTwo256;   // rare path: re-scale own output
            norm_scale += 1;
        }
        osc[0] = (uint16_t)norm_scale;
    }
}

// Multifurcating node (node->degree() > 3, tree/phylokernelnew.h:1407-1631): the running product over the children lives in the
// ORIGINAL state space, the 2^-256 test follows EVERY child on that product (:1557-1605), the inverse-eigenvector transform
// comes last without a further test (:1610-1628).  The destination CLV itself is the workspace for the running product.
template <int NT, bool SAFE>
__device__ __forceinline__ void process_op_multi(const Dev &D, const DevOp &op, long long p) {
    constexpr int NMAX = NT ? NT : 64;
    const int n = NT ? NT : D.n;
    const int C = D.C;
    const ClvIdx at(D, p);
    const long long sb = SAFE ? pbase(p, C) : p;
    double *out = D.partial + (long long)op.dst * D.slot_stride;
    uint16_t *osc = D.scale + (long long)op.dst * D.scale_stride + sb;
    const bool may = D.invar[p] == 0.0;
    for (int c = 0; c < C; c++) {
        for (int x = 0; x < n; x++) out[at(c, x)] = 1.0;
        if (SAFE) osc[c * kWarp] = 0;
    }
    unsigned norm_scale = 0;
    double lv[NMAX];
    for (int k = 0; k < op.nchild; k++) {
        const MultiChild ch = D.mchild[op.mc_off + k];
        const double *leaf = nullptr, *Lp = nullptr;
        const uint16_t *csc = nullptr;
        if (ch.kind == 0) leaf = D.opinfo + ch.leaf_off + (long long)D.tips[(long long)ch.id * D.Ppad + p] * D.B;
        else {
            Lp = D.partial + (long long)ch.id * D.slot_stride;
            csc = D.scale + (long long)ch.id * D.scale_stride + sb;
            if (!SAFE) norm_scale += csc[0];
        }
        double lh_max = 0.0;
        for (int c = 0; c < C; c++) {
            double cmax = 0.0;
            if (Lp) {
                for (int i = 0; i < n; i++) lv[i] = Lp[at(c, i)];
                const double *ep = D.opinfo + ch.e_off + (long long)c * n * n;
                for (int x = 0; x < n; x++) {
                    double v = __ldg(ep + x * n) * lv[0];   // :1542-1546: sequential mul_add
                    for (int i = 1; i < n; i++) v = fma(__ldg(ep + x * n + i), lv[i], v);
                    const double r = out[at(c, x)] * v;
                    out[at(c, x)] = r;
                    cmax = fmax(cmax, fabs(r));
                }
            } else {
                for (int x = 0; x < n; x++) {
                    const double r = out[at(c, x)] * leaf[c * n + x];
                    out[at(c, x)] = r;
                    cmax = fmax(cmax, fabs(r));
                }
            }
            if (SAFE) {
                unsigned sc = osc[c * kWarp] + (csc ? csc[c * kWarp] : 0u);
                if (may && cmax < kScalingThreshold) {
                    for (int x = 0; x < n; x++) out[at(c, x)] *= kTwo256;
                    sc += 1;
                }
                osc[c * kWarp] = (uint16_t)sc;
            } else lh_max = fmax(lh_max, cmax);
        }
        if (!SAFE && may && lh_max < kScalingThreshold) {
            for (int c = 0; c < C; c++)
                for (int x = 0; x < n; x++) out[at(c, x)] *= kTwo256;
            norm_scale += 1;
        }
    }
    if (!SAFE) osc[0] = (uint16_t)norm_scale;
    if (D.clv_state) return;   // the slot keeps the state-space product
    for (int c = 0; c < C; c++) {
        const double *inv = D.inv_evec + (long long)(c / D.denom) * n * n;
        for (int x = 0; x < n; x++) lv[x] = out[at(c, x)];
        for (int i = 0; i < n; i++) out[at(c, i)] = dot4acc<NT>(inv + i * n, lv, n);
    }
}

template <int NT, bool SAFE>
__global__ void __launch_bounds__(kTile) k_partials_generic(Dev D, const DevOp *ops, int op_begin, int op_end, int walk,
                                                            unsigned long long *counter) {
    __shared__ long long s_item;
    const long long ntiles = D.Ppad / kTile;
    WorkItem w;
    while (next_item(counter, &s_item, ntiles, op_begin, op_end, walk, w)) {
        const long long p = w.tile * kTile + threadIdx.x;
        for (int o = w.o0; o < w.o1; o++) {
            if (ops[o].ncase == 3) process_op_multi<NT, SAFE>(D, ops[o], p);
            else process_op_generic<NT, SAFE>(D, ops[o], p);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// 4-state x 4-category kernel (GTR+G4 and every DNA model with 4 rate categories): bandwidth-bound.
// Per op the CTA stages echild(left,right) [4][4][4] and the leaf lookup tables in shared memory; each thread owns
// one pattern: 32 coalesced 8-byte loads (2 children x 16), 192 FP64 FMA-class ops against warp-uniform shared
// operands (LDS.128 broadcasts), 16 coalesced stores + one uint16 scale count.
// ------------------------------------------------------------------------------------------------------------
constexpr int kMaxLeafRows4 = 19;   // DNA: states 0..18 (alignment/alignment.cpp:738)

__device__ __forceinline__ double dotq(const double2 *m, double v0, double v1, double v2, double v3) {
    // productVecMat / dotProductDualVec N==4 form: fma(a1,b1,a0*b0) + fma(a3,b3,a2*b2)  (:235, :335)
    double2 m01 = m[0], m23 = m[1];
    return fma(m01.y, v1, m01.x * v0) + fma(m23.y, v3, m23.x * v2);
}

template <bool SAFE>
__global__ void __launch_bounds__(kTile, 4) k_partials_n4c4(Dev D, const DevOp *ops, int op_begin, int op_end, int walk,
                                                           unsigned long long *counter) {
    __shared__ __align__(16) double s_e[2][64];                       // echild left | right : [c][x][i]
    __shared__ __align__(16) double s_leaf[2][kMaxLeafRows4 * 16];    // [state][c][x]
    __shared__ __align__(16) double s_inv[16];                        // inv_evec[i][x]
    __shared__ long long s_item;
    const int tid = threadIdx.x;
    if (tid < 16) s_inv[tid] = D.inv_evec[tid];
    const long long ntiles = D.Ppad / kTile;
    const int nleaf = (D.su + 1) * 16;
    WorkItem w;
    while (next_item(counter, &s_item, ntiles, op_begin, op_end, walk, w)) {
        const long long p = w.tile * kTile + tid;
        const long long pb = pbase(p, 16);
        const long long sb = SAFE ? pbase(p, 4) : p;
        const double invar = D.invar[p];
        for (int o = w.o0; o < w.o1; o++) {
            const DevOp op = ops[o];
            __syncthreads();   // previous op's readers of s_e/s_leaf are done
            if (tid < 64) s_e[0][tid] = D.opinfo[op.eL + tid];
            else s_e[1][tid - 64] = D.opinfo[op.eR + tid - 64];
            if (op.ncase <= 1)
                for (int k = tid; k < nleaf; k += kTile) s_leaf[0][k] = D.opinfo[op.leafL + k];
            if (op.ncase == 0)
                for (int k = tid; k < nleaf; k += kTile) s_leaf[1][k] = D.opinfo[op.leafR + k];
            __syncthreads();

            double L[16], R[16], out[16];
            unsigned sc[SAFE ? 4 : 1];
            if (op.ncase == 2) {
                const double *Lp = D.partial + (long long)op.lid * D.slot_stride + pb;
#pragma unroll
                for (int k = 0; k < 16; k++) L[k] = Lp[k * kWarp];
            }
            if (op.ncase >= 1) {
                const double *Rp = D.partial + (long long)op.rid * D.slot_stride + pb;
#pragma unroll
                for (int k = 0; k < 16; k++) R[k] = Rp[k * kWarp];
            }
            // scale_num of the children (:1640, :1753-1756, :1926-1931, :1949)
#pragma unroll
            for (int c = 0; c < (SAFE ? 4 : 1); c++) {
                unsigned s = 0;
                if (op.ncase == 2) s += D.scale[(long long)op.lid * D.scale_stride + sb + (SAFE ? c * kWarp : 0)];
                if (op.ncase >= 1) s += D.scale[(long long)op.rid * D.scale_stride + sb + (SAFE ? c * kWarp : 0)];
                sc[c] = s;
            }
            const double *tl = nullptr, *tr = nullptr;
            if (op.ncase <= 1) tl = &s_leaf[0][(int)D.tips[(long long)op.lid * D.Ppad + p] * 16];
            if (op.ncase == 0) tr = &s_leaf[1][(int)D.tips[(long long)op.rid * D.Ppad + p] * 16];

            const double2 *eLq = reinterpret_cast<const double2 *>(s_e[0]);
            const double2 *eRq = reinterpret_cast<const double2 *>(s_e[1]);
            const double2 *invq = reinterpret_cast<const double2 *>(s_inv);
            double lh_max = 0.0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                double tmp[4];
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    double a, b;
                    if (op.ncase == 2) a = dotq(eLq + (c * 4 + x) * 2, L[c * 4], L[c * 4 + 1], L[c * 4 + 2], L[c * 4 + 3]);
                    else a = tl[c * 4 + x];
                    if (op.ncase >= 1) b = dotq(eRq + (c * 4 + x) * 2, R[c * 4], R[c * 4 + 1], R[c * 4 + 2], R[c * 4 + 3]);
                    else b = tr[c * 4 + x];
                    tmp[x] = a * b;
                }
                double cmax = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    double v = dotq(invq + i * 2, tmp[0], tmp[1], tmp[2], tmp[3]);
                    out[c * 4 + i] = v;
                    cmax = fmax(cmax, fabs(v));
                }
                if (SAFE) {
                    if (op.ncase != 0 && cmax < kScalingThreshold && invar == 0.0) {
#pragma unroll
                        for (int i = 0; i < 4; i++) out[c * 4 + i] *= kTwo256;
                        sc[c] += 1;
                    }
                } else lh_max = fmax(lh_max, cmax);
            }
            if (!SAFE && op.ncase != 0 && lh_max < kScalingThreshold && invar == 0.0) {
#pragma unroll
                for (int k = 0; k < 16; k++) out[k] *= kTwo256;
                sc[0] += 1;
            }
            double *Op = D.partial + (long long)op.dst * D.slot_stride + pb;
#pragma unroll
            for (int k = 0; k < 16; k++) Op[k * kWarp] = out[k];
#pragma unroll
            for (int c = 0; c < (SAFE ? 4 : 1); c++)
                D.scale[(long long)op.dst * D.scale_stride + sb + (SAFE ? c * kWarp : 0)] = (uint16_t)sc[c];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// 4-state x 4-category WALKER, v2 (the hot kernel of BASELINE config 2).
//
// What changed against k_partials_n4c4 (profiles/r1_v1_*: latency-bound, 604 instr per warp-op, 96 LDS.128 per
// thread-op, 4-way bank conflicts on the leaf tables, two CTA barriers and three dependent L2 round trips per op):
//   * register chaining — in post-order the parent of an internal child is the NEXT op, so the CLV a thread has
//     just produced is consumed from registers (`cur`) instead of being re-loaded: at most ONE child CLV is loaded
//     per op (the sibling subtree's), none for tip-inner nodes;
//   * factorised operator — echild[c][x][i] = evec[x][i]*exp(eval_i r_c t) (tree/phylokernelnew.h:1016-1031) is
//     applied as evec . (exp (.) L): the per-op data shrinks from 2x64 to 2x16 doubles and evec / inv_evec become
//     kernel-parameter (constant-bank) operands of the FP64 FMAs — no shared-memory traffic for them;
//   * tip children — unambiguous states index a [c][x][state] table (conflict-free, 4 states contiguous), the gap
//     state is the constant 1, IUPAC codes fall back to the direct sum;
//   * one CTA barrier per op; the next op's 1.3 KB staging block is fetched into registers while the current op
//     computes, and stored to the other shared buffer afterwards;
//   * the underflow test max|v| < 2^-256 is an integer compare on the exponent fields (ALU pipe, not FP64).
// ------------------------------------------------------------------------------------------------------------
struct Model4 {
    double evec[16];   // evec[x*4 + i]
    double inv[16];    // inv_evec[i*4 + x]
};

struct WalkJob4 {      // host -> k_walk_info_n4c4
    double lenA, lenB;
    int32_t dst, ncase, aid, bid, loadA, pad;
};

// staging block per op: header (kHdr4 doubles: 8 ints, then 6 byte offsets — the 64-bit products slot * stride / row * Ppad that every
// thread of every tile would otherwise recompute per op, ~11 integer instructions each) | expA[16] | expB[16] | tabA[64] | tabB[64]
// (tables for EVERY state code instead of the four nucleotides — 676 doubles per block, no branch on the tip state — were measured:
// 9 % slower on config 2, the block is re-staged per tile and op and 5.4 KB per 128 patterns is a third of the CLV it produces)
constexpr int kHdr4 = 10;
constexpr int kStage4 = kHdr4 + 16 + 16 + 64 + 64;   // 170 doubles = 1360 bytes (16-byte multiple)
struct StageOff4 { long long dst_clv, b_clv, dst_sc, b_sc, tip_a, tip_b; };   // bytes: CLV slot of dst / sibling, their scale vectors, tip rows
struct StageHdr4 { int32_t dst, ncase, aid, bid, loadA, pad0, pad1, pad2; };
constexpr int kPfAhead4 = 4;   // ops between an L2 prefetch and the use of the data (an op of one tile takes ~2 us at full load)

// child A is the CLV that lives in registers (`cur`: chained from the previous op, or loaded when loadA), child B the
// other one: an internal CLV loaded from memory (ncase 2) or a tip (ncase <= 1); ncase 0: both are tips.
__global__ void __launch_bounds__(64) k_walk_info_n4c4(Dev D, const WalkJob4 *jobs, double *stage, unsigned long long *counter) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0ULL;   // the walker's tile queue (next launch on this stream)
    __shared__ double s_exp[2][16];
    const WalkJob4 job = jobs[blockIdx.x];
    double *st = stage + (size_t)blockIdx.x * kStage4;
    const int t = threadIdx.x;
    if (t < 32) {
        const int ch = t >> 4, k = t & 15, c = k >> 2, i = k & 3;
        const double len_child = D.cat_rate[c] * resolve_len_cat(D, ch ? job.lenB : job.lenA, c);   // :1017
        const double v = exp(D.eval[i] * len_child);                           // :1022
        s_exp[ch][k] = v;
        st[kHdr4 + ch * 16 + k] = v;
    }
    if (t == 1) {
        StageOff4 o;
        o.dst_clv = (long long)job.dst * D.slot_stride * (long long)sizeof(double);
        o.b_clv = job.ncase == 2 ? (long long)job.bid * D.slot_stride * (long long)sizeof(double) : 0;
        o.dst_sc = (long long)job.dst * D.scale_stride * (long long)sizeof(uint16_t);
        o.b_sc = job.ncase == 2 ? (long long)job.bid * D.scale_stride * (long long)sizeof(uint16_t) : 0;
        o.tip_a = job.ncase == 0 ? (long long)job.aid * D.Ppad : 0;
        o.tip_b = job.ncase <= 1 ? (long long)job.bid * D.Ppad : 0;
        *reinterpret_cast<StageOff4 *>(st + 4) = o;
    }
    if (t == 0) {
        // pad0..2: what the op kPfAhead4 ops further on will read from memory that may have left L2 — its tip rows (B, A) and its
        // sibling CLV slot; k_walk_n4c4 turns them into L2 bulk prefetches of its own tile (-1: nothing)
        int pfB = -1, pfA = -1, pfS = -1;
        const int on = (int)blockIdx.x + kPfAhead4;
        if (on < (int)gridDim.x) {
            const WalkJob4 jn = jobs[on];
            if (jn.ncase <= 1) pfB = jn.bid;
            if (jn.ncase == 0) pfA = jn.aid;
            if (jn.ncase == 2) pfS = jn.bid;
        }
        StageHdr4 h{job.dst, job.ncase, job.aid, job.bid, job.loadA, pfB, pfA, pfS};
        *reinterpret_cast<StageHdr4 *>(st) = h;
    }
    __syncthreads();
    // partial_lh_leaves for the 4 unambiguous states, stored [c][x][state] (:1036-1051, Vec4d lanes + horizontal_add)
    const int c = t >> 4, x = (t >> 2) & 3, s = t & 3;
    for (int ch = 0; ch < 2; ch++) {
        const bool is_tip = ch == 0 ? (job.ncase == 0) : (job.ncase <= 1);
        double v = 0.0;
        if (is_tip) {
            const double *tip = D.tip_lh + s * 4;
            double p0 = __dmul_rn(__dmul_rn(D.evec[x * 4 + 0], s_exp[ch][c * 4 + 0]), tip[0]);
            double p1 = __dmul_rn(__dmul_rn(D.evec[x * 4 + 1], s_exp[ch][c * 4 + 1]), tip[1]);
            double p2 = __dmul_rn(__dmul_rn(D.evec[x * 4 + 2], s_exp[ch][c * 4 + 2]), tip[2]);
            double p3 = __dmul_rn(__dmul_rn(D.evec[x * 4 + 3], s_exp[ch][c * 4 + 3]), tip[3]);
            v = __dadd_rn(__dadd_rn(p0, p1), __dadd_rn(p2, p3));
        }
        st[kHdr4 + 32 + ch * 64 + t] = v;
    }
}

__device__ __forceinline__ unsigned abs_hi(double v) { return (unsigned)__double2hiint(v) & 0x7fffffffu; }
constexpr unsigned kThresholdHi = (1023u - 256u) << 20;   // high word of 2^-256: |v| < 2^-256  <=>  abs_hi(v) < this

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// one instruction, one thread: `bytes` (multiple of 16, 16-byte aligned) towards L2
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// TMA bulk store helpers (cp.async.bulk shared::cta -> global, bulk async-groups)
__device__ __forceinline__ void bulk_store_issue(void *gdst, const void *ssrc, unsigned bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// STORE: 0 = one burst of 16 global stores per thread at the end of the op; 1 = four stores after each category;
//        2 = the CLV goes to shared memory category by category and leaves the SM as ONE 16 KB TMA bulk store per (tile, op),
//            issued by one thread after the next op's barrier: no thread ever waits for the LSU to drain its stores (with
//            direct stores ncu shows the next op's first FP64 instruction stalled on the register hand-over of the burst, and
//            the kernel's time ~ compute + store time instead of their maximum).  CLV loads then bypass L1 (ld.global.cg): the
//            bulk store writes through the async proxy, and the slot's lines may sit in this SM's L1 from an earlier read.
// DBG (diagnosis builds, wrong results): bit 0 no CLV stores, bit 1 no sibling loads
template <bool SAFE, int STORE = 2, int DBG = 0, int PF = 20>   // PF: bit 0 = L2 prefetch of tip rows, bit 1 = of sibling CLVs (slower: A/B only), bit 2 = tip states through shared memory
__global__ void __launch_bounds__(kTile, 4) k_walk_n4c4(Dev D, const __grid_constant__ Model4 M, const double *stage, int nops,
                                                       unsigned long long *counter) {
    // 3-deep ring of staging blocks: while op o computes, block o+1 is already resident (its header steers the
    // L1 prefetch of op o+1's operands) and block o+2 is in flight in registers.
    __shared__ __align__(16) double s_stage[3][kStage4];
    __shared__ double s_tip[kMaxLeafRows4 * 4];
    __shared__ long long s_item;
    // PF bit 2: the next op's tip states travel global -> shared with cp.async (no register, hence no scoreboard, is held across the
    // op: with the register form ncu put 27 % of all stall samples on the op's first FP64 instruction, a shared-memory consumer
    // whose scoreboard the just-issued tip-state load of the NEXT op shares)
    constexpr bool TST = STORE == 2, ESTORE = STORE >= 1;
    constexpr bool TSM = (PF & 4) != 0 && !TST;   // (the TMA-store variant's two 16 KB buffers leave no static shared memory for it)
    __shared__ __align__(16) unsigned char s_ts[TSM ? 2 : 1][2][TSM ? kTile : 4];
    // PF bit 3: the sibling CLV of the NEXT op (always written before this op began: the planner chains the previous op's result as
    // child A) travels global -> shared with cp.async while this op computes, 16 KB per tile in the slot's own layout
    constexpr bool XSM = (PF & 8) != 0 && !TST;
    extern __shared__ __align__(16) double s_x_dyn[];   // XSM: [2][16 * kTile] (dynamic: the static arrays are near the 48 KB limit)
    double (*s_x)[16 * kTile] = reinterpret_cast<double (*)[16 * kTile]>(s_x_dyn);
    __shared__ __align__(128) double s_out[TST ? 2 : 1][TST ? 16 * kTile : 2];   // [buffer][warp][k][lane] = the tile's 16 KB of a slot
    const int tid = threadIdx.x;
    const int so = (tid >> 5) * 16 * kWarp + (tid & 31);
    int prev_dst = -1;
    long long prev_tile_off = 0;
    for (int k = tid; k < (D.su + 1) * 4; k += kTile) s_tip[k] = D.tip_lh[k];
    const long long ntiles = D.Ppad / kTile;
    constexpr int kStageVec = kStage4 / 2;   // double2 elements per staging block
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (long long)atomicAdd(counter, 1ULL);
        __syncthreads();
        const long long tile = s_item;
        if (tile >= ntiles) break;
        const long long p = tile * kTile + tid;
        const long long pb = pbase(p, 16);
        const long long sb = SAFE ? pbase(p, 4) : p;
        const bool may_scale = D.invar[p] == 0.0;
        for (int k = tid; k < kStageVec; k += kTile) {
            reinterpret_cast<double2 *>(s_stage[0])[k] = reinterpret_cast<const double2 *>(stage)[k];
            if (nops > 1) reinterpret_cast<double2 *>(s_stage[1])[k] = reinterpret_cast<const double2 *>(stage + kStage4)[k];
        }
        double cur[16];
        unsigned cursc[SAFE ? 4 : 1];
#pragma unroll
        for (int k = 0; k < 16; k++) cur[k] = 0.0;
#pragma unroll
        for (int c = 0; c < (SAFE ? 4 : 1); c++) cursc[c] = 0;
        int sA_next = 0, sB_next = 0;
        const int ts_off = (tid >> 5) * 32 + (tid & 7) * 4;   // TSM: lanes 0..7 of a warp copy the warp's 32 states, 4 bytes each
        {   // tip states of op 0
            const int4 h = *reinterpret_cast<const int4 *>(stage);   // global: the shared copy is not visible yet
            if (XSM && h.y == 2) {
                const double *src = D.partial + (long long)h.w * D.slot_stride + tile * (16LL * kTile);
#pragma unroll
                for (int m = 0; m < 8; m++) cp_async16(&s_x[0][2 * (tid + kTile * m)], src + 2 * (tid + kTile * m));
            }
            if (TSM) {
                if ((tid & 31) < 8) {
                    if (h.y <= 1) cp_async4(&s_ts[0][1][ts_off], D.tips + (long long)h.w * D.Ppad + tile * kTile + ts_off);
                    if (h.y == 0) cp_async4(&s_ts[0][0][ts_off], D.tips + (long long)h.z * D.Ppad + tile * kTile + ts_off);
                }
            } else {
                if (h.y <= 1) sB_next = D.tips[(long long)h.w * D.Ppad + p];
                if (h.y == 0) sA_next = D.tips[(long long)h.z * D.Ppad + p];
            }
        }

        int buf = 0;
        for (int o = 0; o < nops; o++) {
            if (TST) {
                // every bulk store issued so far has completed: its buffer may be rewritten, its CLV is visible in L2
                if (tid == 0) bulk_store_wait_all();
                fence_proxy_async_smem();   // this thread's shared-memory writes of the previous op, for the async proxy
            }
            cp_async_wait_all();   // staging block o (copied asynchronously during op o-2) has landed for this thread
            __syncthreads();
            if (TST && tid == 0 && prev_dst >= 0 && !(DBG & 1))
                bulk_store_issue(D.partial + (long long)prev_dst * D.slot_stride + prev_tile_off, s_out[(o + 1) & 1], 16 * kTile * sizeof(double));
            const double *st = s_stage[buf];
            const int buf1 = buf == 2 ? 0 : buf + 1, buf2 = buf1 == 2 ? 0 : buf1 + 1;
            const int4 h0 = *reinterpret_cast<const int4 *>(st);
            const int4 h1 = *reinterpret_cast<const int4 *>(st + 2);
            const StageOff4 &off = *reinterpret_cast<const StageOff4 *>(st + 4);
            const int h_loadA = h1.x;
            const int dst = h0.x, ncase = h0.y, aid = h0.z, bid = h0.w;
            if (PF && tid == 0) {
                // tip rows are streamed from DRAM once per evaluation, and a sibling CLV written many ops ago may have left L2: under
                // the store stream of this kernel such a read takes longer than an op (ncu: 37 % of all stall samples sat on the
                // first uses of a tip state loaded ONE op ahead) — ask for this tile's part of them kPfAhead4 ops early
                if ((PF & 1) && h1.y >= 0) prefetch_l2_bulk(D.tips + (long long)h1.y * D.Ppad + tile * kTile, kTile);
                if ((PF & 1) && h1.z >= 0) prefetch_l2_bulk(D.tips + (long long)h1.z * D.Ppad + tile * kTile, kTile);
                if ((PF & 2) && h1.w >= 0) prefetch_l2_bulk(D.partial + (long long)h1.w * D.slot_stride + tile * (16LL * kTile), 16 * kTile * sizeof(double));
            }
            // staging block o+2 travels global -> shared with cp.async (no registers held across the op); its ring slot was
            // op o-1's, which every thread has left (barrier above)
            if (o + 2 < nops)
                for (int k = tid; k < kStageVec; k += kTile) cp_async16(&s_stage[buf2][2 * k], stage + (size_t)(o + 2) * kStage4 + 2 * k);
            const int sA = TSM ? (ncase == 0 ? (int)s_ts[o & 1][0][tid] : 0) : sA_next;
            const int sB = TSM ? (ncase <= 1 ? (int)s_ts[o & 1][1][tid] : 0) : sB_next;

            double X[16];
            unsigned scB[SAFE ? 4 : 1];
            if (ncase == 2) {
                const double *Bp = reinterpret_cast<const double *>(reinterpret_cast<const char *>(D.partial) + off.b_clv) + pb;
#pragma unroll
                for (int k = 0; k < 16; k++) X[k] = (DBG & 2) ? cur[15 - k] + 1e-3 : XSM ? s_x[o & 1][so + k * kWarp] : TST ? __ldcg(Bp + k * kWarp) : Bp[k * kWarp];
#pragma unroll
                for (int c = 0; c < (SAFE ? 4 : 1); c++)
                    scB[c] = reinterpret_cast<const uint16_t *>(reinterpret_cast<const char *>(D.scale) + off.b_sc)[sb + (SAFE ? c * kWarp : 0)];
            } else {
#pragma unroll
                for (int c = 0; c < (SAFE ? 4 : 1); c++) scB[c] = 0;
            }
            if (ncase == 0) {
#pragma unroll
                for (int c = 0; c < (SAFE ? 4 : 1); c++) cursc[c] = 0;
            } else if (h_loadA) {
                const double *Ap = D.partial + (long long)aid * D.slot_stride + pb;
#pragma unroll
                for (int k = 0; k < 16; k++) cur[k] = TST ? __ldcg(Ap + k * kWarp) : Ap[k * kWarp];
#pragma unroll
                for (int c = 0; c < (SAFE ? 4 : 1); c++) cursc[c] = D.scale[(long long)aid * D.scale_stride + sb + (SAFE ? c * kWarp : 0)];
            }
            if (o + 1 < nops) {
                // steer the next op: its tip states go to registers, its CLV operands towards L1
                const double *sn = s_stage[buf1];
                const int4 hn = *reinterpret_cast<const int4 *>(sn);
                const int n_loadA = *reinterpret_cast<const int *>(sn + 2);
                if (TSM) {
                    if ((tid & 31) < 8) {
                        const StageOff4 &offn = *reinterpret_cast<const StageOff4 *>(sn + 4);
                        if (hn.y <= 1) cp_async4(&s_ts[(o + 1) & 1][1][ts_off], D.tips + offn.tip_b + tile * kTile + ts_off);
                        if (hn.y == 0) cp_async4(&s_ts[(o + 1) & 1][0][ts_off], D.tips + offn.tip_a + tile * kTile + ts_off);
                    }
                } else {
                    if (hn.y <= 1) sB_next = D.tips[(long long)hn.w * D.Ppad + p];
                    if (hn.y == 0) sA_next = D.tips[(long long)hn.z * D.Ppad + p];
                }
                if (hn.y == 2) {
                    if (XSM) {
                        const double *src = D.partial + (long long)hn.w * D.slot_stride + tile * (16LL * kTile);
#pragma unroll
                        for (int m = 0; m < 8; m++) cp_async16(&s_x[(o + 1) & 1][2 * (tid + kTile * m)], src + 2 * (tid + kTile * m));
                    } else if (!(PF & 16)) {   // PF bit 4: no L1 prefetches (A/B: ncu shows the sibling loads missing L1 regardless)
                        const double *Bn = D.partial + (long long)hn.w * D.slot_stride + pb;
#pragma unroll
                        for (int k = 0; k < 16; k++) if (!TST) prefetch_l1(Bn + k * kWarp);
                    }
                }
                if (hn.y >= 1 && n_loadA && !(PF & 16)) {
                    const double *An = D.partial + (long long)hn.z * D.slot_stride + pb;
#pragma unroll
                    for (int k = 0; k < 16; k++) if (!TST) prefetch_l1(An + k * kWarp);
                }
            }
            const double *expA = st + kHdr4, *expB = st + kHdr4 + 16, *tabA = st + kHdr4 + 32, *tabB = st + kHdr4 + 96;
            double *Op = reinterpret_cast<double *>(reinterpret_cast<char *>(D.partial) + off.dst_clv) + pb;
            unsigned mhi = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                double fa[4], fb[4];
                if (ncase >= 1) {
                    const double2 e01 = *reinterpret_cast<const double2 *>(expA + c * 4), e23 = *reinterpret_cast<const double2 *>(expA + c * 4 + 2);
                    const double a0 = e01.x * cur[c * 4], a1 = e01.y * cur[c * 4 + 1], a2 = e23.x * cur[c * 4 + 2], a3 = e23.y * cur[c * 4 + 3];
#pragma unroll
                    for (int x = 0; x < 4; x++)
                        fa[x] = fma(M.evec[x * 4 + 3], a3, fma(M.evec[x * 4 + 2], a2, fma(M.evec[x * 4 + 1], a1, M.evec[x * 4] * a0)));
                } else if (sA < 4) {
#pragma unroll
                    for (int x = 0; x < 4; x++) fa[x] = tabA[c * 16 + x * 4 + sA];
                } else if (sA == D.su) {
#pragma unroll
                    for (int x = 0; x < 4; x++) fa[x] = 1.0;
                } else {
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        double q0 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 0], expA[c * 4 + 0]), s_tip[sA * 4 + 0]);
                        double q1 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 1], expA[c * 4 + 1]), s_tip[sA * 4 + 1]);
                        double q2 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 2], expA[c * 4 + 2]), s_tip[sA * 4 + 2]);
                        double q3 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 3], expA[c * 4 + 3]), s_tip[sA * 4 + 3]);
                        fa[x] = __dadd_rn(__dadd_rn(q0, q1), __dadd_rn(q2, q3));
                    }
                }
                if (ncase == 2) {
                    const double2 e01 = *reinterpret_cast<const double2 *>(expB + c * 4), e23 = *reinterpret_cast<const double2 *>(expB + c * 4 + 2);
                    const double b0 = e01.x * X[c * 4], b1 = e01.y * X[c * 4 + 1], b2 = e23.x * X[c * 4 + 2], b3 = e23.y * X[c * 4 + 3];
#pragma unroll
                    for (int x = 0; x < 4; x++)
                        fb[x] = fma(M.evec[x * 4 + 3], b3, fma(M.evec[x * 4 + 2], b2, fma(M.evec[x * 4 + 1], b1, M.evec[x * 4] * b0)));
                } else if (sB < 4) {
#pragma unroll
                    for (int x = 0; x < 4; x++) fb[x] = tabB[c * 16 + x * 4 + sB];
                } else if (sB == D.su) {
#pragma unroll
                    for (int x = 0; x < 4; x++) fb[x] = 1.0;
                } else {
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        double q0 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 0], expB[c * 4 + 0]), s_tip[sB * 4 + 0]);
                        double q1 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 1], expB[c * 4 + 1]), s_tip[sB * 4 + 1]);
                        double q2 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 2], expB[c * 4 + 2]), s_tip[sB * 4 + 2]);
                        double q3 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 3], expB[c * 4 + 3]), s_tip[sB * 4 + 3]);
                        fb[x] = __dadd_rn(__dadd_rn(q0, q1), __dadd_rn(q2, q3));
                    }
                }
                const double t0 = fa[0] * fb[0], t1 = fa[1] * fb[1], t2 = fa[2] * fb[2], t3 = fa[3] * fb[3];
                unsigned chi = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const double v = fma(M.inv[i * 4 + 3], t3, fma(M.inv[i * 4 + 2], t2, fma(M.inv[i * 4 + 1], t1, M.inv[i * 4] * t0)));
                    cur[c * 4 + i] = v;
                    chi = max(chi, abs_hi(v));
                }
                if (SAFE) {
                    cursc[c] += scB[c];
                    if (ncase != 0 && may_scale && chi < kThresholdHi) {
#pragma unroll
                        for (int i = 0; i < 4; i++) cur[c * 4 + i] *= kTwo256;
                        cursc[c] += 1;
                    }
                } else mhi = max(mhi, chi);
                // store this category now: the LSU drains four stores per thread while the next category computes, instead
                // of a burst of 16 at the end of the op that the next op's first FP64 instruction then waits behind (ncu:
                // 18 % of all stall samples sat on the register hand-over of that burst)
                if (TST) {
#pragma unroll
                    for (int i = 0; i < 4; i++) s_out[o & 1][so + (c * 4 + i) * kWarp] = cur[c * 4 + i];
                } else if (ESTORE) {
#pragma unroll
                    for (int i = 0; i < 4; i++) Op[(c * 4 + i) * kWarp] = cur[c * 4 + i];
                }
            }
            if (!SAFE) {
                cursc[0] += scB[0];
                if (ncase != 0 && may_scale && mhi < kThresholdHi) {   // rare: all 16 values underflowed; rescale and store again
#pragma unroll
                    for (int k = 0; k < 16; k++) cur[k] *= kTwo256;
                    cursc[0] += 1;
                    if (TST) {
#pragma unroll
                        for (int k = 0; k < 16; k++) s_out[o & 1][so + k * kWarp] = cur[k];
                    } else if (ESTORE) {
#pragma unroll
                        for (int k = 0; k < 16; k++) Op[k * kWarp] = cur[k];
                    }
                }
            }
            if (!ESTORE) {
#pragma unroll
                for (int k = 0; k < 16; k++) Op[k * kWarp] = cur[k];
            }
#pragma unroll
            for (int c = 0; c < (SAFE ? 4 : 1); c++)
                reinterpret_cast<uint16_t *>(reinterpret_cast<char *>(D.scale) + off.dst_sc)[sb + (SAFE ? c * kWarp : 0)] = (uint16_t)cursc[c];
            buf = buf1;
            prev_dst = dst;
            prev_tile_off = tile * (16LL * kTile);
        }
        if (TST) {   // the tile's last CLV
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0 && prev_dst >= 0 && !(DBG & 1)) {
                bulk_store_issue(D.partial + (long long)prev_dst * D.slot_stride + prev_tile_off, s_out[(nops + 1) & 1], 16 * kTile * sizeof(double));
                bulk_store_wait_all();
            }
            prev_dst = -1;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// 20-state models (LG/WAG/JTT... +G/+I+G/+R, C10-C60 mixtures, fused classes): shared pieces of the FP64-MMA walkers
// (b200lk_walk20ss.cuh, b200lk_walk20v2.cuh).  Per pattern, category and child the update is a 20x20 mat-vec: on B200 this sits at
// the ridge between the FP64 pipe (36-37 TFLOP/s measured, DFMA and DMMA share it) and HBM (6.5 TB/s).
// History (config 3, ms per evaluation): generic thread-per-pattern kernel 30.6; DFMA with constant-bank operands 11.4; DMMA in
// eigen space with register-resident B fragments (round 1) 8.4 -> 6.5; state-space CLVs (2 mat-vecs per inner node): DESIGN.md.
// ------------------------------------------------------------------------------------------------------------
struct WalkJob20 {
    double lenA, lenB;
    int32_t dst, ncase, aid, bid;   // ncase: 0 tip-tip, 1 tip(A)-inner(B), 2 inner-inner; ids: taxon row or slot
};

// tip-table rows are padded by 2 doubles: a row stride of C*20 doubles is a multiple of 32 banks for C = 4, which would put the
// same (c, x) entry of every state in the same bank — a 20-way conflict when the lanes of a warp hold different states
__host__ __device__ inline int tab20_stride(int C) { return C * 20 + 2; }

// mma.sync.m8n8k4.f64 (DMMA).  Fragment layouts (PTX ISA), g = lane/4, q = lane%4: A[row g][k q], B[k q][col g],
// C[row g][cols 2q, 2q+1].  M = 8 patterns, K = 4, N = 8 states.
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// value index i that lane (., q) holds in K-slot q of k-tile kt when a CLV block is loaded as A fragments: two k-tiles per
// 16-byte load (values 2q, 2q+1 of state tile kt/2), the tail k-tile = values 16 + q
__host__ __device__ inline int clv_k(int kt, int q) { return kt < 4 ? 8 * (kt >> 1) + 2 * q + (kt & 1) : 16 + q; }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// block reduction of up to 3 doubles per thread into red[blockIdx.x*3 + j]; a 1-CTA kernel sums the per-block
// values in index order, so results are bit-reproducible for a given pattern count.
// ------------------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *red, long long blk = -1) {
    if (blk < 0) blk = blockIdx.x;
    __shared__ double s_red[NV][32];
#pragma unroll
    for (int j = 0; j < NV; j++)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[j] += __shfl_down_sync(0xffffffffu, v[j], off);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int j = 0; j < NV; j++) s_red[j][wid] = v[j];
    __syncthreads();
    if (wid == 0) {
        const int nw = blockDim.x >> 5;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            double x = lane < nw ? s_red[j][lane] : 0.0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
            if (lane == 0) red[blk * NV + j] = x;
        }
    }
}

// block_reduce_store + the final sum by the LAST block to arrive (ticket counter) in a fixed order (256 virtual threads, thread t
// sums blocks t, t+256, ... in order, then a binary tree): bit-identical to the separate final-reduction launch it replaced.  `result` may be mapped host memory: the scalars land in the caller's buffer without a copy.
// (blk, nblocks): this CTA's index among the CTAs that reduce into `result` — the whole grid, or one member of a grouped launch
template <int NV>
__device__ __forceinline__ void block_reduce_finish(double (&v)[NV], double *red, unsigned *ticket, double *result, long long blk = -1, int nblocks = -1) {
    if (blk < 0) { blk = blockIdx.x; nblocks = gridDim.x; }
    block_reduce_store<NV>(v, red, blk);
    __shared__ bool s_last;
    __shared__ double s_fin[NV][256];
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == (unsigned)nblocks - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int vt = threadIdx.x; vt < 256; vt += blockDim.x) {
        double acc[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) acc[j] = 0.0;
        for (int b = vt; b < nblocks; b += 256)
#pragma unroll
            for (int j = 0; j < NV; j++) acc[j] += __ldcg(red + (long long)b * NV + j);
#pragma unroll
        for (int j = 0; j < NV; j++) s_fin[j][vt] = acc[j];
    }
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        for (int vt = threadIdx.x; vt < off; vt += blockDim.x)
#pragma unroll
            for (int j = 0; j < NV; j++) s_fin[j][vt] += s_fin[j][vt + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NV; j++) result[j] = s_fin[j][0];
        *ticket = 0;
    }
}

// ------------------------------------------------------------------------------------------------------------
// computeLikelihoodBranch{,Generic}SIMD root stage (tree/phylokernelnew.h:2753-3104): per-pattern likelihood across
// the evaluation branch, _pattern_lh / _pattern_lh_cat, and the frequency-weighted sum.
// ------------------------------------------------------------------------------------------------------------
// ASC (Lewis' ascertainment-bias correction, tree/phylokernelnew.h:3076-3091): the unobserved constant patterns [Pobs, P)
// contribute their likelihood (un-scaled once if their scale count is non-zero) to prob_const instead of to the tree logL; the
// reduction then carries (tree_lh, prob_const) and k_asc_finish_logl applies logL -= nsite*log(1-prob_const) (:3198-3221).
template <bool SAFE, bool ASC>
__device__ __forceinline__ void root_body(const Dev &D, long long blk, int nblocks, int node_slot, int dad_kind, int dad_id, int save_log, double *red,
                                          unsigned *ticket, double *result, double length, int fuse_vals, double *s_vals_dyn) {
    const double *vals = fuse_vals ? stage_vals(D, length, 0, s_vals_dyn) : D.val;
    const int n = D.n, C = D.C;
    const long long p = blk * kTile + threadIdx.x;
    const long long cb = pbase(p, C);
    const long long sb = SAFE ? cb : p;
    const ClvIdx at(D, p);
    const double *dad = D.partial + (long long)node_slot * D.slot_stride;       // dad_branch->partial_lh
    const uint16_t *dsc = D.scale + (long long)node_slot * D.scale_stride + sb;
    const double *node = nullptr, *tip = nullptr;
    const uint16_t *nsc = nullptr;
    if (dad_kind == 1) {
        node = D.partial + (long long)dad_id * D.slot_stride;
        nsc = D.scale + (long long)dad_id * D.scale_stride + sb;
    } else {
        tip = D.tip_lh + (long long)D.tips[(long long)dad_id * D.Ppad + p] * (n * D.nmix);
    }
    double *lh_cat = D.pattern_lh_cat + cb;
    double lh_ptn = 0.0;
    unsigned min_scale = 0xffffffffu;
    for (int c = 0; c < C; c++) {
        const double *val = vals + c * n;
        const double *tp = tip ? tip + (c / D.denom) * n : nullptr;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int i = 0;
        for (; i + 3 < n; i += 4) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double f = val[i + j] * (node ? node[at(c, i + j)] : __ldg(tp + i + j));
                double d = dad[at(c, i + j)];
                if (j == 0) a0 = fma(f, d, a0); else if (j == 1) a1 = fma(f, d, a1);
                else if (j == 2) a2 = fma(f, d, a2); else a3 = fma(f, d, a3);
            }
        }
        double s = (a0 + a1) + (a2 + a3);
        for (; i < n; i++) {
            double f = val[i] * (node ? node[at(c, i)] : __ldg(tp + i));
            s = fma(f, dad[at(c, i)], s);
        }
        if (SAFE) {
            unsigned sc = dsc[c * kWarp] + (nsc ? nsc[c * kWarp] : 0u);
            min_scale = min(min_scale, sc);
            lh_cat[c * kWarp] = s;       // rescaled in the second pass
        } else {
            lh_cat[c * kWarp] = s;
            lh_ptn += s;
        }
    }
    double vc_min_scale;
    if (SAFE) {
        // align categories to the minimum scale count (:2891-2915, :3028-3053), then sumVec
        double v0 = 0, v1 = 0, v2 = 0, v3 = 0, tail = 0;
        const int c4 = C & ~3;
        for (int c = 0; c < C; c++) {
            unsigned sc = dsc[c * kWarp] + (nsc ? nsc[c * kWarp] : 0u);
            double x = lh_cat[c * kWarp];
            if (sc == min_scale + 1) x *= kScalingThreshold;
            else if (sc > min_scale + 1) x = 0.0;
            lh_cat[c * kWarp] = x;
            if (c < c4) { if ((c & 3) == 0) v0 += x; else if ((c & 3) == 1) v1 += x; else if ((c & 3) == 2) v2 += x; else v3 += x; }
            else tail += x;
        }
        v0 += v2; v1 += v3; v0 += v1;
        lh_ptn = (c4 ? v0 : 0.0) + tail;
        vc_min_scale = (double)min_scale;
    } else {
        vc_min_scale = (double)(dsc[0] + (nsc ? nsc[0] : 0u));
    }
    vc_min_scale *= kLogScalingThreshold;
    lh_ptn = fabs(lh_ptn) + D.invar[p];
    double lg = log(lh_ptn) + vc_min_scale;
    double acc[ASC ? 2 : 1] = {0.0};
    if (p < D.Pobs) {
        if (save_log) D.pattern_lh[p] = lg;
        else { D.pattern_lh[p] = lh_ptn; D.pattern_scaling[p] = vc_min_scale; }
        acc[0] = lg * D.freq[p];
    } else if (ASC && p < D.P) {
        const double lhc = vc_min_scale != 0.0 ? lh_ptn * kScalingThreshold : lh_ptn;
        if (D.asc == 2) D.asc_tmp[p - D.Pobs] = lhc;   // Holder keeps it per pattern (:3088-3089)
        else acc[ASC ? 1 : 0] = lhc;
        D.pattern_lh[p] = 0.0;
    }
    block_reduce_finish<ASC ? 2 : 1>(acc, red, ticket, result, blk, nblocks);
}
template <bool SAFE, bool ASC = false>
__global__ void __launch_bounds__(kTile) k_root(Dev D, int node_slot, int dad_kind, int dad_id, int save_log, double *red,
                                                unsigned *ticket, double *result, double length, int fuse_vals) {
    extern __shared__ double s_vals_dyn[];
    root_body<SAFE, ASC>(D, blockIdx.x, gridDim.x, node_slot, dad_kind, dad_id, save_log, red, ticket, result, length, fuse_vals, s_vals_dyn);
}

// +ASC epilogues: raw = (tree_lh, prob_const) -> result[0] = logL, _pattern_lh -= log(1 - prob_const) (:3198-3221, 3412-3426);
// raw = (df, ddf, prob_const, df_const, ddf_const) -> result[0..1] = (df, ddf) (:2618-2626)
__global__ void __launch_bounds__(256) k_asc_finish_logl(Dev D, const double *raw, double *result) {
    const double corr = log(1.0 - raw[1]);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < D.Pobs; p += (long long)gridDim.x * blockDim.x) D.pattern_lh[p] -= corr;
    if (blockIdx.x == 0 && threadIdx.x == 0) result[0] = raw[0] - (double)D.asc_nsite * corr;
}
__global__ void k_asc_finish_derv(Dev D, const double *raw, double *result) {
    const double all = 1.0 - raw[2], df_frac = raw[3] / all, ddf_frac = raw[4] / all;
    result[0] = raw[0] + (double)D.asc_nsite * df_frac;
    result[1] = raw[1] + (double)D.asc_nsite * (ddf_frac + df_frac * df_frac);
}
// E-step of the reference's EM algorithms (RateFree::optimizeWithEM, model/ratefree.cpp:556-573; ModelMixture::optimizeWithEM,
// model/modelmixture.cpp:2008-2024) on the per-category pattern likelihoods of the last k_root, marginalised as
// PhyloTree::computePatternLhCat(wsl) does (tree/phylotree.cpp:1366-1391): mode 0 = every category (K = C), 1 = per rate category
// (sum over the mixture classes, K = ncat), 2 = per mixture class (sum over the rate categories, K = nmix).
//   post[k][p] = freq_p L_pk / (invar_p*add_invar + sum_k L_pk)   (column-major: a column is a ptn_freq vector for the M-step)
//   red[block][k] = sum over the block's patterns of post[k][p]   (k_em_reduce adds the blocks in order)
__device__ __forceinline__ double em_marginal(const double *lh_cat, int mode, int k, int ncat, int nmix) {
    if (mode == 0) return lh_cat[(long long)k * kWarp];
    if (mode == 1) {
        double v = lh_cat[(long long)k * kWarp];
        for (int m = 1; m < nmix; m++) v += lh_cat[(long long)(m * ncat + k) * kWarp];
        return v;
    }
    double v = lh_cat[(long long)(k * ncat) * kWarp];
    for (int c = 1; c < ncat; c++) v += lh_cat[(long long)(k * ncat + c) * kWarp];
    return v;
}
__global__ void __launch_bounds__(kTile) k_em_posterior(Dev D, int mode, int K, int add_invar, double *post, double *red) {
    extern __shared__ double s_em[];   // [warps][K]
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool live = p < D.Pobs;
    const double *lh_cat = D.pattern_lh_cat + pbase(p, D.C);
    double w = 0.0;
    if (live) {
        double lk = add_invar ? D.invar[p] : 0.0;
        for (int k = 0; k < K; k++) lk += em_marginal(lh_cat, mode, k, D.ncat, D.nmix);
        w = D.freq[p] / lk;
    }
    for (int k = 0; k < K; k++) {
        double v = live ? em_marginal(lh_cat, mode, k, D.ncat, D.nmix) * w : 0.0;
        post[(long long)k * D.Ppad + p] = v;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) s_em[wid * K + k] = v;
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        double v = s_em[k];
        for (int w2 = 1; w2 < nw; w2++) v += s_em[w2 * K + k];
        red[(long long)blockIdx.x * K + k] = v;
    }
}
__global__ void k_em_reduce(int nblocks, int K, const double *red, double *out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double v = 0.0;
    for (int b = 0; b < nblocks; b++) v += red[(long long)b * K + k];
    out[k] = v;
}
// dst ptn_freq := one column of the posteriors (the M-step trees of the EM algorithms: ratefree.cpp:636-640)
__global__ void k_copy_column(long long P, const double *col, double *dst) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) dst[p] = col[p];
}

// robust phylo (:3134-3173): lh / f = pattern log-likelihoods in ascending order with their frequencies, cum = running sum of f.
// Pattern i holds the site ranks [cum[i] - f[i], cum[i]).  keep < 1: sum of the site log-likelihoods from rank
// (int)((1 - keep) * sites) on; median: the site of rank sites/2 (odd) or the mean of ranks sites/2 - 1 and sites/2 (even).
__global__ void __launch_bounds__(kTile) k_robust_phylo(long long P, const double *lh, const double *f, const double *cum, double keep,
                                                        int median, double *red, unsigned *ticket, double *result) {
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    double acc[1] = {0.0};
    if (p < P && f[p] > 0.0) {
        const double sites = cum[P - 1], hi = cum[p], lo = hi - f[p];
        if (!median) {
            const double drop = (double)(int)((1.0 - keep) * sites);
            const double cnt = hi - fmax(lo, drop);
            if (cnt > 0.0) acc[0] = lh[p] * cnt;
        } else {
            const long long s = (long long)sites;
            const double m = (double)(s / 2);
            double w = 0.0;
            if (lo <= m && m < hi) w += (s & 1) ? 1.0 : 0.5;
            if (!(s & 1) && lo <= m - 1.0 && m - 1.0 < hi) w += 0.5;
            if (w > 0.0) acc[0] = lh[p] * w;
        }
    }
    block_reduce_finish<1>(acc, red, ticket, result);
}

// Holder's correction for missing data (+ASC_MIS; :3175-3197, 3385-3410): per observed pattern p the likelihoods of the n
// constant patterns that share p's gaps (asc_tmp[s*step + p]) give _pattern_lh[p] -= log(1 - sum_s) and
// logL = raw[0] - sum_p freq_p log(1 - sum_s).  One thread per pattern, same grid as k_root.
__global__ void __launch_bounds__(kTile) k_asc_holder_logl(Dev D, const double *raw, double *red, unsigned *ticket, double *result) {
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    double acc[1] = {0.0};
    if (p < D.Pobs) {
        double c = D.asc_tmp[p];
        for (int s = 1; s < D.n; s++) c += D.asc_tmp[(long long)s * D.asc_step + p];
        const double prob_variant = log(1.0 - c);
        D.pattern_lh[p] -= prob_variant;
        acc[0] = -prob_variant * D.freq[p];
    }
    if (p == 0) acc[0] += raw[0];
    block_reduce_finish<1>(acc, red, ticket, result);
}
// (:2580-2617): df += sum_p freq_p df_frac, ddf += sum_p freq_p (ddf_frac + df_frac^2) with the per-pattern sums over the n groups
__global__ void __launch_bounds__(kTile) k_asc_holder_derv(Dev D, const double *raw, double *red, unsigned *ticket, double *result) {
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    const long long U = D.P - D.Pobs;
    double acc[2] = {0.0, 0.0};
    if (p < D.Pobs) {
        double c = D.asc_tmp[p], d1 = D.asc_tmp[U + p], d2 = D.asc_tmp[2 * U + p];
        for (int s = 1; s < D.n; s++) {
            const long long u = (long long)s * D.asc_step + p;
            c += D.asc_tmp[u]; d1 += D.asc_tmp[U + u]; d2 += D.asc_tmp[2 * U + u];
        }
        c = 1.0 - c;
        const double df_frac = d1 / c, ddf_frac = d2 / c, freq = D.freq[p];
        acc[0] = freq * df_frac;
        acc[1] = freq * (ddf_frac + df_frac * df_frac);
    }
    if (p == 0) { acc[0] += raw[0]; acc[1] += raw[1]; }
    block_reduce_finish<2>(acc, red, ticket, result);
}

// computeLikelihoodBuffer{,Generic}SIMD (tree/phylokernelnew.h:2091-2230): theta_all and buffer_scale_all
template <bool SAFE>
__device__ __forceinline__ void theta_body(const Dev &D, long long blk, int node_slot, int dad_kind, int dad_id) {
    const int n = D.n, C = D.C, B = D.B;
    const long long p = blk * kTile + threadIdx.x;
    const long long pb = pbase(p, B);
    const long long sb = SAFE ? pbase(p, C) : p;
    const ClvIdx at(D, p);
    const double *dad = D.partial + (long long)node_slot * D.slot_stride;
    const uint16_t *dsc = D.scale + (long long)node_slot * D.scale_stride + sb;
    const double *node = nullptr, *tip = nullptr;
    const uint16_t *nsc = nullptr;
    if (dad_kind == 1) {
        node = D.partial + (long long)dad_id * D.slot_stride;
        nsc = D.scale + (long long)dad_id * D.scale_stride + sb;
    } else {
        tip = D.tip_lh + (long long)D.tips[(long long)dad_id * D.Ppad + p] * (n * D.nmix);
    }
    double *theta = D.theta + pb;
    unsigned min_scale = 0xffffffffu;
    if (SAFE)
        for (int c = 0; c < C; c++) min_scale = min(min_scale, dsc[c * kWarp] + (nsc ? nsc[c * kWarp] : 0u));
    for (int c = 0; c < C; c++) {
        double f = 1.0;
        bool zero = false;
        if (SAFE) {
            unsigned sc = dsc[c * kWarp] + (nsc ? nsc[c * kWarp] : 0u);
            if (sc == min_scale + 1) f = kScalingThreshold;
            else if (sc > min_scale + 1) zero = true;
        }
        const double *tp = tip ? tip + (c / D.denom) * n : nullptr;
        for (int i = 0; i < n; i++) {
            const long long k = (long long)(c * n + i) * kWarp, kc = at(c, i);
            double t = (node ? node[kc] : __ldg(tp + i)) * dad[kc];
            if (SAFE) t = zero ? 0.0 : (f != 1.0 ? t * f : t);
            theta[k] = t;
        }
    }
    double bs = SAFE ? (double)min_scale : (double)(dsc[0] + (nsc ? nsc[0] : 0u));
    D.buffer_scale[p] = bs * kLogScalingThreshold;
}
template <bool SAFE>
__global__ void __launch_bounds__(kTile) k_theta(Dev D, int node_slot, int dad_kind, int dad_id) {
    theta_body<SAFE>(D, blockIdx.x, node_slot, dad_kind, dad_id);
}

// computeLikelihoodDerv{,Generic}SIMD pattern loop (tree/phylokernelnew.h:2463-2505): (lh, df, ddf) = val{0,1,2}.theta
template <int BT, bool ASC>
__device__ __forceinline__ void derv_body(const Dev &D, long long blk, int nblocks, double *red, unsigned *ticket, double *result, double length, int fuse_vals,
                                          double *s_vals_dyn) {
    const double *vals = fuse_vals ? stage_vals(D, length, 1, s_vals_dyn) : D.val;
    const int B = BT ? BT : D.B;
    const long long p = blk * kTile + threadIdx.x;
    const double *theta = D.theta + pbase(p, B);
    const double *v0 = vals, *v1 = vals + B, *v2 = vals + 2 * B;
    // dotProductTriple: two lane accumulators (even / odd k)
    double ad[2] = {0, 0}, bd[2] = {0, 0}, cd[2] = {0, 0};
    int k = 0;
#pragma unroll 8
    for (; k + 1 < B; k += 2) {
        double t0 = theta[(long long)k * kWarp], t1 = theta[(long long)(k + 1) * kWarp];
        ad[0] = fma(v0[k], t0, ad[0]); ad[1] = fma(v0[k + 1], t1, ad[1]);
        bd[0] = fma(v1[k], t0, bd[0]); bd[1] = fma(v1[k + 1], t1, bd[1]);
        cd[0] = fma(v2[k], t0, cd[0]); cd[1] = fma(v2[k + 1], t1, cd[1]);
    }
    double lh = ad[0] + ad[1], df = bd[0] + bd[1], ddf = cd[0] + cd[1];
    if (k < B) {
        double t = theta[(long long)k * kWarp];
        lh = fma(v0[k], t, lh); df = fma(v1[k], t, df); ddf = fma(v2[k], t, ddf);
    } else if ((D.quirks & 1) && (D.n & 1)) {
        // B200LK_QUIRK_DOTTRIPLE_ODD: the reference's dotProductTriple takes its odd-length branch on nstates&1 and, when
        // the block is even, adds the last element a second time (tree/phylokernelnew.h:700-714)
        double t = theta[(long long)(B - 1) * kWarp];
        lh = fma(v0[B - 1], t, lh); df = fma(v1[B - 1], t, df); ddf = fma(v2[B - 1], t, ddf);
    }
    lh = fabs(lh) + D.invar[p];
    double acc[ASC ? 5 : 2] = {0.0, 0.0};
    if (p < D.Pobs) {
        double inv = 1.0 / lh;
        double df_frac = df * inv, ddf_frac = ddf * inv;
        double freq = D.freq[p];
        double tmp1 = df_frac * freq, tmp2 = ddf_frac * freq;
        acc[0] = tmp1;
        acc[1] = fma(-tmp1, df_frac, tmp2);   // nmul_add(tmp1, df_frac, tmp2) :2505
    } else if (ASC && p < D.P) {   // :2506-2535
        const double f = D.buffer_scale[p] != 0.0 ? kScalingThreshold : 1.0;
        if (D.asc == 2) {
            const long long U = D.P - D.Pobs, u = p - D.Pobs;
            D.asc_tmp[u] = lh * f; D.asc_tmp[U + u] = df * f; D.asc_tmp[2 * U + u] = ddf * f;
        } else {
            acc[ASC ? 2 : 0] = lh * f; acc[ASC ? 3 : 0] = df * f; acc[ASC ? 4 : 0] = ddf * f;
        }
    }
    block_reduce_finish<ASC ? 5 : 2>(acc, red, ticket, result, blk, nblocks);
}
template <int BT, bool ASC = false>
__global__ void __launch_bounds__(kTile) k_derv(Dev D, double *red, unsigned *ticket, double *result, double length, int fuse_vals) {
    extern __shared__ double s_vals_dyn[];
    derv_body<BT, ASC>(D, blockIdx.x, gridDim.x, red, ticket, result, length, fuse_vals, s_vals_dyn);
}

// ------------------------------------------------------------------------------------------------------------
// The whole Newton-Raphson iteration of optimizeOneBranch on the device: Optimization::minimizeNewton
// (utils/optimization.cpp:422-499, statement for statement) driving computeFuncDerv (tree/phylotree.cpp:2529-2536) =
// the (df, ddf) pass of computeLikelihoodDerv on the resident theta buffer (k_derv's arithmetic).  One COOPERATIVE launch:
// every NR step is a grid-stride pass over the patterns, per-block partial sums, one grid barrier, and the same fixed-order
// final sum in every CTA — so every thread holds the same (f, df) and walks the same control flow; no host round trip per
// step (a host-driven step costs a launch, a synchronisation and ~10 us of latency; a 60-taxon x 4 000-pattern sweep spent
// 90 us per branch that way).  Results go to mapped host memory: io[0] = optimised length, io[1] = derv evaluations,
// io[2] = status (0 ok, 2 = non-finite derivative: df = ddf = 0 semantics of tree/phylokernelnew.h:2629-2632).
// ------------------------------------------------------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(kTile) k_newton_coop(Dev D, double *red, double *io, double x1, double xguess, double x2, double xacc,
                                                       int maxNRStep, double *len_out, int check_divergence) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double s_vals_dyn[];
    __shared__ double s_fin[2][256];
    const int B = BT ? BT : D.B;
    int nevals = 0;
    auto func_derv = [&](double x, double &f, double &df) {
        const double *vals = stage_vals(D, x, 1, s_vals_dyn);
        const double *v0 = vals, *v1 = vals + B, *v2 = vals + 2 * B;
        double acc[2] = {0.0, 0.0};
        for (long long p = (long long)blockIdx.x * kTile + threadIdx.x; p < D.Ppad; p += (long long)gridDim.x * kTile) {
            const double *theta = D.theta + pbase(p, B);
            double ad[2] = {0, 0}, bd[2] = {0, 0}, cd[2] = {0, 0};
            int k = 0;
#pragma unroll 8
            for (; k + 1 < B; k += 2) {
                double t0 = theta[(long long)k * kWarp], t1 = theta[(long long)(k + 1) * kWarp];
                ad[0] = fma(v0[k], t0, ad[0]); ad[1] = fma(v0[k + 1], t1, ad[1]);
                bd[0] = fma(v1[k], t0, bd[0]); bd[1] = fma(v1[k + 1], t1, bd[1]);
                cd[0] = fma(v2[k], t0, cd[0]); cd[1] = fma(v2[k + 1], t1, cd[1]);
            }
            double lh = ad[0] + ad[1], d1 = bd[0] + bd[1], d2 = cd[0] + cd[1];
            if (k < B) {
                double t = theta[(long long)k * kWarp];
                lh = fma(v0[k], t, lh); d1 = fma(v1[k], t, d1); d2 = fma(v2[k], t, d2);
            } else if ((D.quirks & 1) && (D.n & 1)) {
                double t = theta[(long long)(B - 1) * kWarp];
                lh = fma(v0[B - 1], t, lh); d1 = fma(v1[B - 1], t, d1); d2 = fma(v2[B - 1], t, d2);
            }
            lh = fabs(lh) + D.invar[p];
            if (p < D.P) {
                double inv = 1.0 / lh;
                double df_frac = d1 * inv, ddf_frac = d2 * inv;
                double freq = D.freq[p];
                double tmp1 = df_frac * freq, tmp2 = ddf_frac * freq;
                acc[0] += tmp1;
                acc[1] += fma(-tmp1, df_frac, tmp2);
            }
        }
        double *r = red + (size_t)(nevals & 1) * 2 * gridDim.x;   // double-buffered by evaluation parity: one barrier per step
        block_reduce_store<2>(acc, r);
        __threadfence();
        grid.sync();
        const int nblocks = gridDim.x;
        for (int vt = threadIdx.x; vt < 256; vt += blockDim.x) {
            double a0 = 0.0, a1 = 0.0;
            for (int b = vt; b < nblocks; b += 256) { a0 += __ldcg(r + 2 * b); a1 += __ldcg(r + 2 * b + 1); }
            s_fin[0][vt] = a0; s_fin[1][vt] = a1;
        }
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            for (int vt = threadIdx.x; vt < off; vt += blockDim.x) { s_fin[0][vt] += s_fin[0][vt + off]; s_fin[1][vt] += s_fin[1][vt + off]; }
            __syncthreads();
        }
        f = -s_fin[0][0];      // computeFuncDerv returns -df, -ddf
        df = -s_fin[1][0];
        __syncthreads();
        nevals++;
    };
    // computeLikelihoodFromBuffer's sum at length x (tree/phylokernelnew.h:3313-3340) for the divergence test below; does not
    // touch _pattern_lh
    auto logl_at = [&](double x) -> double {
        const double *vals = stage_vals(D, x, 0, s_vals_dyn);
        double acc[2] = {0.0, 0.0};
        for (long long p = (long long)blockIdx.x * kTile + threadIdx.x; p < D.P; p += (long long)gridDim.x * kTile) {
            const double *theta = D.theta + pbase(p, B);
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            int k = 0;
            for (; k + 3 < B; k += 4) {
                a0 = fma(vals[k], theta[(long long)k * kWarp], a0); a1 = fma(vals[k + 1], theta[(long long)(k + 1) * kWarp], a1);
                a2 = fma(vals[k + 2], theta[(long long)(k + 2) * kWarp], a2); a3 = fma(vals[k + 3], theta[(long long)(k + 3) * kWarp], a3);
            }
            double lh = (a0 + a1) + (a2 + a3);
            for (; k < B; k++) lh = fma(vals[k], theta[(long long)k * kWarp], lh);
            lh = fabs(lh) + D.invar[p];
            acc[0] += (log(fabs(lh)) + D.buffer_scale[p]) * D.freq[p];
        }
        double *r = red + (size_t)(nevals & 1) * 2 * gridDim.x;
        block_reduce_store<2>(acc, r);
        __threadfence();
        grid.sync();
        for (int vt = threadIdx.x; vt < 256; vt += blockDim.x) {
            double a = 0.0;
            for (int b = vt; b < (int)gridDim.x; b += 256) a += __ldcg(r + 2 * b);
            s_fin[0][vt] = a;
        }
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            for (int vt = threadIdx.x; vt < off; vt += blockDim.x) s_fin[0][vt] += s_fin[0][vt + off];
            __syncthreads();
        }
        const double v = s_fin[0][0];
        __syncthreads();
        nevals++;
        return v;
    };
    auto finish = [&](double x, double status) {
        // PhyloTree::optimizeOneBranch, tree/phylotree.cpp:2616-2625: Newton-Raphson diverged — keep the better of the two
        if (check_divergence && status == 0.0 && x > x2 * 0.95) {
            const double opt_lh = logl_at(x), orig_lh = logl_at(xguess);
            if (orig_lh > opt_lh) x = xguess;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            io[0] = x; io[1] = (double)nevals; io[2] = status;
            if (len_out) *len_out = x;
        }
    };
    double df, dx, f, temp, xh, xl, rts, rts_old;
    rts = xguess;
    if (rts < x1) rts = x1;
    if (rts > x2) rts = x2;
    func_derv(rts, f, df);
    if (!isfinite(f) || !isfinite(df)) { finish(rts, 2.0); return; }
    if (df >= 0.0 && fabs(f) < xacc) { finish(rts, 0.0); return; }
    if (f < 0.0) { xl = rts; xh = x2; } else { xh = rts; xl = x1; }
    dx = fabs(xh - xl);
    for (int j = 1; j <= maxNRStep; j++) {
        rts_old = rts;
        if ((df <= 0.0) || (((rts - xh) * df - f) * ((rts - xl) * df - f) >= 0.0)) {
            dx = 0.5 * (xh - xl);
            rts = xl + dx;
            if (xl == rts) { finish(rts, 0.0); return; }
        } else {
            dx = f / df;
            temp = rts;
            rts -= dx;
            if (temp == rts) { finish(rts, 0.0); return; }
        }
        if (fabs(dx) < xacc || (j == maxNRStep)) { finish(rts_old, 0.0); return; }
        func_derv(rts, f, df);
        if (!isfinite(f) || !isfinite(df)) { finish(rts, 2.0); return; }
        if (df > 0.0 && fabs(f) < xacc) { finish(rts, 0.0); return; }
        if (f < 0.0) xl = rts;
        else if (f > 0.0) xh = rts;
    }
    finish(rts, 0.0);
}

// computeLikelihoodFromBuffer{,Generic}SIMD pattern loop (tree/phylokernelnew.h:3313-3368)
template <int BT, bool ASC>
__device__ __forceinline__ void from_buffer_body(const Dev &D, long long blk, int nblocks, double *red, unsigned *ticket, double *result, double length,
                                                 int fuse_vals, double *s_vals_dyn) {
    const double *val0 = fuse_vals ? stage_vals(D, length, 0, s_vals_dyn) : D.val;
    const int B = BT ? BT : D.B;
    const long long p = blk * kTile + threadIdx.x;
    const double *theta = D.theta + pbase(p, B);
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int k = 0;
#pragma unroll 4
    for (; k + 3 < B; k += 4) {
        a0 = fma(val0[k], theta[(long long)k * kWarp], a0);
        a1 = fma(val0[k + 1], theta[(long long)(k + 1) * kWarp], a1);
        a2 = fma(val0[k + 2], theta[(long long)(k + 2) * kWarp], a2);
        a3 = fma(val0[k + 3], theta[(long long)(k + 3) * kWarp], a3);
    }
    double lh = (a0 + a1) + (a2 + a3);
    for (; k < B; k++) lh = fma(val0[k], theta[(long long)k * kWarp], lh);
    lh = fabs(lh) + D.invar[p];
    double acc[ASC ? 2 : 1] = {0.0};
    if (p < D.Pobs) {
        double lg = log(fabs(lh)) + D.buffer_scale[p];
        D.pattern_lh[p] = lg;
        acc[0] = lg * D.freq[p];
    } else if (ASC && p < D.P) {   // :3342-3356
        const double lhc = D.buffer_scale[p] != 0.0 ? lh * kScalingThreshold : lh;
        if (D.asc == 2) D.asc_tmp[p - D.Pobs] = lhc;
        else acc[ASC ? 1 : 0] = lhc;
        D.pattern_lh[p] = 0.0;
    }
    block_reduce_finish<ASC ? 2 : 1>(acc, red, ticket, result, blk, nblocks);
}
template <int BT, bool ASC = false>
__global__ void __launch_bounds__(kTile) k_from_buffer(Dev D, double *red, unsigned *ticket, double *result, double length, int fuse_vals) {
    extern __shared__ double s_vals_dyn[];
    from_buffer_body<BT, ASC>(D, blockIdx.x, gridDim.x, red, ticket, result, length, fuse_vals, s_vals_dyn);
}

// layout converters for read-backs: device warp-interleaved -> reference V=1 layout [p*W + k]
__global__ void k_deinterleave_f64(const double *src, double *dst, long long P, int W) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P * W) return;
    long long p = idx / W;
    int k = (int)(idx - p * W);
    dst[idx] = src[pbase(p, W) + (long long)k * kWarp];
}
// ------------------------------------------------------------------------------------------------------------
// UFBoot RELL (IQTree::saveCurrentTree, tree/iqtree.cpp:3551-3637): rell[b] = sum_p (float)_pattern_lh[p] * boot_sample[b][p]
// for every bootstrap replicate — the reference's dotProductSIMD<float> (tree/phylokernel.h:64-70) over BootValType = float
// arrays, once per candidate tree.  One CTA per replicate streams that replicate's resampled frequencies (16-byte loads);
// _pattern_lh stays on the device (L2-resident across the CTAs).  Products in float as in the reference, accumulated in double
// in a fixed order (thread t takes elements 4(t + 256 j) .. +3), so the result is reproducible and at least as accurate as the
// reference's 8-lane float accumulation.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rell(const double *pattern_lh, const float *boot, long long nptn, long long stride, float *out) {
    const float *row = boot + (long long)blockIdx.x * stride;
    double acc = 0.0;
    const long long n4 = nptn >> 2;
    for (long long k = threadIdx.x; k < n4; k += 256) {
        const float4 f = __ldcs(reinterpret_cast<const float4 *>(row) + k);
        const double2 l0 = *reinterpret_cast<const double2 *>(pattern_lh + 4 * k), l1 = *reinterpret_cast<const double2 *>(pattern_lh + 4 * k + 2);
        acc += (double)((float)l0.x * f.x);
        acc += (double)((float)l0.y * f.y);
        acc += (double)((float)l1.x * f.z);
        acc += (double)((float)l1.y * f.w);
    }
    for (long long p = (n4 << 2) + threadIdx.x; p < nptn; p += 256) acc += (double)((float)pattern_lh[p] * row[p]);
    __shared__ double s[256];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) s[threadIdx.x] += s[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = (float)s[0];
}

// UFBoot bookkeeping (IQTree::saveCurrentTree, tree/iqtree.cpp:3593-3632) for all replicates at once: one thread per replicate.
//   better = rell > boot_logl + eps;  a tie (within eps) wins with probability 1 / (boot_counts + 1) — the caller's uniform number;
//   a winner resets (clear win) or bumps (tie) boot_counts, boot_logl = max(boot_logl, rell), boot_orig_logl = cur_logl.
// updated[b] = 1 tells the host to remember the current tree for replicate b (boot_trees[b] = tree_str); *nupdated counts them.
__global__ void k_rell_update(int nboot, const float *rell, double cur_logl, double eps, const double *uniform, double *boot_logl, int *boot_counts,
                              double *boot_orig_logl, unsigned char *updated, int *nupdated) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nboot) return;
    const double r = (double)rell[b], cur = boot_logl[b];
    bool better = r > cur + eps;
    if (!better && r > cur - eps) better = (uniform ? uniform[b] : 2.0) <= 1.0 / (double)(boot_counts[b] + 1);
    if (better) {
        if (r <= cur + eps) boot_counts[b]++;
        else boot_counts[b] = 1;
        boot_logl[b] = fmax(cur, r);
        boot_orig_logl[b] = cur_logl;
        atomicAdd(nupdated, 1);
    }
    updated[b] = better ? 1 : 0;
}
__global__ void k_boot_reset(int nboot, double *boot_logl, int *boot_counts, double *boot_orig_logl) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nboot) return;
    boot_logl[b] = -1.7976931348623157e308;   // -DBL_MAX (tree/iqtree.cpp:375-378)
    boot_orig_logl[b] = -1.7976931348623157e308;
    boot_counts[b] = 0;
}

__global__ void k_deinterleave_clv(Dev D, const double *src, double *dst) {   // one CLV slot -> [p][c*n + i]
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= D.P * D.B) return;
    long long p = idx / D.B;
    int k = (int)(idx - p * D.B);
    const ClvIdx at(D, p);
    dst[idx] = src[at(k / D.n, k % D.n)];
}
__global__ void k_deinterleave_u16(const uint16_t *src, uint16_t *dst, long long P, int W) {
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P * W) return;
    long long p = idx / W;
    int k = (int)(idx - p * W);
    dst[idx] = src[pbase(p, W) + (long long)k * kWarp];
}

}  // namespace b200lk
