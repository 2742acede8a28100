// b200lk_nonrev.cuh — the kernels for NON-REVERSIBLE models: the device side of tree/phylokernelnonrev.h (selected in the reference when
// !model->useRevKernel(): a non-reversible model — UNREST, the Lie-Markov family, NONREV protein matrices — or any model under
// --kernel-nonrev; model/modelsubst.h:73-75, pointers assigned through setLikelihoodKernel, tree/phylokernelfma.cpp).
//
// What differs from the reversible path (b200lk_kernels.cuh):
//   * CLVs live in STATE space in every slot; per (child, category) the operator is the transition matrix P_m(rate_c * len) itself,
//     TRANSPOSED when the child's branch points towards the root (computePartialInfo's non-reversible branch,
//     tree/phylokernelnew.h:932-1008).  The matrices are built ON THE DEVICE from the complex eigensystem the reference's default
//     technique uses, P = Re(cevec . diag(exp(ceval t)) . cinv_evec) (ModelMarkov::computeTransMatrixNonrev,
//     model/modelmarkov.cpp:477-486), uploaded once with the model: an evaluation still uploads only its op list;
//   * the virtual root leaf of a rooted tree (child kind 2) contributes the state frequencies, without a matrix (:959-964);
//   * tip_partial_lh are the plain state indicators (tree/phylotreesse.cpp:353-375);
//   * the 2^-256 test is on max(value), signed (phylokernelnonrev.h:446-447, 520), and the root-internal node case never rescales
//     (:353-388);
//   * the branch likelihood (:1062-1528) and the derivatives (:586-1056) read both ends' CLVs at every call — no theta buffer, no
//     from-buffer form (the reference has none: computeLikelihoodFromBufferPointer stays NULL); derivative matrices Q.P and Q.Q.P
//     (model/modelmarkov.cpp:735-747).
// Thread per pattern, whole op list per 128-pattern tile (the walker schedule of the generic kernel); CLV addressing through ClvIdx,
// so every slot layout of the engine serves.  Parity-green first path, not a tuned one (the reversible walkers' register chaining and
// DMMA tiles carry over — the operator is a plain matrix here too — and are the next step).
#pragma once
#include "b200lk_kernels.cuh"

namespace b200lk {

constexpr int kNrKindMask = 3, kNrKindRoot = 2, kNrTransposed = 4;

struct NrModel {          // device pointers into the engine's non-reversible model block
    const double *ceval;  // [nmix][n][2]
    const double *cevec;  // [nmix][n][n][2]
    const double *cinv;   // [nmix][n][n][2]
    const double *qmat;   // [nmix][n][n]
    const double *sfreq;  // [nmix][n]
    const double *tip;    // [su+1][n] state indicators
    int rooted;
};

struct NrChildJob {
    double length;
    int64_t e_off;      // [C][n][n] operator block inside opinfo
    int64_t leaf_off;   // leaf table [su+1][B] (root leaf: one row), -1 for an internal child
    int32_t kind;       // 0 tip, 1 slot, 2 root leaf; | 4 transposed
    int32_t pad_;
};

// P_m(t)[i][j] for one (i, j): sum over k of Re(cevec[i][k] * e_k * cinv[k][j]); (er, ei) = exp(ceval_k t) staged by the caller
__device__ __forceinline__ double nr_pij(const double *V, const double *W, const double *er, const double *ei, int n, int i, int j) {
    double s = 0.0;
    for (int k = 0; k < n; k++) {
        const double vr = V[(size_t)(i * n + k) * 2], vi = V[(size_t)(i * n + k) * 2 + 1];
        const double ar = vr * er[k] - vi * ei[k], ai = vr * ei[k] + vi * er[k];
        s += ar * W[(size_t)(k * n + j) * 2] - ai * W[(size_t)(k * n + j) * 2 + 1];
    }
    return s;
}

// computePartialInfo, non-reversible branch: one CTA per (op, child)
__global__ void __launch_bounds__(256) k_nonrev_info(Dev D, NrModel M, const NrChildJob *jobs) {
    extern __shared__ double s_e[];   // [C][n][2]: exp(ceval_k * rate_c * len)
    const NrChildJob job = jobs[blockIdx.x];
    const int n = D.n, C = D.C, B = D.B, nn = n * n;
    const double len = resolve_len(D, job.length);
    for (int k = threadIdx.x; k < C * n; k += blockDim.x) {
        const int c = k / n, i = k - c * n, m = c / D.denom;
        const double t = D.cat_rate[c % D.ncat] * len;
        const double a = exp(M.ceval[(size_t)(m * n + i) * 2] * t);
        double sn, cs;
        sincos(M.ceval[(size_t)(m * n + i) * 2 + 1] * t, &sn, &cs);
        s_e[2 * k] = a * cs; s_e[2 * k + 1] = a * sn;
    }
    __syncthreads();
    double *E = D.opinfo + job.e_off;
    const bool tr = (job.kind & kNrTransposed) != 0;
    for (int k = threadIdx.x; k < C * nn; k += blockDim.x) {
        const int c = k / nn, rem = k - c * nn, i = rem / n, j = rem - i * n, m = c / D.denom;
        // de-interleave the staged exponentials on the fly
        double s = 0.0;
        const double *V = M.cevec + (size_t)m * nn * 2, *W = M.cinv + (size_t)m * nn * 2, *ex = s_e + (size_t)c * n * 2;
        for (int q = 0; q < n; q++) {
            const double vr = V[(size_t)(i * n + q) * 2], vi = V[(size_t)(i * n + q) * 2 + 1];
            const double ar = vr * ex[2 * q] - vi * ex[2 * q + 1], ai = vr * ex[2 * q + 1] + vi * ex[2 * q];
            s += ar * W[(size_t)(q * n + j) * 2] - ai * W[(size_t)(q * n + j) * 2 + 1];
        }
        E[(size_t)c * nn + (tr ? j * n + i : i * n + j)] = s;   // transposed for a root-ward branch (:945-949)
    }
    if (job.leaf_off < 0) return;
    __syncthreads();
    double *leaf = D.opinfo + job.leaf_off;
    if ((job.kind & kNrKindMask) == kNrKindRoot) {   // :959-964
        for (int k = threadIdx.x; k < B; k += blockDim.x) leaf[k] = M.sfreq[(size_t)((k / n) / D.denom) * n + k % n];
        return;
    }
    const int total = (D.su + 1) * B;
    for (int k = threadIdx.x; k < total; k += blockDim.x) {
        const int state = k / B, cx = k - state * B;
        if (state == D.su) { leaf[k] = 1.0; continue; }
        const double *ep = E + (size_t)cx * n, *tip = M.tip + (size_t)state * n;
        double v = 0.0;
        for (int i = 0; i < n; i++) v += ep[i] * tip[i];
        leaf[k] = v;
    }
}

// DevOp.ncase for this path: 0 tip-tip (either may be the root leaf), 1 tip-internal, 2 internal-internal, 4 root-internal (no
// underflow test); lid / rid = -1 for the root leaf (its one table row)
template <int NT, bool SAFE>
__device__ __forceinline__ void process_op_nonrev(const Dev &D, const DevOp &op, long long p) {
    constexpr int NMAX = NT ? NT : 64;
    const int n = NT ? NT : D.n;
    const int C = D.C, B = D.B;
    const ClvIdx at(D, p);
    const long long sb = SAFE ? pbase(p, C) : p;
    double *out = D.partial + (long long)op.dst * D.slot_stride;
    uint16_t *osc = D.scale + (long long)op.dst * D.scale_stride + sb;
    const bool l_int = op.ncase == 2, r_int = op.ncase >= 1;
    const double *Lp = l_int ? D.partial + (long long)op.lid * D.slot_stride : nullptr;
    const double *Rp = r_int ? D.partial + (long long)op.rid * D.slot_stride : nullptr;
    const uint16_t *lsc = l_int ? D.scale + (long long)op.lid * D.scale_stride + sb : nullptr;
    const uint16_t *rsc = r_int ? D.scale + (long long)op.rid * D.scale_stride + sb : nullptr;
    const double *leafL = nullptr, *leafR = nullptr;
    if (!l_int) leafL = D.opinfo + op.leafL + (op.lid < 0 ? 0 : (long long)D.tips[(long long)op.lid * D.Ppad + p] * B);
    if (!r_int) leafR = D.opinfo + op.leafR + (op.rid < 0 ? 0 : (long long)D.tips[(long long)op.rid * D.Ppad + p] * B);
    const double *eL = D.opinfo + op.eL, *eR = D.opinfo + op.eR;
    const bool check = (op.ncase == 1 || op.ncase == 2) && D.invar[p] == 0.0;
    double lh_max = 0.0;
    unsigned norm_scale = 0;
    if (!SAFE) norm_scale = (lsc ? lsc[0] : 0u) + (rsc ? rsc[0] : 0u);
    double lv[NMAX], rv[NMAX], ov[NMAX];
    for (int c = 0; c < C; c++) {
        if (Lp) {
#pragma unroll
            for (int i = 0; i < n; i++) lv[i] = Lp[at(c, i)];
        }
        if (Rp) {
#pragma unroll
            for (int i = 0; i < n; i++) rv[i] = Rp[at(c, i)];
        }
        double cmax = 0.0;
#pragma unroll
        for (int x = 0; x < n; x++) {
            const double a = Lp ? dot4acc<NT>(eL + (long long)(c * n + x) * n, lv, n) : leafL[c * n + x];
            const double b = Rp ? dot4acc<NT>(eR + (long long)(c * n + x) * n, rv, n) : leafR[c * n + x];
            const double v = a * b;
            ov[x] = v;
            cmax = fmax(cmax, v);   // signed maximum, from 0 (:446-447, 520)
        }
        if (SAFE) {
            unsigned sc = (lsc ? lsc[c * kWarp] : 0u) + (rsc ? rsc[c * kWarp] : 0u);
            if (check && cmax < kScalingThreshold) {
#pragma unroll
                for (int i = 0; i < n; i++) ov[i] *= kTwo256;
                sc += 1;
            }
            osc[c * kWarp] = (uint16_t)sc;
        } else lh_max = fmax(lh_max, cmax);
#pragma unroll
        for (int i = 0; i < n; i++) out[at(c, i)] = ov[i];
    }
    if (!SAFE) {
        if (check && lh_max < kScalingThreshold) {
            for (int c = 0; c < C; c++)
                for (int i = 0; i < n; i++) out[at(c, i)] *= kTwo256;
            norm_scale += 1;
        }
        osc[0] = (uint16_t)norm_scale;
    }
}

template <int NT, bool SAFE>
__global__ void __launch_bounds__(kTile) k_nonrev_partials(Dev D, const DevOp *ops, int nops, unsigned long long *counter) {
    __shared__ long long s_item;
    const long long ntiles = D.Ppad / kTile;
    for (;;) {   // (the host zeroes the counter before the launch)
        __syncthreads();
        if (threadIdx.x == 0) s_item = (long long)atomicAdd(counter, 1ULL);
        __syncthreads();
        const long long tile = s_item;
        if (tile >= ntiles) break;
        const long long p = tile * kTile + threadIdx.x;
        for (int o = 0; o < nops; o++) process_op_nonrev<NT, SAFE>(D, ops[o], p);
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_nonrev_walk4<C, SAFE>: the 4-state non-reversible partial-likelihood kernel with the walker's structure (cf. k_walk_n4c4): a thread
// keeps its pattern through the whole post-order op list and the CLV it has just produced stays in registers for its parent (the
// planner's order makes an internal child the previous op's result nearly always), so an op reads at most ONE child CLV from memory
// and tip-inner nodes none; the op's two 4 x 4 x C transition matrices (built, transposed where the branch points root-ward, by
// k_nonrev_info) are staged in shared memory one op ahead, one double per thread; tip states are loaded one op ahead.
// Arithmetic and summation order are process_op_nonrev's (dot4acc: two pair sums), the 2^-256 test is the signed maximum.
// ------------------------------------------------------------------------------------------------------------
template <int C, bool SAFE>
__global__ void __launch_bounds__(kTile, 4) k_nonrev_walk4(Dev D, const DevOp *ops, int nops, unsigned long long *counter) {
    constexpr int W = 16 * C, NS = SAFE ? C : 1;
    __shared__ double s_P[2][2][W];   // [buffer][left / right][c][x][y]
    __shared__ long long s_item;
    const int tid = threadIdx.x;
    const long long ntiles = D.Ppad / kTile;
    // an internal child's matrices [c][x][y]; for a leaf child the rows of its table for the first four state codes (the
    // nucleotides; row 0 of the root leaf's one-row table), transposed to [c][x][state] so that lanes with different states hit
    // different banks — the other codes (ambiguity, gap) are gathered from the table in global memory
    auto stage_mats = [&](const DevOp &op, int buf) {
        for (int t = tid; t < 2 * W; t += kTile) {
            const int ch = t / W, k = t - ch * W;
            const bool internal = ch == 0 ? op.ncase == 2 : op.ncase >= 1;
            if (internal) s_P[buf][ch][k] = D.opinfo[(ch ? op.eR : op.eL) + k];
            else {
                const int s = k / (4 * C), j = k - s * (4 * C);   // k = state * 4C + (c * 4 + x)
                const bool root = (ch ? op.rid : op.lid) < 0;      // one row only
                s_P[buf][ch][j * 4 + s] = (!root || s == 0) ? D.opinfo[(ch ? op.leafR : op.leafL) + k] : 0.0;
            }
        }
    };
    for (;;) {   // (the host zeroes the counter before the launch)
        __syncthreads();
        if (tid == 0) s_item = (long long)atomicAdd(counter, 1ULL);
        __syncthreads();
        const long long tile = s_item;
        if (tile >= ntiles) break;
        const long long p = tile * kTile + tid;
        const long long pb = pbase(p, 4 * C);
        const long long sb = SAFE ? pbase(p, C) : p;
        const bool may_scale = D.invar[p] == 0.0;
        double cur[4 * C];
        unsigned cursc[NS];
#pragma unroll
        for (int k = 0; k < 4 * C; k++) cur[k] = 0.0;
#pragma unroll
        for (int c = 0; c < NS; c++) cursc[c] = 0;
        int cur_slot = -1;   // the slot whose CLV is in `cur`
        DevOp op_next = ops[0];
        stage_mats(op_next, 0);
        auto tip_state = [&](const DevOp &op, int which) -> int {   // state of a leaf child (0 for the root leaf / an internal child)
            const int id = which ? op.rid : op.lid;
            const bool leaf = which ? op.ncase == 0 : op.ncase != 2;
            return leaf && id >= 0 ? (int)D.tips[(long long)id * D.Ppad + p] : 0;
        };
        int sL_next = tip_state(op_next, 0), sR_next = tip_state(op_next, 1);
        for (int o = 0; o < nops; o++) {
            __syncthreads();   // the matrices of op o are staged; everyone has left op o - 1 (its buffer is rewritten below)
            const DevOp op = op_next;
            const int sL = sL_next, sR = sR_next;
            if (o + 1 < nops) {
                op_next = ops[o + 1];
                stage_mats(op_next, (o + 1) & 1);
                sL_next = tip_state(op_next, 0); sR_next = tip_state(op_next, 1);
            }
            const bool l_int = op.ncase == 2, r_int = op.ncase >= 1;
            const bool l_reg = l_int && op.lid == cur_slot;          // left child chained
            const bool r_reg = r_int && !l_reg;                       // otherwise the right child lives in `cur` (chained, or loaded now)
            double X[4 * C];
            unsigned scX[NS];
#pragma unroll
            for (int c = 0; c < NS; c++) scX[c] = 0;
            if (r_reg && op.rid != cur_slot) {
                const double *Rp = D.partial + (long long)op.rid * D.slot_stride + pb;
#pragma unroll
                for (int k = 0; k < 4 * C; k++) cur[k] = Rp[k * kWarp];
#pragma unroll
                for (int c = 0; c < NS; c++) cursc[c] = D.scale[(long long)op.rid * D.scale_stride + sb + (SAFE ? c * kWarp : 0)];
            }
            if (!r_int) {
#pragma unroll
                for (int c = 0; c < NS; c++) cursc[c] = 0;
            }
            if (l_int) {   // both internal: the one that is not in registers comes from memory
                const int xid = l_reg ? op.rid : op.lid;
                const double *Xp = D.partial + (long long)xid * D.slot_stride + pb;
#pragma unroll
                for (int k = 0; k < 4 * C; k++) X[k] = Xp[k * kWarp];
#pragma unroll
                for (int c = 0; c < NS; c++) scX[c] = D.scale[(long long)xid * D.scale_stride + sb + (SAFE ? c * kWarp : 0)];
            }
            const double *leafL = l_int ? nullptr : D.opinfo + op.leafL + (op.lid < 0 ? 0 : (long long)sL * (4 * C));
            const double *leafR = r_int ? nullptr : D.opinfo + op.leafR + (op.rid < 0 ? 0 : (long long)sR * (4 * C));
            const double *PL = s_P[o & 1][0], *PR = s_P[o & 1][1];
            const bool check = (op.ncase == 1 || op.ncase == 2) && may_scale;
            double *Op = D.partial + (long long)op.dst * D.slot_stride + pb;
            double lh_max = 0.0;
#pragma unroll
            for (int c = 0; c < C; c++) {
                double lv[4], rv[4];
#pragma unroll
                for (int y = 0; y < 4; y++) {
                    lv[y] = l_reg ? cur[c * 4 + y] : X[c * 4 + y];      // (read only when l_int)
                    rv[y] = r_reg ? cur[c * 4 + y] : X[c * 4 + y];      // (read only when r_int)
                }
                double v[4], cmax = 0.0;
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    double a, b;
                    if (l_int) {
                        const double *m = PL + (c * 4 + x) * 4;
                        a = (m[0] * lv[0] + m[1] * lv[1]) + (m[2] * lv[2] + m[3] * lv[3]);
                    } else a = sL < 4 ? PL[(c * 4 + x) * 4 + sL] : leafL[c * 4 + x];
                    if (r_int) {
                        const double *m = PR + (c * 4 + x) * 4;
                        b = (m[0] * rv[0] + m[1] * rv[1]) + (m[2] * rv[2] + m[3] * rv[3]);
                    } else b = sR < 4 ? PR[(c * 4 + x) * 4 + sR] : leafR[c * 4 + x];
                    v[x] = a * b;
                    cmax = fmax(cmax, v[x]);   // signed maximum, from 0 (:446-447, 520)
                }
                if (SAFE) {
                    unsigned sc = cursc[SAFE ? c : 0] + scX[SAFE ? c : 0];
                    if (check && cmax < kScalingThreshold) {
#pragma unroll
                        for (int x = 0; x < 4; x++) v[x] *= kTwo256;
                        sc += 1;
                    }
                    cursc[SAFE ? c : 0] = sc;
                } else lh_max = fmax(lh_max, cmax);
#pragma unroll
                for (int x = 0; x < 4; x++) { cur[c * 4 + x] = v[x]; Op[(c * 4 + x) * kWarp] = v[x]; }
            }
            if (!SAFE) {
                cursc[0] += scX[0];
                if (check && lh_max < kScalingThreshold) {   // rare: rescale and store again
#pragma unroll
                    for (int k = 0; k < 4 * C; k++) { cur[k] *= kTwo256; Op[k * kWarp] = cur[k]; }
                    cursc[0] += 1;
                }
            }
#pragma unroll
            for (int c = 0; c < NS; c++) D.scale[(long long)op.dst * D.scale_stride + sb + (SAFE ? c * kWarp : 0)] = (uint16_t)cursc[c];
            cur_slot = op.dst;
        }
    }
}

// The evaluation branch's matrices, weighted as phylokernelnonrev.h:628-659 / 1103-1124: tm (and d1 = Q.P, d2 = Q.Q.P) [c][i][x],
// times prop (prop*rate, prop*rate^2), rows times the state frequencies on an unrooted tree; and, for a leaf at the far end, the
// tables T[state][c][i] = sum_x tm[c][i][x] tip[state][x] (:1160-1176, 693-724) — for the root leaf one row freq*prop (:1150-1158).
// One CTA per category.  bm: tm | d1 | d2 (B*n each) | T0 | T1 | T2 ((su+1)*B each).
__global__ void __launch_bounds__(256) k_nonrev_branch_info(Dev D, NrModel M, double length, int dad_kind, int with_derv, double *bm) {
    extern __shared__ double s_buf[];   // er[n] | ei[n] | P[n*n] | D1[n*n]
    const int n = D.n, B = D.B, nn = n * n, c = blockIdx.x, m = c / D.denom, r = c % D.ncat;
    double *er = s_buf, *ei = er + n, *P = ei + n, *D1 = P + nn;
    const double rate = D.cat_rate[r], t = rate * resolve_len(D, length), prop = D.cat_prop[r] * D.mix_weight[m];
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const double a = exp(M.ceval[(size_t)(m * n + k) * 2] * t);
        double sn, cs;
        sincos(M.ceval[(size_t)(m * n + k) * 2 + 1] * t, &sn, &cs);
        er[k] = a * cs; ei[k] = a * sn;
    }
    __syncthreads();
    const double *V = M.cevec + (size_t)m * nn * 2, *W = M.cinv + (size_t)m * nn * 2, *Q = M.qmat + (size_t)m * nn;
    for (int k = threadIdx.x; k < nn; k += blockDim.x) P[k] = nr_pij(V, W, er, ei, n, k / n, k % n);
    __syncthreads();
    const size_t tab = (size_t)(D.su + 1) * B;
    double *tm = bm + (size_t)c * nn, *d1 = bm + (size_t)B * n + (size_t)c * nn, *d2 = bm + (size_t)2 * B * n + (size_t)c * nn;
    if (with_derv) {
        for (int k = threadIdx.x; k < nn; k += blockDim.x) {
            const int i = k / n, j = k - i * n;
            double s = 0.0;
            for (int q = 0; q < n; q++) s += Q[i * n + q] * P[q * n + j];
            D1[k] = s;
        }
        __syncthreads();
    }
    const double prop_rate = prop * rate, prop_rate_2 = prop_rate * rate;
    for (int k = threadIdx.x; k < nn; k += blockDim.x) {
        const int i = k / n, j = k - i * n;
        const double f = M.rooted ? 1.0 : M.sfreq[(size_t)m * n + i];
        double v0 = P[k] * prop;
        if (!M.rooted) v0 *= f;
        tm[k] = v0;
        if (with_derv) {
            double s = 0.0;
            for (int q = 0; q < n; q++) s += Q[i * n + q] * D1[q * n + j];
            double v1 = D1[k] * prop_rate, v2 = s * prop_rate_2;
            if (!M.rooted) { v1 *= f; v2 *= f; }
            d1[k] = v1; d2[k] = v2;
        }
    }
    if (dad_kind == 1) return;
    __syncthreads();
    double *T0 = bm + (size_t)3 * B * n, *T1 = T0 + tab, *T2 = T1 + tab;
    if (dad_kind == kNrKindRoot) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            T0[c * n + i] = M.sfreq[(size_t)m * n + i] * prop;
            if (with_derv) { T1[c * n + i] = 0.0; T2[c * n + i] = 0.0; }   // (the reference leaves them uninitialised, :677-691; never used)
        }
        return;
    }
    for (int k = threadIdx.x; k < (D.su + 1) * n; k += blockDim.x) {
        const int state = k / n, i = k - state * n;
        const double *tip = M.tip + (size_t)state * n;
        double a = 0.0, b = 0.0, g = 0.0;
        for (int x = 0; x < n; x++) {
            a += tm[i * n + x] * tip[x];
            if (with_derv) { b += d1[i * n + x] * tip[x]; g += d2[i * n + x] * tip[x]; }
        }
        T0[(size_t)state * B + c * n + i] = a;
        if (with_derv) { T1[(size_t)state * B + c * n + i] = b; T2[(size_t)state * B + c * n + i] = g; }
    }
}

// computeNonrevLikelihoodBranch (:1062-1528): node_slot = dad_branch->partial_lh (the root-ward end), dad = leaf / root leaf / slot
template <bool SAFE>
__global__ void __launch_bounds__(kTile) k_nonrev_root(Dev D, int node_slot, int dad_kind, int dad_id, const double *bm, double *red, unsigned *ticket,
                                                       double *result) {
    const int n = D.n, C = D.C, B = D.B;
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    const long long cb = pbase(p, C);
    const long long sb = SAFE ? cb : p;
    const ClvIdx at(D, p);
    const double *dad = D.partial + (long long)node_slot * D.slot_stride;
    const uint16_t *dsc = D.scale + (long long)node_slot * D.scale_stride + sb;
    const double *node = nullptr, *tab = nullptr;
    const uint16_t *nsc = nullptr;
    if (dad_kind == 1) {
        node = D.partial + (long long)dad_id * D.slot_stride;
        nsc = D.scale + (long long)dad_id * D.scale_stride + sb;
    } else {
        const long long st = dad_kind == kNrKindRoot ? 0 : D.tips[(long long)dad_id * D.Ppad + p];
        tab = bm + (size_t)3 * B * n + st * B;
    }
    double *lh_cat = D.pattern_lh_cat + cb;
    double lh_ptn = 0.0;
    unsigned min_scale = 0xffffffffu;
    double nv[64];
    for (int c = 0; c < C; c++) {
        double s = 0.0;
        if (node) {
            for (int x = 0; x < n; x++) nv[x] = node[at(c, x)];
            const double *tm = bm + (size_t)c * n * n;
            for (int i = 0; i < n; i++) {
                double a0 = 0.0, a1 = 0.0;
                int x = 0;
                for (; x + 1 < n; x += 2) { a0 = fma(__ldg(tm + i * n + x), nv[x], a0); a1 = fma(__ldg(tm + i * n + x + 1), nv[x + 1], a1); }
                if (x < n) a0 = fma(__ldg(tm + i * n + x), nv[x], a0);
                s = fma(dad[at(c, i)], a0 + a1, s);
            }
        } else {
            for (int i = 0; i < n; i++) s = fma(__ldg(tab + c * n + i), dad[at(c, i)], s);
        }
        lh_cat[c * kWarp] = s;
        if (SAFE) min_scale = min(min_scale, (unsigned)dsc[c * kWarp] + (nsc ? nsc[c * kWarp] : 0u));
        else lh_ptn += s;
    }
    double vc_min_scale;
    if (SAFE) {
        for (int c = 0; c < C; c++) {
            const unsigned sc = dsc[c * kWarp] + (nsc ? nsc[c * kWarp] : 0u);
            double x = lh_cat[c * kWarp];
            if (sc == min_scale + 1) x *= kScalingThreshold;
            else if (sc > min_scale + 1) x = 0.0;
            lh_cat[c * kWarp] = x;
            lh_ptn += x;
        }
        vc_min_scale = (double)min_scale;
    } else vc_min_scale = (double)(dsc[0] + (nsc ? nsc[0] : 0u));
    vc_min_scale *= kLogScalingThreshold;
    lh_ptn = lh_ptn + D.invar[p];
    const double lg = log(lh_ptn) + vc_min_scale;
    double acc[1] = {0.0};
    if (p < D.P) {
        D.pattern_lh[p] = lg;
        acc[0] = lg * D.freq[p];
    }
    block_reduce_finish<1>(acc, red, ticket, result);
}

// computeNonrevLikelihoodDerv (:586-1056)
template <bool SAFE>
__global__ void __launch_bounds__(kTile) k_nonrev_derv(Dev D, int node_slot, int dad_kind, int dad_id, const double *bm, double *red, unsigned *ticket,
                                                       double *result) {
    const int n = D.n, C = D.C, B = D.B;
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    const long long sb = SAFE ? pbase(p, C) : p;
    const ClvIdx at(D, p);
    const double *dad = D.partial + (long long)node_slot * D.slot_stride;
    const uint16_t *dsc = D.scale + (long long)node_slot * D.scale_stride + sb;
    const double *node = nullptr, *T0 = nullptr;
    const uint16_t *nsc = nullptr;
    const size_t tab = (size_t)(D.su + 1) * B;
    if (dad_kind == 1) {
        node = D.partial + (long long)dad_id * D.slot_stride;
        nsc = D.scale + (long long)dad_id * D.scale_stride + sb;
    } else {
        const long long st = dad_kind == kNrKindRoot ? 0 : D.tips[(long long)dad_id * D.Ppad + p];
        T0 = bm + (size_t)3 * B * n + st * B;
    }
    // SAFE_LH: each end's categories are aligned to that end's minimum scale count (:706-722, 880-917)
    unsigned dmin = 0xffffffffu, nmin = 0xffffffffu;
    if (SAFE) {
        for (int c = 0; c < C; c++) { dmin = min(dmin, (unsigned)dsc[c * kWarp]); if (nsc) nmin = min(nmin, (unsigned)nsc[c * kWarp]); }
    }
    double lh = 0.0, df = 0.0, ddf = 0.0;
    double nv[64];
    for (int c = 0; c < C; c++) {
        double fd = 1.0, fn = 1.0;
        if (SAFE) {
            const unsigned sd = dsc[c * kWarp];
            fd = sd == dmin ? 1.0 : (sd == dmin + 1 ? kScalingThreshold : 0.0);
            if (nsc) { const unsigned sn = nsc[c * kWarp]; fn = sn == nmin ? 1.0 : (sn == nmin + 1 ? kScalingThreshold : 0.0); }
        }
        if (node) {
            for (int x = 0; x < n; x++) { const double v = node[at(c, x)]; nv[x] = SAFE ? (fn == 0.0 ? 0.0 : v * fn) : v; }
            const double *tm = bm + (size_t)c * n * n, *d1 = tm + (size_t)B * n, *d2 = d1 + (size_t)B * n;
            for (int i = 0; i < n; i++) {
                double a = 0.0, b = 0.0, g = 0.0;
                for (int x = 0; x < n; x++) {
                    a = fma(__ldg(tm + i * n + x), nv[x], a);
                    b = fma(__ldg(d1 + i * n + x), nv[x], b);
                    g = fma(__ldg(d2 + i * n + x), nv[x], g);
                }
                double dv = dad[at(c, i)];
                if (SAFE) dv = fd == 0.0 ? 0.0 : dv * fd;
                lh = fma(dv, a, lh); df = fma(dv, b, df); ddf = fma(dv, g, ddf);
            }
        } else {
            for (int i = 0; i < n; i++) {
                double dv = dad[at(c, i)];
                if (SAFE) dv = fd == 0.0 ? 0.0 : dv * fd;
                lh = fma(__ldg(T0 + c * n + i), dv, lh);
                df = fma(__ldg(T0 + tab + c * n + i), dv, df);
                ddf = fma(__ldg(T0 + 2 * tab + c * n + i), dv, ddf);
            }
        }
    }
    double acc[2] = {0.0, 0.0};
    if (p < D.P) {
        lh = 1.0 / (lh + D.invar[p]);
        const double df_frac = df * lh, ddf_frac = ddf * lh, f = D.freq[p];
        const double tmp1 = df_frac * f, tmp2 = ddf_frac * f;
        acc[0] = tmp1;
        acc[1] = fma(-tmp1, df_frac, tmp2);
    }
    block_reduce_finish<2>(acc, red, ticket, result);
}

}  // namespace b200lk
