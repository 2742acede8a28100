// b200lk_walk4.cuh — the 4-state pattern-tile walker for ANY number of rate categories (JC/HKY/GTR without +G, +G8, +R2..+R10,
// +I+R ...): the reference's 4-state specialisation serves every ncat (tree/phylokernelfma.cpp:17-20, KERNEL_FIX_STATES), and most of
// ModelFinder's DNA candidates (main/phylotesting.cpp:1572-1700) do not have exactly four categories.
//
// Same design as k_walk_n4c4 (b200lk_kernels.cuh), templated on the category count C:
//   * a thread keeps ONE pattern for the whole post-order op list; the CLV it has just produced (4C doubles in registers) is the
//     first operand of the next op whenever that op is its parent (register chaining), so at most one child CLV is loaded per op;
//   * the operator is applied factorised, evec . (exp(eval r_c t) (.) L), with evec / inv_evec as kernel-parameter (constant-bank)
//     operands of the FP64 FMAs; per op only 2 x 4C exponentials and the [c][x][state] tip tables are staged in shared memory;
//   * a category's four values are stored as soon as they are final; the sibling CLV is loaded category by category, one category
//     ahead of its use, so the register file holds 4C + 8 doubles of CLV data whatever C is;
//   * the 2^-256 test is an integer compare on the exponent field.
#pragma once
#include "b200lk_kernels.cuh"

namespace b200lk {

// staging block per op (doubles): header[4] | expA[4C] | expB[4C] | tabA[16C] | tabB[16C]
__host__ __device__ constexpr int stage4_doubles(int C) { return 4 + 40 * C; }

// one CTA per op; child A is the CLV that lives in registers, child B the other one (WalkJob4 as for k_walk_info_n4c4)
__device__ __forceinline__ void walk_info4_body(const Dev &D, const WalkJob4 &job, double *st, double *s_exp4) {
    const int C = D.C;
    for (int t = threadIdx.x; t < 8 * C; t += blockDim.x) {
        const int ch = t / (4 * C), k = t - ch * 4 * C, c = k >> 2, i = k & 3;
        const double len_child = D.cat_rate[c % D.ncat] * resolve_len_cat(D, ch ? job.lenB : job.lenA, c % D.ncat);   // :1017
        const double v = exp(D.eval[i] * len_child);                                                    // :1022
        s_exp4[t] = v;
        st[4 + t] = v;
    }
    if (threadIdx.x == 0) {
        StageHdr4 h{job.dst, job.ncase, job.aid, job.bid, job.loadA, 0, 0, 0};
        *reinterpret_cast<StageHdr4 *>(st) = h;
    }
    __syncthreads();
    // partial_lh_leaves for the 4 unambiguous states, stored [c][x][state] (:1036-1051, Vec4d lanes + horizontal_add)
    for (int t = threadIdx.x; t < 32 * C; t += blockDim.x) {
        const int ch = t / (16 * C), k = t - ch * 16 * C, c = k >> 4, x = (k >> 2) & 3, s = k & 3;
        const bool is_tip = ch == 0 ? (job.ncase == 0) : (job.ncase <= 1);
        double v = 0.0;
        if (is_tip) {
            const double *tip = D.tip_lh + s * 4, *e = s_exp4 + ch * 4 * C + c * 4;
            const double p0 = __dmul_rn(__dmul_rn(D.evec[x * 4 + 0], e[0]), tip[0]);
            const double p1 = __dmul_rn(__dmul_rn(D.evec[x * 4 + 1], e[1]), tip[1]);
            const double p2 = __dmul_rn(__dmul_rn(D.evec[x * 4 + 2], e[2]), tip[2]);
            const double p3 = __dmul_rn(__dmul_rn(D.evec[x * 4 + 3], e[3]), tip[3]);
            v = __dadd_rn(__dadd_rn(p0, p1), __dadd_rn(p2, p3));
        }
        st[4 + 8 * C + t] = v;
    }
}
__global__ void __launch_bounds__(128) k_walk_info_n4(Dev D, const WalkJob4 *jobs, double *stage, unsigned long long *counter) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0ULL;
    extern __shared__ double s_exp4[];   // [2][4C]
    walk_info4_body(D, jobs[blockIdx.x], stage + (size_t)blockIdx.x * stage4_doubles(D.C), s_exp4);
}

template <int C>
__host__ __device__ constexpr int walk4_min_blocks() { return C <= 4 ? 4 : C <= 8 ? 3 : 2; }

// One tile (kTile patterns) of one engine: the whole post-order op list.  D and M may live in the constant bank (the single-engine
// kernel's parameters) or in shared memory / registers (a member of a grouped launch).
template <int C, bool SAFE>
__device__ __forceinline__ void walk4_tile(const Dev &D, const Model4 &M, const double *stage, const int nops, const long long tile,
                                           double (*s_stage)[stage4_doubles(C)], const double *s_tip, unsigned char (*s_ts)[2][kTile]) {
    constexpr int KS = stage4_doubles(C), KSV = KS / 2;   // doubles / 16-byte chunks per staging block
    constexpr int NS = SAFE ? C : 1;
    const int tid = threadIdx.x;
    auto stage_copy = [&](double *dst, const double *src) {   // one staging block, asynchronously
        for (int k = tid; k < KSV; k += kTile) cp_async16(dst + 2 * k, src + 2 * k);
    };
    // one child value from a tip state: the table for the 4 unambiguous states, 1 for a gap, the direct sum for IUPAC codes
    auto tip_value = [&](const double *tab, const double *ex, int s, int c, int x) -> double {
        if (s < 4) return tab[c * 16 + x * 4 + s];
        if (s == D.su) return 1.0;
        const double q0 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 0], ex[c * 4 + 0]), s_tip[s * 4 + 0]);
        const double q1 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 1], ex[c * 4 + 1]), s_tip[s * 4 + 1]);
        const double q2 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 2], ex[c * 4 + 2]), s_tip[s * 4 + 2]);
        const double q3 = __dmul_rn(__dmul_rn(M.evec[x * 4 + 3], ex[c * 4 + 3]), s_tip[s * 4 + 3]);
        return __dadd_rn(__dadd_rn(q0, q1), __dadd_rn(q2, q3));
    };
    const long long p = tile * kTile + tid;
    const long long pb = pbase(p, 4 * C);
    const long long sb = SAFE ? pbase(p, C) : p;
    const bool may_scale = D.invar[p] == 0.0;
    stage_copy(s_stage[0], stage);
    if (nops > 1) stage_copy(s_stage[1], stage + KS);
    double cur[4 * C];
    unsigned cursc[NS];
#pragma unroll
    for (int k = 0; k < 4 * C; k++) cur[k] = 0.0;
#pragma unroll
    for (int c = 0; c < NS; c++) cursc[c] = 0;
    // the next op's tip states travel global -> shared with cp.async, lanes 0..7 of a warp 4 bytes each: loaded into registers one op
    // ahead, the load shared a scoreboard with the op's first shared-memory consumer and stalled it (see k_walk_n4c4)
    const int ts_off = (tid >> 5) * 32 + (tid & 7) * 4;
    {
        const int4 h = *reinterpret_cast<const int4 *>(stage);   // global: the shared copy has not landed yet
        if ((tid & 31) < 8) {
            if (h.y <= 1) cp_async4(&s_ts[0][1][ts_off], D.tips + (long long)h.w * D.Ppad + tile * kTile + ts_off);
            if (h.y == 0) cp_async4(&s_ts[0][0][ts_off], D.tips + (long long)h.z * D.Ppad + tile * kTile + ts_off);
        }
    }
    int buf = 0;
    for (int o = 0; o < nops; o++) {
        cp_async_wait_all();
        __syncthreads();   // block o has landed for everyone; everyone has left op o-1 (its ring slot is reused below)
        const double *st = s_stage[buf];
        const int buf1 = buf == 2 ? 0 : buf + 1, buf2 = buf1 == 2 ? 0 : buf1 + 1;
        const int4 h0 = *reinterpret_cast<const int4 *>(st);
        const int h_loadA = *reinterpret_cast<const int *>(st + 2);
        const int dst = h0.x, ncase = h0.y, aid = h0.z, bid = h0.w;
        if (o + 2 < nops) stage_copy(s_stage[buf2], stage + (size_t)(o + 2) * KS);
        const int sA = ncase == 0 ? (int)s_ts[o & 1][0][tid] : 0, sB = ncase <= 1 ? (int)s_ts[o & 1][1][tid] : 0;
        const double *Bp = D.partial + (long long)bid * D.slot_stride + pb;
        const uint16_t *Bs = D.scale + (long long)bid * D.scale_stride + sb;
        unsigned scB[NS];
#pragma unroll
        for (int c = 0; c < NS; c++) scB[c] = ncase == 2 ? (unsigned)Bs[SAFE ? c * kWarp : 0] : 0u;
        if (ncase == 0) {
#pragma unroll
            for (int c = 0; c < NS; c++) cursc[c] = 0;
        } else if (h_loadA) {
            const double *Ap = D.partial + (long long)aid * D.slot_stride + pb;
#pragma unroll
            for (int k = 0; k < 4 * C; k++) cur[k] = Ap[k * kWarp];
#pragma unroll
            for (int c = 0; c < NS; c++) cursc[c] = D.scale[(long long)aid * D.scale_stride + sb + (SAFE ? c * kWarp : 0)];
        }
        // the sibling CLV travels one category ahead of its use
        double X[4], Xn[4];
        if (ncase == 2) {
#pragma unroll
            for (int i = 0; i < 4; i++) Xn[i] = Bp[i * kWarp];
        }
        if (o + 1 < nops) {   // steer the next op: tip states to registers, CLV operands towards L1
            const double *sn = s_stage[buf1];
            const int4 hn = *reinterpret_cast<const int4 *>(sn);
            const int n_loadA = *reinterpret_cast<const int *>(sn + 2);
            if ((tid & 31) < 8) {
                if (hn.y <= 1) cp_async4(&s_ts[(o + 1) & 1][1][ts_off], D.tips + (long long)hn.w * D.Ppad + tile * kTile + ts_off);
                if (hn.y == 0) cp_async4(&s_ts[(o + 1) & 1][0][ts_off], D.tips + (long long)hn.z * D.Ppad + tile * kTile + ts_off);
            }
            // (no L1 prefetch of the next op's CLV operands: the loads missed L1 regardless and the prefetch instructions alone cost
            // 4.6 % of k_walk_n4c4's time, gpurun_out/r2_25 variants 0 / 27)
            (void)n_loadA;
        }
        const double *expA = st + 4, *expB = st + 4 + 4 * C, *tabA = st + 4 + 8 * C, *tabB = st + 4 + 24 * C;
        double *Op = D.partial + (long long)dst * D.slot_stride + pb;
        unsigned mhi = 0;
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (ncase == 2) {
#pragma unroll
                for (int i = 0; i < 4; i++) X[i] = Xn[i];
                if (c + 1 < C) {
#pragma unroll
                    for (int i = 0; i < 4; i++) Xn[i] = Bp[((c + 1) * 4 + i) * kWarp];
                }
            }
            double fa[4], fb[4];
            if (ncase >= 1) {
                const double2 e01 = *reinterpret_cast<const double2 *>(expA + c * 4), e23 = *reinterpret_cast<const double2 *>(expA + c * 4 + 2);
                const double a0 = e01.x * cur[c * 4], a1 = e01.y * cur[c * 4 + 1], a2 = e23.x * cur[c * 4 + 2], a3 = e23.y * cur[c * 4 + 3];
#pragma unroll
                for (int x = 0; x < 4; x++)
                    fa[x] = fma(M.evec[x * 4 + 3], a3, fma(M.evec[x * 4 + 2], a2, fma(M.evec[x * 4 + 1], a1, M.evec[x * 4] * a0)));
            } else {
#pragma unroll
                for (int x = 0; x < 4; x++) fa[x] = tip_value(tabA, expA, sA, c, x);
            }
            if (ncase == 2) {
                const double2 e01 = *reinterpret_cast<const double2 *>(expB + c * 4), e23 = *reinterpret_cast<const double2 *>(expB + c * 4 + 2);
                const double b0 = e01.x * X[0], b1 = e01.y * X[1], b2 = e23.x * X[2], b3 = e23.y * X[3];
#pragma unroll
                for (int x = 0; x < 4; x++)
                    fb[x] = fma(M.evec[x * 4 + 3], b3, fma(M.evec[x * 4 + 2], b2, fma(M.evec[x * 4 + 1], b1, M.evec[x * 4] * b0)));
            } else {
#pragma unroll
                for (int x = 0; x < 4; x++) fb[x] = tip_value(tabB, expB, sB, c, x);
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
                cursc[SAFE ? c : 0] += scB[SAFE ? c : 0];
                if (ncase != 0 && may_scale && chi < kThresholdHi) {
#pragma unroll
                    for (int i = 0; i < 4; i++) cur[c * 4 + i] *= kTwo256;
                    cursc[SAFE ? c : 0] += 1;
                }
            } else mhi = max(mhi, chi);
#pragma unroll
            for (int i = 0; i < 4; i++) Op[(c * 4 + i) * kWarp] = cur[c * 4 + i];
        }
        if (!SAFE) {
            cursc[0] += scB[0];
            if (ncase != 0 && may_scale && mhi < kThresholdHi) {   // rare: every value underflowed; rescale and store again
#pragma unroll
                for (int k = 0; k < 4 * C; k++) { cur[k] *= kTwo256; Op[k * kWarp] = cur[k]; }
                cursc[0] += 1;
            }
        }
#pragma unroll
        for (int c = 0; c < NS; c++) D.scale[(long long)dst * D.scale_stride + sb + (SAFE ? c * kWarp : 0)] = (uint16_t)cursc[c];
        buf = buf1;
    }
}

template <int C, bool SAFE>
__global__ void __launch_bounds__(kTile, walk4_min_blocks<C>()) k_walk_n4(Dev D, const __grid_constant__ Model4 M, const double *stage, int nops,
                                                                         unsigned long long *counter) {
    __shared__ __align__(16) double s_stage[3][stage4_doubles(C)];
    __shared__ double s_tip[kMaxLeafRows4 * 4];
    __shared__ __align__(16) unsigned char s_ts[2][2][kTile];
    __shared__ long long s_item;
    const int tid = threadIdx.x;
    for (int k = tid; k < (D.su + 1) * 4; k += kTile) s_tip[k] = D.tip_lh[k];
    const long long ntiles = D.Ppad / kTile;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (long long)atomicAdd(counter, 1ULL);
        __syncthreads();
        const long long tile = s_item;
        if (tile >= ntiles) break;
        walk4_tile<C, SAFE>(D, M, stage, nops, tile, s_stage, s_tip, s_ts);
    }
}

}  // namespace b200lk
