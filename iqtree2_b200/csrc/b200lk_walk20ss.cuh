// b200lk_walk20ss.cuh — the 20-state FP64-MMA walker on STATE-SPACE conditional likelihood vectors.
//
// The reference keeps partial_lh in eigen coordinates, W = inv_evec . t with t[x] = (P_A t_A)[x] * (P_B t_B)[x], and pays THREE
// 20x20 mat-vecs per inner-inner node: evec.(exp (.) W_A), evec.(exp (.) W_B) and the back transform inv_evec . t
// (tree/phylokernelnew.h:1907-2032).  This kernel keeps t itself in the CLV slots and folds the back transform of a child
// into the child pass of its parent:
//     (evec . diag(exp(eval r_c len)) . inv_evec) . t_child = P_c(len) . t_child          — ONE mat-vec per internal child,
// with the per-(branch, category) transition matrices P_c(len) built on the device by the info kernel (20^3 FMAs each) and
// staged through shared memory in B-fragment order.  DMMAs per (8 patterns x 1 category): inner-inner 45 -> 30, tip-inner
// 30 -> 15, tip-tip 15 -> 0 (a product of two table rows): half the FP64 work of BASELINE configs 3 and 4.
//
// What the reference's callers see is unchanged:
//   * partial_lh read-backs, the root/theta kernels and b200lk_get_partial go through k_clv20_to_eigen (W = inv_evec . t, one
//     15-DMMA pass into a scratch slot; 2 of ~200 CLVs per evaluation);
//   * scale_num stays BIT-EXACT although the reference's underflow test is on max|W| < 2^-256, not on max|t|:
//         max|t| / ||evec||_inf  <=  max|W|  <=  ||inv_evec||_inf . max|t|
//     decides every pattern whose max|t| lies outside a band of ~2^6..2^12 around the threshold; patterns inside the band
//     (a few per cent of the nodes at which a pattern rescales at all) take the exact path: W of that 8-pattern tile is
//     computed from the just-stored t with the same 15 DMMAs the back pass used to cost everywhere.
//
// Fragment conventions are those of the DMMA helpers in b200lk_kernels.cuh: M = 8 patterns (row g = lane/4), K = 4 (slot q = lane%4),
// N = 8 states; CLVs fragment-major (clv20_off), loaded back as A fragments in the K order clv_k(kt, q).
#pragma once
#include "b200lk_kernels.cuh"

namespace b200lk {

// ---- staging block of one (op, class block): header[4] | part A | part B ; a part is the tip table [(su+1)][TS] of a tip child
// or the cpb transition matrices of an internal child, 400 doubles each in B-fragment order:
//     [kt 0..4][lane 0..31][nt 0..1]  (320)  |  [kt 0..4][lane 0..15] for nt = 2 (states 16..19 live in lanes g < 4)  (80)
// element (kt, nt, lane = 4g+q) = P[x = 8nt+g][y = clv_k(kt, q)].
constexpr int kMat20 = 400;
__host__ __device__ inline size_t ss_tab_doubles(int cpb, int su) { return (size_t)(su + 1) * tab20_stride(cpb); }
__host__ __device__ inline size_t ss_part_doubles(int cpb, int su, bool tip) { return tip ? ss_tab_doubles(cpb, su) : (size_t)cpb * kMat20; }
__host__ __device__ inline size_t ss_stage_doubles(int cpb, int su) {
    const size_t a = ss_tab_doubles(cpb, su), b = (size_t)cpb * kMat20;
    return 4 + 2 * (a > b ? a : b);
}
// doubles of the block an op of this node case needs (a contiguous prefix)
__host__ __device__ inline size_t ss_stage_used(int cpb, int su, int ncase) {
    return 4 + ss_part_doubles(cpb, su, ncase <= 1) + ss_part_doubles(cpb, su, ncase == 0);
}

// One CTA per (op, class block): exp(eval r_c len), the tip tables (the reference's partial_lh_leaves)
// (tree/phylokernelnew.h:1036-1056), and P_c(len) = evec . diag(exp) . inv_evec for internal children.
__device__ __forceinline__ void walk_info20ss_body(const Dev &D, const WalkJob20 &job, double *st, const int m, double *s_exp20) {
    const int nblk = D.nmix, cpb = D.C / nblk, B = cpb * 20, TS = tab20_stride(cpb);
    for (int k = threadIdx.x; k < 2 * B; k += blockDim.x) {
        const int ch = k / B, kk = k - ch * B, r = kk / 20, i = kk - r * 20;
        const int c = m * cpb + r;
        const double len_child = D.cat_rate[c % D.ncat] * resolve_len_cat(D, ch ? job.lenB : job.lenA, c % D.ncat);   // :1017
        s_exp20[k] = exp(D.eval[m * 20 + i] * len_child);                                              // :1022
    }
    if (threadIdx.x == 0) {
        int32_t *h = reinterpret_cast<int32_t *>(st);
        h[0] = job.dst; h[1] = job.ncase; h[2] = job.aid; h[3] = job.bid; h[4] = m; h[5] = h[6] = h[7] = 0;
    }
    // the class's eigenvectors and tip vectors in shared memory: every output element is a 20-term sum over them, and with the operands
    // in global memory the kernel was bound by L1 latency (0.77 of config 5's 9.3 ms per evaluation: 10 000 CTAs of this kernel)
    constexpr int kTipRows = 24;
    __shared__ double s_ev[400], s_iv[400], s_tl[kTipRows * 20];
    const int rows = D.su + 1;
    const bool tip_sm = rows <= kTipRows;
    for (int k = threadIdx.x; k < 400; k += blockDim.x) { s_ev[k] = D.evec[(size_t)m * 400 + k]; s_iv[k] = D.inv_evec[(size_t)m * 400 + k]; }
    if (tip_sm)
        for (int k = threadIdx.x; k < rows * 20; k += blockDim.x) s_tl[k] = D.tip_lh[((size_t)(k / 20) * D.nmix + m) * 20 + k % 20];
    __syncthreads();
    const double *ev = s_ev, *iv = s_iv;
    double *part = st + 4;
    for (int ch = 0; ch < 2; ch++) {
        const bool is_tip = ch == 0 ? (job.ncase <= 1) : (job.ncase == 0);
        const double *ex = s_exp20 + ch * B;
        if (is_tip) {
            for (int k = threadIdx.x; k < rows * TS; k += blockDim.x) {
                const int state = k / TS, cx = k - state * TS;
                if (cx >= B) { part[k] = 0.0; continue; }
                const int r = cx / 20, x = cx - r * 20;
                if (state == D.su) { part[k] = 1.0; continue; }
                const double *tip = tip_sm ? s_tl + state * 20 : D.tip_lh + ((size_t)state * D.nmix + m) * 20;
                const double *evx = ev + x * 20, *e = ex + r * 20;
                double a0 = (evx[0] * e[0]) * tip[0], a1 = (evx[1] * e[1]) * tip[1], a2 = (evx[2] * e[2]) * tip[2], a3 = (evx[3] * e[3]) * tip[3];
#pragma unroll
                for (int i = 4; i < 20; i += 4) {
                    a0 = fma(evx[i] * e[i], tip[i], a0); a1 = fma(evx[i + 1] * e[i + 1], tip[i + 1], a1);
                    a2 = fma(evx[i + 2] * e[i + 2], tip[i + 2], a2); a3 = fma(evx[i + 3] * e[i + 3], tip[i + 3], a3);
                }
                part[k] = (a0 + a1) + (a2 + a3);
            }
            part += rows * TS;
        } else {
            for (int k = threadIdx.x; k < cpb * kMat20; k += blockDim.x) {
                const int r = k / kMat20, j = k - r * kMat20;
                int kt, nt, ln;
                if (j < 320) { kt = j >> 6; ln = (j & 63) >> 1; nt = j & 1; }
                else { kt = (j - 320) >> 4; ln = (j - 320) & 15; nt = 2; }
                const int x = 8 * nt + (ln >> 2), y = clv_k(kt, ln & 3);
                const double *evx = ev + x * 20, *e = ex + r * 20;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll
                for (int i = 0; i < 20; i += 2) {
                    a0 = fma(evx[i] * e[i], iv[i * 20 + y], a0);
                    a1 = fma(evx[i + 1] * e[i + 1], iv[(i + 1) * 20 + y], a1);
                }
                part[k] = a0 + a1;
            }
            part += cpb * kMat20;
        }
    }
}
__global__ void __launch_bounds__(256) k_walk_info_n20ss(Dev D, const WalkJob20 *jobs, double *stage, unsigned long long *counter) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0ULL;
    extern __shared__ double s_exp20[];   // [2][cpb*20]
    const int nblk = D.nmix, cpb = D.C / nblk;
    const int op = blockIdx.x / nblk, m = blockIdx.x - op * nblk;
    walk_info20ss_body(D, jobs[op], stage + (size_t)blockIdx.x * ss_stage_doubles(cpb, D.su), m, s_exp20);
}

// inv_evec as B fragments of the transform W = inv_evec . t, per class: [m][(kt*3 + nt)][lane] = inv_evec_m[i = 8nt+g][x = clv_k(kt,q)]
__global__ void k_inv_frag20(Dev D, double *frag) {
    const int m = blockIdx.x;
    for (int k = threadIdx.x; k < 480; k += blockDim.x) {
        const int ln = k & 31, f = k >> 5, kt = f / 3, nt = f - 3 * kt, i = 8 * nt + (ln >> 2);
        frag[(size_t)m * 480 + k] = i < 20 ? D.inv_evec[(size_t)m * 400 + i * 20 + clv_k(kt, ln & 3)] : 0.0;
    }
}

// max over i of |(inv_evec . t)[p][i]| as exponent word, for the 8 patterns of one (M-tile, category) block of 160 doubles that
// THIS warp has just written (or wrote in an earlier class block); every lane returns the value of its pattern g.
__device__ __forceinline__ unsigned eigen_max20(const double *blk, const double *frag, int lane) {
    const int g = lane >> 2, q = lane & 3, lo2 = lane * 2;
    const double2 v0 = *reinterpret_cast<const double2 *>(blk + lo2), v1 = *reinterpret_cast<const double2 *>(blk + 64 + lo2);
    const double f[5] = {v0.x, v0.y, v1.x, v1.y, blk[128 + lane]};
    double o[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int kt = 0; kt < 5; kt++)
#pragma unroll
        for (int nt = 0; nt < 3; nt++) dmma_m8n8k4(o[nt][0], o[nt][1], f[kt], __ldg(frag + (kt * 3 + nt) * kWarp + lane));
    unsigned chi = 0u;
#pragma unroll
    for (int nt = 0; nt < 3; nt++) chi = max(chi, max(abs_hi(o[nt][0]), abs_hi(o[nt][1])));
    chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 1));
    chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 2));
    return chi;
}

// multiply what this lane stored of one (M-tile, category) block by 2^256
__device__ __forceinline__ void rescale20(double *blk, int lo2, int lot, int q) {
    double2 *p0 = reinterpret_cast<double2 *>(blk + lo2), *p1 = reinterpret_cast<double2 *>(blk + 64 + lo2);
    double2 v0 = *p0, v1 = *p1;
    v0.x *= kTwo256; v0.y *= kTwo256; v1.x *= kTwo256; v1.y *= kTwo256;
    *p0 = v0; *p1 = v1;
    if (q < 2) {
        double2 *p2 = reinterpret_cast<double2 *>(blk + lot);
        double2 v2 = *p2;
        v2.x *= kTwo256; v2.y *= kTwo256;
        *p2 = v2;
    }
}

struct Walk20ssOp {
    const double *Aq, *Bq;   // children: category 0 of the whole CLV in M-tile 0 of the warp's 32-pattern block
    double *Oq;              // destination, same position
    const uint16_t *Asc, *Bsc;
    uint16_t *Osc;           // scale_num: NORM [p], SAFE [c*32 + p%32] inside the warp's block (column g folded in)
    unsigned stA, stB;       // tip states of the lane's pattern in the 4 M-tiles, one byte each
    unsigned may_mask;       // bit mt: ptn_invar == 0
    int first_class, last_class;
    int Ctot, cpb;           // categories of the whole CLV / per class block
    unsigned *s_mhi;         // MIX: per-lane running maximum across the class blocks of one op, [4][threads]
    const double *inv_frag;  // D.inv_frag20
};

// NORM_LH decision for M-tile mt given the pattern's max|t| exponent word over ALL categories (every lane of the pattern holds
// it): tree/phylokernelnew.h:2015-2029.  Two steps, because with a category split the warps of a block must all have taken the
// decision (the exact path READS every category) before any of them rescales what it wrote.
__device__ __forceinline__ bool norm_test20(const Dev &D, const Walk20ssOp &W, int mt, unsigned mhi, int lane) {
    const bool may = (W.may_mask >> mt) & 1u;
    bool lo = may && mhi < D.ss_lo;
    const bool band = may && !lo && mhi < D.ss_hi;
    if (__any_sync(0xffffffffu, band)) {   // exact: max|W| over every category of the pattern
        const double *tile = W.Oq + (size_t)mt * W.Ctot * kClv20Step;
        __syncwarp();
        unsigned chi = 0u;
        for (int cc = 0; cc < W.Ctot; cc++) chi = max(chi, eigen_max20(tile + (size_t)cc * kClv20Step, W.inv_frag + (size_t)(cc / W.cpb) * 480, lane));
        if (band) lo = chi < kThresholdHi;
    }
    return lo;
}
// `mine(cc)` says whether this warp wrote category cc (category split: not all)
template <typename Mine>
__device__ __forceinline__ void norm_apply20(const Walk20ssOp &W, int mt, bool lo, int lane, Mine mine) {
    if (!lo) return;
    const int g = lane >> 2, q = lane & 3, lo2 = lane * 2, lot = 128 + g * 4 + 2 * q;
    double *tile = W.Oq + (size_t)mt * W.Ctot * kClv20Step;
    for (int cc = 0; cc < W.Ctot; cc++) if (mine(cc)) rescale20(tile + (size_t)cc * kClv20Step, lo2, lot, q);
}
template <typename Mine>
__device__ __forceinline__ bool norm_decide20(const Dev &D, const Walk20ssOp &W, int mt, unsigned mhi, int lane, Mine mine) {
    const bool lo = norm_test20(D, W, mt, mhi, lane);
    norm_apply20(W, mt, lo, lane, mine);
    return lo;
}

// One (op, class block) for the warp's 32 patterns: categories c0, c0+cstep, ... (C of them; c0 = 0, cstep = 1 unless the CTA's
// warps split the categories) in the OUTER loop — the two transition matrices of a category are loaded from the staging block
// into registers as B fragments once and serve the four M-tiles of the inner loop.
template <bool SAFE, bool MIX, int NCASE, bool CSP>
__device__ __forceinline__ void walk20ss_op(const Dev &D, const int C, const int c0, const int cstep, const Walk20ssOp &W, const double *sA,
                                            const double *sB, const int TS, const int lane, unsigned (&mhi_out)[4], unsigned (&sc_out)[4]) {
    const int g = lane >> 2, q = lane & 3;
    const int lo2 = lane * 2, lot = 128 + g * 4 + 2 * q;
    const long long mt_stride = (long long)W.Ctot * kClv20Step;
    const int coff = (W.first_class ? 0 : 0);   // (class offset is folded into Aq/Bq/Oq by the caller)
    (void)coff;
    // raw CLV fragments of the next (category, M-tile) step are in flight while this one runs its MMAs
    double nA[5], nB[5];
    int pf_j = 0, pf_mt = 0;
    auto fetch = [&]() {
        if (pf_j < C) {
            const long long off = pf_mt * mt_stride + (long long)(c0 + pf_j * cstep) * kClv20Step;
            if (NCASE == 2) {
                const double *p = W.Aq + off;
                const double2 v0 = *reinterpret_cast<const double2 *>(p + lo2), v1 = *reinterpret_cast<const double2 *>(p + 64 + lo2);
                nA[0] = v0.x; nA[1] = v0.y; nA[2] = v1.x; nA[3] = v1.y; nA[4] = p[128 + lane];
            }
            if (NCASE >= 1) {
                const double *p = W.Bq + off;
                const double2 v0 = *reinterpret_cast<const double2 *>(p + lo2), v1 = *reinterpret_cast<const double2 *>(p + 64 + lo2);
                nB[0] = v0.x; nB[1] = v0.y; nB[2] = v1.x; nB[3] = v1.y; nB[4] = p[128 + lane];
            }
            if (++pf_mt == 4) { pf_mt = 0; pf_j++; }
        }
    };
    fetch();
    unsigned mhi[4] = {0u, 0u, 0u, 0u};
    const int l16 = lane & 15;
    for (int j = 0; j < C; j++) {
        const int c = c0 + j * cstep;   // category inside the class block
        double PA[5][3], PB[5][3];
        if (NCASE == 2) {
            const double *m = sA + (size_t)c * kMat20;
#pragma unroll
            for (int kt = 0; kt < 5; kt++) {
                const double2 v = *reinterpret_cast<const double2 *>(m + kt * 64 + 2 * lane);
                PA[kt][0] = v.x; PA[kt][1] = v.y;
                const double t = m[320 + kt * 16 + l16];
                PA[kt][2] = lane < 16 ? t : 0.0;
            }
        }
        if (NCASE >= 1) {
            const double *m = sB + (size_t)c * kMat20;
#pragma unroll
            for (int kt = 0; kt < 5; kt++) {
                const double2 v = *reinterpret_cast<const double2 *>(m + kt * 64 + 2 * lane);
                PB[kt][0] = v.x; PB[kt][1] = v.y;
                const double t = m[320 + kt * 16 + l16];
                PB[kt][2] = lane < 16 ? t : 0.0;
            }
        }
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
            double fA[5], fB[5];
#pragma unroll
            for (int k = 0; k < 5; k++) { fA[k] = nA[k]; fB[k] = nB[k]; }
            fetch();
            double a[3][2], b[3][2];
            if (NCASE == 2) {
#pragma unroll
                for (int nt = 0; nt < 3; nt++) { a[nt][0] = 0.0; a[nt][1] = 0.0; }
            } else {
                const double *tA = sA + (size_t)((W.stA >> (8 * mt)) & 0xffu) * TS + c * 20 + 2 * q;
#pragma unroll
                for (int nt = 0; nt < 3; nt++) {
                    double2 t = make_double2(0.0, 0.0);
                    if (nt < 2 || q < 2) t = *reinterpret_cast<const double2 *>(tA + 8 * nt);
                    a[nt][0] = t.x; a[nt][1] = t.y;
                }
            }
            if (NCASE >= 1) {
#pragma unroll
                for (int nt = 0; nt < 3; nt++) { b[nt][0] = 0.0; b[nt][1] = 0.0; }
            } else {
                const double *tB = sB + (size_t)((W.stB >> (8 * mt)) & 0xffu) * TS + c * 20 + 2 * q;
#pragma unroll
                for (int nt = 0; nt < 3; nt++) {
                    double2 t = make_double2(0.0, 0.0);
                    if (nt < 2 || q < 2) t = *reinterpret_cast<const double2 *>(tB + 8 * nt);
                    b[nt][0] = t.x; b[nt][1] = t.y;
                }
            }
#pragma unroll
            for (int kt = 0; kt < 5; kt++) {
#pragma unroll
                for (int nt = 0; nt < 3; nt++) {
                    if (NCASE == 2) dmma_m8n8k4(a[nt][0], a[nt][1], fA[kt], PA[kt][nt]);
                    if (NCASE >= 1) dmma_m8n8k4(b[nt][0], b[nt][1], fB[kt], PB[kt][nt]);
                }
            }
            double out[3][2];
#pragma unroll
            for (int nt = 0; nt < 3; nt++) { out[nt][0] = a[nt][0] * b[nt][0]; out[nt][1] = a[nt][1] * b[nt][1]; }
            unsigned chi = 0u;
#pragma unroll
            for (int nt = 0; nt < 3; nt++) chi = max(chi, max(abs_hi(out[nt][0]), abs_hi(out[nt][1])));
            double *po = W.Oq + mt * mt_stride + (long long)c * kClv20Step;
            *reinterpret_cast<double2 *>(po + lo2) = make_double2(out[0][0], out[0][1]);
            *reinterpret_cast<double2 *>(po + 64 + lo2) = make_double2(out[1][0], out[1][1]);
            if (q < 2) *reinterpret_cast<double2 *>(po + lot) = make_double2(out[2][0], out[2][1]);
            if (SAFE) {   // per (pattern, category) decision (tree/phylokernelnew.h:1994-2005)
                unsigned sc = 0;
                if (NCASE == 2) sc += W.Asc[c * kWarp + 8 * mt];
                if (NCASE >= 1) sc += W.Bsc[c * kWarp + 8 * mt];
                if (NCASE != 0) {
                    chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 1));
                    chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 2));
                    const bool may = (W.may_mask >> mt) & 1u;
                    bool lo = may && chi < D.ss_lo;
                    const bool band = may && !lo && chi < D.ss_hi;
                    if (__any_sync(0xffffffffu, band)) {
                        __syncwarp();
                        const unsigned ce = eigen_max20(po, W.inv_frag, lane);   // caller points inv_frag at this class
                        if (band) lo = ce < kThresholdHi;
                    }
                    if (lo) { rescale20(po, lo2, lot, q); sc += 1; }
                }
                if (q == 0) W.Osc[c * kWarp + 8 * mt] = (uint16_t)sc;
            } else mhi[mt] = max(mhi[mt], chi);
        }
    }
    if (!SAFE) {
#pragma unroll
        for (int mt = 0; mt < 4; mt++) {
            unsigned m = mhi[mt];
            m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
            unsigned sc = 0;
            if (NCASE == 2) sc += W.Asc[8 * mt];
            if (NCASE >= 1) sc += W.Bsc[8 * mt];
            if (CSP) { mhi_out[mt] = m; sc_out[mt] = sc; continue; }   // decided after the CTA barrier (walk20ss_csplit_finish)
            if (MIX) {   // fold in the classes walked before this one (each lane owns its slot: no synchronisation)
                unsigned *slot = W.s_mhi + mt * blockDim.x + threadIdx.x;
                if (!W.first_class) m = max(m, *slot);
                if (!W.last_class) *slot = m;
            }
            if (!MIX || W.last_class) {
                if (NCASE != 0 && norm_decide20(D, W, mt, m, lane, [](int) { return true; })) sc += 1;
                if (q == 0) W.Osc[8 * mt] = (uint16_t)sc;
            }
        }
    }
}

// NW warps per CTA; tile = up to NW 32-pattern blocks (balanced tiles) or, with CSP (small problems), ONE
// block whose categories the warps split.  Staging: one TMA bulk load per (op, class block) into a 2-deep ring, completion on
// an mbarrier per buffer.
template <bool SAFE, bool MIX, int NW, int MINB, bool CSP>
__global__ void __launch_bounds__(NW * 32, MINB) k_walk_n20ss(Dev D, const double *stage, int nops, unsigned long long *counter, int ntiles) {
    static_assert(!(CSP && MIX), "category split serves single-class models");
    constexpr int NT = NW * 32;
    extern __shared__ __align__(16) double s_dyn20[];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ long long s_item;
    __shared__ unsigned s_mhi[(MIX || CSP) ? 4 * NT : 1];
    unsigned seq = 0;
    if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2;
    const int nblk = MIX ? D.nmix : 1, cpb = D.C / nblk;
    const int rows = D.su + 1, TS = tab20_stride(cpb);
    const size_t stage_sz = ss_stage_doubles(cpb, D.su);
    const int nvo = nops * nblk;
    const long long nblocks = D.Ppad >> 5;
    auto issue = [&](int v, int ncase, unsigned sq) {   // thread 0: staging block v -> buffer sq & 1
        const unsigned bytes = (unsigned)(ss_stage_used(cpb, D.su, ncase) * sizeof(double));
        double *sdst = s_dyn20 + (size_t)(sq & 1) * stage_sz;
        const unsigned b = (unsigned)__cvta_generic_to_shared(&s_bar[sq & 1]), d = (unsigned)__cvta_generic_to_shared(sdst);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                     "l"(stage + (size_t)v * stage_sz), "r"(bytes), "r"(b) : "memory");
    };
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = (long long)atomicAdd(counter, 1ULL);
        __syncthreads();
        const long long tile = s_item;
        if (tile >= ntiles) break;
        const long long b0 = CSP ? tile : tile * nblocks / ntiles, b1 = CSP ? tile + 1 : (tile + 1) * nblocks / ntiles;
        const long long p0 = CSP ? b0 * 32 : (b0 + warp) * 32;
        const int csp_ncat = CSP ? (D.C - warp + NW - 1) / NW : 0;
        const bool active = CSP ? csp_ncat > 0 : b0 + warp < b1;
        const long long wbase = (p0 >> 5) * (long long)D.B * kWarp;
        const long long sbase = SAFE ? (p0 >> 5) * (long long)D.C * kWarp : p0;
        unsigned may_mask = 0;
        if (active) {
#pragma unroll
            for (int mt = 0; mt < 4; mt++) if (D.invar[p0 + 8 * mt + g] == 0.0) may_mask |= 1u << mt;
        }
        auto tip_states = [&](const int4 &hd, unsigned &sA, unsigned &sB) {
            sA = sB = 0;
            if (!active) return;
            if (hd.y <= 1) {
#pragma unroll
                for (int mt = 0; mt < 4; mt++) sA |= (unsigned)D.tips[(long long)hd.z * D.Ppad + p0 + 8 * mt + g] << (8 * mt);
            }
            if (hd.y == 0) {
#pragma unroll
                for (int mt = 0; mt < 4; mt++) sB |= (unsigned)D.tips[(long long)hd.w * D.Ppad + p0 + 8 * mt + g] << (8 * mt);
            }
        };
        // headers travel two blocks ahead in registers (they steer the copy and the tip-state loads of the next block)
        int4 h_next = *reinterpret_cast<const int4 *>(stage);
        int4 h_next2 = nvo > 1 ? *reinterpret_cast<const int4 *>(stage + stage_sz) : h_next;
        if (tid == 0) issue(0, h_next.y, seq);
        unsigned stA_next, stB_next;
        tip_states(h_next, stA_next, stB_next);
        int m = 0;
        for (int v = 0; v < nvo; v++, seq++) {
            const double *sb = s_dyn20 + (size_t)(seq & 1) * stage_sz;
            const int4 h = h_next;
            const unsigned stA = stA_next, stB = stB_next;
            h_next = h_next2;
            if (v + 2 < nvo) h_next2 = *reinterpret_cast<const int4 *>(stage + (size_t)(v + 2) * stage_sz);
            mbar_wait(&s_bar[seq & 1], (seq >> 1) & 1);
            __syncthreads();   // block v has landed for everyone; everyone is done with the other buffer
            if (v + 1 < nvo) {
                if (tid == 0) issue(v + 1, h_next.y, seq + 1);
                if (!MIX || m + 1 == nblk) tip_states(h_next, stA_next, stB_next);
            }
            unsigned mh[4] = {0, 0, 0, 0}, sc[4] = {0, 0, 0, 0};
            Walk20ssOp W;
            const int dst = h.x, ncase = h.y, aid = h.z, bid = h.w;
            if (active) {
                const int class_off = m * cpb * kClv20Step;
                W.Aq = D.partial + (long long)aid * D.slot_stride + wbase + class_off;
                W.Bq = D.partial + (long long)bid * D.slot_stride + wbase + class_off;
                W.Oq = D.partial + (long long)dst * D.slot_stride + wbase + class_off;
                const int sc_off = SAFE ? m * cpb * kWarp : 0;
                W.Asc = D.scale + (long long)aid * D.scale_stride + sbase + sc_off + g;
                W.Bsc = D.scale + (long long)bid * D.scale_stride + sbase + sc_off + g;
                W.Osc = D.scale + (long long)dst * D.scale_stride + sbase + sc_off + g;
                W.may_mask = may_mask; W.stA = stA; W.stB = stB;
                W.first_class = m == 0; W.last_class = m + 1 == nblk;
                W.Ctot = D.C; W.cpb = cpb; W.s_mhi = s_mhi;
                W.inv_frag = D.inv_frag20 + (SAFE ? (size_t)m * 480 : 0);
                const double *sA = sb + 4, *sB = sA + ss_part_doubles(cpb, D.su, ncase <= 1);
                const int nc = CSP ? csp_ncat : cpb, c0 = CSP ? warp : 0, cstep = CSP ? NW : 1;
                // NORM + MIX: the exact path and the rescale run over the whole CLV: point Oq at class 0
                if (MIX && !SAFE) {
                    // walk20ss_op addresses categories relative to the class block; norm_decide20 needs the whole CLV
                }
                if (ncase == 2) walk20ss_op<SAFE, MIX, 2, CSP>(D, nc, c0, cstep, W, sA, sB, TS, lane, mh, sc);
                else if (ncase == 1) walk20ss_op<SAFE, MIX, 1, CSP>(D, nc, c0, cstep, W, sA, sB, TS, lane, mh, sc);
                else walk20ss_op<SAFE, MIX, 0, CSP>(D, nc, c0, cstep, W, sA, sB, TS, lane, mh, sc);
            }
            if (CSP && !SAFE) {
                // category split, NORM_LH: the pattern's maximum spans the categories of all warps
                if (active) {
#pragma unroll
                    for (int mt = 0; mt < 4; mt++) s_mhi[(warp * 4 + mt) * kWarp + lane] = mh[mt];
                }
                __syncthreads();   // also orders the warps' CLV stores for the exact path
                bool lo4[4] = {false, false, false, false};
                if (active) {
                    const int nwarps = D.C < NW ? D.C : NW;
#pragma unroll
                    for (int mt = 0; mt < 4; mt++) {
                        unsigned mm = 0;
                        for (int w = 0; w < nwarps; w++) mm = max(mm, s_mhi[(w * 4 + mt) * kWarp + lane]);
                        lo4[mt] = ncase != 0 && norm_test20(D, W, mt, mm, lane);
                    }
                }
                __syncthreads();   // every warp has decided (the exact path read all categories) before anyone rescales
                if (active) {
                    const int q = lane & 3;
#pragma unroll
                    for (int mt = 0; mt < 4; mt++) {
                        norm_apply20(W, mt, lo4[mt], lane, [&](int cc) { return cc % NW == warp; });
                        if (warp == 0 && q == 0) W.Osc[8 * mt] = (uint16_t)(sc[mt] + (lo4[mt] ? 1u : 0u));
                    }
                }
            }
            if (MIX && ++m == nblk) m = 0;
        }
    }
}

// W = inv_evec . t for one CLV slot (same fragment-major layout) + its scale counts: what the root / theta kernels and
// b200lk_get_partial read.  One warp per 32-pattern block.
__device__ __forceinline__ void clv20_to_eigen_block(const Dev &D, int src, int dst, long long blk) {
    if (blk >= (D.Ppad >> 5)) return;
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    const int lo2 = lane * 2, lot = 128 + g * 4 + 2 * q;
    const int cpb = D.C / D.nmix;
    const double *S = D.partial + (long long)src * D.slot_stride;
    double *O = D.partial + (long long)dst * D.slot_stride;
    const long long base = blk * (long long)D.B * kWarp;
    for (int m = 0; m < D.nmix; m++) {
        double Bi[5][3];
#pragma unroll
        for (int kt = 0; kt < 5; kt++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) Bi[kt][nt] = __ldg(D.inv_frag20 + (size_t)m * 480 + (kt * 3 + nt) * kWarp + lane);
        for (int mt = 0; mt < 4; mt++)
            for (int r = 0; r < cpb; r++) {
                const long long off = base + ((long long)mt * D.C + m * cpb + r) * kClv20Step;
                const double *p = S + off;
                const double2 v0 = *reinterpret_cast<const double2 *>(p + lo2), v1 = *reinterpret_cast<const double2 *>(p + 64 + lo2);
                const double f[5] = {v0.x, v0.y, v1.x, v1.y, p[128 + lane]};
                double o[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                for (int kt = 0; kt < 5; kt++)
#pragma unroll
                    for (int nt = 0; nt < 3; nt++) dmma_m8n8k4(o[nt][0], o[nt][1], f[kt], Bi[kt][nt]);
                double *po = O + off;
                *reinterpret_cast<double2 *>(po + lo2) = make_double2(o[0][0], o[0][1]);
                *reinterpret_cast<double2 *>(po + 64 + lo2) = make_double2(o[1][0], o[1][1]);
                if (q < 2) *reinterpret_cast<double2 *>(po + lot) = make_double2(o[2][0], o[2][1]);
            }
    }
    const long long sb = blk * (long long)D.S * kWarp;
    const uint16_t *ss = D.scale + (long long)src * D.scale_stride + sb;
    uint16_t *os = D.scale + (long long)dst * D.scale_stride + sb;
    for (int k = lane; k < D.S * kWarp; k += kWarp) os[k] = ss[k];
}
__global__ void __launch_bounds__(128) k_clv20_to_eigen(Dev D, int src, int dst) {
    const long long nblocks = D.Ppad >> 5;
    for (long long blk = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); blk < nblocks; blk += (long long)gridDim.x * (blockDim.x >> 5))
        clv20_to_eigen_block(D, src, dst, blk);
}

}  // namespace b200lk
