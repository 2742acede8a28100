// b200lk_walkgen.cuh — the FP64-MMA pattern-tile walker for GENERAL state counts (8 < n <= 64, n != 20): codon models (61 states
// -> 64: 4 % padding), morphological / binary-coded data, 5- to 19-state and 21- to 60-state user models.  The reference serves
// them with computePartialLikelihoodGenericSIMD (tree/phylokernelnew.h:1289-2039; instantiations tree/phylokernelfma.cpp:47-180)
// and always with SAFE_LH for n not in {4, 20} (tree/phylotreesse.cpp:95-96); round 1 ran them thread-per-pattern with three
// 64-double local arrays per thread.
//
// Same design as the 20-state walkers (b200lk_walk20ss.cuh, b200lk_walk20v2.cuh), templated on NT8 = ceil(n / 8) state tiles:
//   * STATE-SPACE CLVs in a fragment-major layout ("f8"): a 32-pattern block is [M-tile 4][category C][NT8 tiles][8 patterns][8
//     values]; lane 4g+q of a warp owns values 2q, 2q+1 of pattern g in every tile = the 16-byte chunk number `lane` of the tile;
//     padded values are exact zeros.  Two mat-vecs per inner node: t = (P_A t_A) (.) (P_B t_B), P_c(len) = evec exp(eval r_c len)
//     inv_evec built on the device per (op, category) and staged in shared memory in B-fragment order (one TMA bulk load);
//   * the transition matrices are too large for registers (n8^2 / 32 doubles per lane): their B fragments are read from shared
//     memory at each use, two per 16-byte load — half of the shared-memory bandwidth at the full DMMA rate;
//   * the children are processed one after the other (A fragments of one child + both accumulator sets in registers);
//   * per-warp cp.async rings feed the CLV fragments K = 3 steps ahead (a lane copies, reads and — one op earlier — stored the
//     same chunks: no cross-lane dependency), staging blocks are handed over with an mbarrier + release counter, no CTA barrier;
//   * a virtual op is (op, category): the staging block holds that category's two matrices / tip tables; NORM_LH carries the
//     pattern's maximum over the categories of an op in a per-lane shared-memory slot;
//   * exact scale counts as for 20 states: the max|t| bounds decide, the band takes W = inv_evec . t of the tile (thread loop).
#pragma once
#include "b200lk_walk20v2.cuh"

namespace b200lk {

constexpr int kGenNW = 4;       // warps per CTA
// ring depth (steps; a virtual op has 4 steps): 3, except for 8 state tiles (57..64 states: codon models) where two staged 64 x 64
// matrices per staging block leave room for 2
__host__ __device__ constexpr int gen_ring_depth(int n8) { return n8 >= 64 ? 2 : 3; }

// element (tile nt, pattern g, value j) inside a step block of NT8*64 doubles
__host__ __device__ __forceinline__ int f8_off(int i, int g) { return (i >> 3) * 64 + g * 8 + (i & 7); }

// staging block of one (op, category): header[4] | part A | part B; a part is the tip table [(su+1)][n8] of a tip child or the
// transition matrix of an internal child in B-fragment order [kt][nt/2][lane][2]:
//   element (kt, nt, lane = 4g+q) = P[x = 8nt+g][y = 8(kt/2) + 2q + kt%2], zero outside n x n.     (n8^2 doubles)
// doubles of a matrix in that order: KT * ceil(NT8/2) * 64 (= n8^2 for an even number of state tiles)
__host__ __device__ inline size_t gen_mat_doubles(int n8) { return (size_t)(n8 / 4) * ((n8 / 8 + 1) / 2) * 64; }
__host__ __device__ inline size_t gen_part_doubles(int n8, int su, bool tip) { return tip ? (size_t)(su + 1) * n8 : gen_mat_doubles(n8); }
__host__ __device__ inline size_t gen_stage_doubles(int n8, int su) {
    const size_t a = (size_t)(su + 1) * n8, b = gen_mat_doubles(n8);
    return 4 + 2 * (a > b ? a : b);
}
__host__ __device__ inline size_t gen_stage_used(int n8, int su, int ncase) {
    return 4 + gen_part_doubles(n8, su, ncase <= 1) + gen_part_doubles(n8, su, ncase == 0);
}
__host__ __device__ inline size_t gen_smem_doubles(int n8, int su) { return 2 * gen_stage_doubles(n8, su) + (size_t)kGenNW * gen_ring_depth(n8) * 2 * n8 * 8; }

// One CTA per (op, category): exp(eval r_c len), the tip table (partial_lh_leaves, tree/phylokernelnew.h:1036-1056) or P_c(len)
__device__ __forceinline__ void walk_info_gen_body(const Dev &D, const WalkJob20 &job, double *st, const int c, double *s_exp) {
    const int n = D.n, n8 = D.n8, m = c / D.denom;
    for (int k = threadIdx.x; k < 2 * n; k += blockDim.x) {
        const int ch = k / n, i = k - ch * n;
        const double len_child = D.cat_rate[c % D.ncat] * resolve_len_cat(D, ch ? job.lenB : job.lenA, c % D.ncat);   // :1017
        s_exp[k] = exp(D.eval[m * n + i] * len_child);                                                 // :1022
    }
    if (threadIdx.x == 0) {
        int32_t *h = reinterpret_cast<int32_t *>(st);
        h[0] = job.dst; h[1] = job.ncase; h[2] = job.aid; h[3] = job.bid; h[4] = c; h[5] = h[6] = h[7] = 0;
    }
    __syncthreads();
    const int rows = D.su + 1;
    const double *ev = D.evec + (size_t)m * n * n, *iv = D.inv_evec + (size_t)m * n * n;
    double *part = st + 4;
    for (int ch = 0; ch < 2; ch++) {
        const bool is_tip = ch == 0 ? (job.ncase <= 1) : (job.ncase == 0);
        const double *ex = s_exp + ch * n;
        if (is_tip) {
            for (int k = threadIdx.x; k < rows * n8; k += blockDim.x) {
                const int state = k / n8, x = k - state * n8;
                double v = 0.0;
                if (x < n) {
                    if (state == D.su) v = 1.0;
                    else {
                        const double *tip = D.tip_lh + ((size_t)state * D.nmix + m) * n, *evx = ev + (size_t)x * n;
                        if ((n & 3) == 0) {   // Vec4d lane accumulators + horizontal_add
                            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                            for (int i = 0; i < n; i += 4) {
                                a0 = fma(evx[i] * ex[i], tip[i], a0); a1 = fma(evx[i + 1] * ex[i + 1], tip[i + 1], a1);
                                a2 = fma(evx[i + 2] * ex[i + 2], tip[i + 2], a2); a3 = fma(evx[i + 3] * ex[i + 3], tip[i + 3], a3);
                            }
                            v = (a0 + a1) + (a2 + a3);
                        } else {              // scalar branch (:1085-1099)
                            v = (evx[0] * ex[0]) * tip[0];
                            for (int i = 1; i < n; i++) v += (evx[i] * ex[i]) * tip[i];
                        }
                    }
                }
                part[k] = v;
            }
            part += (size_t)rows * n8;
        } else {
            for (int k = threadIdx.x; k < (int)gen_mat_doubles(n8); k += blockDim.x) {
                // k = ((kt * (NT8/2 rounded up) + ntp) * 32 + lane) * 2 + j   with nt = 2 ntp + j
                const int j = k & 1, ln = (k >> 1) & 31, r = k >> 6, ntp_n = (n8 / 8 + 1) / 2, kt = r / ntp_n, ntp = r - kt * ntp_n;
                const int nt = 2 * ntp + j, x = 8 * nt + (ln >> 2), y = 8 * (kt >> 1) + 2 * (ln & 3) + (kt & 1);
                double v = 0.0;
                if (x < n && y < n) {
                    const double *evx = ev + (size_t)x * n;
                    double a0 = 0.0, a1 = 0.0;
                    int i = 0;
                    for (; i + 1 < n; i += 2) {
                        a0 = fma(evx[i] * ex[i], iv[(size_t)i * n + y], a0);
                        a1 = fma(evx[i + 1] * ex[i + 1], iv[(size_t)(i + 1) * n + y], a1);
                    }
                    if (i < n) a0 = fma(evx[i] * ex[i], iv[(size_t)i * n + y], a0);
                    v = a0 + a1;
                }
                part[k] = v;
            }
            part += gen_mat_doubles(n8);
        }
    }
}
__global__ void __launch_bounds__(256) k_walk_info_gen(Dev D, const WalkJob20 *jobs, double *stage, unsigned long long *counter) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0ULL;
    extern __shared__ double s_expg[];   // [2][n]
    const int op = blockIdx.x / D.C, c = blockIdx.x - op * D.C;
    walk_info_gen_body(D, jobs[op], stage + (size_t)blockIdx.x * gen_stage_doubles(D.n8, D.su), c, s_expg);
}

// exact max_i |(inv_evec . t)[p][i]| (exponent word) for the 8 patterns of one (M-tile, category) block this warp has stored; the 4
// lanes of a pattern split the rows i.  Rare path (patterns whose max|t| lies in the band around 2^-256).
__device__ __forceinline__ unsigned eigen_max_gen(const Dev &D, const double *blk, int m, int lane) {
    const int g = lane >> 2, q = lane & 3, n = D.n;
    const double *iv = D.inv_evec + (size_t)m * n * n;
    unsigned chi = 0u;
    for (int i = q; i < n; i += 4) {
        double w = 0.0;
        for (int x = 0; x < n; x++) w = fma(__ldg(iv + (size_t)i * n + x), blk[f8_off(x, g)], w);
        chi = max(chi, abs_hi(w));
    }
    chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 1));
    chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 2));
    return chi;
}

template <int NT8>
__device__ __forceinline__ void rescale_gen(double *blk, int lane) {
#pragma unroll
    for (int nt = 0; nt < NT8; nt++) {
        double2 *p = reinterpret_cast<double2 *>(blk + nt * 64 + 2 * lane);
        double2 v = *p;
        v.x *= kTwo256; v.y *= kTwo256;
        *p = v;
    }
}

struct GenShared {
    unsigned long long full[2];
    unsigned release[2];
    unsigned mhi[4 * kGenNW * 32];   // NORM_LH: per-lane running maximum over the categories of an op, [warp][M-tile][lane]
};

// One tile (32-pattern blocks [b0, b1), at most kGenNW) of one engine: the whole op list, nvo = nops * C virtual ops
template <bool SAFE, int NT8>
__device__ __forceinline__ void walkgen_tile(const Dev &D, const double *stage, const int nops, const long long b0, const long long b1, double *s_dyn,
                                             GenShared &sh, unsigned &seq) {
    constexpr int N8 = NT8 * 8, STEP = N8 * 8, KT = NT8 * 2, NTP = (NT8 + 1) / 2, kGenK = gen_ring_depth(N8);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const int C = D.C;
    const size_t stage_sz = gen_stage_doubles(N8, D.su);
    double *ring = s_dyn + 2 * stage_sz + (size_t)warp * kGenK * 2 * STEP;
    const int nvo = nops * C;
    const long long mt_stride = (long long)C * STEP;
    auto stage_issue = [&](int v, int ncase, unsigned sq) {
        const unsigned bytes = (unsigned)(gen_stage_used(N8, D.su, ncase) * sizeof(double));
        mbar_expect(&sh.full[sq & 1], bytes);
        bulk_g2s(s_dyn + (size_t)(sq & 1) * stage_sz, stage + (size_t)v * stage_sz, bytes, &sh.full[sq & 1]);
    };
    const bool active = b0 + warp < b1;
    const long long p0 = (b0 + warp) * 32;
    const long long wbase = (p0 >> 5) * (long long)C * N8 * kWarp;
    const long long sbase = SAFE ? (p0 >> 5) * (long long)C * kWarp : p0;
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
    int4 h_next = *reinterpret_cast<const int4 *>(stage);
    int4 h_next2 = nvo > 1 ? *reinterpret_cast<const int4 *>(stage + stage_sz) : h_next;
    if (tid == 0) {
        stage_issue(0, h_next.y, seq);
        if (nvo > 1) stage_issue(1, h_next2.y, seq + 1);
    }
    unsigned stA_next, stB_next;
    tip_states(h_next, stA_next, stB_next);
    // ring: step mt of virtual op iv (category iv % C) into stage k; a lane copies the chunks it will read
    auto ring_copy = [&](const int4 &hd, int c, int mt, int k) {
        if (hd.y != 0) {
            const long long off = wbase + mt * mt_stride + (long long)c * STEP + 2 * lane;
            double *dst = ring + (size_t)k * 2 * STEP + 2 * lane;
            if (hd.y == 2) {
                const double *src = D.partial + (long long)hd.z * D.slot_stride + off;
#pragma unroll
                for (int nt = 0; nt < NT8; nt++) cp_async16(dst + nt * 64, src + nt * 64);
            }
            const double *src = D.partial + (long long)hd.w * D.slot_stride + off;
#pragma unroll
            for (int nt = 0; nt < NT8; nt++) cp_async16(dst + STEP + nt * 64, src + nt * 64);
        }
        cp_async_commit();
    };
    int iv = 0, is = 0, ik = 0;
    auto issue_next = [&](const int4 &h_now, const int4 &h_nxt, int v_now) {
        if (iv < nvo) {
            ring_copy(iv == v_now ? h_now : h_nxt, iv % C, is, ik);
            if (++is == 4) { is = 0; iv++; }
        } else cp_async_commit();
        ik = ik + 1 == kGenK ? 0 : ik + 1;
    };
    if (active)
        for (int k = 0; k < kGenK; k++) issue_next(h_next, h_next2, 0);
    int ck = 0;
    for (int v = 0; v < nvo; v++, seq++) {
        const double *sb = s_dyn + (size_t)(seq & 1) * stage_sz;
        const int4 h = h_next;
        const unsigned stA = stA_next, stB = stB_next;
        h_next = h_next2;
        if (v + 2 < nvo) h_next2 = *reinterpret_cast<const int4 *>(stage + (size_t)(v + 2) * stage_sz);
        mbar_wait(&sh.full[seq & 1], (seq >> 1) & 1);
        const int c = v % C;   // category of this virtual op (class-major, rate-minor: c = m * ncat + r)
        if (v + 1 < nvo && c + 1 == C) tip_states(h_next, stA_next, stB_next);
        const int dst = h.x, ncase = h.y;
        if (active) {
            const int m = c / D.denom;
            double *Oq = D.partial + (long long)dst * D.slot_stride + wbase + (long long)c * STEP;
            const int sc_off = SAFE ? c * kWarp : 0;
            const uint16_t *Asc = D.scale + (long long)h.z * D.scale_stride + sbase + sc_off + g;
            const uint16_t *Bsc = D.scale + (long long)h.w * D.scale_stride + sbase + sc_off + g;
            uint16_t *Osc = D.scale + (long long)dst * D.scale_stride + sbase + sc_off + g;
            const double *sA = sb + 4, *sB = sA + gen_part_doubles(N8, D.su, ncase <= 1);
            bool rescaled = false;
#pragma unroll 1
            for (int mt = 0; mt < 4; mt++) {
                const double *rs = ring + (size_t)ck * 2 * STEP + 2 * lane;
                cp_async_wait_group<kGenK - 1>();   // this lane's chunks of the step have landed (it reads no one else's)
                double acc[2][NT8][2];
#pragma unroll
                for (int ch = 0; ch < 2; ch++) {
                    const bool internal = ch == 0 ? ncase == 2 : ncase >= 1;
                    if (internal) {
                        double a[KT];
#pragma unroll
                        for (int nt = 0; nt < NT8; nt++) {
                            const double2 t = *reinterpret_cast<const double2 *>(rs + ch * STEP + nt * 64);
                            a[2 * nt] = t.x; a[2 * nt + 1] = t.y;
                        }
#pragma unroll
                        for (int nt = 0; nt < NT8; nt++) { acc[ch][nt][0] = 0.0; acc[ch][nt][1] = 0.0; }
                        const double *mat = (ch == 0 ? sA : sB) + 2 * lane;
#pragma unroll
                        for (int kt = 0; kt < KT; kt++) {
#pragma unroll
                            for (int ntp = 0; ntp < NTP; ntp++) {
                                const double2 bf = *reinterpret_cast<const double2 *>(mat + (size_t)(kt * NTP + ntp) * 64);
                                dmma_m8n8k4(acc[ch][2 * ntp][0], acc[ch][2 * ntp][1], a[kt], bf.x);
                                if (2 * ntp + 1 < NT8) dmma_m8n8k4(acc[ch][2 * ntp + 1][0], acc[ch][2 * ntp + 1][1], a[kt], bf.y);
                            }
                        }
                    } else {
                        const unsigned st = ((ch == 0 ? stA : stB) >> (8 * mt)) & 0xffu;
                        const double *tab = (ch == 0 ? sA : sB) + (size_t)st * N8 + 2 * q;
#pragma unroll
                        for (int nt = 0; nt < NT8; nt++) {
                            const double2 t = *reinterpret_cast<const double2 *>(tab + 8 * nt);
                            acc[ch][nt][0] = t.x; acc[ch][nt][1] = t.y;
                        }
                    }
                }
                unsigned chi = 0u;
                double *po = Oq + mt * mt_stride;
#pragma unroll
                for (int nt = 0; nt < NT8; nt++) {
                    const double o0 = acc[0][nt][0] * acc[1][nt][0], o1 = acc[0][nt][1] * acc[1][nt][1];
                    chi = max(chi, max(abs_hi(o0), abs_hi(o1)));
                    *reinterpret_cast<double2 *>(po + nt * 64 + 2 * lane) = make_double2(o0, o1);
                }
                chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 1));
                chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 2));
                if (SAFE) {   // per (pattern, category) decision (tree/phylokernelnew.h:1994-2005)
                    unsigned sc = 0;
                    if (ncase == 2) sc += Asc[8 * mt];
                    if (ncase >= 1) sc += Bsc[8 * mt];
                    if (ncase != 0) {
                        const bool may = (may_mask >> mt) & 1u;
                        bool lo = may && chi < D.ss_lo;
                        const bool band = may && !lo && chi < D.ss_hi;
                        if (__any_sync(0xffffffffu, band)) {
                            __syncwarp();
                            const unsigned ce = eigen_max_gen(D, po, m, lane);
                            if (band) lo = ce < kThresholdHi;
                        }
                        if (__any_sync(0xffffffffu, lo)) {
                            __syncwarp();
                            if (lo) {   // the 4 lanes of the pattern rescale their chunks: all chunks of pattern g belong to lanes 4g..4g+3
                                rescale_gen<NT8>(po, lane);
                                sc += 1;
                            }
                        }
                    }
                    if (q == 0) Osc[8 * mt] = (uint16_t)sc;
                } else {      // NORM_LH: running maximum over the categories of this op
                    unsigned *slot = sh.mhi + (warp * 4 + mt) * kWarp + lane;
                    if (c != 0) chi = max(chi, *slot);
                    if (c + 1 != C) *slot = chi;
                    else {
                        unsigned sc = 0;
                        if (ncase == 2) sc += Asc[8 * mt];
                        if (ncase >= 1) sc += Bsc[8 * mt];
                        if (ncase != 0) {
                            const bool may = (may_mask >> mt) & 1u;
                            bool lo = may && chi < D.ss_lo;
                            const bool band = may && !lo && chi < D.ss_hi;
                            double *all = D.partial + (long long)dst * D.slot_stride + wbase + mt * mt_stride;   // category 0 of this M-tile
                            if (__any_sync(0xffffffffu, band)) {
                                __syncwarp();
                                unsigned ce = 0u;
                                for (int cc = 0; cc < C; cc++) ce = max(ce, eigen_max_gen(D, all + (size_t)cc * STEP, cc / D.denom, lane));
                                if (band) lo = ce < kThresholdHi;
                            }
                            if (__any_sync(0xffffffffu, lo)) {
                                __syncwarp();
                                if (lo) {
                                    for (int cc = 0; cc < C; cc++) rescale_gen<NT8>(all + (size_t)cc * STEP, lane);
                                    sc += 1;
                                    rescaled = true;
                                }
                            }
                        }
                        if (q == 0) Osc[8 * mt] = (uint16_t)sc;
                    }
                }
                __syncwarp();
                issue_next(h, h_next, v);
                ck = ck + 1 == kGenK ? 0 : ck + 1;
            }
            // a NORM_LH rescale rewrote blocks that the ring already holds for the next virtual op (its first K steps)
            if (!SAFE && __any_sync(0xffffffffu, rescaled) && v + 1 < nvo) {
                cp_async_wait_group<0>();
                __syncwarp();
                int k = ck;
                for (int s2 = 0; s2 < kGenK; s2++) {
                    ring_copy(h_next, (v + 1) % C, s2, k);
                    k = k + 1 == kGenK ? 0 : k + 1;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            const unsigned nrel = atomicAdd(&sh.release[seq & 1], 1u);
            if (nrel == kGenNW - 1) {
                sh.release[seq & 1] = 0;
                if (v + 2 < nvo) stage_issue(v + 2, h_next2.y, seq + 2);
            }
        }
    }
    cp_async_wait_group<0>();
}

template <bool SAFE, int NT8>
__global__ void __launch_bounds__(kGenNW * 32) k_walk_gen(Dev D, const double *stage, int nops, unsigned long long *counter, int ntiles) {
    extern __shared__ __align__(16) double s_dyn[];
    __shared__ __align__(8) GenShared sh;
    __shared__ long long s_item;
    if (threadIdx.x == 0) { mbar_init(&sh.full[0], 1); mbar_init(&sh.full[1], 1); sh.release[0] = sh.release[1] = 0; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const long long nblocks = D.Ppad >> 5;
    unsigned seq = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = (long long)atomicAdd(counter, 1ULL);
        __syncthreads();
        const long long tile = s_item;
        if (tile >= ntiles) break;
        walkgen_tile<SAFE, NT8>(D, stage, nops, tile * nblocks / ntiles, (tile + 1) * nblocks / ntiles, s_dyn, sh, seq);
    }
}

// partial_lh = inv_evec . t for one CLV slot of an f8 engine (the root / theta kernels and read-backs): thread per (pattern, category)
__global__ void __launch_bounds__(128) k_clv_gen_to_eigen(Dev D, int src, int dst) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int C = D.C, n = D.n;
    if (idx < D.Ppad * C) {
        const long long p = idx / C;
        const int c = (int)(idx - p * C);
        const ClvIdx at(D, p);
        const double *S = D.partial + (long long)src * D.slot_stride;
        double *O = D.partial + (long long)dst * D.slot_stride;
        const double *iv = D.inv_evec + (size_t)(c / D.denom) * n * n;
        double t[64];
        for (int x = 0; x < n; x++) t[x] = S[at(c, x)];
        int i = 0;
        for (; i + 3 < n; i += 4) {   // four rows at a time: one pass over t, four independent chains (each row's sum keeps its order)
            double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
            const double *r0 = iv + (size_t)i * n, *r1 = r0 + n, *r2 = r1 + n, *r3 = r2 + n;
            for (int x = 0; x < n; x++) {
                const double tx = t[x];
                w0 = fma(__ldg(r0 + x), tx, w0); w1 = fma(__ldg(r1 + x), tx, w1);
                w2 = fma(__ldg(r2 + x), tx, w2); w3 = fma(__ldg(r3 + x), tx, w3);
            }
            O[at(c, i)] = w0; O[at(c, i + 1)] = w1; O[at(c, i + 2)] = w2; O[at(c, i + 3)] = w3;
        }
        for (; i < n; i++) {
            double w = 0.0;
            for (int x = 0; x < n; x++) w = fma(__ldg(iv + (size_t)i * n + x), t[x], w);
            O[at(c, i)] = w;
        }
        for (int i = n; i < D.n8; i++) O[at(c, i)] = 0.0;
    }
    // scale counts: same layout in both slots
    const long long ns = D.Ppad * D.S;
    for (long long k = idx; k < ns; k += (long long)gridDim.x * blockDim.x)
        D.scale[(long long)dst * D.scale_stride + k] = D.scale[(long long)src * D.scale_stride + k];
}

}  // namespace b200lk
