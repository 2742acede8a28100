// b200lk_walk20v2.cuh — the 20-state state-space walker, second generation: asynchronous-copy pipelines instead of register
// prefetch and CTA barriers.
//
// ncu on k_walk_n20ss (profiles/r2/k_walk_n20ss_ncu_summary.txt; config 3, 5.88 ms): the FP64-MMA pipe is busy 37 % of the
// time; 40 % of all stall samples are long-scoreboard waits on the CLV fragments that each lane prefetches ONE step ahead into
// registers, 17 % sit at the per-op __syncthreads, and the 168-register cap (3 CTAs x 4 warps per SM) spills.  Here:
//   * ONE CTA per SM with NW = 12 warps: a staging block (the op's transition matrices / tip tables, 25-31 KB; one TMA bulk load)
//     is fetched once per 12 pattern blocks instead of once per 4;
//   * every warp owns a K = 4 deep ring in shared memory for the CLV fragments of its next steps, fed with cp.async (16 bytes per
//     lane, L1 bypassed) K steps ahead — across op boundaries too: no registers hold data in flight and nothing waits on a global
//     load.  The fragment-major layout makes the 8 patterns x 20 values of a step one contiguous 1280-byte run per child in which
//     lane l owns the 16-byte chunks l and l + 32: the copy is coalesced, the A-fragment reads are conflict-free, and the chunks a
//     lane reads are the ones it copied (and, one op earlier, stored) itself, except for the 4-value tail;
//   * no CTA barrier: staging blocks live in a 2-deep ring with a `full` mbarrier (TMA completion) and a release counter per
//     buffer — the LAST warp to leave block v issues the TMA of block v+2 — so the warps of a CTA drift apart by up to an op.
// Arithmetic, staging blocks and the exact-scale-count logic are those of b200lk_walk20ss.cuh.
//
// Hazards of prefetching across ops (a child of op v+1 is very often the CLV that op v is writing):
//   * the block (M-tile, c) of op v+1's child is requested K steps before its use, i.e. at the earliest at the end of the step
//     of op v that WROTE it (K <= steps per op); cp.async reads through the generic proxy, so program order (own chunks) and the
//     __syncwarp before the request (the tail chunk, stored by a neighbour lane) are all the ordering it needs.  (A first version
//     used cp.async.bulk: the async proxy needs fence.proxy.async after the stores, which cost a membar stall of 4.8 issue slots
//     per instruction — profiles/r2/k_walk_n20v2_tma_ring_ncu.txt.)
//   * NORM_LH decides the 2^256 rescale of a pattern at the END of the op, after the first K steps of the next op have been
//     requested: when any lane of the warp rescales, the ring is drained and those steps are requested again.
#pragma once
#include "b200lk_walk20ss.cuh"

namespace b200lk {

constexpr int kRingK = 4;                            // ring depth in steps
constexpr int kRingStage = 2 * kClv20Step;           // doubles per stage: child A | child B
__host__ __device__ inline size_t v2_smem_doubles(int cpb, int su, int nw, int sd) { return (size_t)sd * ss_stage_doubles(cpb, su) + (size_t)nw * kRingK * kRingStage; }

__device__ __forceinline__ void bulk_g2s(void *sdst, const void *gsrc, unsigned bytes, unsigned long long *bar) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared memory of a v2 CTA (the kernels below declare it; the tile function is shared with the grouped launch)
// SD = depth of the staging ring (3 where shared memory allows: a warp may then run two ops ahead of the slowest warp of its CTA
// before it has to wait for a staging buffer; ncu on the 2-deep ring showed 16 % of the stall samples in that wait)
template <int NW, bool NEED_MHI, int SD>
struct V2Shared {
    unsigned long long full[SD];             // staging buffer b has landed (TMA completion)
    unsigned release[SD];                    // warps that have left the block in staging buffer b
    unsigned mhi[NEED_MHI ? 4 * NW * 32 : 1];
};

// One tile (the 32-pattern blocks [b0, b1), at most NW / CS of them) of one engine: the whole op list.
// CS = category split: CS warps share a block and take the categories cs, cs + CS, ... of a single-class model (the host picks
// CS <= categories, so every warp of an active block has work and meets its block's named barrier).
// seq: staging blocks consumed so far by this CTA (mbarrier phases; carried from tile to tile).
template <bool SAFE, bool MIX, int NW, int CS, int SD>
__device__ __forceinline__ void walk20v2_tile(const Dev &D, const double *stage, const int nops, const long long b0, const long long b1, double *s_dyn,
                                              V2Shared<NW, (MIX || CS > 1), SD> &sh, unsigned &seq) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, q = lane & 3;
    const int bs = warp / CS, cs = warp % CS;
    const int nblk = MIX ? D.nmix : 1, cpb = D.C / nblk;
    const int TS = tab20_stride(cpb);
    const size_t stage_sz = ss_stage_doubles(cpb, D.su);
    double *ring = s_dyn + SD * stage_sz + (size_t)warp * kRingK * kRingStage;
    const int nvo = nops * nblk;
    const int lo2 = lane * 2, lot = 128 + g * 4 + 2 * q, l16 = lane & 15;
    const int ncw = CS > 1 ? (cpb - cs + CS - 1) / CS : cpb;   // categories of this warp per class block
    const int S = 4 * ncw;                                       // steps per virtual op
    const long long mt_stride = (long long)D.C * kClv20Step;
    auto stage_issue = [&](int v, int ncase, unsigned sq) {   // one thread: staging block v -> buffer sq % SD
        const unsigned bytes = (unsigned)(ss_stage_used(cpb, D.su, ncase) * sizeof(double));
        mbar_expect(&sh.full[sq % SD], bytes);
        bulk_g2s(s_dyn + (size_t)(sq % SD) * stage_sz, stage + (size_t)v * stage_sz, bytes, &sh.full[sq % SD]);
    };
    const bool active = b0 + bs < b1 && ncw > 0;
    const long long p0 = (b0 + bs) * 32;
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
    int4 h_next = *reinterpret_cast<const int4 *>(stage);
    int4 h_next2 = nvo > 1 ? *reinterpret_cast<const int4 *>(stage + stage_sz) : h_next;
    if (tid == 0) {
        stage_issue(0, h_next.y, seq);
        if (nvo > 1) stage_issue(1, h_next2.y, seq + 1);
        if (SD > 2 && nvo > 2) stage_issue(2, reinterpret_cast<const int4 *>(stage + 2 * stage_sz)->y, seq + 2);
    }
    unsigned stA_next, stB_next;
    tip_states(h_next, stA_next, stB_next);

    // ---- the warp's CLV ring ------------------------------------------------------------------------------------------------
    // copy step s of a virtual op with header hd (class vclass) into ring stage k: every lane its chunks; ONE commit group per step
    auto ring_copy = [&](const int4 &hd, int vclass, int s, int k) {
        if (hd.y != 0) {
            const int j = s >> 2, mt = s & 3, c = (CS > 1 ? cs + j * CS : j) + vclass * cpb;
            const long long off = wbase + mt * mt_stride + (long long)c * kClv20Step + lo2;
            double *dst = ring + (size_t)k * kRingStage + lo2;
            if (hd.y == 2) {
                const double *src = D.partial + (long long)hd.z * D.slot_stride + off;
                cp_async16(dst, src);
                cp_async16(dst + 64, src + 64);
                if (lane < 16) cp_async16(dst + 128, src + 128);
            }
            const double *src = D.partial + (long long)hd.w * D.slot_stride + off;
            cp_async16(dst + kClv20Step, src);
            cp_async16(dst + kClv20Step + 64, src + 64);
            if (lane < 16) cp_async16(dst + kClv20Step + 128, src + 128);
        }
        cp_async_commit();
    };
    int iv = 0, is = 0, ik = 0;   // issue cursor: virtual op, step inside it, ring stage
    // the issue cursor is at most one virtual op ahead of the consumer (K <= S)
    auto issue_next = [&](const int4 &h_now, const int4 &h_nxt, int v_now) {
        if (iv < nvo) {
            ring_copy(iv == v_now ? h_now : h_nxt, MIX ? iv % nblk : 0, is, ik);
            if (++is == S) { is = 0; iv++; }
        } else cp_async_commit();   // keep one group per step so that wait_group<K-1> always means "this step has landed"
        ik = ik + 1 == kRingK ? 0 : ik + 1;
    };
    if (active) {
        __syncwarp();
        for (int k = 0; k < kRingK; k++) issue_next(h_next, h_next2, 0);
    }

    int m = 0, ck = 0;   // class of the current virtual op; consume cursor (ring stage)
    for (int v = 0; v < nvo; v++, seq++) {
        const double *sb = s_dyn + (size_t)(seq % SD) * stage_sz;
        const int4 h = h_next;
        const unsigned stA = stA_next, stB = stB_next;
        h_next = h_next2;
        if (v + 2 < nvo) h_next2 = *reinterpret_cast<const int4 *>(stage + (size_t)(v + 2) * stage_sz);
        mbar_wait(&sh.full[seq % SD], (seq / SD) & 1);   // staging block v has landed
        if (v + 1 < nvo && (!MIX || m + 1 == nblk)) tip_states(h_next, stA_next, stB_next);
        const int dst = h.x, ncase = h.y;
        if (active) {
            unsigned mhi[4] = {0u, 0u, 0u, 0u};
            unsigned scn[4] = {0u, 0u, 0u, 0u};
            bool rescaled = false;
            const int class_off = m * cpb * kClv20Step;
            double *Oq = D.partial + (long long)dst * D.slot_stride + wbase + class_off;
            const int sc_off = SAFE ? m * cpb * kWarp : 0;
            const uint16_t *Asc = D.scale + (long long)h.z * D.scale_stride + sbase + sc_off + g;
            const uint16_t *Bsc = D.scale + (long long)h.w * D.scale_stride + sbase + sc_off + g;
            uint16_t *Osc = D.scale + (long long)dst * D.scale_stride + sbase + sc_off + g;
            const double *sA = sb + 4, *sB = sA + ss_part_doubles(cpb, D.su, ncase <= 1);
            const double *inv_frag = D.inv_frag20 + (size_t)m * 480;
            if (!SAFE) {
#pragma unroll
                for (int mt = 0; mt < 4; mt++) {
                    if (ncase == 2) scn[mt] += Asc[8 * mt];
                    if (ncase >= 1) scn[mt] += Bsc[8 * mt];
                }
            }
            for (int j = 0; j < ncw; j++) {
                const int c = CS > 1 ? cs + j * CS : j;   // category inside the class block
                double PA[5][3], PB[5][3];
                if (ncase == 2) {
                    const double *mm = sA + (size_t)c * kMat20;
#pragma unroll
                    for (int kt = 0; kt < 5; kt++) {
                        const double2 t = *reinterpret_cast<const double2 *>(mm + kt * 64 + 2 * lane);
                        PA[kt][0] = t.x; PA[kt][1] = t.y;
                        const double u = mm[320 + kt * 16 + l16];
                        PA[kt][2] = lane < 16 ? u : 0.0;
                    }
                }
                if (ncase >= 1) {
                    const double *mm = sB + (size_t)c * kMat20;
#pragma unroll
                    for (int kt = 0; kt < 5; kt++) {
                        const double2 t = *reinterpret_cast<const double2 *>(mm + kt * 64 + 2 * lane);
                        PB[kt][0] = t.x; PB[kt][1] = t.y;
                        const double u = mm[320 + kt * 16 + l16];
                        PB[kt][2] = lane < 16 ? u : 0.0;
                    }
                }
#pragma unroll 1
                for (int mt = 0; mt < 4; mt++) {
                    double a[3][2], b[3][2];
                    const double *rs = ring + (size_t)ck * kRingStage;
                    cp_async_wait_group<kRingK - 1>();   // this lane's chunks of the step have landed ...
                    __syncwarp();                         // ... and the neighbour's tail chunk
                    if (ncase == 2) {
                        const double2 v0 = *reinterpret_cast<const double2 *>(rs + lo2), v1 = *reinterpret_cast<const double2 *>(rs + 64 + lo2);
                        const double f4 = rs[128 + lane];
#pragma unroll
                        for (int nt = 0; nt < 3; nt++) { a[nt][0] = 0.0; a[nt][1] = 0.0; }
#pragma unroll
                        for (int nt = 0; nt < 3; nt++) {
                            dmma_m8n8k4(a[nt][0], a[nt][1], v0.x, PA[0][nt]);
                            dmma_m8n8k4(a[nt][0], a[nt][1], v0.y, PA[1][nt]);
                            dmma_m8n8k4(a[nt][0], a[nt][1], v1.x, PA[2][nt]);
                            dmma_m8n8k4(a[nt][0], a[nt][1], v1.y, PA[3][nt]);
                            dmma_m8n8k4(a[nt][0], a[nt][1], f4, PA[4][nt]);
                        }
                    } else {
                        const double *tA = sA + (size_t)((stA >> (8 * mt)) & 0xffu) * TS + c * 20 + 2 * q;
#pragma unroll
                        for (int nt = 0; nt < 3; nt++) {
                            double2 t = make_double2(0.0, 0.0);
                            if (nt < 2 || q < 2) t = *reinterpret_cast<const double2 *>(tA + 8 * nt);
                            a[nt][0] = t.x; a[nt][1] = t.y;
                        }
                    }
                    if (ncase >= 1) {
                        const double *rb = rs + kClv20Step;
                        const double2 v0 = *reinterpret_cast<const double2 *>(rb + lo2), v1 = *reinterpret_cast<const double2 *>(rb + 64 + lo2);
                        const double f4 = rb[128 + lane];
#pragma unroll
                        for (int nt = 0; nt < 3; nt++) { b[nt][0] = 0.0; b[nt][1] = 0.0; }
#pragma unroll
                        for (int nt = 0; nt < 3; nt++) {
                            dmma_m8n8k4(b[nt][0], b[nt][1], v0.x, PB[0][nt]);
                            dmma_m8n8k4(b[nt][0], b[nt][1], v0.y, PB[1][nt]);
                            dmma_m8n8k4(b[nt][0], b[nt][1], v1.x, PB[2][nt]);
                            dmma_m8n8k4(b[nt][0], b[nt][1], v1.y, PB[3][nt]);
                            dmma_m8n8k4(b[nt][0], b[nt][1], f4, PB[4][nt]);
                        }
                    } else {
                        const double *tB = sB + (size_t)((stB >> (8 * mt)) & 0xffu) * TS + c * 20 + 2 * q;
#pragma unroll
                        for (int nt = 0; nt < 3; nt++) {
                            double2 t = make_double2(0.0, 0.0);
                            if (nt < 2 || q < 2) t = *reinterpret_cast<const double2 *>(tB + 8 * nt);
                            b[nt][0] = t.x; b[nt][1] = t.y;
                        }
                    }
                    double out[3][2];
#pragma unroll
                    for (int nt = 0; nt < 3; nt++) { out[nt][0] = a[nt][0] * b[nt][0]; out[nt][1] = a[nt][1] * b[nt][1]; }
                    unsigned chi = 0u;
#pragma unroll
                    for (int nt = 0; nt < 3; nt++) chi = max(chi, max(abs_hi(out[nt][0]), abs_hi(out[nt][1])));
                    double *po = Oq + mt * mt_stride + (long long)c * kClv20Step;
                    *reinterpret_cast<double2 *>(po + lo2) = make_double2(out[0][0], out[0][1]);
                    *reinterpret_cast<double2 *>(po + 64 + lo2) = make_double2(out[1][0], out[1][1]);
                    if (q < 2) *reinterpret_cast<double2 *>(po + lot) = make_double2(out[2][0], out[2][1]);
                    if (SAFE) {   // per (pattern, category) decision (tree/phylokernelnew.h:1994-2005)
                        unsigned sc = 0;
                        if (ncase == 2) sc += Asc[c * kWarp + 8 * mt];
                        if (ncase >= 1) sc += Bsc[c * kWarp + 8 * mt];
                        if (ncase != 0) {
                            chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 1));
                            chi = max(chi, __shfl_xor_sync(0xffffffffu, chi, 2));
                            const bool may = (may_mask >> mt) & 1u;
                            bool lo = may && chi < D.ss_lo;
                            const bool band = may && !lo && chi < D.ss_hi;
                            if (__any_sync(0xffffffffu, band)) {
                                __syncwarp();
                                const unsigned ce = eigen_max20(po, inv_frag, lane);
                                if (band) lo = ce < kThresholdHi;
                            }
                            if (lo) { rescale20(po, lo2, lot, q); sc += 1; }
                        }
                        if (q == 0) Osc[c * kWarp + 8 * mt] = (uint16_t)sc;
                    } else {
                        if (mt == 0) mhi[0] = max(mhi[0], chi);
                        else if (mt == 1) mhi[1] = max(mhi[1], chi);
                        else if (mt == 2) mhi[2] = max(mhi[2], chi);
                        else mhi[3] = max(mhi[3], chi);
                    }
                    // this stage is free again: request the step K ahead into it.  The stores above may be its source (the next
                    // op's child is usually this op's CLV): the tail chunk was stored by a neighbour lane.
                    __syncwarp();
                    issue_next(h, h_next, v);
                    ck = ck + 1 == kRingK ? 0 : ck + 1;
                }
            }
            // ---- NORM_LH: the pattern's decision (tree/phylokernelnew.h:2015-2029) ---------------------------------------
            if (!SAFE) {
                Walk20ssOp W;
                W.Oq = D.partial + (long long)dst * D.slot_stride + wbase;   // class 0: norm_decide20 walks the whole CLV
                W.may_mask = may_mask; W.Ctot = D.C; W.cpb = cpb; W.inv_frag = D.inv_frag20;
#pragma unroll
                for (int mt = 0; mt < 4; mt++) {
                    unsigned mm = mhi[mt];
                    mm = max(mm, __shfl_xor_sync(0xffffffffu, mm, 1));
                    mm = max(mm, __shfl_xor_sync(0xffffffffu, mm, 2));
                    mhi[mt] = mm;
                }
                if (CS > 1) {   // the maximum spans the categories of the block's CS warps
#pragma unroll
                    for (int mt = 0; mt < 4; mt++) sh.mhi[(warp * 4 + mt) * kWarp + lane] = mhi[mt];
                    named_bar_sync(1 + bs, CS * 32);   // also orders the CS warps' CLV stores for the exact path
                    bool lo4[4];
#pragma unroll
                    for (int mt = 0; mt < 4; mt++) {
                        unsigned mm = 0;
                        for (int w = 0; w < CS; w++) mm = max(mm, sh.mhi[((bs * CS + w) * 4 + mt) * kWarp + lane]);
                        lo4[mt] = ncase != 0 && norm_test20(D, W, mt, mm, lane);
                    }
                    named_bar_sync(1 + bs, CS * 32);   // every warp of the block has decided (the exact path reads all categories) before anyone rescales
#pragma unroll
                    for (int mt = 0; mt < 4; mt++) {
                        norm_apply20(W, mt, lo4[mt], lane, [&](int cc) { return cc % CS == cs; });
                        rescaled = rescaled || lo4[mt];
                        if (cs == 0 && q == 0) Osc[8 * mt] = (uint16_t)(scn[mt] + (lo4[mt] ? 1u : 0u));
                    }
                    named_bar_sync(1 + bs, CS * 32);   // scale counts and rescaled blocks are visible to the block's warps; sh.mhi may be rewritten
                } else {
#pragma unroll
                    for (int mt = 0; mt < 4; mt++) {
                        unsigned mm = mhi[mt];
                        if (MIX) {   // fold in the classes walked before this one (each lane owns its slot)
                            unsigned *slot = sh.mhi + (warp * 4 + mt) * kWarp + lane;
                            if (m != 0) mm = max(mm, *slot);
                            if (m + 1 != nblk) *slot = mm;
                        }
                        if (!MIX || m + 1 == nblk) {
                            unsigned s = scn[mt];
                            if (ncase != 0 && norm_decide20(D, W, mt, mm, lane, [](int) { return true; })) { s += 1; rescaled = true; }
                            if (q == 0) Osc[8 * mt] = (uint16_t)s;
                        }
                    }
                }
                // a rescale rewrote CLV blocks that the ring already holds for the next virtual op: drain it, take those steps again
                if (__any_sync(0xffffffffu, rescaled) && v + 1 < nvo) {
                    cp_async_wait_group<0>();
                    __syncwarp();
                    int k = ck;
                    for (int s2 = 0; s2 < kRingK; s2++) {   // steps 0 .. K-1 of virtual op v+1 (all of them were requested: K <= S)
                        ring_copy(h_next, MIX ? (v + 1) % nblk : 0, s2, k);
                        k = k + 1 == kRingK ? 0 : k + 1;
                    }
                }
            }
        }
        // leave staging block v: the last warp out requests block v+SD into its buffer
        __syncwarp();
        if (lane == 0) {
            const unsigned n = atomicAdd(&sh.release[seq % SD], 1u);
            if (n == NW - 1) {
                sh.release[seq % SD] = 0;
                if (v + SD < nvo) stage_issue(v + SD, SD == 2 ? h_next2.y : reinterpret_cast<const int4 *>(stage + (size_t)(v + SD) * stage_sz)->y, seq + SD);
            }
        }
        if (MIX && ++m == nblk) m = 0;
    }
    cp_async_wait_group<0>();   // (only empty groups are left)
}

template <bool SAFE, bool MIX, int NW, int CS, int SD>
__global__ void __launch_bounds__(NW * 32, 1) k_walk_n20v2(Dev D, const double *stage, int nops, unsigned long long *counter, int ntiles) {
    static_assert(!(CS > 1 && MIX), "category split serves single-class models");
    static_assert(NW % CS == 0, "warps per CTA must be a multiple of the category split");
    extern __shared__ __align__(16) double s_dyn[];
    __shared__ __align__(8) V2Shared<NW, (MIX || CS > 1), SD> sh;
    __shared__ long long s_item;
    if (threadIdx.x == 0)
        for (int k = 0; k < SD; k++) { mbar_init(&sh.full[k], 1); sh.release[k] = 0; }
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
        walk20v2_tile<SAFE, MIX, NW, CS, SD>(D, stage, nops, tile * nblocks / ntiles, (tile + 1) * nblocks / ntiles, s_dyn, sh, seq);
    }
}

}  // namespace b200lk
