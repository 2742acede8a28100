// b200lk_site.cuh — kernels for SITE-SPECIFIC models (model->isSiteSpecificModel(): ModelSet, i.e. `-fs` site-frequency files and
// the PMSF approximation of the C10–C60 mixtures): the device side of the SITE_MODEL template branch of tree/phylokernelnew.h
// (partial :1362-1374, 1653-1683, 1772-1801, 1937-1968; buffer :2129-2140; derivative :2306-2311, 2469-2487; branch :2708-2713,
// 2833-2848, 2993-3008; from-buffer :3271-3276, 3316-3327; instantiations tree/phylokernelfma.cpp:43-100).
//
// Every pattern p has its OWN eigensystem (eval_p, evec_p, inv_evec_p), so nothing can be staged per branch: the exponentials
// exp(eval_p[i] * rate_c * len) are taken per pattern, category and call, the tip vectors inv_evec_p . indicator(state)
// (computeTipPartialLikelihood's site branch, tree/phylotreesse.cpp:257-351) are gathered from inv_evec_p on the fly — the reference
// keeps a [taxon][pattern][state] table for them; here a tip costs one strided column of a matrix the thread streams anyway — and
// ncat_mix = ncat.  The per-pattern matrices live warp-interleaved like the CLVs (element k of pattern p at pbase(p, n^2) + 32 k):
// a warp's 32 patterns read 256 contiguous bytes per matrix element.  Thread per pattern; this path is bound by streaming 2 n^2
// doubles per pattern, category and node (13 x the CLV traffic at 20 states) — as the reference's is.
#pragma once
#include "b200lk_kernels.cuh"

namespace b200lk {

struct SiteModel {
    const double *eval;   // pbase(p, n)   + 32 i
    const double *evec;   // pbase(p, n^2) + 32 (x n + i)
    const double *inv;    // pbase(p, n^2) + 32 (i n + x)
};

// tip_partial_lh[taxon][p][.] of the reference, on the fly: inv_evec_p . indicator(state) (tree/phylotreesse.cpp:271-331)
__device__ __forceinline__ void site_tip(const Dev &D, const double *inv /* of pattern p */, int seq_type, int state, int n, double *out) {
    if (state < n) {
        for (int i = 0; i < n; i++) out[i] = inv[(long long)(i * n + state) * kWarp];
    } else if (state == D.su || (seq_type != 1 && seq_type != 2)) {
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            for (int x = 0; x < n; x++) s += inv[(long long)(i * n + x) * kWarp];
            out[i] = s;
        }
    } else if (seq_type == 1) {
        const int cstate = state - n + 1;
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            for (int x = 0; x < n; x++) if (cstate & (1 << x)) s += inv[(long long)(i * n + x) * kWarp];
            out[i] = s;
        }
    } else {
        const int ambi = state - n == 0 ? 4 + 8 : (state - n == 1 ? 32 + 64 : 512 + 1024);
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            for (int x = 0; x < 11; x++) if (ambi & (1 << x)) s += inv[(long long)(i * n + x) * kWarp];
            out[i] = s;
        }
    }
}

template <int NT, bool SAFE>
__device__ __forceinline__ void process_op_site(const Dev &D, const SiteModel &M, int seq_type, const DevOp &op, long long p) {
    constexpr int NMAX = NT ? NT : 64;
    const int n = NT ? NT : D.n;
    const int C = D.C;
    const ClvIdx at(D, p);
    const long long sb = SAFE ? pbase(p, C) : p;
    const double *eval = M.eval + pbase(p, n), *evec = M.evec + pbase(p, n * n), *inv = M.inv + pbase(p, n * n);
    double *out = D.partial + (long long)op.dst * D.slot_stride;
    uint16_t *osc = D.scale + (long long)op.dst * D.scale_stride + sb;
    const double *Lp = op.ncase == 2 ? D.partial + (long long)op.lid * D.slot_stride : nullptr;
    const double *Rp = op.ncase >= 1 ? D.partial + (long long)op.rid * D.slot_stride : nullptr;
    const uint16_t *lsc = Lp ? D.scale + (long long)op.lid * D.scale_stride + sb : nullptr;
    const uint16_t *rsc = Rp ? D.scale + (long long)op.rid * D.scale_stride + sb : nullptr;
    // DevOp.eL / eR carry the two child lengths (as doubles, bit-cast) for this path: nothing is staged per op
    const double lenL = resolve_len(D, __longlong_as_double(op.eL)), lenR = resolve_len(D, __longlong_as_double(op.eR));
    double tl[NMAX], tr[NMAX];   // tip vectors (eigen space); reused as the scaled child vectors
    double el[NMAX], er[NMAX], tmp[NMAX];
    if (!Lp) site_tip(D, inv, seq_type, D.tips[(long long)op.lid * D.Ppad + p], n, tl);
    if (!Rp) site_tip(D, inv, seq_type, D.tips[(long long)op.rid * D.Ppad + p], n, tr);
    const bool check = (op.ncase != 0) && (D.invar[p] == 0.0);
    double lh_max = 0.0;
    unsigned norm_scale = 0;
    if (!SAFE) norm_scale = (lsc ? lsc[0] : 0u) + (rsc ? rsc[0] : 0u);
    for (int c = 0; c < C; c++) {
        const double rl = D.cat_rate[c] * lenL, rr = D.cat_rate[c] * lenR;   // :1365-1366
#pragma unroll
        for (int i = 0; i < n; i++) {
            const double ev = eval[(long long)i * kWarp];
            el[i] = exp(ev * rl) * (Lp ? Lp[at(c, i)] : tl[i]);
            er[i] = exp(ev * rr) * (Rp ? Rp[at(c, i)] : tr[i]);
        }
#pragma unroll
        for (int x = 0; x < n; x++) {
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
            const double *row = evec + (long long)(x * n) * kWarp;
            int i = 0;
            for (; i + 1 < n; i += 2) {
                const double u0 = row[(long long)i * kWarp], u1 = row[(long long)(i + 1) * kWarp];
                a0 = fma(u0, el[i], a0); a1 = fma(u1, el[i + 1], a1);
                b0 = fma(u0, er[i], b0); b1 = fma(u1, er[i + 1], b1);
            }
            if (i < n) { const double u0 = row[(long long)i * kWarp]; a0 = fma(u0, el[i], a0); b0 = fma(u0, er[i], b0); }
            tmp[x] = (a0 + a1) * (b0 + b1);
        }
        double cmax = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) {
            double s0 = 0.0, s1 = 0.0;
            const double *row = inv + (long long)(i * n) * kWarp;
            int x = 0;
            for (; x + 1 < n; x += 2) { s0 = fma(row[(long long)x * kWarp], tmp[x], s0); s1 = fma(row[(long long)(x + 1) * kWarp], tmp[x + 1], s1); }
            if (x < n) s0 = fma(row[(long long)x * kWarp], tmp[x], s0);
            el[i] = s0 + s1;   // output staging
            cmax = fmax(cmax, fabs(el[i]));
        }
        if (SAFE) {
            unsigned sc = (lsc ? lsc[c * kWarp] : 0u) + (rsc ? rsc[c * kWarp] : 0u);
            if (check && cmax < kScalingThreshold) {
#pragma unroll
                for (int i = 0; i < n; i++) el[i] *= kTwo256;
                sc += 1;
            }
            osc[c * kWarp] = (uint16_t)sc;
        } else lh_max = fmax(lh_max, cmax);
#pragma unroll
        for (int i = 0; i < n; i++) out[at(c, i)] = el[i];
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
__global__ void __launch_bounds__(kTile) k_site_partials(Dev D, SiteModel M, int seq_type, const DevOp *ops, int nops, unsigned long long *counter) {
    __shared__ long long s_item;
    const long long ntiles = D.Ppad / kTile;
    for (;;) {   // (the host zeroes the counter before the launch)
        __syncthreads();
        if (threadIdx.x == 0) s_item = (long long)atomicAdd(counter, 1ULL);
        __syncthreads();
        const long long tile = s_item;
        if (tile >= ntiles) break;
        const long long p = tile * kTile + threadIdx.x;
        for (int o = 0; o < nops; o++) process_op_site<NT, SAFE>(D, M, seq_type, ops[o], p);
    }
}

// ------------------------------------------------------------------------------------------------------------
// k_site_walk20: the 20-state site-specific partial-likelihood kernel on the FP64-MMA pipe (PMSF on protein alignments — what
// site-specific models are used for).  A pattern's mat-vecs cannot share a matrix with other patterns, but they share it across the
// CATEGORIES and the TWO CHILDREN of a node: for one pattern and node, the child pass is evec_p [20 x 20] . E [20 x 8] with
// E[i][2c + child] = exp(eval_p[i] rate_c len_child) child[c][i] — 3 M-tiles x 5 K-steps = 15 DMMAs m8n8k4 — the product of the two
// children's columns is lane-local in the accumulators, and the back pass inv_p [20 x 20] . tmp [20 x 4] is 15 more (half of the N
// dimension idle).  A CTA owns one 32-pattern group for the whole post-order walk and keeps its patterns' matrices (evec, inv_evec,
// eval, row sums of inv_evec for gap tips: 841 doubles per pattern, 215 KB) in SHARED MEMORY: they are read from HBM once per
// evaluation instead of once per node and category (13 x the CLV traffic in k_site_partials).  A warp keeps its patterns through all
// ops, so a CLV written at one op is read back by the same warp at the parent (L2); per pattern and op a lane takes 5 exponentials
// per 4-category chunk.  Rows 20..23 of the padded M-tiles are zeros made in registers, never stored.
// ------------------------------------------------------------------------------------------------------------
constexpr int kSitePS = 841;   // evec [400] | inv_evec [400] | eval [20] | inv row sums [20] | 1 pad (odd stride: conflict-free staging)

__device__ __forceinline__ double site_tip20(const double *sm, int su, int seq_type, int state, int i) {
    if (state < 20) return sm[400 + i * 20 + state];
    if (state == su || seq_type != 2) return sm[820 + i];
    const int ambi = state - 20 == 0 ? 4 + 8 : (state - 20 == 1 ? 32 + 64 : 512 + 1024);
    double s = 0.0;
    for (int x = 0; x < 11; x++) if (ambi & (1 << x)) s += sm[400 + i * 20 + x];
    return s;
}

template <bool SAFE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_site_walk20(Dev D, SiteModel M, int seq_type, const DevOp *ops, int nops, unsigned long long *counter) {
    extern __shared__ double s_m[];   // [32 patterns][kSitePS]
    __shared__ long long s_item;
    constexpr int PPW = 32 / NW;   // patterns per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int C = D.C;
    const long long ngroups = D.Ppad / 32;
    for (;;) {   // (the host zeroes the counter before the launch)
        __syncthreads();
        if (threadIdx.x == 0) s_item = (long long)atomicAdd(counter, 1ULL);
        __syncthreads();
        const long long grp = s_item;
        if (grp >= ngroups) break;
        const long long p0 = grp * 32;
        {   // element k of the group's 32 patterns is one 256-byte line
            const double *ge = M.evec + pbase(p0, 400), *gi = M.inv + pbase(p0, 400), *gv = M.eval + pbase(p0, 20);
            for (int k = warp; k < 400; k += NW) {
                s_m[lane * kSitePS + k] = ge[(long long)k * kWarp + lane];
                s_m[lane * kSitePS + 400 + k] = gi[(long long)k * kWarp + lane];
            }
            for (int k = warp; k < 20; k += NW) s_m[lane * kSitePS + 800 + k] = gv[(long long)k * kWarp + lane];
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 32 * 20; t += NW * 32) {   // tip vector of the gap state: inv_evec_p . 1 (tree/phylotreesse.cpp:283-289)
            const int pp = t / 20, i = t - pp * 20;
            const double *row = s_m + pp * kSitePS + 400 + i * 20;
            double s = 0.0;
            for (int x = 0; x < 20; x++) s += row[x];
            s_m[pp * kSitePS + 820 + i] = s;
        }
        __syncthreads();
        // CLV element offsets of this lane for each of its patterns, fixed for the whole walk: the 5 values it loads per category
        // (i = 4 ks + q), the 3 it stores (i = 8 mt + g), the category stride
        int offL[PPW][5], offS[PPW][3], cstride;
        {
            const ClvIdx at0(D, p0);
            cstride = (int)(at0(1, 0) - at0(0, 0));
        }
#pragma unroll
        for (int pw = 0; pw < PPW; pw++) {
            const ClvIdx at(D, p0 + warp * PPW + pw);
            const long long b0 = at(0, 0);
#pragma unroll
            for (int ks = 0; ks < 5; ks++) offL[pw][ks] = (int)(at(0, 4 * ks + q) - b0);
#pragma unroll
            for (int mt = 0; mt < 3; mt++) offS[pw][mt] = 8 * mt + g < 20 ? (int)(at(0, 8 * mt + g) - b0) : 0;
        }
        // per pattern of this warp: may it rescale at all (ptn_invar == 0), and the tip states of the lane's child — loaded one op ahead
        bool may_scale[PPW];
#pragma unroll
        for (int pw = 0; pw < PPW; pw++) may_scale[pw] = D.invar[p0 + warp * PPW + pw] == 0.0;
        const int ch = g & 1, cc = g >> 1;   // the lane's column of E: child = g & 1 (0 = left), category cb + (g >> 1)
        auto tip_states = [&](const DevOp &op, int (&st)[PPW]) {
            const bool tip = op.ncase == 0 || (op.ncase == 1 && ch == 0);
            const int cid = ch ? op.rid : op.lid;
#pragma unroll
            for (int pw = 0; pw < PPW; pw++) st[pw] = tip ? (int)D.tips[(long long)cid * D.Ppad + p0 + warp * PPW + pw] : 0;
        };
        DevOp op_next = ops[0];
        int st_next[PPW];
        tip_states(op_next, st_next);
        for (int o = 0; o < nops; o++) {
            const DevOp op = op_next;
            int st_now[PPW];
#pragma unroll
            for (int pw = 0; pw < PPW; pw++) st_now[pw] = st_next[pw];
            if (o + 1 < nops) { op_next = ops[o + 1]; tip_states(op_next, st_next); }
            const bool tip = op.ncase == 0 || (op.ncase == 1 && ch == 0);
            const int cid = ch ? op.rid : op.lid;
            const double len = resolve_len(D, __longlong_as_double(ch ? op.eR : op.eL));
            const double *src = tip ? nullptr : D.partial + (long long)cid * D.slot_stride;
            double *out = D.partial + (long long)op.dst * D.slot_stride;
#pragma unroll
            for (int pw = 0; pw < PPW; pw++) {
                const int pl = warp * PPW + pw;
                const long long p = p0 + pl;
                const double *sm = s_m + pl * kSitePS;
                const long long pb = ClvIdx(D, p)(0, 0);
                const double *srcp = src ? src + pb : nullptr;
                double *outp = out + pb;
                const long long sb = SAFE ? pbase(p, C) : p;
                double tv[5];
                if (tip) {
                    const int state = st_now[pw];
#pragma unroll
                    for (int ks = 0; ks < 5; ks++) tv[ks] = site_tip20(sm, D.su, seq_type, state, 4 * ks + q);
                }
                const bool check = (op.ncase != 0) && may_scale[pw];
                const uint16_t *lsc = op.ncase == 2 ? D.scale + (long long)op.lid * D.scale_stride + sb : nullptr;
                const uint16_t *rsc = op.ncase >= 1 ? D.scale + (long long)op.rid * D.scale_stride + sb : nullptr;
                uint16_t *osc = D.scale + (long long)op.dst * D.scale_stride + sb;
                double lh_max = 0.0;
                for (int cb = 0; cb < C; cb += 4) {
                    const int c = cb + cc;
                    const bool cv = c < C;
                    const double rl = cv ? D.cat_rate[c] * len : 0.0;   // :1365-1366
                    double e[5];
#pragma unroll
                    for (int ks = 0; ks < 5; ks++) {
                        const int i = 4 * ks + q;
                        e[ks] = cv ? exp(sm[800 + i] * rl) * (srcp ? srcp[c * cstride + offL[pw][ks]] : tv[ks]) : 0.0;
                    }
                    double a[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                    for (int ks = 0; ks < 5; ks++)
#pragma unroll
                        for (int mt = 0; mt < 3; mt++) {
                            const int x = 8 * mt + g;
                            const double A = x < 20 ? sm[x * 20 + 4 * ks + q] : 0.0;
                            dmma_m8n8k4(a[mt][0], a[mt][1], A, e[ks]);
                        }
                    // a[mt] = (left, right) sums of state x = 8 mt + g, category cb + q: the product is lane-local
                    double t[3];
#pragma unroll
                    for (int mt = 0; mt < 3; mt++) t[mt] = a[mt][0] * a[mt][1];
                    double b[5];
#pragma unroll
                    for (int ks = 0; ks < 5; ks++) {   // B[k = q][col = g] = tmp[x = 4 ks + q][category cb + g], held by lane 4 (x % 8) + g in t[x / 8]
                        const double v = __shfl_sync(0xffffffffu, t[ks >> 1], 4 * ((4 * ks + q) & 7) + (g & 3));
                        b[ks] = g < 4 ? v : 0.0;
                    }
                    double r[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
                    for (int ks = 0; ks < 5; ks++)
#pragma unroll
                        for (int mt = 0; mt < 3; mt++) {
                            const int i = 8 * mt + g;
                            const double A = i < 20 ? sm[400 + i * 20 + 4 * ks + q] : 0.0;
                            dmma_m8n8k4(r[mt][0], r[mt][1], A, b[ks]);
                        }
                    // r[mt][j] = out[i = 8 mt + g][category cb + 2 q + j]  (q < 2)
                    double m0 = fmax(fmax(fabs(r[0][0]), fabs(r[1][0])), fabs(r[2][0]));
                    double m1 = fmax(fmax(fabs(r[0][1]), fabs(r[1][1])), fabs(r[2][1]));
                    const int c0 = cb + 2 * q, c1 = c0 + 1;
                    const bool v0 = q < 2 && c0 < C, v1 = q < 2 && c1 < C;
                    if (!v0) m0 = 0.0;
                    if (!v1) m1 = 0.0;
                    if (SAFE) {
#pragma unroll
                        for (int off = 4; off < 32; off <<= 1) {   // over the rows g of the same column
                            m0 = fmax(m0, __shfl_xor_sync(0xffffffffu, m0, off));
                            m1 = fmax(m1, __shfl_xor_sync(0xffffffffu, m1, off));
                        }
                        unsigned s0 = 0, s1 = 0;
                        if (v0) s0 = (lsc ? lsc[c0 * kWarp] : 0u) + (rsc ? rsc[c0 * kWarp] : 0u);
                        if (v1) s1 = (lsc ? lsc[c1 * kWarp] : 0u) + (rsc ? rsc[c1 * kWarp] : 0u);
                        if (check && v0 && m0 < kScalingThreshold) { r[0][0] *= kTwo256; r[1][0] *= kTwo256; r[2][0] *= kTwo256; s0 += 1; }
                        if (check && v1 && m1 < kScalingThreshold) { r[0][1] *= kTwo256; r[1][1] *= kTwo256; r[2][1] *= kTwo256; s1 += 1; }
                        if (g == 0 && v0) osc[c0 * kWarp] = (uint16_t)s0;
                        if (g == 0 && v1) osc[c1 * kWarp] = (uint16_t)s1;
                    } else lh_max = fmax(lh_max, fmax(m0, m1));
#pragma unroll
                    for (int mt = 0; mt < 3; mt++) {
                        const int i = 8 * mt + g;
                        if (i < 20) {
                            if (v0) outp[c0 * cstride + offS[pw][mt]] = r[mt][0];
                            if (v1) outp[c1 * cstride + offS[pw][mt]] = r[mt][1];
                        }
                    }
                }
                if (!SAFE) {
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) lh_max = fmax(lh_max, __shfl_xor_sync(0xffffffffu, lh_max, off));
                    unsigned sc = (lsc ? lsc[0] : 0u) + (rsc ? rsc[0] : 0u);
                    if (check && lh_max < kScalingThreshold) {   // every lane re-scales what it stored itself
                        for (int cb = 0; cb < C; cb += 4) {
                            const int c0 = cb + 2 * q, c1 = c0 + 1;
#pragma unroll
                            for (int mt = 0; mt < 3; mt++) {
                                const int i = 8 * mt + g;
                                if (i < 20 && q < 2) {
                                    if (c0 < C) outp[c0 * cstride + offS[pw][mt]] *= kTwo256;
                                    if (c1 < C) outp[c1 * cstride + offS[pw][mt]] *= kTwo256;
                                }
                            }
                        }
                        sc += 1;
                    }
                    if (lane == 0) osc[0] = (uint16_t)sc;
                }
                __syncwarp();   // the parent op reads these values through other lanes of this warp
            }
        }
    }
}

// computeLikelihoodBranch, SITE_MODEL (:2833-2848 leaf at the far end, :2993-3008 two CLVs): lh_cat[c] = prop_c * sum_i
// exp(eval_p[i] rate_c len) node[c][i] dad[c][i] (dotProductExp), then the normal model's scaling alignment and logarithm
template <bool SAFE>
__global__ void __launch_bounds__(kTile) k_site_root(Dev D, SiteModel M, int seq_type, int node_slot, int dad_kind, int dad_id, double length, int save_log,
                                                     double *red, unsigned *ticket, double *result) {
    const int n = D.n, C = D.C;
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    const long long cb = pbase(p, C);
    const long long sb = SAFE ? cb : p;
    const ClvIdx at(D, p);
    const double *eval = M.eval + pbase(p, n);
    const double *dad = D.partial + (long long)node_slot * D.slot_stride;
    const uint16_t *dsc = D.scale + (long long)node_slot * D.scale_stride + sb;
    const double *node = nullptr;
    const uint16_t *nsc = nullptr;
    double tip[64];
    if (dad_kind == 1) {
        node = D.partial + (long long)dad_id * D.slot_stride;
        nsc = D.scale + (long long)dad_id * D.scale_stride + sb;
    } else site_tip(D, M.inv + pbase(p, n * n), seq_type, D.tips[(long long)dad_id * D.Ppad + p], n, tip);
    const double len = resolve_len(D, length);
    double *lh_cat = D.pattern_lh_cat + cb;
    double lh_ptn = 0.0;
    unsigned min_scale = 0xffffffffu;
    for (int c = 0; c < C; c++) {
        const double t = D.cat_rate[c] * len;
        double s = 0.0;
        for (int i = 0; i < n; i++) {
            const double f = (node ? node[at(c, i)] : tip[i]) * dad[at(c, i)];
            const double ex = exp(eval[(long long)i * kWarp] * t);
            s = i == 0 ? ex * f : fma(ex, f, s);
        }
        s *= D.cat_prop[c];
        lh_cat[c * kWarp] = s;
        if (SAFE) min_scale = min(min_scale, (unsigned)dsc[c * kWarp] + (nsc ? nsc[c * kWarp] : 0u));
        else lh_ptn += s;
    }
    double vc_min_scale;
    if (SAFE) {
        double v0 = 0, v1 = 0, v2 = 0, v3 = 0, tail = 0;
        const int c4 = C & ~3;
        for (int c = 0; c < C; c++) {
            const unsigned sc = dsc[c * kWarp] + (nsc ? nsc[c * kWarp] : 0u);
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
    } else vc_min_scale = (double)(dsc[0] + (nsc ? nsc[0] : 0u));
    vc_min_scale *= kLogScalingThreshold;
    lh_ptn = fabs(lh_ptn) + D.invar[p];
    const double lg = log(lh_ptn) + vc_min_scale;
    double acc[1] = {0.0};
    if (p < D.P) {
        if (save_log) D.pattern_lh[p] = lg;
        else { D.pattern_lh[p] = lh_ptn; D.pattern_scaling[p] = vc_min_scale; }
        acc[0] = lg * D.freq[p];
    }
    block_reduce_finish<1>(acc, red, ticket, result);
}

// computeLikelihoodBuffer with per-pattern tip vectors (:2129-2140)
template <bool SAFE>
__global__ void __launch_bounds__(kTile) k_site_theta(Dev D, SiteModel M, int seq_type, int node_slot, int dad_kind, int dad_id) {
    const int n = D.n, C = D.C, B = D.B;
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    const long long sb = SAFE ? pbase(p, C) : p;
    const ClvIdx at(D, p);
    const double *dad = D.partial + (long long)node_slot * D.slot_stride;
    const uint16_t *dsc = D.scale + (long long)node_slot * D.scale_stride + sb;
    const double *node = nullptr;
    const uint16_t *nsc = nullptr;
    double tip[64];
    if (dad_kind == 1) {
        node = D.partial + (long long)dad_id * D.slot_stride;
        nsc = D.scale + (long long)dad_id * D.scale_stride + sb;
    } else site_tip(D, M.inv + pbase(p, n * n), seq_type, D.tips[(long long)dad_id * D.Ppad + p], n, tip);
    double *theta = D.theta + pbase(p, B);
    unsigned min_scale = 0xffffffffu;
    if (SAFE)
        for (int c = 0; c < C; c++) min_scale = min(min_scale, dsc[c * kWarp] + (nsc ? nsc[c * kWarp] : 0u));
    for (int c = 0; c < C; c++) {
        double f = 1.0;
        bool zero = false;
        if (SAFE) {
            const unsigned sc = dsc[c * kWarp] + (nsc ? nsc[c * kWarp] : 0u);
            if (sc == min_scale + 1) f = kScalingThreshold;
            else if (sc > min_scale + 1) zero = true;
        }
        for (int i = 0; i < n; i++) {
            double t = (node ? node[at(c, i)] : tip[i]) * dad[at(c, i)];
            if (SAFE) t = zero ? 0.0 : (f != 1.0 ? t * f : t);
            theta[(long long)(c * n + i) * kWarp] = t;
        }
    }
    const double bs = SAFE ? (double)min_scale : (double)(dsc[0] + (nsc ? nsc[0] : 0u));
    D.buffer_scale[p] = bs * kLogScalingThreshold;
}

// computeLikelihoodDerv, SITE_MODEL (:2469-2487)
__global__ void __launch_bounds__(kTile) k_site_derv(Dev D, SiteModel M, double length, double *red, unsigned *ticket, double *result) {
    const int n = D.n, C = D.C, B = D.B;
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    const double *eval = M.eval + pbase(p, n), *theta = D.theta + pbase(p, B);
    const double len = resolve_len(D, length);
    double lh = 0.0, df = 0.0, ddf = 0.0;
    for (int c = 0; c < C; c++) {
        double lc = 0.0, dc = 0.0, ddc = 0.0;
        const double rate = D.cat_rate[c];
        for (int i = 0; i < n; i++) {
            const double cof = eval[(long long)i * kWarp] * rate;
            const double val = exp(cof * len) * theta[(long long)(c * n + i) * kWarp];
            const double val1 = cof * val;
            lc += val; dc += val1; ddc = fma(cof, val1, ddc);
        }
        const double prop = D.cat_prop[c];
        lh = fma(prop, lc, lh); df = fma(prop, dc, df); ddf = fma(prop, ddc, ddf);
    }
    lh = fabs(lh) + D.invar[p];
    double acc[2] = {0.0, 0.0};
    if (p < D.P) {
        const double inv = 1.0 / lh, df_frac = df * inv, ddf_frac = ddf * inv, f = D.freq[p];
        const double tmp1 = df_frac * f, tmp2 = ddf_frac * f;
        acc[0] = tmp1;
        acc[1] = fma(-tmp1, df_frac, tmp2);
    }
    block_reduce_finish<2>(acc, red, ticket, result);
}

// computeLikelihoodFromBuffer, SITE_MODEL (:3316-3327)
__global__ void __launch_bounds__(kTile) k_site_from_buffer(Dev D, SiteModel M, double length, double *red, unsigned *ticket, double *result) {
    const int n = D.n, C = D.C, B = D.B;
    const long long p = (long long)blockIdx.x * kTile + threadIdx.x;
    const double *eval = M.eval + pbase(p, n), *theta = D.theta + pbase(p, B);
    const double len = resolve_len(D, length);
    double lh = 0.0;
    for (int c = 0; c < C; c++) {
        const double t = D.cat_rate[c] * len;
        double s = 0.0;
        for (int i = 0; i < n; i++) {
            const double ex = exp(eval[(long long)i * kWarp] * t), th = theta[(long long)(c * n + i) * kWarp];
            s = i == 0 ? ex * th : fma(ex, th, s);
        }
        lh = fma(s, D.cat_prop[c], lh);
    }
    lh = fabs(lh) + D.invar[p];
    const double lg = log(fabs(lh)) + D.buffer_scale[p];
    double acc[1] = {0.0};
    if (p < D.P) {
        D.pattern_lh[p] = lg;
        acc[0] = lg * D.freq[p];
    }
    block_reduce_finish<1>(acc, red, ticket, result);
}

}  // namespace b200lk
