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
