/*
 * lk_oracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, scalar, V=1 layout) of the
 * reference's reversible-model likelihood kernel, IQ-TREE 2.4.0 tree/phylokernelnew.h.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  The product path (iqtree2_b200/csrc, libb200lk.so) never links or calls it.
 *
 * Pinning status: the reference ships NO golden vectors / known-answer tests for this path
 * (SURVEY §4, §8c).  This restatement is pinned (a) against an independent brute-force Felsenstein
 * pruning with P(t)=expm(Qrt) (tests/test_oracle.py, "Route C") and (b) against the REAL reference
 * kernel templates compiled from /root/reference into oracle/_ref (oracle/build_ref.sh, "Route A")
 * when that build is available; see DESIGN.md §Oracle for which of the two held at commit time.
 *
 * Each function cites the reference file:line it follows.  All paths are relative to /root/reference.
 * Summation orders follow the reference helper templates (phylokernelnew.h:82-842) lane-for-lane, so
 * oracle-vs-reference noise is at the last-ulp level; the product is only required to agree to 1e-11
 * relative per pattern (BASELINE.json north_star) and bit-exactly on scale_num.
 *
 * Buffer layouts are the reference's with SIMD width V=1 (SURVEY §8a "Layout crib"):
 *   partial_lh / theta : [p*B + c*n + i]           B = n*ncat_mix
 *   scale_num NORM_LH  : [p]   (uint16)            SAFE_LH: [p*C + c]
 *   echildren          : [c*n*n + x*n + i]         partial_lh_leaves: [state*B + c*n + x]
 *   tip_partial_lh     : [state*nmix*n + m*n + i]
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* tree/phylotree.h:74-80 */
#define SCALING_THRESHOLD 8.636168555094444625386e-78
#define SCALING_THRESHOLD_EXP 256
#define LOG_SCALING_THRESHOLD (-177.4456782233459932741)

typedef struct {
    int32_t nstates, ncat, nmix, fused_mix_rate, safe_numeric;
    int32_t seq_type;      /* 0 generic, 1 DNA, 2 protein */
    int32_t state_unknown; /* aln->STATE_UNKNOWN */
    int32_t ntaxa;
    int64_t nptn;
    int32_t nslots;
    int32_t device; /* unused here; keeps the struct identical to b200lk_config */
} orc_config;

typedef struct { int32_t kind; int32_t id; double length; } orc_child; /* kind 0 = tip (id = taxon), 1 = slot */
/* nextra > 0: a multifurcating node (node->degree() > 3, phylokernelnew.h:1407-1631): its third, fourth ... children in
 * neighbour order are taken, in op order, from the array handed to orc_set_extra_children before the call */
typedef struct { int32_t dst_slot; int32_t nextra; orc_child left, right; } orc_op;

typedef struct {
    orc_config cfg;
    int C, B, denom;
    uint8_t *tips;    /* [ntaxa][nptn] */
    double *ptn_freq, *ptn_invar;
    double *eval, *evec, *inv_evec, *mix_weight, *cat_rate, *cat_prop;
    double *tip_partial_lh; /* [(SU+1)][nmix][n] */
    double *partial;        /* [nslots][nptn*B] */
    uint16_t *scale;        /* [nslots][nptn*S] */
    double *theta, *buffer_scale_all, *pattern_lh, *pattern_lh_cat, *pattern_scaling;
    int theta_computed;
    int quirks; /* orc_set_reference_quirks */
    double robust_keep; int robust_median; /* orc_set_robust_phylo: Params::robust_phy_keep (on if 0 < keep < 1) / robust_median */
    /* +ASC, Lewis' correction (phylokernelnew.h:3076-3091, 3198-3221, 2506-2535, 2618-2626, 3342-3356, 3412-3426): the last
     * asc_nunobs patterns are model_factory->unobserved_ptns; asc_nsite = aln->getNSite() */
    int asc_type; int64_t asc_nunobs, asc_nsite;
    orc_child *extra; /* pending extra children of multifurcating ops (consumed by the next call that takes ops) */
    int nextra;
    /* non-reversible kernel (lk_oracle_nonrev.h): complex eigensystem [nmix][n](re,im) / [nmix][n][n](re,im), rate matrices, frequencies */
    int nonrev, rooted;
    double *ceval, *cevec, *cinv, *qmat, *sfreq;
    /* site-specific models (lk_oracle_site.h): one eigensystem per pattern, [p][n] / [p][n][n] */
    int site;
    double *s_eval, *s_evec, *s_inv;
    /* mixed branch lengths / heterotachy (lk_oracle_mixlen.h): nmixlen = ncat lengths per branch, [rows][ncat]; a `length` argument
     * is then the ROW of the branch in this table */
    int mixlen, mixlen_rows;
    double *mixlen_tab;
    /* end points of the branch the theta buffer was built for are irrelevant after the build */
    char err[256];
} orc_engine;

/* child->getLength(cat_id[c]) / dad_branch->getLength(mycat) (phylokernelnew.h:1017, 1069, 2325, 2348, 2724, 2738, 3286, 3301) */
static inline double len_of(const orc_engine *e, double length, int r) {
    return e->mixlen ? e->mixlen_tab[(size_t)(int)length * e->cfg.ncat + r] : length;
}

static size_t S_of(const orc_engine *e) { return e->cfg.safe_numeric ? (size_t)e->C : 1; }

/* ---------------------------------------------------------------------------------------------
 * helper dot products — phylokernelnew.h:82-842 (scalar lanes of the VectorClass templates)
 * ------------------------------------------------------------------------------------------- */

/* dotProductVec, JAMES_VERSION branch — phylokernelnew.h:89-126 */
static inline double dot_vec(const double *A, const double *B, size_t N) {
    if (N == 4) return (A[0] * B[0] + A[1] * B[1]) + (A[2] * B[2] + A[3] * B[3]);
    size_t cruft = N & 3;
    if (N == cruft) {
        double X = 0;
        for (size_t i = 0; i < cruft; i++) X = fma(A[i], B[i], X);
        return X;
    }
    double X = A[0] * B[0], v1 = A[1] * B[1], v2 = A[2] * B[2], v3 = A[3] * B[3];
    size_t stop = N - cruft, k = 4;
    for (; k < stop; k += 4) {
        X = fma(A[k], B[k], X);
        v1 = fma(A[k + 1], B[k + 1], v1);
        v2 = fma(A[k + 2], B[k + 2], v2);
        v3 = fma(A[k + 3], B[k + 3], v3);
    }
    if (cruft > 0) {
        X = fma(A[k], B[k], X);
        if (cruft > 1) {
            v1 = fma(A[k + 1], B[k + 1], v1);
            if (cruft > 2) v2 = fma(A[k + 2], B[k + 2], v2);
        }
    }
    return (X + v1) + (v2 + v3);
}

/* dotProductDualVec — phylokernelnew.h:214-314 : (A.B)*(C.D) */
static inline double dot_dual(const double *A, const double *B, const double *C, const double *D, size_t N) {
    switch (N) {
    case 1: return (A[0] * B[0]) * (C[0] * D[0]);
    case 2: return fma(A[1], B[1], A[0] * B[0]) * fma(C[1], D[1], C[0] * D[0]);
    case 3: return fma(A[2], B[2], fma(A[1], B[1], A[0] * B[0])) * fma(C[2], D[2], fma(C[1], D[1], C[0] * D[0]));
    case 4: return (fma(A[1], B[1], A[0] * B[0]) + fma(A[3], B[3], A[2] * B[2])) *
                   (fma(C[1], D[1], C[0] * D[0]) + fma(C[3], D[3], C[2] * D[2]));
    default: break;
    }
    size_t i, j;
    switch (N % 4) {
    case 0: {
        double AB[4], CD[4];
        for (j = 0; j < 4; j++) { AB[j] = A[j] * B[j]; CD[j] = C[j] * D[j]; }
        for (i = 4; i < N; i += 4)
            for (j = 0; j < 4; j++) { AB[j] = fma(A[i + j], B[i + j], AB[j]); CD[j] = fma(C[i + j], D[i + j], CD[j]); }
        return ((AB[0] + AB[1]) + (AB[2] + AB[3])) * ((CD[0] + CD[1]) + CD[2] + CD[3]);
    }
    case 1: {
        size_t Nm1 = N - 1;
        double AB[4], CD[4];
        for (j = 0; j < 4; j++) { AB[j] = A[j] * B[j]; CD[j] = C[j] * D[j]; }
        for (i = 4; i < Nm1; i += 4)
            for (j = 0; j < 4; j++) { AB[j] = fma(A[i + j], B[i + j], AB[j]); CD[j] = fma(C[i + j], D[i + j], CD[j]); }
        return fma(A[Nm1], B[Nm1], (AB[0] + AB[1]) + (AB[2] + AB[3])) * fma(C[Nm1], D[Nm1], (CD[0] + CD[1]) + CD[2] + CD[3]);
    }
    case 2: {
        double AB[2], CD[2];
        for (j = 0; j < 2; j++) { AB[j] = A[j] * B[j]; CD[j] = C[j] * D[j]; }
        for (i = 2; i < N; i += 2)
            for (j = 0; j < 2; j++) { AB[j] = fma(A[i + j], B[i + j], AB[j]); CD[j] = fma(C[i + j], D[i + j], CD[j]); }
        return (AB[0] + AB[1]) * (CD[0] + CD[1]);
    }
    default: {
        double AB[2], CD[2];
        for (j = 0; j < 2; j++) { AB[j] = A[j] * B[j]; CD[j] = C[j] * D[j]; }
        for (i = 2; i < N - 1; i += 2)
            for (j = 0; j < 2; j++) { AB[j] = fma(A[i + j], B[i + j], AB[j]); CD[j] = fma(C[i + j], D[i + j], CD[j]); }
        AB[0] = fma(A[N - 1], B[N - 1], AB[0] + AB[1]);
        CD[0] = fma(C[N - 1], D[N - 1], CD[0] + CD[1]);
        return AB[0] * CD[0];
    }
    }
}

/* one row of productVecMat — phylokernelnew.h:326-567: X[i] = sum_x A[x]*M[i*N+x] */
static inline double pvm_row(const double *A, const double *M, size_t N) {
    size_t x, j;
    if (N == 4) return fma(A[1], M[1], A[0] * M[0]) + fma(A[3], M[3], A[2] * M[2]);
    switch (N) {
    case 1: return A[0] * M[0];
    case 2: return fma(A[1], M[1], A[0] * M[0]);
    case 3: return fma(A[2], M[2], fma(A[1], M[1], A[0] * M[0]));
    default: break;
    }
    switch (N % 4) {
    case 0: {
        double V[4];
        for (j = 0; j < 4; j++) V[j] = A[j] * M[j];
        for (x = 4; x < N; x += 4)
            for (j = 0; j < 4; j++) V[j] = fma(A[x + j], M[x + j], V[j]);
        return (V[0] + V[1]) + (V[2] + V[3]);
    }
    case 1: {
        size_t Nm1 = N - 1;
        double V[4];
        for (j = 0; j < 4; j++) V[j] = A[j] * M[j];
        for (x = 4; x < Nm1; x += 4)
            for (j = 0; j < 4; j++) V[j] = fma(A[x + j], M[x + j], V[j]);
        return fma(A[Nm1], M[Nm1], (V[0] + V[1]) + (V[2] + V[3]));
    }
    case 2: {
        double V[2];
        for (j = 0; j < 2; j++) V[j] = A[j] * M[j];
        for (x = 2; x < N; x += 2)
            for (j = 0; j < 2; j++) V[j] = fma(A[x + j], M[x + j], V[j]);
        return V[0] + V[1];
    }
    default: {
        double V[2];
        for (j = 0; j < 2; j++) V[j] = A[j] * M[j];
        for (x = 2; x < N - 1; x += 2)
            for (j = 0; j < 2; j++) V[j] = fma(A[x + j], M[x + j], V[j]);
        return fma(A[N - 1], M[N - 1], V[0] + V[1]);
    }
    }
}

/* productVecMat with Xmax — phylokernelnew.h:450-567 */
static inline void prod_vec_mat(const double *A, const double *M, double *X, double *Xmax, size_t N) {
    for (size_t i = 0; i < N; i++) {
        X[i] = pvm_row(A, M + i * N, N);
        if (Xmax) { double a = fabs(X[i]); if (a > *Xmax) *Xmax = a; }
    }
}

/* dotProduct3Vec — phylokernelnew.h:737-793 : sum A*B*C */
static inline double dot3(const double *A, const double *B, const double *C, size_t N) {
    size_t i, j;
    switch (N % 4) {
    case 0: {
        double V[4];
        for (j = 0; j < 4; j++) V[j] = A[j] * B[j] * C[j];
        for (i = 4; i < N; i += 4)
            for (j = 0; j < 4; j++) V[j] = fma(A[i + j] * B[i + j], C[i + j], V[j]);
        return (V[0] + V[1]) + (V[2] + V[3]);
    }
    case 1: {
        double V[4];
        size_t Nm1 = N - 1;
        for (j = 0; j < 4; j++) V[j] = A[j] * B[j] * C[j];
        for (i = 4; i < Nm1; i += 4)
            for (j = 0; j < 4; j++) V[j] = fma(A[i + j] * B[i + j], C[i + j], V[j]);
        return fma(A[Nm1] * B[Nm1], C[Nm1], (V[0] + V[1]) + (V[2] + V[3]));
    }
    case 2: {
        double V[2];
        for (j = 0; j < 2; j++) V[j] = A[j] * B[j] * C[j];
        for (i = 2; i < N; i += 2)
            for (j = 0; j < 2; j++) V[j] = fma(A[i + j] * B[i + j], C[i + j], V[j]);
        return V[0] + V[1];
    }
    default: {
        double V[2];
        for (j = 0; j < 2; j++) V[j] = A[j] * B[j] * C[j];
        for (i = 2; i < N - 1; i += 2)
            for (j = 0; j < 2; j++) V[j] = fma(A[i + j] * B[i + j], C[i + j], V[j]);
        return fma(A[N - 1] * B[N - 1], C[N - 1], V[0] + V[1]);
    }
    }
}

/* dotProductTriple<ADD=false> — phylokernelnew.h:666-725.  NOTE: the reference picks its odd-length branch on
 * `nstates & 1`, not on N (= block = nstates*ncat_mix).  With odd nstates and an EVEN block (e.g. 61-state codon
 * x 4 categories, 5 states x 6) the loop `for (i = 2; i < N-1; i += 2)` visits index N-1 and the trailing mul_add
 * adds it again: the last element is counted twice (verified against the compiled reference, oracle/_ref).
 * `odd` selects the branch: the caller passes nstates&1 to reproduce the reference bit for bit (quirk bit 0,
 * orc_set_reference_quirks) or N&1 for the intended mathematics (every element once). */
static inline void dot_triple(const double *A, const double *B, const double *C, const double *D, size_t N, int odd,
                              double *X, double *Y, double *Z) {
    double AD[2], BD[2], CD[2];
    size_t i, j;
    for (j = 0; j < 2; j++) { AD[j] = A[j] * D[j]; BD[j] = B[j] * D[j]; CD[j] = C[j] * D[j]; }
    size_t stop = odd ? N - 1 : N;
    for (i = 2; i < stop; i += 2)
        for (j = 0; j < 2; j++) {
            AD[j] = fma(A[i + j], D[i + j], AD[j]);
            BD[j] = fma(B[i + j], D[i + j], BD[j]);
            CD[j] = fma(C[i + j], D[i + j], CD[j]);
        }
    if (!odd) {
        *X = AD[0] + AD[1]; *Y = BD[0] + BD[1]; *Z = CD[0] + CD[1];
    } else {
        *X = fma(A[N - 1], D[N - 1], AD[0] + AD[1]);
        *Y = fma(B[N - 1], D[N - 1], BD[0] + BD[1]);
        *Z = fma(C[N - 1], D[N - 1], CD[0] + CD[1]);
    }
}

/* sumVec<append=true> — phylokernelnew.h:43-69 */
static inline double sum_vec_append(const double *A, double X, size_t N) {
    size_t cruft = N & 3;
    if (cruft < N) {
        double V[4] = {0, 0, 0, 0};
        size_t k = 0;
        for (; k < N - cruft; k += 4) { V[0] += A[k]; V[1] += A[k + 1]; V[2] += A[k + 2]; V[3] += A[k + 3]; }
        V[0] += V[2]; V[1] += V[3]; V[0] += V[1];
        X += V[0];
        if (cruft == 0) return X;
        A += k;
    }
    for (size_t k = 0; k < cruft; k++) X += A[k];
    return X;
}

/* ---------------------------------------------------------------------------------------------
 * engine life-cycle
 * ------------------------------------------------------------------------------------------- */
orc_engine *orc_create(const orc_config *cfg) {
    orc_engine *e = (orc_engine *)calloc(1, sizeof(orc_engine));
    e->cfg = *cfg;
    int n = cfg->nstates;
    e->C = cfg->fused_mix_rate ? cfg->ncat : cfg->ncat * cfg->nmix; /* phylokernelnew.h:1322-1323 */
    e->denom = cfg->fused_mix_rate ? 1 : cfg->ncat;                 /* :1325 */
    e->B = n * e->C;
    size_t P = (size_t)cfg->nptn;
    e->tips = (uint8_t *)calloc((size_t)cfg->ntaxa * P + 1, 1);
    e->ptn_freq = (double *)calloc(P + 1, sizeof(double));
    e->ptn_invar = (double *)calloc(P + 1, sizeof(double));
    e->eval = (double *)calloc((size_t)cfg->nmix * n, sizeof(double));
    e->evec = (double *)calloc((size_t)cfg->nmix * n * n, sizeof(double));
    e->inv_evec = (double *)calloc((size_t)cfg->nmix * n * n, sizeof(double));
    e->mix_weight = (double *)calloc(cfg->nmix, sizeof(double));
    e->cat_rate = (double *)calloc(cfg->ncat, sizeof(double));
    e->cat_prop = (double *)calloc(cfg->ncat, sizeof(double));
    e->tip_partial_lh = (double *)calloc((size_t)(cfg->state_unknown + 1) * cfg->nmix * n, sizeof(double));
    e->partial = (double *)calloc((size_t)cfg->nslots * P * e->B + 1, sizeof(double));
    e->scale = (uint16_t *)calloc((size_t)cfg->nslots * P * S_of(e) + 1, sizeof(uint16_t));
    e->theta = (double *)calloc(P * e->B + 1, sizeof(double));
    e->buffer_scale_all = (double *)calloc(P + 1, sizeof(double));
    e->pattern_lh = (double *)calloc(P + 1, sizeof(double));
    e->pattern_scaling = (double *)calloc(P + 1, sizeof(double));
    e->pattern_lh_cat = (double *)calloc(P * e->C + 1, sizeof(double));
    if (!e->partial || !e->scale || !e->theta) { free(e); return NULL; }
    return e;
}

void orc_destroy(orc_engine *e) {
    if (!e) return;
    free(e->tips); free(e->ptn_freq); free(e->ptn_invar); free(e->eval); free(e->evec); free(e->inv_evec);
    free(e->mix_weight); free(e->cat_rate); free(e->cat_prop); free(e->tip_partial_lh); free(e->partial);
    free(e->scale); free(e->theta); free(e->buffer_scale_all); free(e->pattern_lh); free(e->pattern_scaling);
    free(e->pattern_lh_cat); free(e->extra);
    free(e->ceval); free(e->cevec); free(e->cinv); free(e->qmat); free(e->sfreq);
    free(e->s_eval); free(e->s_evec); free(e->s_inv); free(e->mixlen_tab);
    free(e);
}

const char *orc_last_error(orc_engine *e) { return e->err; }

int orc_set_tip_states(orc_engine *e, const uint8_t *states) {
    memcpy(e->tips, states, (size_t)e->cfg.ntaxa * e->cfg.nptn);
    return 0;
}
int orc_set_pattern_freq(orc_engine *e, const double *f) { memcpy(e->ptn_freq, f, sizeof(double) * e->cfg.nptn); return 0; }
int orc_set_pattern_invar(orc_engine *e, const double *f) { memcpy(e->ptn_invar, f, sizeof(double) * e->cfg.nptn); return 0; }
int orc_set_rates(orc_engine *e, const double *rate, const double *prop) {
    memcpy(e->cat_rate, rate, sizeof(double) * e->cfg.ncat);
    memcpy(e->cat_prop, prop, sizeof(double) * e->cfg.ncat);
    return 0;
}

/* 0/1 tip likelihood of state — model/modelsubst.cpp:197-207, modeldna.cpp:462-475, modelprotein.cpp:1149-1166 */
static void tip_indicator(const orc_engine *e, int state, double *lk) {
    int n = e->cfg.nstates;
    if (state < n) {
        memset(lk, 0, sizeof(double) * n);
        lk[state] = 1.0;
    } else if (e->cfg.seq_type == 1 && state < 18) {
        memset(lk, 0, sizeof(double) * n);
        int cstate = state - n + 1;
        for (int i = 0; i < n; i++) if (cstate & (1 << i)) lk[i] = 1.0;
    } else if (e->cfg.seq_type == 2 && state < n + 3) {
        static const int ambi_aa[] = {2, 3, 5, 6, 9, 10};
        memset(lk, 0, sizeof(double) * n);
        int cstate = state - n;
        lk[ambi_aa[cstate * 2]] = 1.0;
        lk[ambi_aa[cstate * 2 + 1]] = 1.0;
    } else {
        for (int i = 0; i < n; i++) lk[i] = 1.0;
    }
}

/* computeTipPartialLikelihood — tree/phylotreesse.cpp:367-375 + ModelSubst::multiplyWithInvEigenvector
 * (model/modelsubst.cpp:182-195) */
int orc_set_model(orc_engine *e, const double *eval, const double *evec, const double *inv_evec, const double *mix_weight) {
    int n = e->cfg.nstates, nmix = e->cfg.nmix;
    memcpy(e->eval, eval, sizeof(double) * nmix * n);
    memcpy(e->evec, evec, sizeof(double) * nmix * n * n);
    memcpy(e->inv_evec, inv_evec, sizeof(double) * nmix * n * n);
    memcpy(e->mix_weight, mix_weight, sizeof(double) * nmix);
    double *saved = (double *)malloc(sizeof(double) * n);
    for (int state = 0; state <= e->cfg.state_unknown; state++) {
        double *lk = e->tip_partial_lh + (size_t)state * nmix * n;
        tip_indicator(e, state, saved);
        for (int m = 0; m < nmix; m++) {
            const double *iv = e->inv_evec + (size_t)m * n * n;
            for (int i = 0; i < n; i++) {
                double acc = 0.0;
                for (int j = 0; j < n; j++) acc += iv[i * n + j] * saved[j];
                lk[m * n + i] = acc;
            }
        }
    }
    free(saved);
    e->theta_computed = 0;
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * computePartialInfo — phylokernelnew.h:900-1110 (reversible branch)
 * ------------------------------------------------------------------------------------------- */
static void partial_info_child(const orc_engine *e, const orc_child *ch, double *echild, double *leaf /* or NULL */) {
    int n = e->cfg.nstates, C = e->C, B = e->B, nmix = e->cfg.nmix, ncat = e->cfg.ncat;
    int tip_block = n * nmix;
    double *expchild = (double *)malloc(sizeof(double) * n);
    for (int c = 0; c < C; c++) {
        int m = c / e->denom, r = c % ncat;
        double len_child = e->cat_rate[r] * len_of(e, ch->length, r); /* :1017 */
        const double *eval_ptr = e->eval + (size_t)m * n;
        const double *evec_ptr = e->evec + (size_t)m * n * n;
        for (int i = 0; i < n; i++) expchild[i] = exp(eval_ptr[i] * len_child); /* :1022 / :1073 */
        for (int x = 0; x < n; x++)
            for (int i = 0; i < n; i++)
                echild[(size_t)c * n * n + x * n + i] = evec_ptr[x * n + i] * expchild[i]; /* :1027 / :1077 */
    }
    free(expchild);
    if (!leaf) return;
    int su = e->cfg.state_unknown;
    for (int state = 0; state <= su; state++) {
        for (int c = 0; c < C; c++) {
            int m = c / e->denom;
            const double *tip = e->tip_partial_lh + (size_t)state * tip_block + m * n;
            for (int x = 0; x < n; x++) {
                const double *ep = echild + (size_t)c * n * n + x * n;
                double v;
                if (n % 4 == 0) {
                    /* vectorised (Vec4d) branch :1036-1051: lane accumulators then horizontal_add */
                    double a0 = ep[0] * tip[0], a1 = ep[1] * tip[1], a2 = ep[2] * tip[2], a3 = ep[3] * tip[3];
                    for (int i = 4; i < n; i += 4) {
                        a0 = fma(ep[i], tip[i], a0); a1 = fma(ep[i + 1], tip[i + 1], a1);
                        a2 = fma(ep[i + 2], tip[i + 2], a2); a3 = fma(ep[i + 3], tip[i + 3], a3);
                    }
                    v = (a0 + a1) + (a2 + a3);
                } else {
                    /* non-vectorised branch :1085-1099 */
                    v = ep[0] * tip[0];
                    for (int i = 1; i < n; i++) v += ep[i] * tip[i];
                }
                leaf[(size_t)state * B + c * n + x] = v;
            }
        }
    }
    for (int x = 0; x < B; x++) leaf[(size_t)su * B + x] = 1.0; /* :1052-1055 */
}

static inline int tip_state(const orc_engine *e, int taxon, size_t p) { return e->tips[(size_t)taxon * e->cfg.nptn + p]; }

/* ---------------------------------------------------------------------------------------------
 * computePartialLikelihood(Generic)SIMD — phylokernelnew.h:1289-2039 (binary nodes, normal models)
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    orc_op op;            /* after the leaf swap: a single leaf child is `left` (:1395-1405) */
    double *echildren;    /* owns eleft+eright */
    double *leaves;       /* owns leaf tables */
    double *eleft, *eright, *leafL, *leafR;
    /* multifurcating node: all children in neighbour order, echild / leaf table per child (NULL for internal children) */
    int nchild;
    orc_child *child;
    double **mechild, **mleaf;
} orc_opinfo;

/* per-op precompute = computePartialInfo for one TraversalInfo (:1243-1260) */
static void opinfo_init_multi(const orc_engine *e, const orc_op *op_in, const orc_child *extra, orc_opinfo *oi) {
    int n = e->cfg.nstates, B = e->B;
    size_t info_sz = (size_t)B * n, leaf_sz = (size_t)(e->cfg.state_unknown + 1) * B;
    oi->op = *op_in;
    oi->nchild = 2 + op_in->nextra;
    oi->child = (orc_child *)malloc(sizeof(orc_child) * oi->nchild);
    oi->child[0] = op_in->left; oi->child[1] = op_in->right;
    for (int k = 0; k < op_in->nextra; k++) oi->child[2 + k] = extra[k];
    oi->mechild = (double **)calloc(oi->nchild, sizeof(double *));
    oi->mleaf = (double **)calloc(oi->nchild, sizeof(double *));
    oi->echildren = oi->leaves = NULL;
    for (int k = 0; k < oi->nchild; k++) {
        oi->mechild[k] = (double *)malloc(sizeof(double) * info_sz);
        if (oi->child[k].kind == 0) oi->mleaf[k] = (double *)malloc(sizeof(double) * leaf_sz);
        partial_info_child(e, &oi->child[k], oi->mechild[k], oi->mleaf[k]);
    }
}

static void opinfo_init(const orc_engine *e, const orc_op *op_in, orc_opinfo *oi) {
    int n = e->cfg.nstates, B = e->B;
    oi->op = *op_in;
    oi->nchild = 2; oi->child = NULL; oi->mechild = oi->mleaf = NULL;
    size_t info_sz = (size_t)B * n;
    oi->echildren = (double *)malloc(sizeof(double) * 2 * info_sz);
    size_t leaf_sz = (size_t)(e->cfg.state_unknown + 1) * B;
    oi->leaves = (double *)malloc(sizeof(double) * 2 * leaf_sz);
    /* echildren in neighbour order; partial_lh_leaves only for leaf children, in neighbour order (:1012-1058) */
    double *leaf_ptr = oi->leaves;
    oi->leafL = oi->leafR = NULL;
    if (oi->op.left.kind == 0) { oi->leafL = leaf_ptr; leaf_ptr += leaf_sz; }
    if (oi->op.right.kind == 0) { oi->leafR = leaf_ptr; leaf_ptr += leaf_sz; }
    partial_info_child(e, &oi->op.left, oi->echildren, oi->leafL);
    partial_info_child(e, &oi->op.right, oi->echildren + info_sz, oi->leafR);
    oi->eleft = oi->echildren;
    oi->eright = oi->echildren + info_sz;
    if (oi->op.left.kind != 0 && oi->op.right.kind == 0) {
        orc_child t = oi->op.left; oi->op.left = oi->op.right; oi->op.right = t;
        double *et = oi->eleft; oi->eleft = oi->eright; oi->eright = et;
        oi->leafL = oi->leafR; oi->leafR = NULL;
    }
}
static void opinfo_free(orc_opinfo *oi) {
    free(oi->echildren); free(oi->leaves);
    if (oi->child) {
        for (int k = 0; k < oi->nchild; k++) { free(oi->mechild[k]); free(oi->mleaf[k]); }
        free(oi->child); free(oi->mechild); free(oi->mleaf);
    }
}

#define ORC_MAX_STATES 256

/* multifurcating node, normal models — phylokernelnew.h:1407-1631: the running product over the children lives in the
 * ORIGINAL state space (partial_lh_all), the 2^-256 test follows EVERY child (:1557-1605) on that product, and the
 * inverse-eigenvector transform comes last without a further test (:1610-1628) */
static void compute_partial_multi(orc_engine *e, const orc_opinfo *oi, size_t lo, size_t hi) {
    int n = e->cfg.nstates, C = e->C, B = e->B, safe = e->cfg.safe_numeric;
    size_t P = (size_t)e->cfg.nptn, S = S_of(e);
    double *dad_lh = e->partial + (size_t)oi->op.dst_slot * P * B;
    uint16_t *dad_sc = e->scale + (size_t)oi->op.dst_slot * P * S;
    double *all = (double *)malloc(sizeof(double) * B);
    for (size_t p = lo; p < hi; p++) {
        for (int i = 0; i < B; i++) all[i] = 1.0;
        uint16_t *sc = dad_sc + p * S;
        memset(sc, 0, sizeof(uint16_t) * S);
        for (int k = 0; k < oi->nchild; k++) {
            const orc_child *ch = &oi->child[k];
            if (ch->kind == 0) {
                const double *child_lh = oi->mleaf[k] + (size_t)tip_state(e, ch->id, p) * B;
                for (int c = 0; c < B; c++) all[c] *= child_lh[c]; /* :1520-1523 */
            } else {
                const double *plc = e->partial + (size_t)ch->id * P * B + p * B;
                const uint16_t *csc = e->scale + (size_t)ch->id * P * S + p * S;
                if (!safe) sc[0] = (uint16_t)(sc[0] + csc[0]); /* :1529-1532 */
                const double *ep = oi->mechild[k];
                for (int c = 0; c < C; c++) {
                    if (safe) sc[c] = (uint16_t)(sc[c] + csc[c]); /* :1536-1539 */
                    for (int x = 0; x < n; x++) {
                        double v = ep[0] * plc[c * n]; /* :1542-1546: sequential mul_add */
                        for (int i = 1; i < n; i++) v = fma(ep[i], plc[c * n + i], v);
                        ep += n;
                        all[c * n + x] *= v;
                    }
                }
            }
            if (safe) { /* :1558-1580 */
                for (int c = 0; c < C; c++) {
                    double lh_max = 0.0;
                    for (int x = 0; x < n; x++) lh_max = fmax(lh_max, fabs(all[c * n + x]));
                    if (lh_max < SCALING_THRESHOLD && e->ptn_invar[p] == 0.0) {
                        for (int x = 0; x < n; x++) all[c * n + x] = ldexp(all[c * n + x], SCALING_THRESHOLD_EXP);
                        sc[c] += 1;
                    }
                }
            } else { /* :1581-1603 */
                double lh_max = 0.0;
                for (int x = 0; x < B; x++) lh_max = fmax(lh_max, fabs(all[x]));
                if (lh_max < SCALING_THRESHOLD && e->ptn_invar[p] == 0.0) {
                    for (int x = 0; x < B; x++) all[x] = ldexp(all[x], SCALING_THRESHOLD_EXP);
                    sc[0] += 1;
                }
            }
        }
        double *out = dad_lh + p * B;
        for (int c = 0; c < C; c++) {
            const double *inv = e->inv_evec + (size_t)(c / e->denom) * n * n;
            prod_vec_mat(all + c * n, inv, out + c * n, NULL, n); /* :1622-1624 */
        }
    }
    free(all);
}

/* one TraversalInfo over the pattern range [lo, hi) — the body of computePartialLikelihood(info, lo, hi, packet) */
static void compute_partial_range(orc_engine *e, const orc_opinfo *oi, size_t lo, size_t hi) {
    if (oi->child) { compute_partial_multi(e, oi, lo, hi); return; }
    int n = e->cfg.nstates, C = e->C, B = e->B, safe = e->cfg.safe_numeric;
    size_t P = (size_t)e->cfg.nptn, S = S_of(e);
    const orc_op *op = &oi->op;
    double *dad_lh = e->partial + (size_t)op->dst_slot * P * B;
    uint16_t *dad_sc = e->scale + (size_t)op->dst_slot * P * S;
    double tmpv[ORC_MAX_STATES];

    if (op->left.kind == 0 && op->right.kind == 0) {
        /* TIP-TIP :1632-1746 — scale_num zeroed, never rescaled */
        memset(dad_sc + lo * S, 0, sizeof(uint16_t) * (hi - lo) * S);
        for (size_t p = lo; p < hi; p++) {
            const double *vl = oi->leafL + (size_t)tip_state(e, op->left.id, p) * B;
            const double *vr = oi->leafR + (size_t)tip_state(e, op->right.id, p) * B;
            double *out = dad_lh + p * B;
            for (int c = 0; c < C; c++) {
                const double *inv = e->inv_evec + (size_t)(c / e->denom) * n * n;
                for (int x = 0; x < n; x++) tmpv[x] = vl[c * n + x] * vr[c * n + x];
                prod_vec_mat(tmpv, inv, out + c * n, NULL, n);
            }
        }
    } else if (op->left.kind == 0) {
        /* TIP-INNER :1748-1905 */
        const double *right_lh = e->partial + (size_t)op->right.id * P * B;
        const uint16_t *right_sc = e->scale + (size_t)op->right.id * P * S;
        memcpy(dad_sc + lo * S, right_sc + lo * S, sizeof(uint16_t) * (hi - lo) * S);
        for (size_t p = lo; p < hi; p++) {
            const double *vl = oi->leafL + (size_t)tip_state(e, op->left.id, p) * B;
            const double *pr = right_lh + p * B;
            double *out = dad_lh + p * B;
            double lh_max = 0.0;
            for (int c = 0; c < C; c++) {
                if (safe) lh_max = 0.0;
                const double *inv = e->inv_evec + (size_t)(c / e->denom) * n * n;
                const double *er = oi->eright + (size_t)c * n * n;
                for (int x = 0; x < n; x++) {
                    double vright = dot_vec(er + x * n, pr + c * n, n);
                    tmpv[x] = vl[c * n + x] * vright;
                }
                prod_vec_mat(tmpv, inv, out + c * n, &lh_max, n);
                if (safe && lh_max < SCALING_THRESHOLD && e->ptn_invar[p] == 0.0) {
                    for (int i = 0; i < n; i++) out[c * n + i] = ldexp(out[c * n + i], SCALING_THRESHOLD_EXP);
                    dad_sc[p * C + c] += 1;
                }
            }
            if (!safe && lh_max < SCALING_THRESHOLD && e->ptn_invar[p] == 0.0) {
                for (int i = 0; i < B; i++) out[i] = ldexp(out[i], SCALING_THRESHOLD_EXP);
                dad_sc[p] += 1;
            }
        }
    } else {
        /* INNER-INNER :1907-2032 */
        const double *left_lh = e->partial + (size_t)op->left.id * P * B;
        const double *right_lh = e->partial + (size_t)op->right.id * P * B;
        const uint16_t *left_sc = e->scale + (size_t)op->left.id * P * S;
        const uint16_t *right_sc = e->scale + (size_t)op->right.id * P * S;
        for (size_t p = lo; p < hi; p++) {
            const double *pl = left_lh + p * B, *pr = right_lh + p * B;
            double *out = dad_lh + p * B;
            double lh_max = 0.0;
            if (!safe) dad_sc[p] = (uint16_t)(left_sc[p] + right_sc[p]);
            for (int c = 0; c < C; c++) {
                if (safe) { lh_max = 0.0; dad_sc[p * C + c] = (uint16_t)(left_sc[p * C + c] + right_sc[p * C + c]); }
                const double *inv = e->inv_evec + (size_t)(c / e->denom) * n * n;
                const double *el = oi->eleft + (size_t)c * n * n, *er = oi->eright + (size_t)c * n * n;
                for (int x = 0; x < n; x++) tmpv[x] = dot_dual(el + x * n, pl + c * n, er + x * n, pr + c * n, n);
                prod_vec_mat(tmpv, inv, out + c * n, &lh_max, n);
                if (safe && lh_max < SCALING_THRESHOLD && e->ptn_invar[p] == 0.0) {
                    for (int i = 0; i < n; i++) out[c * n + i] = ldexp(out[c * n + i], SCALING_THRESHOLD_EXP);
                    dad_sc[p * C + c] += 1;
                }
            }
            if (!safe && lh_max < SCALING_THRESHOLD && e->ptn_invar[p] == 0.0) {
                for (int i = 0; i < B; i++) out[i] = ldexp(out[i], SCALING_THRESHOLD_EXP);
                dad_sc[p] += 1;
            }
        }
    }
}

/* The reference walks the WHOLE traversal list per pattern packet (phylokernelnew.h:1270-1277, 2087-2089,
 * 2821-2824): packets are independent pattern ranges, so children written by a packet are still cache-hot
 * when the same packet computes their parent.  Same structure here (packet = ORC_PACKET patterns). */
#define ORC_PACKET 512
static int nonrev_update_partials(orc_engine *e, const orc_op *ops, int nops);
static int nonrev_branch_logl(orc_engine *e, int node_slot, int dad_kind, int dad_id, double length, double *logl_out);
static int nonrev_branch_derv(orc_engine *e, int node_slot, int dad_kind, int dad_id, double length, double *df_out, double *ddf_out);
static int site_update_partials(orc_engine *e, const orc_op *ops, int nops);
static int site_branch_logl(orc_engine *e, int node_slot, int dad_kind, int dad_id, double length, int save_log_value, double *logl_out);
static int site_branch_derv(orc_engine *e, int node_slot, int dad_kind, int dad_id, double length, double *df_out, double *ddf_out);
static int site_logl_from_buffer(orc_engine *e, double length, double *logl_out);
int orc_update_partials(orc_engine *e, const orc_op *ops, int nops) {
    if (e->nonrev) return nonrev_update_partials(e, ops, nops);
    if (e->site) return site_update_partials(e, ops, nops);
    if (nops <= 0) return 0;
    if (e->cfg.nstates > ORC_MAX_STATES) { snprintf(e->err, sizeof e->err, "nstates > %d", ORC_MAX_STATES); return -1; }
    orc_opinfo *info = (orc_opinfo *)malloc(sizeof(orc_opinfo) * nops);
    /* extra children of multifurcating ops: consumed in op order */
    int *xoff = (int *)malloc(sizeof(int) * (nops + 1));
    xoff[0] = 0;
    for (int k = 0; k < nops; k++) xoff[k + 1] = xoff[k] + (ops[k].nextra > 0 ? ops[k].nextra : 0);
    if (xoff[nops] > e->nextra) {
        snprintf(e->err, sizeof e->err, "multifurcating ops need %d extra children, %d set", xoff[nops], e->nextra);
        free(xoff); free(info);
        return -1;
    }
#pragma omp parallel for schedule(static) if (nops >= 3)
    for (int k = 0; k < nops; k++) {
        if (ops[k].nextra > 0) opinfo_init_multi(e, &ops[k], e->extra + xoff[k], &info[k]);
        else opinfo_init(e, &ops[k], &info[k]);
    }
    free(xoff);
    e->nextra = 0;
    size_t P = (size_t)e->cfg.nptn;
    long npacket = (long)((P + ORC_PACKET - 1) / ORC_PACKET);
#pragma omp parallel for schedule(dynamic, 1)
    for (long pk = 0; pk < npacket; pk++) {
        size_t lo = (size_t)pk * ORC_PACKET, hi = lo + ORC_PACKET;
        if (hi > P) hi = P;
        for (int k = 0; k < nops; k++) compute_partial_range(e, &info[k], lo, hi);
    }
    for (int k = 0; k < nops; k++) opinfo_free(&info[k]);
    free(info);
    return 0;
}

/* val[c][i] = exp(eval_i * rate_c * len) * prop_c * weight_m — phylokernelnew.h:2714-2745 ;
 * with derivatives :2322-2358 */
static void compute_vals(const orc_engine *e, double len_branch, double *val0, double *val1, double *val2) {
    int n = e->cfg.nstates, C = e->C, ncat = e->cfg.ncat;
    for (int c = 0; c < C; c++) {
        int m = c / e->denom, r = c % ncat;
        double prop = e->cat_prop[r] * e->mix_weight[m];
        const double *eval_ptr = e->eval + (size_t)m * n;
        if (val1) {
            double myrate = e->cat_rate[r];
            for (int i = 0; i < n; i++) {
                double cof = eval_ptr[i] * myrate;
                double val = exp(cof * len_of(e, len_branch, r)) * prop;
                double v1 = cof * val;
                val0[c * n + i] = val; val1[c * n + i] = v1; val2[c * n + i] = cof * v1;
            }
        } else {
            double len = e->cat_rate[r] * len_of(e, len_branch, r);
            for (int i = 0; i < n; i++) val0[c * n + i] = exp(eval_ptr[i] * len) * prop;
        }
    }
}

static int cmp_double(const void *a, const void *b) {
    const double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : x > y ? 1 : 0;
}

/* Holder's ascertainment-bias correction for missing data (+ASC_MIS, ASC_VARIANT_MISSING; phylokernelnew.h:3175-3197,
 * 3385-3410): the unobserved block holds nstates groups of `step` constant patterns, group s = every observed pattern with
 * its non-gap characters replaced by state s (Alignment::getUnobservedConstPatterns, alignment/alignment.cpp:4158-4180;
 * step = the padded pattern count there, anything >= Pobs here).  pattern_lh[Pobs + s*step + p] holds their likelihoods.
 * Applies _pattern_lh[p] -= log(1 - sum_s const_lh) and returns sum_p freq_p * log(1 - ...), to be subtracted from tree_lh. */
static double holder_logl_correction(orc_engine *e, size_t Pobs) {
    const int n = e->cfg.nstates;
    const size_t P = (size_t)e->cfg.nptn, step = (size_t)e->asc_nunobs / (size_t)n;
    double sum_corr = 0.0;
    for (size_t p = 0; p < Pobs; p++) {
        double c = e->pattern_lh[Pobs + p];
        for (int s = 1; s < n; s++) c += e->pattern_lh[Pobs + (size_t)s * step + p];
        const double prob_variant = log(1.0 - c);
        e->pattern_lh[p] -= prob_variant;
        sum_corr += prob_variant * e->ptn_freq[p];
    }
    for (size_t p = Pobs; p < P; p++) e->pattern_lh[p] = 0.0;
    return sum_corr;
}

/* ---------------------------------------------------------------------------------------------
 * computeLikelihoodBranch(Generic)SIMD — phylokernelnew.h:2641-3223 (+ASC: Lewis and Holder; no robust-phylo)
 *   node end must be a slot; dad end is a tip (dad_kind 0) or a slot (dad_kind 1).
 * ------------------------------------------------------------------------------------------- */
int orc_branch_logl(orc_engine *e, const orc_op *ops, int nops, int node_slot, int dad_kind, int dad_id,
                    double length, int save_log_value, double *logl_out) {
    if (e->nonrev) {
        if (orc_update_partials(e, ops, nops) < 0) return -1;
        return nonrev_branch_logl(e, node_slot, dad_kind, dad_id, length, logl_out);
    }
    if (e->site) {
        if (orc_update_partials(e, ops, nops) < 0) return -1;
        return site_branch_logl(e, node_slot, dad_kind, dad_id, length, save_log_value, logl_out);
    }
    orc_update_partials(e, ops, nops);
    int n = e->cfg.nstates, C = e->C, B = e->B, safe = e->cfg.safe_numeric, nmix = e->cfg.nmix;
    size_t P = (size_t)e->cfg.nptn, S = S_of(e);
    double *val = (double *)malloc(sizeof(double) * B);
    compute_vals(e, length, val, NULL, NULL);
    /* naming follows the reference after its leaf swap: dad_branch->partial_lh is the slot side */
    const double *dad_lh = e->partial + (size_t)node_slot * P * B;
    const uint16_t *dad_sc = e->scale + (size_t)node_slot * P * S;
    double tree_lh = 0.0, prob_const = 0.0;
    const size_t Pobs = e->asc_type ? P - (size_t)e->asc_nunobs : P;   /* orig_nptn */
    if (dad_kind == 0) {
        /* TIP-INNER :2753-2965 : partial_lh_node[state][c][i] = val[c][i]*tip_partial_lh[state][m][i] */
        int su = e->cfg.state_unknown;
        double *plnode = (double *)malloc(sizeof(double) * (su + 1) * B);
        for (int state = 0; state <= su; state++)
            for (int c = 0; c < C; c++) {
                int m = c / e->denom;
                for (int i = 0; i < n; i++)
                    plnode[(size_t)state * B + c * n + i] = val[c * n + i] * e->tip_partial_lh[(size_t)state * nmix * n + m * n + i];
            }
#pragma omp parallel for schedule(static) reduction(+ : tree_lh, prob_const)
        for (size_t p = 0; p < P; p++) {
            const double *lh_node = plnode + (size_t)tip_state(e, dad_id, p) * B;
            const double *pd = dad_lh + p * B;
            double *lh_cat = e->pattern_lh_cat + p * C;
            double lh_ptn = 0.0;
            for (int c = 0; c < C; c++) {
                lh_cat[c] = dot_vec(lh_node + c * n, pd + c * n, n);
                if (!safe) lh_ptn += lh_cat[c];
            }
            double min_scale_d;
            if (safe) {
                const uint16_t *sd = dad_sc + p * C;
                uint16_t min_scale = sd[0];
                for (int c = 1; c < C; c++) if (sd[c] < min_scale) min_scale = sd[c];
                min_scale_d = min_scale;
                for (int c = 0; c < C; c++) {
                    if (sd[c] == min_scale + 1) lh_cat[c] *= SCALING_THRESHOLD;
                    else if (sd[c] > min_scale + 1) lh_cat[c] = 0.0;
                }
                lh_ptn = sum_vec_append(lh_cat, lh_ptn, C);
            } else min_scale_d = dad_sc[p];
            min_scale_d *= LOG_SCALING_THRESHOLD;
            lh_ptn = fabs(lh_ptn) + e->ptn_invar[p];
            if (p >= Pobs) { /* unobserved constant pattern: :3076-3091 */
                if (min_scale_d != 0.0) lh_ptn *= SCALING_THRESHOLD;
                if (e->asc_type == 2) e->pattern_lh[p] = lh_ptn; /* Holder: kept per pattern (:3088-3089) */
                else prob_const += lh_ptn;
                continue;
            }
            double lg = log(lh_ptn) + min_scale_d;
            if (save_log_value) e->pattern_lh[p] = lg;
            else { e->pattern_lh[p] = lh_ptn; e->pattern_scaling[p] = min_scale_d; }
            tree_lh += lg * e->ptn_freq[p];
        }
        free(plnode);
    } else {
        /* INNER-INNER :2966-3104 */
        const double *node_lh = e->partial + (size_t)dad_id * P * B;
        const uint16_t *node_sc = e->scale + (size_t)dad_id * P * S;
#pragma omp parallel for schedule(static) reduction(+ : tree_lh, prob_const)
        for (size_t p = 0; p < P; p++) {
            const double *pd = dad_lh + p * B, *pn = node_lh + p * B;
            double *lh_cat = e->pattern_lh_cat + p * C;
            double lh_ptn = 0.0;
            for (int c = 0; c < C; c++) {
                lh_cat[c] = dot3(val + c * n, pn + c * n, pd + c * n, n);
                if (!safe) lh_ptn += lh_cat[c];
            }
            double min_scale_d;
            if (safe) {
                const uint16_t *sd = dad_sc + p * C, *sn = node_sc + p * C;
                uint16_t min_scale = (uint16_t)(sd[0] + sn[0]);
                for (int c = 1; c < C; c++) { uint16_t s = (uint16_t)(sd[c] + sn[c]); if (s < min_scale) min_scale = s; }
                min_scale_d = min_scale;
                for (int c = 0; c < C; c++) {
                    uint16_t s = (uint16_t)(sd[c] + sn[c]);
                    if (s == min_scale + 1) lh_cat[c] *= SCALING_THRESHOLD;
                    else if (s > min_scale + 1) lh_cat[c] = 0.0;
                }
                lh_ptn = sum_vec_append(lh_cat, lh_ptn, C);
            } else min_scale_d = (double)(dad_sc[p] + node_sc[p]);
            min_scale_d *= LOG_SCALING_THRESHOLD;
            lh_ptn = fabs(lh_ptn) + e->ptn_invar[p];
            if (p >= Pobs) { /* unobserved constant pattern: :3076-3091 */
                if (min_scale_d != 0.0) lh_ptn *= SCALING_THRESHOLD;
                if (e->asc_type == 2) e->pattern_lh[p] = lh_ptn; /* Holder: kept per pattern (:3088-3089) */
                else prob_const += lh_ptn;
                continue;
            }
            double lg = log(lh_ptn) + min_scale_d;
            if (save_log_value) e->pattern_lh[p] = lg;
            else { e->pattern_lh[p] = lh_ptn; e->pattern_scaling[p] = min_scale_d; }
            tree_lh += lg * e->ptn_freq[p];
        }
    }
    free(val);
    int status = 0;
    if (!isfinite(tree_lh)) {
        /* NaN repair :3115-3132 */
        status = 1;
        tree_lh = 0.0;
        for (size_t p = 0; p < Pobs; p++) {
            double ptn_lh = e->pattern_lh[p];
            if (!isfinite(ptn_lh)) {
                if (save_log_value) e->pattern_lh[p] = LOG_SCALING_THRESHOLD * 4;
                else { e->pattern_lh[p] = 1.0; e->pattern_scaling[p] = LOG_SCALING_THRESHOLD * 4; }
                ptn_lh = LOG_SCALING_THRESHOLD * 4;
            }
            tree_lh += ptn_lh * e->ptn_freq[p];
        }
    }
    /* robust phylogeny project (:3134-3173): the log-likelihood of the best sites / of the median site */
    const int robust_keep_on = e->robust_keep > 0.0 && e->robust_keep < 1.0;
    if (robust_keep_on || e->robust_median) {
        if (e->asc_type) { snprintf(e->err, sizeof e->err, "+ASC not supported with robust phylo"); return -1; }
        size_t sites = 0;
        for (size_t p = 0; p < P; p++) sites += (size_t)e->ptn_freq[p];
        double *site_lh = (double *)malloc(sizeof(double) * (sites ? sites : 1));
        size_t site = 0;
        for (size_t p = 0; p < P; p++)
            for (size_t i = 0; i < (size_t)e->ptn_freq[p]; i++) site_lh[site++] = e->pattern_lh[p];
        qsort(site_lh, sites, sizeof(double), cmp_double); /* the reference's nth_element leaves the kept part unordered */
        if (robust_keep_on) {
            const size_t sites_drop = (size_t)(int)((1.0 - e->robust_keep) * (double)sites);
            tree_lh = 0.0;
            for (size_t i = sites_drop; i < sites; i++) tree_lh += site_lh[i];
        } else if (sites & 1) tree_lh = site_lh[sites / 2];
        else tree_lh = 0.5 * (site_lh[sites / 2] + site_lh[sites / 2 - 1]);
        free(site_lh);
    }
    if (e->asc_type == 1) { /* Lewis :3198-3221 */
        double corr = log(1.0 - prob_const);
        for (size_t p = 0; p < Pobs; p++) e->pattern_lh[p] -= corr;
        for (size_t p = Pobs; p < P; p++) e->pattern_lh[p] = 0.0;
        tree_lh -= (double)e->asc_nsite * corr;
    } else if (e->asc_type == 2) tree_lh -= holder_logl_correction(e, Pobs); /* :3175-3197 */
    *logl_out = tree_lh;
    return status;
}

/* computeLikelihoodBuffer(Generic)SIMD — phylokernelnew.h:2048-2231 */
static void build_theta(orc_engine *e, int node_slot, int dad_kind, int dad_id) {
    int n = e->cfg.nstates, C = e->C, B = e->B, safe = e->cfg.safe_numeric, nmix = e->cfg.nmix;
    size_t P = (size_t)e->cfg.nptn, S = S_of(e);
    const double *dad_lh = e->partial + (size_t)node_slot * P * B;
    const uint16_t *dad_sc = e->scale + (size_t)node_slot * P * S;
    if (dad_kind == 0) {
        int tip_block = n * nmix;
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < P; p++) {
            const double *tip = e->tip_partial_lh + (size_t)tip_state(e, dad_id, p) * tip_block;
            const double *pd = dad_lh + p * B;
            double *theta = e->theta + p * B;
            for (int c = 0; c < C; c++) {
                const double *lh_tip = tip + (c / e->denom) * n;
                for (int i = 0; i < n; i++) theta[c * n + i] = lh_tip[i] * pd[c * n + i];
            }
            if (safe) {
                const uint16_t *sd = dad_sc + p * C;
                uint16_t min_scale = sd[0];
                for (int c = 1; c < C; c++) if (sd[c] < min_scale) min_scale = sd[c];
                e->buffer_scale_all[p] = min_scale;
                for (int c = 0; c < C; c++) {
                    if (sd[c] == min_scale + 1) for (int x = 0; x < n; x++) theta[c * n + x] *= SCALING_THRESHOLD;
                    else if (sd[c] > min_scale + 1) for (int x = 0; x < n; x++) theta[c * n + x] = 0.0;
                }
            } else e->buffer_scale_all[p] = dad_sc[p];
            e->buffer_scale_all[p] *= LOG_SCALING_THRESHOLD;
        }
    } else {
        const double *node_lh = e->partial + (size_t)dad_id * P * B;
        const uint16_t *node_sc = e->scale + (size_t)dad_id * P * S;
#pragma omp parallel for schedule(static)
        for (size_t p = 0; p < P; p++) {
            const double *pd = dad_lh + p * B, *pn = node_lh + p * B;
            double *theta = e->theta + p * B;
            for (int i = 0; i < B; i++) theta[i] = pn[i] * pd[i];
            if (safe) {
                const uint16_t *sd = dad_sc + p * C, *sn = node_sc + p * C;
                uint16_t min_scale = (uint16_t)(sd[0] + sn[0]);
                for (int c = 1; c < C; c++) { uint16_t s = (uint16_t)(sd[c] + sn[c]); if (s < min_scale) min_scale = s; }
                e->buffer_scale_all[p] = min_scale;
                for (int c = 0; c < C; c++) {
                    uint16_t s = (uint16_t)(sd[c] + sn[c]);
                    if (s == min_scale + 1) for (int x = 0; x < n; x++) theta[c * n + x] *= SCALING_THRESHOLD;
                    else if (s > min_scale + 1) for (int x = 0; x < n; x++) theta[c * n + x] = 0.0;
                }
            } else e->buffer_scale_all[p] = (double)(dad_sc[p] + node_sc[p]);
            e->buffer_scale_all[p] *= LOG_SCALING_THRESHOLD;
        }
    }
    e->theta_computed = 1;
}

int orc_theta_reset(orc_engine *e) { e->theta_computed = 0; return 0; }
/* bit 0: reproduce dotProductTriple's double-counted last element for odd nstates x even block (see dot_triple) */
int orc_set_reference_quirks(orc_engine *e, int mask) { e->quirks = mask; return 0; }

/* --robust-phy keep / --robust-median (utils/tools.cpp:4374-4392): orc_branch_logl returns the sum over the best keep*sites sites /
 * the median site log-likelihood.  keep = 1 and median = 0 switch it off. */
int orc_set_robust_phylo(orc_engine *e, double keep, int median) {
    if (!(keep > 0.0 && keep <= 1.0) || (keep < 1.0 && median)) { snprintf(e->err, sizeof e->err, "bad robust-phylo arguments"); return -1; }
    e->robust_keep = keep; e->robust_median = median;
    return 0;
}

int orc_set_ascertainment(orc_engine *e, int asc_type, int64_t nunobs, int64_t nsite) {
    if (asc_type < 0 || asc_type > 2 || nunobs < 0 || nunobs > e->cfg.nptn) { snprintf(e->err, sizeof e->err, "bad +ASC arguments"); return -1; }
    /* 2 = Holder: nstates groups of `step` >= (observed patterns) constant patterns */
    if (asc_type == 2 && (nunobs % e->cfg.nstates != 0 || nunobs / e->cfg.nstates < e->cfg.nptn - nunobs)) { snprintf(e->err, sizeof e->err, "bad +ASC_MIS block"); return -1; }
    e->asc_type = asc_type; e->asc_nunobs = asc_type ? nunobs : 0; e->asc_nsite = nsite;
    e->theta_computed = 0;
    return 0;
}

int orc_set_extra_children(orc_engine *e, const orc_child *children, int n) {
    free(e->extra);
    e->extra = NULL; e->nextra = 0;
    if (n <= 0) return 0;
    e->extra = (orc_child *)malloc(sizeof(orc_child) * n);
    memcpy(e->extra, children, sizeof(orc_child) * n);
    e->nextra = n;
    return 0;
}

/* computeLikelihoodDerv(Generic)SIMD — phylokernelnew.h:2233-2633 (joint branch length, no +ASC) */
int orc_branch_derv(orc_engine *e, const orc_op *ops, int nops, int node_slot, int dad_kind, int dad_id,
                    double length, double *df_out, double *ddf_out) {
    if (e->nonrev) {
        if (orc_update_partials(e, ops, nops) < 0) return -1;
        return nonrev_branch_derv(e, node_slot, dad_kind, dad_id, length, df_out, ddf_out);
    }
    if (e->site) {
        if (orc_update_partials(e, ops, nops) < 0) return -1;
        return site_branch_derv(e, node_slot, dad_kind, dad_id, length, df_out, ddf_out);
    }
    if (e->mixlen) { snprintf(e->err, sizeof e->err, "mixlen: computeLikelihoodDerv returns a gradient and a Hessian (orc_branch_derv_mixlen)"); return -1; }
    orc_update_partials(e, ops, nops);
    if (!e->theta_computed) build_theta(e, node_slot, dad_kind, dad_id);
    int B = e->B;
    size_t P = (size_t)e->cfg.nptn;
    double *val0 = (double *)malloc(sizeof(double) * 3 * B), *val1 = val0 + B, *val2 = val1 + B;
    compute_vals(e, length, val0, val1, val2);
    double all_df = 0.0, all_ddf = 0.0, prob_const = 0.0, df_const = 0.0, ddf_const = 0.0;
    const size_t Pobs = e->asc_type ? P - (size_t)e->asc_nunobs : P;
    const int odd = (e->quirks & 1) ? (e->cfg.nstates & 1) : (B & 1);
    double *const_df = NULL, *const_ddf = NULL; /* Holder: per unobserved pattern (:2371-2373, 2527-2530) */
    if (e->asc_type == 2) { const_df = (double *)calloc(2 * (P - Pobs) + 1, sizeof(double)); const_ddf = const_df + (P - Pobs); }
#pragma omp parallel for schedule(static) reduction(+ : all_df, all_ddf, prob_const, df_const, ddf_const)
    for (size_t p = 0; p < P; p++) {
        double lh_ptn, df_ptn, ddf_ptn;
        dot_triple(val0, val1, val2, e->theta + p * B, B, odd, &lh_ptn, &df_ptn, &ddf_ptn);
        lh_ptn = fabs(lh_ptn) + e->ptn_invar[p];
        if (p >= Pobs) { /* :2506-2535 */
            if (e->buffer_scale_all[p] != 0.0) { lh_ptn *= SCALING_THRESHOLD; df_ptn *= SCALING_THRESHOLD; ddf_ptn *= SCALING_THRESHOLD; }
            if (e->asc_type == 2) { e->pattern_lh[p] = lh_ptn; const_df[p - Pobs] = df_ptn; const_ddf[p - Pobs] = ddf_ptn; }
            else { prob_const += lh_ptn; df_const += df_ptn; ddf_const += ddf_ptn; }
            continue;
        }
        lh_ptn = 1.0 / lh_ptn;
        double df_frac = df_ptn * lh_ptn, ddf_frac = ddf_ptn * lh_ptn;
        double freq = e->ptn_freq[p];
        double tmp1 = df_frac * freq, tmp2 = ddf_frac * freq;
        all_df += tmp1;
        all_ddf += fma(-tmp1, df_frac, tmp2); /* nmul_add(tmp1, df_frac, tmp2) :2505 */
    }
    free(val0);
    if (e->asc_type == 1) { /* :2618-2626 */
        double all = 1.0 - prob_const, df_frac = df_const / all, ddf_frac = ddf_const / all;
        all_df += (double)e->asc_nsite * df_frac;
        all_ddf += (double)e->asc_nsite * (ddf_frac + df_frac * df_frac);
    } else if (e->asc_type == 2) { /* Holder :2580-2617 */
        const int n = e->cfg.nstates;
        const size_t step = (size_t)e->asc_nunobs / (size_t)n;
        double sum_df = 0.0, sum_ddf = 0.0;
        for (size_t p = 0; p < Pobs; p++) {
            double c = e->pattern_lh[Pobs + p], d1 = const_df[p], d2 = const_ddf[p];
            for (int s = 1; s < n; s++) { c += e->pattern_lh[Pobs + (size_t)s * step + p]; d1 += const_df[(size_t)s * step + p]; d2 += const_ddf[(size_t)s * step + p]; }
            c = 1.0 - c;
            const double df_frac = d1 / c, ddf_frac = d2 / c;
            sum_df += e->ptn_freq[p] * df_frac;
            sum_ddf += e->ptn_freq[p] * (ddf_frac + df_frac * df_frac);
        }
        for (size_t p = Pobs; p < P; p++) e->pattern_lh[p] = 0.0;
        all_df += sum_df;
        all_ddf += sum_ddf;
        free(const_df);
    }
    *df_out = all_df;
    *ddf_out = all_ddf;
    if (!isfinite(all_df)) { *df_out = *ddf_out = 0.0; return 1; } /* :2629-2632 */
    return 0;
}

/* computeLikelihoodFromBuffer(Generic)SIMD — phylokernelnew.h:3232-3428 */
int orc_logl_from_buffer(orc_engine *e, double length, double *logl_out) {
    if (!e->theta_computed) { snprintf(e->err, sizeof e->err, "theta not computed"); return -1; }
    if (e->site) return site_logl_from_buffer(e, length, logl_out);
    int B = e->B;
    size_t P = (size_t)e->cfg.nptn;
    double *val0 = (double *)malloc(sizeof(double) * B);
    compute_vals(e, length, val0, NULL, NULL);
    double tree_lh = 0.0, prob_const = 0.0;
    const size_t Pobs = e->asc_type ? P - (size_t)e->asc_nunobs : P;
#pragma omp parallel for schedule(static) reduction(+ : tree_lh, prob_const)
    for (size_t p = 0; p < P; p++) {
        double lh_ptn = dot_vec(val0, e->theta + p * B, B); /* dotProductVec over the whole block :3329 */
        lh_ptn = fabs(lh_ptn) + e->ptn_invar[p];
        if (p >= Pobs) { /* :3342-3356 */
            if (e->buffer_scale_all[p] != 0.0) lh_ptn *= SCALING_THRESHOLD;
            if (e->asc_type == 2) e->pattern_lh[p] = lh_ptn; /* :3352-3353 */
            else prob_const += lh_ptn;
            continue;
        }
        lh_ptn = log(fabs(lh_ptn)) + e->buffer_scale_all[p];
        e->pattern_lh[p] = lh_ptn;
        tree_lh += lh_ptn * e->ptn_freq[p];
    }
    free(val0);
    int status = 0;
    if (!isfinite(tree_lh)) {
        status = 1;
        tree_lh = 0.0;
        for (size_t p = 0; p < Pobs; p++) {
            if (!isfinite(e->pattern_lh[p])) e->pattern_lh[p] = LOG_SCALING_THRESHOLD * 4;
            tree_lh += e->pattern_lh[p] * e->ptn_freq[p];
        }
    }
    if (e->asc_type == 1) { /* :3412-3426 */
        double corr = log(1.0 - prob_const);
        for (size_t p = 0; p < Pobs; p++) e->pattern_lh[p] -= corr;
        for (size_t p = Pobs; p < P; p++) e->pattern_lh[p] = 0.0;
        tree_lh -= (double)e->asc_nsite * corr;
    } else if (e->asc_type == 2) tree_lh -= holder_logl_correction(e, Pobs); /* :3385-3410 */
    *logl_out = tree_lh;
    return status;
}

/* ---------------------------------------------------------------------------------------------
 * read-backs (reference layouts)
 * ------------------------------------------------------------------------------------------- */
int orc_get_pattern_lh(orc_engine *e, double *out) { memcpy(out, e->pattern_lh, sizeof(double) * e->cfg.nptn); return 0; }
int orc_get_pattern_scaling(orc_engine *e, double *out) { memcpy(out, e->pattern_scaling, sizeof(double) * e->cfg.nptn); return 0; }
int orc_get_pattern_lh_cat(orc_engine *e, double *out) { memcpy(out, e->pattern_lh_cat, sizeof(double) * e->cfg.nptn * e->C); return 0; }
/* PhyloTree::computePatternLhCat(wsl) after computeLikelihoodBranch (tree/phylotree.cpp:1338-1393; SiteLoglType utils/tools.h:555:
 * 2 = WSL_RATECAT, 3 = WSL_MIXTURE, 4 = WSL_MIXTURE_RATECAT): the per-category pattern likelihoods of the last orc_branch_logl,
 * for an unfused mixture summed over the mixture classes (per rate category) or over the rate categories (per class).
 * Returns K, the number of categories per pattern; out (may be NULL) is [nptn][K]. */
int orc_pattern_lh_cat_wsl(orc_engine *e, int wsl, double *out) {
    const int C = e->C, ncat = e->cfg.ncat, nmix = e->cfg.nmix;
    const int marginal = nmix > 1 && !e->cfg.fused_mix_rate && (wsl == 2 || wsl == 3);
    const int K = !marginal ? C : wsl == 2 ? ncat : nmix;
    if (wsl < 2 || wsl > 4) { snprintf(e->err, sizeof e->err, "wsl: 2 = RATECAT, 3 = MIXTURE, 4 = MIXTURE_RATECAT"); return -1; }
    if (!out) return K;
    for (size_t p = 0; p < (size_t)e->cfg.nptn; p++) {
        const double *lh_cat = e->pattern_lh_cat + p * C;
        double *res = out + p * K;
        if (!marginal) memcpy(res, lh_cat, sizeof(double) * C);
        else if (wsl == 3) { /* :1368-1378 */
            for (int m = 0; m < nmix; m++) {
                double lh = lh_cat[m * ncat];
                for (int c = 1; c < ncat; c++) lh += lh_cat[m * ncat + c];
                res[m] = lh;
            }
        } else { /* :1379-1391 */
            memcpy(res, lh_cat, sizeof(double) * ncat);
            for (int m = 1; m < nmix; m++)
                for (int c = 0; c < ncat; c++) res[c] += lh_cat[m * ncat + c];
        }
    }
    return K;
}

/* The E-step of RateFree::optimizeWithEM (model/ratefree.cpp:556-573; add_invar = 1) and ModelMixture::optimizeWithEM
 * (model/modelmixture.cpp:2008-2024; add_invar = 0 unless optimize_linked_gtr) on the categories of orc_pattern_lh_cat_wsl:
 * posterior[p][k] = ptn_freq[p] * L[p][k] / (ptn_invar[p]*add_invar + sum_k L[p][k]), new_prop[k] = sum_p posterior[p][k]. */
int orc_em_posterior(orc_engine *e, int wsl, int add_invar, double *new_prop, double *posterior) {
    const int K = orc_pattern_lh_cat_wsl(e, wsl, NULL);
    if (K < 0) return K;
    const size_t P = (size_t)e->cfg.nptn;
    double *post = posterior ? posterior : (double *)malloc(sizeof(double) * P * K);
    orc_pattern_lh_cat_wsl(e, wsl, post);
    for (int k = 0; k < K; k++) new_prop[k] = 0.0;
    for (size_t p = 0; p < P; p++) {
        double *this_lk_cat = post + p * K;
        double lk_ptn = add_invar ? e->ptn_invar[p] : 0.0;
        for (int k = 0; k < K; k++) lk_ptn += this_lk_cat[k];
        lk_ptn = e->ptn_freq[p] / lk_ptn;
        for (int k = 0; k < K; k++) { this_lk_cat[k] *= lk_ptn; new_prop[k] += this_lk_cat[k]; }
    }
    if (!posterior) free(post);
    return K;
}

int orc_get_theta(orc_engine *e, double *out) { memcpy(out, e->theta, sizeof(double) * e->cfg.nptn * e->B); return 0; }
int orc_get_buffer_scale_all(orc_engine *e, double *out) { memcpy(out, e->buffer_scale_all, sizeof(double) * e->cfg.nptn); return 0; }
int orc_get_scale_num(orc_engine *e, int slot, uint16_t *out) {
    size_t sz = (size_t)e->cfg.nptn * S_of(e);
    memcpy(out, e->scale + (size_t)slot * sz, sizeof(uint16_t) * sz);
    return 0;
}
/* partial_lh of one slot in the V=1 reference layout [p*B + c*n + i] */
int orc_get_partial(orc_engine *e, int slot, double *out) {
    size_t sz = (size_t)e->cfg.nptn * e->B;
    memcpy(out, e->partial + (size_t)slot * sz, sizeof(double) * sz);
    return 0;
}
int orc_get_tip_partial_lh(orc_engine *e, double *out) {
    memcpy(out, e->tip_partial_lh, sizeof(double) * (e->cfg.state_unknown + 1) * e->cfg.nmix * e->cfg.nstates);
    return 0;
}
#include "lk_oracle_nonrev.h"
#include "lk_oracle_site.h"
#include "lk_oracle_mixlen.h"

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
