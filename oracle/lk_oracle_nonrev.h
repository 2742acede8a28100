/* lk_oracle_nonrev.h — TEST INFRASTRUCTURE (included at the end of lk_oracle.c): CPU restatement of the reference's kernel for
 * NON-REVERSIBLE models, tree/phylokernelnonrev.h (selected when !model->useRevKernel(): a non-reversible model, or any model with
 * --kernel-nonrev, model/modelsubst.h:73-75; pointers assigned in tree/phylokernelfma.cpp through setLikelihoodKernel).
 *
 *   - CLVs live in STATE space; per child and category the operator is the transition matrix P_m(rate_c * len) itself
 *     (computePartialInfo's non-reversible branch, tree/phylokernelnew.h:932-1008), TRANSPOSED when the child's branch points
 *     towards the root (child->direction == TOWARD_ROOT, :939-951);
 *   - the virtual root leaf of a rooted tree (MTree::isRootLeaf, tree/mtree.h:135) contributes the state frequencies (:959-964);
 *   - tip_partial_lh[state][x] are the plain state indicators (tree/phylotreesse.cpp:353-375 with nmixtures = 1);
 *   - the underflow test is on max(value) (signed), not on max|value| (phylokernelnonrev.h:446-447, 520);
 *   - branch likelihood :1062-1528, derivatives :586-1056 (no theta buffer: both ends' CLVs are read at every call).
 * Transition matrices: P = Re(cevec . diag(exp(ceval t)) . cinv_evec), the reference's default technique for non-reversible models
 * (ModelMarkov::computeTransMatrixNonrev, model/modelmarkov.cpp:477-486); derivatives Q.P and Q.Q.P (:735-747).
 */

#define ORC_KIND_MASK 3
#define ORC_KIND_ROOT 2         /* child / dad kind: the virtual root leaf */
#define ORC_CHILD_TRANSPOSED 4  /* flag or-ed into orc_child.kind: the child's branch points towards the root */

int orc_set_model_nonrev(orc_engine *e, const double *ceval, const double *cevec, const double *cinv_evec, const double *rate_matrix,
                         const double *state_freq, const double *mix_weight, int rooted) {
    int n = e->cfg.nstates, nmix = e->cfg.nmix, su = e->cfg.state_unknown;
    if (n > ORC_MAX_STATES) { snprintf(e->err, sizeof e->err, "nstates > %d", ORC_MAX_STATES); return -1; }
    free(e->ceval); free(e->cevec); free(e->cinv); free(e->qmat); free(e->sfreq);
    e->ceval = (double *)malloc(sizeof(double) * 2 * nmix * n);
    e->cevec = (double *)malloc(sizeof(double) * 2 * nmix * n * n);
    e->cinv = (double *)malloc(sizeof(double) * 2 * nmix * n * n);
    e->qmat = (double *)malloc(sizeof(double) * nmix * n * n);
    e->sfreq = (double *)malloc(sizeof(double) * nmix * n);
    memcpy(e->ceval, ceval, sizeof(double) * 2 * nmix * n);
    memcpy(e->cevec, cevec, sizeof(double) * 2 * nmix * n * n);
    memcpy(e->cinv, cinv_evec, sizeof(double) * 2 * nmix * n * n);
    memcpy(e->qmat, rate_matrix, sizeof(double) * nmix * n * n);
    memcpy(e->sfreq, state_freq, sizeof(double) * nmix * n);
    memcpy(e->mix_weight, mix_weight, sizeof(double) * nmix);
    e->nonrev = 1;
    e->rooted = rooted != 0;
    /* computeTipPartialLikelihood, non-reversible: the indicators themselves, one block (tree/phylotreesse.cpp:353-375) */
    for (int state = 0; state <= su; state++) tip_indicator(e, state, e->tip_partial_lh + (size_t)state * n);
    e->theta_computed = 0;
    return 0;
}

/* ModelMarkov::computeTransMatrixNonrev (model/modelmarkov.cpp:477-486) */
static void nonrev_trans_matrix(const orc_engine *e, int m, double t, double *mat) {
    int n = e->cfg.nstates;
    const double *ev = e->ceval + (size_t)m * n * 2, *V = e->cevec + (size_t)m * n * n * 2, *W = e->cinv + (size_t)m * n * n * 2;
    double er[ORC_MAX_STATES], ei[ORC_MAX_STATES];
    for (int k = 0; k < n; k++) {
        double a = exp(ev[2 * k] * t);
        er[k] = a * cos(ev[2 * k + 1] * t);
        ei[k] = a * sin(ev[2 * k + 1] * t);
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int k = 0; k < n; k++) {
                double vr = V[(size_t)(i * n + k) * 2], vi = V[(size_t)(i * n + k) * 2 + 1];
                double ar = vr * er[k] - vi * ei[k], ai = vr * ei[k] + vi * er[k];
                s += ar * W[(size_t)(k * n + j) * 2] - ai * W[(size_t)(k * n + j) * 2 + 1];
            }
            mat[i * n + j] = s;
        }
}

/* ModelMarkov::computeTransDerv, non-reversible (model/modelmarkov.cpp:735-747): Q.P and Q.Q.P */
static void nonrev_trans_derv(const orc_engine *e, int m, double t, double *mat, double *d1, double *d2) {
    int n = e->cfg.nstates;
    const double *Q = e->qmat + (size_t)m * n * n;
    nonrev_trans_matrix(e, m, t, mat);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int k = 0; k < n; k++) s += Q[i * n + k] * mat[k * n + j];
            d1[i * n + j] = s;
        }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int k = 0; k < n; k++) s += Q[i * n + k] * d1[k * n + j];
            d2[i * n + j] = s;
        }
}

/* computePartialInfo, non-reversible branch (phylokernelnew.h:932-1008) for one child */
static void nonrev_partial_info_child(const orc_engine *e, const orc_child *ch, double *echild, double *leaf /* or NULL */) {
    int n = e->cfg.nstates, C = e->C, B = e->B, ncat = e->cfg.ncat, su = e->cfg.state_unknown;
    int kind = ch->kind & ORC_KIND_MASK;
    double *mat = (double *)malloc(sizeof(double) * n * n);
    for (int c = 0; c < C; c++) {
        int m = c / e->denom;
        double len_child = e->cat_rate[c % ncat] * ch->length;
        double *ep = echild + (size_t)c * n * n;
        if (ch->kind & ORC_CHILD_TRANSPOSED) {
            nonrev_trans_matrix(e, m, len_child, mat);
            for (int i = 0; i < n; i++)
                for (int x = 0; x < n; x++) ep[i * n + x] = mat[x * n + i]; /* :945-949 */
        } else nonrev_trans_matrix(e, m, len_child, ep);
    }
    free(mat);
    if (!leaf) return;
    if (kind == ORC_KIND_ROOT) { /* :959-964: state frequencies; only row 0 of the table is ever used */
        for (int c = 0; c < C; c++) memcpy(leaf + (size_t)c * n, e->sfreq + (size_t)(c / e->denom) * n, sizeof(double) * n);
        return;
    }
    for (int state = 0; state < su; state++) {
        const double *tip = e->tip_partial_lh + (size_t)state * n;
        for (int x = 0; x < B; x++) {
            const double *ep = echild + (size_t)x * n;
            double v;
            if (n % 4 == 0) { /* vectorised (Vec4d) branch :967-979: lane accumulators, no FMA, then horizontal_add */
                double a0 = ep[0] * tip[0], a1 = ep[1] * tip[1], a2 = ep[2] * tip[2], a3 = ep[3] * tip[3];
                for (int i = 4; i < n; i += 4) { a0 += ep[i] * tip[i]; a1 += ep[i + 1] * tip[i + 1]; a2 += ep[i + 2] * tip[i + 2]; a3 += ep[i + 3] * tip[i + 3]; }
                v = (a0 + a1) + (a2 + a3);
            } else { /* :981-994 */
                v = 0.0;
                for (int i = 0; i < n; i++) v += ep[i] * tip[i];
            }
            leaf[(size_t)state * B + x] = v;
        }
    }
    for (int x = 0; x < B; x++) leaf[(size_t)su * B + x] = 1.0; /* :996-1000 */
}

typedef struct {
    orc_op op; /* after the swaps of phylokernelnonrev.h:93-101, 259-272: a single (or the root) leaf child is `left` */
    double *echildren, *leaves, *eleft, *eright, *leafL, *leafR;
} orc_nr_opinfo;

static void nonrev_opinfo_init(const orc_engine *e, const orc_op *op_in, orc_nr_opinfo *oi) {
    int n = e->cfg.nstates, B = e->B;
    size_t info_sz = (size_t)B * n, leaf_sz = (size_t)(e->cfg.state_unknown + 1) * B;
    oi->op = *op_in;
    oi->echildren = (double *)malloc(sizeof(double) * 2 * info_sz);
    oi->leaves = (double *)malloc(sizeof(double) * 2 * leaf_sz);
    oi->leafL = (oi->op.left.kind & ORC_KIND_MASK) != 1 ? oi->leaves : NULL;
    oi->leafR = (oi->op.right.kind & ORC_KIND_MASK) != 1 ? oi->leaves + leaf_sz : NULL;
    nonrev_partial_info_child(e, &oi->op.left, oi->echildren, oi->leafL);
    nonrev_partial_info_child(e, &oi->op.right, oi->echildren + info_sz, oi->leafR);
    oi->eleft = oi->echildren; oi->eright = oi->echildren + info_sz;
    int lk = oi->op.left.kind & ORC_KIND_MASK, rk = oi->op.right.kind & ORC_KIND_MASK;
    int swap = (lk == 1 && rk != 1);                /* :93-101 the leaf becomes `left` */
    if (lk != 1 && rk == ORC_KIND_ROOT) swap = 1;   /* :259-272 the root leaf becomes `left` */
    if (swap) {
        orc_child t = oi->op.left; oi->op.left = oi->op.right; oi->op.right = t;
        double *et = oi->eleft; oi->eleft = oi->eright; oi->eright = et;
        et = oi->leafL; oi->leafL = oi->leafR; oi->leafR = et;
    }
}

static inline int nr_state(const orc_engine *e, const orc_child *ch, size_t p) {
    return (ch->kind & ORC_KIND_MASK) == ORC_KIND_ROOT ? 0 : tip_state(e, ch->id, p);
}

static void nonrev_compute_partial_range(orc_engine *e, const orc_nr_opinfo *oi, size_t lo, size_t hi) {
    int n = e->cfg.nstates, C = e->C, B = e->B, safe = e->cfg.safe_numeric;
    size_t P = (size_t)e->cfg.nptn, S = S_of(e);
    const orc_op *op = &oi->op;
    int lk = op->left.kind & ORC_KIND_MASK, rk = op->right.kind & ORC_KIND_MASK;
    double *dad_lh = e->partial + (size_t)op->dst_slot * P * B;
    uint16_t *dad_sc = e->scale + (size_t)op->dst_slot * P * S;
    if (lk != 1 && rk != 1) {
        /* TIP-TIP :244-350 (one of them may be the root leaf) — scale_num zero */
        memset(dad_sc + lo * S, 0, sizeof(uint16_t) * (hi - lo) * S);
        for (size_t p = lo; p < hi; p++) {
            const double *vl = oi->leafL + (size_t)nr_state(e, &op->left, p) * B;
            const double *vr = oi->leafR + (size_t)nr_state(e, &op->right, p) * B;
            double *out = dad_lh + p * B;
            for (int i = 0; i < B; i++) out[i] = vl[i] * vr[i];
        }
    } else if (lk == ORC_KIND_ROOT) {
        /* ROOT-INTERNAL :353-388 — scale counts of the right subtree, no underflow test */
        const double *right_lh = e->partial + (size_t)op->right.id * P * B;
        const uint16_t *right_sc = e->scale + (size_t)op->right.id * P * S;
        memcpy(dad_sc + lo * S, right_sc + lo * S, sizeof(uint16_t) * (hi - lo) * S);
        for (size_t p = lo; p < hi; p++) {
            const double *pr = right_lh + p * B;
            double *out = dad_lh + p * B;
            for (int c = 0; c < C; c++)
                for (int x = 0; x < n; x++)
                    out[c * n + x] = oi->leafL[c * n + x] * dot_vec(oi->eright + (size_t)(c * n + x) * n, pr + c * n, n);
        }
    } else if (lk == 0) {
        /* TIP-INTERNAL :390-488 */
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
                for (int x = 0; x < n; x++) {
                    double v = vl[c * n + x] * dot_vec(oi->eright + (size_t)(c * n + x) * n, pr + c * n, n);
                    out[c * n + x] = v;
                    if (v > lh_max) lh_max = v;
                }
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
        /* INTERNAL-INTERNAL :490-574 */
        const double *left_lh = e->partial + (size_t)op->left.id * P * B, *right_lh = e->partial + (size_t)op->right.id * P * B;
        const uint16_t *left_sc = e->scale + (size_t)op->left.id * P * S, *right_sc = e->scale + (size_t)op->right.id * P * S;
        for (size_t p = lo; p < hi; p++) {
            const double *pl = left_lh + p * B, *pr = right_lh + p * B;
            double *out = dad_lh + p * B;
            double lh_max = 0.0;
            if (!safe) dad_sc[p] = (uint16_t)(left_sc[p] + right_sc[p]);
            for (int c = 0; c < C; c++) {
                if (safe) { lh_max = 0.0; dad_sc[p * C + c] = (uint16_t)(left_sc[p * C + c] + right_sc[p * C + c]); }
                for (int x = 0; x < n; x++) {
                    double v = dot_dual(oi->eleft + (size_t)(c * n + x) * n, pl + c * n, oi->eright + (size_t)(c * n + x) * n, pr + c * n, n);
                    out[c * n + x] = v;
                    if (v > lh_max) lh_max = v;
                }
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

static int nonrev_update_partials(orc_engine *e, const orc_op *ops, int nops) {
    if (nops <= 0) return 0;
    for (int k = 0; k < nops; k++)
        if (ops[k].nextra > 0) { snprintf(e->err, sizeof e->err, "non-reversible kernel: multifurcating nodes are not restated"); return -1; }
    orc_nr_opinfo *info = (orc_nr_opinfo *)malloc(sizeof(orc_nr_opinfo) * nops);
#pragma omp parallel for schedule(static) if (nops >= 3)
    for (int k = 0; k < nops; k++) nonrev_opinfo_init(e, &ops[k], &info[k]);
    size_t P = (size_t)e->cfg.nptn;
    long npacket = (long)((P + ORC_PACKET - 1) / ORC_PACKET);
#pragma omp parallel for schedule(dynamic, 1)
    for (long pk = 0; pk < npacket; pk++) {
        size_t lo = (size_t)pk * ORC_PACKET, hi = lo + ORC_PACKET;
        if (hi > P) hi = P;
        for (int k = 0; k < nops; k++) nonrev_compute_partial_range(e, &info[k], lo, hi);
    }
    for (int k = 0; k < nops; k++) { free(info[k].echildren); free(info[k].leaves); }
    free(info);
    return 0;
}

/* trans / derv1 / derv2 of the evaluation branch, [c][i][x], weighted as phylokernelnonrev.h:628-659 / 1103-1124 */
static void nonrev_branch_matrices(const orc_engine *e, double length, double *tm, double *d1, double *d2) {
    int n = e->cfg.nstates, C = e->C, ncat = e->cfg.ncat, nn = n * n;
    for (int c = 0; c < C; c++) {
        int m = c / e->denom, r = c % ncat;
        double cat_rate = e->cat_rate[r], len = cat_rate * length, prop = e->cat_prop[r] * e->mix_weight[m];
        double *t0 = tm + (size_t)c * nn, *t1 = d1 ? d1 + (size_t)c * nn : NULL, *t2 = d2 ? d2 + (size_t)c * nn : NULL;
        if (d1) {
            nonrev_trans_derv(e, m, len, t0, t1, t2);
            double prop_rate = prop * cat_rate, prop_rate_2 = prop_rate * cat_rate;
            for (int i = 0; i < nn; i++) { t0[i] *= prop; t1[i] *= prop_rate; t2[i] *= prop_rate_2; }
        } else {
            nonrev_trans_matrix(e, m, len, t0);
            for (int i = 0; i < nn; i++) t0[i] *= prop;
        }
        if (!e->rooted) {
            const double *f = e->sfreq + (size_t)m * n;
            for (int i = 0; i < n; i++)
                for (int x = 0; x < n; x++) {
                    t0[i * n + x] *= f[i];
                    if (d1) { t1[i * n + x] *= f[i]; t2[i * n + x] *= f[i]; }
                }
        }
    }
}

/* min over the categories of a SAFE_LH scale vector; v (a copy of the CLV) is aligned to it (:706-722) */
static unsigned nonrev_align(const orc_engine *e, const uint16_t *sc, double *v) {
    int n = e->cfg.nstates, C = e->C;
    unsigned min_scale = sc[0];
    for (int c = 1; c < C; c++) if (sc[c] < min_scale) min_scale = sc[c];
    for (int c = 0; c < C; c++) {
        if (sc[c] == min_scale + 1) for (int x = 0; x < n; x++) v[c * n + x] *= SCALING_THRESHOLD;
        else if (sc[c] > min_scale + 1) for (int x = 0; x < n; x++) v[c * n + x] = 0.0;
    }
    return min_scale;
}

/* computeNonrevLikelihoodBranch — phylokernelnonrev.h:1062-1528.  node_slot = dad_branch->partial_lh (the root-ward end after the
 * reference's swap :1071-1078); dad = the other end: a leaf (kind 0), the root leaf (kind 2) or a CLV slot (kind 1). */
static int nonrev_branch_logl(orc_engine *e, int node_slot, int dad_kind, int dad_id, double length, double *logl_out) {
    int n = e->cfg.nstates, C = e->C, B = e->B, safe = e->cfg.safe_numeric, su = e->cfg.state_unknown;
    size_t P = (size_t)e->cfg.nptn, S = S_of(e);
    if (e->asc_type) { snprintf(e->err, sizeof e->err, "non-reversible kernel: +ASC is not restated"); return -1; }
    double *tm = (double *)malloc(sizeof(double) * (size_t)B * n);
    nonrev_branch_matrices(e, length, tm, NULL, NULL);
    const double *dad_lh = e->partial + (size_t)node_slot * P * B;
    const uint16_t *dad_sc = e->scale + (size_t)node_slot * P * S;
    double tree_lh = 0.0;
    dad_kind &= ORC_KIND_MASK;
    if (dad_kind != 1) {
        double *plnode = (double *)calloc((size_t)(su + 1) * B, sizeof(double));
        if (dad_kind == ORC_KIND_ROOT) { /* :1150-1158 */
            for (int c = 0; c < C; c++) {
                int m = c / e->denom;
                double prop = e->cat_prop[c % e->cfg.ncat] * e->mix_weight[m];
                for (int i = 0; i < n; i++) plnode[c * n + i] = e->sfreq[(size_t)m * n + i] * prop;
            }
        } else { /* :1160-1176 */
            for (int state = 0; state <= su; state++) {
                const double *tip = e->tip_partial_lh + (size_t)state * n;
                for (int ci = 0; ci < B; ci++) {
                    double v = 0.0;
                    for (int x = 0; x < n; x++) v += tm[(size_t)ci * n + x] * tip[x];
                    plnode[(size_t)state * B + ci] = v;
                }
            }
        }
#pragma omp parallel for schedule(static) reduction(+ : tree_lh)
        for (size_t p = 0; p < P; p++) {
            const double *lh_node = plnode + (size_t)(dad_kind == ORC_KIND_ROOT ? 0 : tip_state(e, dad_id, p)) * B;
            const double *pd = dad_lh + p * B;
            double *lh_cat = e->pattern_lh_cat + p * C;
            double lh_ptn = 0.0, min_scale_d;
            for (int c = 0; c < C; c++) {
                lh_cat[c] = dot_vec(lh_node + c * n, pd + c * n, n);
                if (!safe) lh_ptn += lh_cat[c];
            }
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
            lh_ptn = lh_ptn + e->ptn_invar[p];
            double lg = log(lh_ptn) + min_scale_d;
            e->pattern_lh[p] = lg;
            tree_lh += lg * e->ptn_freq[p];
        }
        free(plnode);
    } else {
        const double *node_lh = e->partial + (size_t)dad_id * P * B;
        const uint16_t *node_sc = e->scale + (size_t)dad_id * P * S;
#pragma omp parallel for schedule(static) reduction(+ : tree_lh)
        for (size_t p = 0; p < P; p++) {
            const double *pd = dad_lh + p * B, *pn = node_lh + p * B;
            double *lh_cat = e->pattern_lh_cat + p * C;
            double lh_ptn = 0.0, min_scale_d;
            for (int c = 0; c < C; c++) { /* :1317-1334 */
                double acc = 0.0;
                for (int i = 0; i < n; i++) acc = fma(pd[c * n + i], dot_vec(tm + (size_t)(c * n + i) * n, pn + c * n, n), acc);
                lh_cat[c] = acc;
                if (!safe) lh_ptn += acc;
            }
            if (safe) { /* :1369-1392 */
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
            lh_ptn = lh_ptn + e->ptn_invar[p];
            double lg = log(lh_ptn) + min_scale_d;
            e->pattern_lh[p] = lg;
            tree_lh += lg * e->ptn_freq[p];
        }
    }
    free(tm);
    int status = 0;
    if (!isfinite(tree_lh)) { /* :1476-1494 */
        status = 1;
        tree_lh = 0.0;
        for (size_t p = 0; p < P; p++) {
            if (!isfinite(e->pattern_lh[p])) e->pattern_lh[p] = LOG_SCALING_THRESHOLD * 4;
            tree_lh += e->pattern_lh[p] * e->ptn_freq[p];
        }
    }
    *logl_out = tree_lh;
    return status;
}

/* computeNonrevLikelihoodDerv — phylokernelnonrev.h:586-1056 */
static int nonrev_branch_derv(orc_engine *e, int node_slot, int dad_kind, int dad_id, double length, double *df_out, double *ddf_out) {
    int n = e->cfg.nstates, C = e->C, B = e->B, safe = e->cfg.safe_numeric, su = e->cfg.state_unknown;
    size_t P = (size_t)e->cfg.nptn, S = S_of(e);
    if (e->asc_type) { snprintf(e->err, sizeof e->err, "non-reversible kernel: +ASC is not restated"); return -1; }
    double *tm = (double *)malloc(sizeof(double) * 3 * (size_t)B * n), *d1 = tm + (size_t)B * n, *d2 = d1 + (size_t)B * n;
    nonrev_branch_matrices(e, length, tm, d1, d2);
    const double *dad_lh = e->partial + (size_t)node_slot * P * B;
    const uint16_t *dad_sc = e->scale + (size_t)node_slot * P * S;
    double all_df = 0.0, all_ddf = 0.0;
    dad_kind &= ORC_KIND_MASK;
    double *tab = NULL;
    if (dad_kind != 1) {
        tab = (double *)calloc((size_t)3 * (su + 1) * B, sizeof(double));
        double *t0 = tab, *t1 = tab + (size_t)(su + 1) * B, *t2 = t1 + (size_t)(su + 1) * B;
        if (dad_kind == ORC_KIND_ROOT) {
            /* :677-691 — as written the derivative tables of the root leaf are never initialised before `*= prop` (lh_derv1 and
             * lh_derv2 are read uninitialised); a branch to the root leaf has no length to optimise and the reference never calls
             * computeLikelihoodDerv there: zeros here */
            for (int c = 0; c < C; c++) {
                int m = c / e->denom;
                double prop = e->cat_prop[c % e->cfg.ncat] * e->mix_weight[m];
                for (int i = 0; i < n; i++) t0[c * n + i] = e->sfreq[(size_t)m * n + i] * prop;
            }
        } else { /* :693-724 */
            for (int state = 0; state <= su; state++) {
                const double *tip = e->tip_partial_lh + (size_t)state * n;
                for (int ci = 0; ci < B; ci++) {
                    double a = 0.0, b = 0.0, c2 = 0.0;
                    for (int x = 0; x < n; x++) { a += tm[(size_t)ci * n + x] * tip[x]; b += d1[(size_t)ci * n + x] * tip[x]; c2 += d2[(size_t)ci * n + x] * tip[x]; }
                    t0[(size_t)state * B + ci] = a; t1[(size_t)state * B + ci] = b; t2[(size_t)state * B + ci] = c2;
                }
            }
        }
    }
    const double *node_lh = dad_kind == 1 ? e->partial + (size_t)dad_id * P * B : NULL;
    const uint16_t *node_sc = dad_kind == 1 ? e->scale + (size_t)dad_id * P * S : NULL;
#pragma omp parallel for schedule(static) reduction(+ : all_df, all_ddf)
    for (size_t p = 0; p < P; p++) {
        double bufd[ORC_MAX_STATES * 16], bufn[ORC_MAX_STATES * 16];
        double *vd = B <= ORC_MAX_STATES * 16 ? bufd : (double *)malloc(sizeof(double) * B);
        double *vn = B <= ORC_MAX_STATES * 16 ? bufn : (double *)malloc(sizeof(double) * B);
        memcpy(vd, dad_lh + p * B, sizeof(double) * B);
        if (safe) nonrev_align(e, dad_sc + p * C, vd);
        double lh = 0.0, df = 0.0, ddf = 0.0;
        if (dad_kind != 1) {
            size_t st = (size_t)(dad_kind == ORC_KIND_ROOT ? 0 : tip_state(e, dad_id, p)) * B;
            const double *t0 = tab + st, *t1 = tab + (size_t)(su + 1) * B + st, *t2 = tab + (size_t)2 * (su + 1) * B + st;
            dot_triple(t0, t1, t2, vd, (size_t)B, (e->quirks & 1) ? (n & 1) : (B & 1), &lh, &df, &ddf); /* :813-817, over the whole block */
        } else {
            memcpy(vn, node_lh + p * B, sizeof(double) * B);
            if (safe) nonrev_align(e, node_sc + p * C, vn);
            for (int c = 0; c < C; c++)
                for (int i = 0; i < n; i++) { /* :951-966 */
                    double a, b, c2;
                    const size_t row = (size_t)(c * n + i) * n;
                    dot_triple(tm + row, d1 + row, d2 + row, vn + c * n, (size_t)n, n & 1, &a, &b, &c2);
                    lh = fma(vd[c * n + i], a, lh); df = fma(vd[c * n + i], b, df); ddf = fma(vd[c * n + i], c2, ddf);
                }
        }
        lh = lh + e->ptn_invar[p];
        lh = 1.0 / lh;
        double df_frac = df * lh, ddf_frac = ddf * lh, f = e->ptn_freq[p];
        double tmp1 = df_frac * f, tmp2 = ddf_frac * f;
        all_df += tmp1;
        all_ddf += fma(-tmp1, df_frac, tmp2); /* nmul_add(tmp1, df_frac, tmp2) */
        if (vd != bufd) { free(vd); free(vn); }
    }
    free(tm); free(tab);
    int status = 0;
    if (!isfinite(all_df)) { status = 2; all_df = all_ddf = 0.0; } /* :1051-1055 */
    *df_out = all_df; *ddf_out = all_ddf;
    return status;
}
