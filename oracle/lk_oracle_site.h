/* lk_oracle_site.h — TEST INFRASTRUCTURE (included at the end of lk_oracle.c): CPU restatement of the reference's SITE_MODEL template
 * branch (tree/phylokernelnew.h: partial :1362-1374, 1653-1683, 1772-1801, 1937-1968; buffer :2129-2140; derivative :2306-2311,
 * 2469-2487; branch :2708-2713, 2833-2848, 2993-3008; from-buffer :3271-3276, 3316-3327) — site-specific models
 * (model->isSiteSpecificModel(): ModelSet, i.e. `-fs` site-frequency files / PMSF): every pattern p has its own eigensystem
 * (eval_p, evec_p, inv_evec_p); nothing is precomputed per branch, exp(eval_p[i] * rate_c * len) is taken per pattern, category and
 * call; tip vectors are inv_evec_p . indicator(state) per (taxon, pattern) (computeTipPartialLikelihood, tree/phylotreesse.cpp:257-351);
 * ncat_mix = ncat (no mixture classes).  Arrays here are plain [p][i], [p][x][i], [p][i][x] (the reference interleaves them by its
 * SIMD width; the harness does that conversion). */

int orc_set_site_model(orc_engine *e, const double *eval, const double *evec, const double *inv_evec) {
    int n = e->cfg.nstates;
    size_t P = (size_t)e->cfg.nptn;
    if (e->cfg.nmix != 1) { snprintf(e->err, sizeof e->err, "site-specific models have no mixture classes"); return -1; }
    if (n > ORC_MAX_STATES) { snprintf(e->err, sizeof e->err, "nstates > %d", ORC_MAX_STATES); return -1; }
    free(e->s_eval); free(e->s_evec); free(e->s_inv);
    e->s_eval = (double *)malloc(sizeof(double) * P * n);
    e->s_evec = (double *)malloc(sizeof(double) * P * n * n);
    e->s_inv = (double *)malloc(sizeof(double) * P * n * n);
    memcpy(e->s_eval, eval, sizeof(double) * P * n);
    memcpy(e->s_evec, evec, sizeof(double) * P * n * n);
    memcpy(e->s_inv, inv_evec, sizeof(double) * P * n * n);
    e->site = 1;
    e->theta_computed = 0;
    return 0;
}

/* tip_partial_lh[taxon][p][i] — tree/phylotreesse.cpp:271-331 */
static void site_tip(const orc_engine *e, int taxon, size_t p, double *out) {
    int n = e->cfg.nstates, su = e->cfg.state_unknown;
    const double *inv = e->s_inv + p * n * n;
    int state = tip_state(e, taxon, p);
    if (state < n) { for (int i = 0; i < n; i++) out[i] = inv[i * n + state]; return; }
    if (state == su) {
        for (int i = 0; i < n; i++) { double s = 0.0; for (int x = 0; x < n; x++) s += inv[i * n + x]; out[i] = s; }
        return;
    }
    if (e->cfg.seq_type == 1) {
        int cstate = state - n + 1;
        for (int i = 0; i < n; i++) { double s = 0.0; for (int x = 0; x < n; x++) if (cstate & (1 << x)) s += inv[i * n + x]; out[i] = s; }
    } else {
        static const int ambi_aa[3] = {4 + 8, 32 + 64, 512 + 1024};
        int cstate = state - n;
        for (int i = 0; i < n; i++) { double s = 0.0; for (int x = 0; x < 11; x++) if (ambi_aa[cstate] & (1 << x)) s += inv[i * n + x]; out[i] = s; }
    }
}

static void site_partial_range(orc_engine *e, const orc_op *op_in, size_t lo, size_t hi) {
    int n = e->cfg.nstates, C = e->C, B = e->B, safe = e->cfg.safe_numeric;
    size_t P = (size_t)e->cfg.nptn, S = S_of(e);
    orc_op op = *op_in;
    if (op.left.kind != 0 && op.right.kind == 0) { orc_child t = op.left; op.left = op.right; op.right = t; } /* :1395-1405 */
    double *dad_lh = e->partial + (size_t)op.dst_slot * P * B;
    uint16_t *dad_sc = e->scale + (size_t)op.dst_slot * P * S;
    const double *left_lh = op.left.kind == 1 ? e->partial + (size_t)op.left.id * P * B : NULL;
    const double *right_lh = op.right.kind == 1 ? e->partial + (size_t)op.right.id * P * B : NULL;
    const uint16_t *left_sc = left_lh ? e->scale + (size_t)op.left.id * P * S : NULL;
    const uint16_t *right_sc = right_lh ? e->scale + (size_t)op.right.id * P * S : NULL;
    double vl[ORC_MAX_STATES], vr[ORC_MAX_STATES], el[ORC_MAX_STATES], er[ORC_MAX_STATES], tmp[ORC_MAX_STATES];
    for (size_t p = lo; p < hi; p++) {
        const double *eval = e->s_eval + p * n, *evec = e->s_evec + p * n * n, *inv = e->s_inv + p * n * n;
        double *out = dad_lh + p * B;
        if (!left_lh) site_tip(e, op.left.id, p, vl);
        if (!right_lh) site_tip(e, op.right.id, p, vr);
        const int check = (left_lh || right_lh) && e->ptn_invar[p] == 0.0; /* tip-tip never rescales (:1640) */
        double lh_max = 0.0;
        if (!safe) dad_sc[p] = (uint16_t)((left_sc ? left_sc[p] : 0) + (right_sc ? right_sc[p] : 0));
        for (int c = 0; c < C; c++) {
            const double len_l = e->cat_rate[c] * op.left.length, len_r = e->cat_rate[c] * op.right.length; /* :1365-1366 */
            const double *pl = left_lh ? left_lh + p * B + c * n : vl, *pr = right_lh ? right_lh + p * B + c * n : vr;
            if (safe) { lh_max = 0.0; dad_sc[p * C + c] = (uint16_t)((left_sc ? left_sc[p * C + c] : 0) + (right_sc ? right_sc[p * C + c] : 0)); }
            for (int i = 0; i < n; i++) { el[i] = exp(eval[i] * len_l) * pl[i]; er[i] = exp(eval[i] * len_r) * pr[i]; }
            for (int x = 0; x < n; x++) tmp[x] = dot_dual(evec + x * n, el, evec + x * n, er, n);
            prod_vec_mat(tmp, inv, out + c * n, &lh_max, n);
            if (safe && check && lh_max < SCALING_THRESHOLD) {
                for (int i = 0; i < n; i++) out[c * n + i] = ldexp(out[c * n + i], SCALING_THRESHOLD_EXP);
                dad_sc[p * C + c] += 1;
            }
        }
        if (!safe && check && lh_max < SCALING_THRESHOLD) {
            for (int i = 0; i < B; i++) out[i] = ldexp(out[i], SCALING_THRESHOLD_EXP);
            dad_sc[p] += 1;
        }
    }
}

static int site_update_partials(orc_engine *e, const orc_op *ops, int nops) {
    if (nops <= 0) return 0;
    for (int k = 0; k < nops; k++)
        if (ops[k].nextra > 0) { snprintf(e->err, sizeof e->err, "site-specific models: multifurcating nodes are not restated"); return -1; }
    size_t P = (size_t)e->cfg.nptn;
    long npacket = (long)((P + ORC_PACKET - 1) / ORC_PACKET);
#pragma omp parallel for schedule(dynamic, 1)
    for (long pk = 0; pk < npacket; pk++) {
        size_t lo = (size_t)pk * ORC_PACKET, hi = lo + ORC_PACKET;
        if (hi > P) hi = P;
        for (int k = 0; k < nops; k++) site_partial_range(e, &ops[k], lo, hi);
    }
    e->theta_computed = 0;
    return 0;
}

/* dotProductExp — phylokernelnew.h:811-818 / 835-842 */
static inline double dot_exp3(const double *A, const double *B, const double *C, double D, int N) {
    double X = exp(A[0] * D) * B[0] * C[0];
    for (int i = 1; i < N; i++) X = fma(exp(A[i] * D), B[i] * C[i], X);
    return X;
}
static inline double dot_exp2(const double *A, const double *B, double D, int N) {
    double X = exp(A[0] * D) * B[0];
    for (int i = 1; i < N; i++) X = fma(exp(A[i] * D), B[i], X);
    return X;
}

static int site_branch_logl(orc_engine *e, int node_slot, int dad_kind, int dad_id, double length, int save_log_value, double *logl_out) {
    int n = e->cfg.nstates, C = e->C, B = e->B, safe = e->cfg.safe_numeric;
    size_t P = (size_t)e->cfg.nptn, S = S_of(e);
    if (e->asc_type) { snprintf(e->err, sizeof e->err, "site-specific models: +ASC is not restated"); return -1; }
    const double *dad_lh = e->partial + (size_t)node_slot * P * B;
    const uint16_t *dad_sc = e->scale + (size_t)node_slot * P * S;
    const double *node_lh = dad_kind == 1 ? e->partial + (size_t)dad_id * P * B : NULL;
    const uint16_t *node_sc = dad_kind == 1 ? e->scale + (size_t)dad_id * P * S : NULL;
    double tree_lh = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : tree_lh)
    for (size_t p = 0; p < P; p++) {
        double tip[ORC_MAX_STATES];
        const double *eval = e->s_eval + p * n, *pd = dad_lh + p * B;
        if (!node_lh) site_tip(e, dad_id, p, tip);
        double *lh_cat = e->pattern_lh_cat + p * C;
        double lh_ptn = 0.0, min_scale_d;
        for (int c = 0; c < C; c++) { /* :2833-2848 / 2993-3008 */
            const double *pn = node_lh ? node_lh + p * B + c * n : tip;
            lh_cat[c] = dot_exp3(eval, pn, pd + c * n, e->cat_rate[c] * length, n) * e->cat_prop[c];
            if (!safe) lh_ptn += lh_cat[c];
        }
        if (safe) {
            uint16_t min_scale = 0xffff;
            for (int c = 0; c < C; c++) { uint16_t s = (uint16_t)(dad_sc[p * C + c] + (node_sc ? node_sc[p * C + c] : 0)); if (s < min_scale) min_scale = s; }
            min_scale_d = min_scale;
            for (int c = 0; c < C; c++) {
                uint16_t s = (uint16_t)(dad_sc[p * C + c] + (node_sc ? node_sc[p * C + c] : 0));
                if (s == min_scale + 1) lh_cat[c] *= SCALING_THRESHOLD;
                else if (s > min_scale + 1) lh_cat[c] = 0.0;
            }
            lh_ptn = sum_vec_append(lh_cat, lh_ptn, C);
        } else min_scale_d = (double)(dad_sc[p] + (node_sc ? node_sc[p] : 0));
        min_scale_d *= LOG_SCALING_THRESHOLD;
        lh_ptn = fabs(lh_ptn) + e->ptn_invar[p];
        double lg = log(lh_ptn) + min_scale_d;
        if (save_log_value) e->pattern_lh[p] = lg;
        else { e->pattern_lh[p] = lh_ptn; e->pattern_scaling[p] = min_scale_d; }
        tree_lh += lg * e->ptn_freq[p];
    }
    int status = 0;
    if (!isfinite(tree_lh)) {
        status = 1;
        tree_lh = 0.0;
        for (size_t p = 0; p < P; p++) {
            double ptn_lh = e->pattern_lh[p];
            if (!isfinite(ptn_lh)) {
                if (save_log_value) e->pattern_lh[p] = LOG_SCALING_THRESHOLD * 4;
                else { e->pattern_lh[p] = 1.0; e->pattern_scaling[p] = LOG_SCALING_THRESHOLD * 4; }
                ptn_lh = LOG_SCALING_THRESHOLD * 4;
            }
            tree_lh += ptn_lh * e->ptn_freq[p];
        }
    }
    *logl_out = tree_lh;
    return status;
}

/* computeLikelihoodBuffer with per-pattern tip vectors (:2129-2140); the internal-internal form is the normal model's */
static void site_build_theta(orc_engine *e, int node_slot, int dad_kind, int dad_id) {
    int n = e->cfg.nstates, C = e->C, B = e->B, safe = e->cfg.safe_numeric;
    size_t P = (size_t)e->cfg.nptn, S = S_of(e);
    const double *dad_lh = e->partial + (size_t)node_slot * P * B;
    const uint16_t *dad_sc = e->scale + (size_t)node_slot * P * S;
    const double *node_lh = dad_kind == 1 ? e->partial + (size_t)dad_id * P * B : NULL;
    const uint16_t *node_sc = dad_kind == 1 ? e->scale + (size_t)dad_id * P * S : NULL;
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < P; p++) {
        double tip[ORC_MAX_STATES];
        if (!node_lh) site_tip(e, dad_id, p, tip);
        const double *pd = dad_lh + p * B;
        double *theta = e->theta + p * B;
        for (int c = 0; c < C; c++)
            for (int i = 0; i < n; i++) theta[c * n + i] = (node_lh ? node_lh[p * B + c * n + i] : tip[i]) * pd[c * n + i];
        if (safe) {
            uint16_t min_scale = 0xffff;
            for (int c = 0; c < C; c++) { uint16_t s = (uint16_t)(dad_sc[p * C + c] + (node_sc ? node_sc[p * C + c] : 0)); if (s < min_scale) min_scale = s; }
            e->buffer_scale_all[p] = min_scale;
            for (int c = 0; c < C; c++) {
                uint16_t s = (uint16_t)(dad_sc[p * C + c] + (node_sc ? node_sc[p * C + c] : 0));
                if (s == min_scale + 1) for (int x = 0; x < n; x++) theta[c * n + x] *= SCALING_THRESHOLD;
                else if (s > min_scale + 1) for (int x = 0; x < n; x++) theta[c * n + x] = 0.0;
            }
        } else e->buffer_scale_all[p] = (double)(dad_sc[p] + (node_sc ? node_sc[p] : 0));
        e->buffer_scale_all[p] *= LOG_SCALING_THRESHOLD;
    }
    e->theta_computed = 1;
}

/* computeLikelihoodDerv, SITE_MODEL (:2469-2487) */
static int site_branch_derv(orc_engine *e, int node_slot, int dad_kind, int dad_id, double length, double *df_out, double *ddf_out) {
    if (!e->theta_computed) site_build_theta(e, node_slot, dad_kind, dad_id);
    int n = e->cfg.nstates, C = e->C, B = e->B;
    size_t P = (size_t)e->cfg.nptn;
    double all_df = 0.0, all_ddf = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : all_df, all_ddf)
    for (size_t p = 0; p < P; p++) {
        const double *eval = e->s_eval + p * n, *theta = e->theta + p * B;
        double lh_ptn = 0.0, df_ptn = 0.0, ddf_ptn = 0.0;
        for (int c = 0; c < C; c++) {
            double lh_cat = 0.0, df_cat = 0.0, ddf_cat = 0.0;
            for (int i = 0; i < n; i++) {
                double cof = eval[i] * e->cat_rate[c];
                double val = exp(cof * length) * theta[c * n + i];
                double val1 = cof * val;
                lh_cat += val;
                df_cat += val1;
                ddf_cat = fma(cof, val1, ddf_cat);
            }
            lh_ptn = fma(e->cat_prop[c], lh_cat, lh_ptn);
            df_ptn = fma(e->cat_prop[c], df_cat, df_ptn);
            ddf_ptn = fma(e->cat_prop[c], ddf_cat, ddf_ptn);
        }
        lh_ptn = fabs(lh_ptn) + e->ptn_invar[p];
        lh_ptn = 1.0 / lh_ptn;
        double df_frac = df_ptn * lh_ptn, ddf_frac = ddf_ptn * lh_ptn, freq = e->ptn_freq[p];
        double tmp1 = df_frac * freq, tmp2 = ddf_frac * freq;
        all_df += tmp1;
        all_ddf += fma(-tmp1, df_frac, tmp2);
    }
    *df_out = all_df; *ddf_out = all_ddf;
    if (!isfinite(all_df)) { *df_out = *ddf_out = 0.0; return 1; }
    return 0;
}

/* computeLikelihoodFromBuffer, SITE_MODEL (:3316-3327) */
static int site_logl_from_buffer(orc_engine *e, double length, double *logl_out) {
    int n = e->cfg.nstates, C = e->C, B = e->B;
    size_t P = (size_t)e->cfg.nptn;
    double tree_lh = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : tree_lh)
    for (size_t p = 0; p < P; p++) {
        const double *eval = e->s_eval + p * n, *theta = e->theta + p * B;
        double lh_ptn = 0.0;
        for (int c = 0; c < C; c++) lh_ptn = fma(dot_exp2(eval, theta + c * n, e->cat_rate[c] * length, n), e->cat_prop[c], lh_ptn);
        lh_ptn = fabs(lh_ptn) + e->ptn_invar[p];
        lh_ptn = log(fabs(lh_ptn)) + e->buffer_scale_all[p];
        e->pattern_lh[p] = lh_ptn;
        tree_lh += lh_ptn * e->ptn_freq[p];
    }
    int status = 0;
    if (!isfinite(tree_lh)) {
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
