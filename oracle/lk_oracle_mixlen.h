/* lk_oracle_mixlen.h — TEST INFRASTRUCTURE (included at the end of lk_oracle.c): CPU restatement of the reference's mixed-branch-length
 * (heterotachy, GHOST `+H` / `*H`) paths of tree/phylokernelnew.h: PhyloTree::isMixlen() trees carry nmixlen = ncat lengths per
 * branch (PhyloNeighborMixlen::getLength(c), tree/phylonodemixlen.h:64-69).
 *   - every place that reads a branch length takes the category's own (computePartialInfo :1017, 1069; computeLikelihoodDerv's val0/1/2
 *     :2325, 2348; computeLikelihoodBranch :2724, 2738; computeLikelihoodFromBuffer :3286, 3301) — len_of() in lk_oracle.c;
 *   - computeLikelihoodDerv returns the GRADIENT over the nmixlen lengths, the HESSIAN, and the log-likelihood in df[nmixlen]
 *     (:2391-2460, 2556-2568) — orc_branch_derv_mixlen;
 *   - computeLikelihoodDervMixlen (:3431-3616): (df, ddf) of the CURRENT class's length alone, as the EM algorithm of
 *     PhyloTreeMixlen::optimizeOneBranch uses it (ptn_freq then holds the class's posterior weights) — orc_branch_derv_mixlen_cur.
 * +ASC is not supported there (the reference asserts, :2449). */

/* table [rows][ncat]; rows = 0 switches the mode off.  Afterwards `length` in ops and branch calls is the row of the branch. */
int orc_set_mixlen(orc_engine *e, const double *table, int rows) {
    free(e->mixlen_tab);
    e->mixlen_tab = NULL;
    e->mixlen = 0; e->mixlen_rows = 0; /* (CLVs and the theta buffer are the caller's to invalidate, as after any length change) */
    if (rows <= 0) return 0;
    if (e->site || e->nonrev) { snprintf(e->err, sizeof e->err, "mixlen: not with site-specific or non-reversible models"); return -1; }
    size_t sz = (size_t)rows * e->cfg.ncat;
    e->mixlen_tab = (double *)malloc(sizeof(double) * sz);
    memcpy(e->mixlen_tab, table, sizeof(double) * sz);
    e->mixlen = e->cfg.ncat; /* ASSERT(nmixlen == ncat) :2367 */
    e->mixlen_rows = rows;
    return 0;
}

/* :2391-2460 + 2556-2568.  df [nmixlen + 1] (the last entry is the log-likelihood), ddf [nmixlen][nmixlen] */
int orc_branch_derv_mixlen(orc_engine *e, const orc_op *ops, int nops, int node_slot, int dad_kind, int dad_id, double length,
                           double *df_out, double *ddf_out) {
    if (!e->mixlen) { snprintf(e->err, sizeof e->err, "mixlen mode is off"); return -1; }
    if (e->asc_type) { snprintf(e->err, sizeof e->err, "mixlen: +ASC not supported (phylokernelnew.h:2449)"); return -1; }
    if (orc_update_partials(e, ops, nops) < 0) return -1;
    if (!e->theta_computed) build_theta(e, node_slot, dad_kind, dad_id);
    const int n = e->cfg.nstates, C = e->C, B = e->B, ncat = e->cfg.ncat, K = e->mixlen;
    const size_t P = (size_t)e->cfg.nptn;
    double *val0 = (double *)malloc(sizeof(double) * 3 * B), *val1 = val0 + B, *val2 = val1 + B;
    compute_vals(e, length, val0, val1, val2);
    double all_lh = 0.0;
    double *g = (double *)calloc((size_t)K * (K + 1), sizeof(double)), *h = g + K;
    double *df_ptn = (double *)malloc(sizeof(double) * 2 * K), *ddf_ptn = df_ptn + K;
    for (size_t p = 0; p < P; p++) {
        for (int i = 0; i < K; i++) df_ptn[i] = ddf_ptn[i] = 0.0;
        double lh_ptn = 0.0;
        const double *theta = e->theta + p * B;
        for (int c = 0; c < C; c++) {
            const int i = c % ncat; /* cat_id[c] */
            double x, y, z;
            dot_triple(val0 + c * n, val1 + c * n, val2 + c * n, theta + c * n, (size_t)n, n & 1, &x, &y, &z);
            lh_ptn += x; df_ptn[i] += y; ddf_ptn[i] += z; /* ADD = true */
        }
        lh_ptn = fabs(lh_ptn) + e->ptn_invar[p];
        const double freq = e->ptn_freq[p], inv = 1.0 / lh_ptn;
        for (int i = 0; i < K; i++) {
            df_ptn[i] *= inv;
            ddf_ptn[i] *= inv;
            g[i] = fma(df_ptn[i], freq, g[i]);
        }
        for (int i = 0; i < K; i++) {
            h[i * K + i] += fma(-df_ptn[i], df_ptn[i], ddf_ptn[i]) * freq; /* nmul_add(df, df, ddf) * freq */
            for (int c = 0; c < K; c++)
                if (c != i) h[i * K + c] -= df_ptn[i] * df_ptn[c] * freq;
        }
        all_lh = fma(log(lh_ptn) + e->buffer_scale_all[p], freq, all_lh);
    }
    for (int i = 0; i < K; i++) df_out[i] = g[i];
    df_out[K] = all_lh;
    for (int i = 0; i < K * K; i++) ddf_out[i] = h[i];
    free(val0); free(g); free(df_ptn);
    return 0;
}

/* :3431-3616 (no +ASC here) */
int orc_branch_derv_mixlen_cur(orc_engine *e, const orc_op *ops, int nops, int node_slot, int dad_kind, int dad_id, double length,
                               int cur_mixlen, double *df_out, double *ddf_out) {
    if (!e->mixlen) { snprintf(e->err, sizeof e->err, "mixlen mode is off"); return -1; }
    if (e->asc_type) { snprintf(e->err, sizeof e->err, "mixlen: +ASC is not restated"); return -1; }
    if (cur_mixlen < 0 || cur_mixlen >= e->mixlen) { snprintf(e->err, sizeof e->err, "cur_mixlen out of range"); return -1; }
    if (orc_update_partials(e, ops, nops) < 0) return -1;
    if (!e->theta_computed) build_theta(e, node_slot, dad_kind, dad_id);
    const int n = e->cfg.nstates, B = e->B;
    const int nmix = e->cfg.fused_mix_rate ? 1 : e->cfg.nmix;
    const size_t P = (size_t)e->cfg.nptn;
    double *val0 = (double *)malloc(sizeof(double) * 3 * nmix * n), *val1 = val0 + nmix * n, *val2 = val1 + nmix * n;
    const double len = len_of(e, length, cur_mixlen); /* :3508 */
    for (int c = 0; c < nmix; c++) {
        const int cur_mix = e->cfg.fused_mix_rate ? cur_mixlen : c;
        const double *eval_ptr = e->eval + (size_t)cur_mix * n;
        const double prop = e->mix_weight[cur_mix];
        for (int i = 0; i < n; i++) {
            const double cof = eval_ptr[i], val = exp(cof * len) * prop, v1 = cof * val;
            val0[c * n + i] = val; val1[c * n + i] = v1; val2[c * n + i] = cof * v1;
        }
    }
    double all_df = 0.0, all_ddf = 0.0;
    for (size_t p = 0; p < P; p++) {
        const double *theta = e->theta + p * B + (size_t)cur_mixlen * n * nmix; /* :3548 */
        double lh_ptn = 0.0, df_ptn = 0.0, ddf_ptn = 0.0;
        for (int c = 0; c < nmix; c++) {
            double x, y, z;
            dot_triple(val0 + c * n, val1 + c * n, val2 + c * n, theta + c * n, (size_t)n, n & 1, &x, &y, &z);
            lh_ptn += x; df_ptn += y; ddf_ptn += z;
        }
        lh_ptn = fabs(lh_ptn) + e->ptn_invar[p];
        const double freq = e->ptn_freq[p], inv = 1.0 / lh_ptn;
        df_ptn *= inv;
        ddf_ptn *= inv;
        all_df = fma(df_ptn, freq, all_df);
        all_ddf += fma(-df_ptn, df_ptn, ddf_ptn) * freq;
    }
    free(val0);
    *df_out = all_df;
    *ddf_out = all_ddf;
    if (!isfinite(all_df)) { *df_out = *ddf_out = 0.0; return 1; } /* :3607-3610 */
    return 0;
}
