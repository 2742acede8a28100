/* b200lk_over_oracle.c — TEST INFRASTRUCTURE.  The subset of the C ABI of include/b200lk.h that the C++ host mirrors call
 * (iqtree2_b200/csrc/phylotree_b200.cpp, supertree_b200.cpp, phylotree_capi.cpp), implemented over the CPU oracle
 * (lk_oracle.c), so that the HOST LOGIC of the product — traversal planning, optimizeOneBranch / optimizeAllBranches,
 * getBestNNIForBran, optimizeTreeLengthScaling, the partition drivers — runs in the `-m "not gpu"` tests against the Python
 * restatement and the compiled reference.  Built by oracle/Makefile into oracle/_build/libb200pt_over_oracle.so together with
 * those C++ files; nothing in the product links it (the product's libb200lk.so has no CPU path).  b200lk_config / b200lk_op /
 * b200lk_child have the layout of orc_config / orc_op / orc_child by construction. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b200lk.h"

typedef struct orc_engine orc_engine;
orc_engine *orc_create(const void *cfg);
void orc_destroy(orc_engine *e);
const char *orc_last_error(orc_engine *e);
int orc_set_tip_states(orc_engine *e, const uint8_t *states);
int orc_set_pattern_freq(orc_engine *e, const double *f);
int orc_set_pattern_invar(orc_engine *e, const double *f);
int orc_set_rates(orc_engine *e, const double *rate, const double *prop);
int orc_set_model(orc_engine *e, const double *eval, const double *evec, const double *inv_evec, const double *mix_weight);
int orc_set_mixlen(orc_engine *e, const double *table, int rows);
int orc_branch_derv_mixlen(orc_engine *e, const void *ops, int nops, int node_slot, int dad_kind, int dad_id, double length, double *df, double *ddf);
int orc_branch_derv_mixlen_cur(orc_engine *e, const void *ops, int nops, int node_slot, int dad_kind, int dad_id, double length, int cur, double *df, double *ddf);
int orc_set_site_model(orc_engine *e, const double *eval, const double *evec, const double *inv_evec);
int orc_set_model_nonrev(orc_engine *e, const double *ceval, const double *cevec, const double *cinv_evec, const double *rate_matrix,
                         const double *state_freq, const double *mix_weight, int rooted);
int orc_update_partials(orc_engine *e, const void *ops, int nops);
int orc_branch_logl(orc_engine *e, const void *ops, int nops, int node_slot, int dad_kind, int dad_id, double length, int save_log_value, double *logl_out);
int orc_branch_derv(orc_engine *e, const void *ops, int nops, int node_slot, int dad_kind, int dad_id, double length, double *df_out, double *ddf_out);
int orc_logl_from_buffer(orc_engine *e, double length, double *logl_out);
int orc_theta_reset(orc_engine *e);
int orc_set_robust_phylo(orc_engine *e, double keep, int median);
int orc_set_ascertainment(orc_engine *e, int asc_type, int64_t nunobs, int64_t nsite);
int orc_set_extra_children(orc_engine *e, const void *children, int n);
int orc_get_pattern_lh(orc_engine *e, double *out);
int orc_get_pattern_lh_cat(orc_engine *e, double *out);
int orc_get_scale_num(orc_engine *e, int slot, uint16_t *out);

struct b200lk_engine {
    orc_engine *o;
    double logl, df, ddf;
    int status;
    int64_t calls;
};

static char g_err[256] = "";
/* fault injection for the host mirrors' error paths (tests/test_host_mirror_cpu.py): the n-th next call that takes an op list
 * is refused BEFORE the oracle runs, as the CUDA engine refuses an invalid op or a failed allocation */
static int g_fail_in = 0;
void b200lk_test_fail_op_call_in(int n) { g_fail_in = n; }
static int refused(const b200lk_op *ops, int32_t nops) {
    (void)ops;
    if (g_fail_in > 0 && nops > 0 && --g_fail_in == 0) { snprintf(g_err, sizeof g_err, "injected failure"); return 1; }
    return 0;
}
static int wrap(b200lk_engine *e, int rc) {   /* the oracle: < 0 error, 1 = repaired underflow */
    e->calls++;
    return rc < 0 ? B200LK_ERR_INVALID : rc;
}

int32_t b200lk_abi_version(void) { return 1; }
int b200lk_create(const b200lk_config *cfg, b200lk_engine **out) {
    b200lk_engine *e = (b200lk_engine *)calloc(1, sizeof *e);
    e->o = orc_create(cfg);
    if (!e->o) { snprintf(g_err, sizeof g_err, "orc_create failed"); free(e); return B200LK_ERR_INVALID; }
    *out = e;
    return 0;
}
void b200lk_destroy(b200lk_engine *e) { if (e) { orc_destroy(e->o); free(e); } }
const char *b200lk_last_error(const b200lk_engine *e) { return e ? orc_last_error(e->o) : g_err; }
int b200lk_set_schedule(b200lk_engine *e, int32_t s) { (void)e; (void)s; return 0; }
int b200lk_set_tip_states(b200lk_engine *e, const uint8_t *s) { return wrap(e, orc_set_tip_states(e->o, s)); }
int b200lk_set_pattern_freq(b200lk_engine *e, const double *f) { return wrap(e, orc_set_pattern_freq(e->o, f)); }
int b200lk_set_pattern_invar(b200lk_engine *e, const double *f) { return wrap(e, orc_set_pattern_invar(e->o, f)); }
int b200lk_set_model(b200lk_engine *e, const double *eval, const double *evec, const double *inv_evec, const double *w) {
    return wrap(e, orc_set_model(e->o, eval, evec, inv_evec, w));
}
int b200lk_set_model_nonrev(b200lk_engine *e, const double *ceval, const double *cevec, const double *cinv_evec, const double *rate_matrix,
                            const double *state_freq, const double *mix_weight, int32_t rooted) {
    return wrap(e, orc_set_model_nonrev(e->o, ceval, cevec, cinv_evec, rate_matrix, state_freq, mix_weight, rooted));
}
int b200lk_set_site_model(b200lk_engine *e, const double *eval, const double *evec, const double *inv_evec) {
    return wrap(e, orc_set_site_model(e->o, eval, evec, inv_evec));
}
int b200lk_set_mixlen(b200lk_engine *e, const double *table, int32_t nrows) { return wrap(e, orc_set_mixlen(e->o, table, nrows)); }
int b200lk_branch_derv_mixlen(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind, int32_t dad_id,
                              double length, double *df, double *ddf) {
    if (refused(ops, nops)) return B200LK_ERR_INVALID;
    return wrap(e, orc_branch_derv_mixlen(e->o, ops, nops, node_slot, dad_kind, dad_id, length, df, ddf));
}
int b200lk_branch_derv_mixlen_cur(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind, int32_t dad_id,
                                  double length, int32_t cur, double *df, double *ddf) {
    if (refused(ops, nops)) return B200LK_ERR_INVALID;
    return wrap(e, orc_branch_derv_mixlen_cur(e->o, ops, nops, node_slot, dad_kind, dad_id, length, cur, df, ddf));
}
int b200lk_set_rates(b200lk_engine *e, const double *r, const double *p) { return wrap(e, orc_set_rates(e->o, r, p)); }
int b200lk_set_ascertainment(b200lk_engine *e, int32_t t, int64_t n, int64_t s) { return wrap(e, orc_set_ascertainment(e->o, t, n, s)); }
int b200lk_set_robust_phylo(b200lk_engine *e, double keep, int32_t median) { return wrap(e, orc_set_robust_phylo(e->o, keep, median)); }
int b200lk_set_extra_children(b200lk_engine *e, const b200lk_child *c, int32_t n) { return wrap(e, orc_set_extra_children(e->o, c, n)); }
int b200lk_theta_reset(b200lk_engine *e) { return orc_theta_reset(e->o); }
int b200lk_update_partials(b200lk_engine *e, const b200lk_op *ops, int32_t nops) {
    if (refused(ops, nops)) return B200LK_ERR_INVALID;
    return wrap(e, orc_update_partials(e->o, ops, nops));
}
int b200lk_branch_logl(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind, int32_t dad_id,
                       double length, int32_t save_log_value, double *logl) {
    if (refused(ops, nops)) return B200LK_ERR_INVALID;
    return wrap(e, orc_branch_logl(e->o, ops, nops, node_slot, dad_kind, dad_id, length, save_log_value, logl));
}
int b200lk_branch_derv(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind, int32_t dad_id,
                       double length, double *df, double *ddf) {
    if (refused(ops, nops)) return B200LK_ERR_INVALID;
    if (nops > 0) orc_theta_reset(e->o);
    return wrap(e, orc_branch_derv(e->o, ops, nops, node_slot, dad_kind, dad_id, length, df, ddf));
}
int b200lk_logl_from_buffer(b200lk_engine *e, double length, double *logl) { return wrap(e, orc_logl_from_buffer(e->o, length, logl)); }
/* the asynchronous forms run at once; b200lk_finish hands the stored scalars back */
int b200lk_enqueue_branch_logl(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind, int32_t dad_id,
                               double length, int32_t save_log_value) {
    e->status = b200lk_branch_logl(e, ops, nops, node_slot, dad_kind, dad_id, length, save_log_value, &e->logl);
    return e->status < 0 ? e->status : 0;
}
int b200lk_enqueue_branch_derv(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind, int32_t dad_id,
                               double length) {
    e->status = b200lk_branch_derv(e, ops, nops, node_slot, dad_kind, dad_id, length, &e->df, &e->ddf);
    return e->status < 0 ? e->status : 0;
}
int b200lk_enqueue_logl_from_buffer(b200lk_engine *e, double length) {
    e->status = b200lk_logl_from_buffer(e, length, &e->logl);
    return e->status < 0 ? e->status : 0;
}
int b200lk_finish(b200lk_engine *e, double *logl, double *df, double *ddf) {
    if (logl) *logl = e->logl;
    if (df) *df = e->df;
    if (ddf) *ddf = e->ddf;
    return e->status;
}
/* minimizeNewton (utils/optimization.cpp:422-499) over the oracle's derivative pass: what the product runs on the device */
int b200lk_newton_branch(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind, int32_t dad_id,
                         double x1, double xguess, double x2, double xacc, int32_t maxNRStep, double *length_out, int32_t *steps_out) {
    int nsteps = 0, first = 1, rc;
    double df, dx, f, temp, xh, xl, rts, rts_old;
#define FUNC_DERV(x)                                                                                                  \
    do {                                                                                                              \
        double d1, d2;                                                                                                \
        rc = b200lk_branch_derv(e, first ? ops : NULL, first ? nops : 0, node_slot, dad_kind, dad_id, (x), &d1, &d2); \
        first = 0;                                                                                                    \
        if (rc < 0) return rc;                                                                                        \
        f = -d1; df = -d2; nsteps++;                                                                                  \
    } while (0)
#define DONE(x) do { *length_out = (x); if (steps_out) *steps_out = nsteps; return 0; } while (0)
    rts = xguess;
    if (rts < x1) rts = x1;
    if (rts > x2) rts = x2;
    FUNC_DERV(rts);
    if (!isfinite(f) || !isfinite(df)) { *length_out = rts; if (steps_out) *steps_out = nsteps; return B200LK_WARN_DERV_UNDERFLOW; }
    if (df >= 0.0 && fabs(f) < xacc) DONE(rts);
    if (f < 0.0) { xl = rts; xh = x2; } else { xh = rts; xl = x1; }
    dx = fabs(xh - xl);
    for (int j = 1; j <= maxNRStep; j++) {
        rts_old = rts;
        if ((df <= 0.0) || (((rts - xh) * df - f) * ((rts - xl) * df - f) >= 0.0)) {
            dx = 0.5 * (xh - xl);
            rts = xl + dx;
            if (xl == rts) DONE(rts);
        } else {
            dx = f / df;
            temp = rts;
            rts -= dx;
            if (temp == rts) DONE(rts);
        }
        if (fabs(dx) < xacc || (j == maxNRStep)) DONE(rts_old);
        FUNC_DERV(rts);
        if (!isfinite(f) || !isfinite(df)) { *length_out = rts; if (steps_out) *steps_out = nsteps; return B200LK_WARN_DERV_UNDERFLOW; }
        if (df > 0.0 && fabs(f) < xacc) DONE(rts);
        if (f < 0.0) xl = rts;
        else if (f > 0.0) xh = rts;
    }
    DONE(rts);
#undef FUNC_DERV
#undef DONE
}
/* whole queued sweeps need the device length table: not available here, the host mirror falls back to branch by branch */
int b200lk_enqueue_newton_branch(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind, int32_t dad_id,
                                 double x1, double xguess, double x2, double xacc, int32_t maxNRStep, int32_t len_slot) {
    (void)e; (void)ops; (void)nops; (void)node_slot; (void)dad_kind; (void)dad_id; (void)x1; (void)xguess; (void)x2; (void)xacc;
    (void)maxNRStep; (void)len_slot;
    return B200LK_ERR_STATE;
}
int b200lk_sweep_finish(b200lk_engine *e, int32_t n, double *len, int32_t *steps, int32_t *status) {
    (void)e; (void)n; (void)len; (void)steps; (void)status;
    return B200LK_ERR_STATE;
}
int b200lk_get_pattern_lh(b200lk_engine *e, double *out) { return orc_get_pattern_lh(e->o, out); }
int b200lk_get_pattern_lh_cat(b200lk_engine *e, double *out) { return orc_get_pattern_lh_cat(e->o, out); }
int b200lk_get_scale_num(b200lk_engine *e, int32_t slot, uint16_t *out) { return orc_get_scale_num(e->o, slot, out); }
int64_t b200lk_launch_count(const b200lk_engine *e) { return e->calls; }
int64_t b200lk_device_bytes(const b200lk_engine *e) { (void)e; return 0; }

/* ---- grouped evaluation (include/b200lk.h b200lk_group_*): the checker simply loops over the members ----------------------- */
struct b200lk_group { b200lk_engine **m; int n; int64_t calls; char err[256]; };
int b200lk_group_create(b200lk_engine *const *engines, int32_t n, b200lk_group **out) {
    b200lk_group *g = (b200lk_group *)calloc(1, sizeof *g);
    g->m = (b200lk_engine **)calloc((size_t)n, sizeof *g->m);
    for (int k = 0; k < n; k++) g->m[k] = engines[k];
    g->n = n;
    *out = g;
    return 0;
}
void b200lk_group_destroy(b200lk_group *g) { if (g) { free(g->m); free(g); } }
const char *b200lk_group_last_error(const b200lk_group *g) { return g ? g->err : g_err; }
int64_t b200lk_group_launch_count(const b200lk_group *g) { int64_t t = g->calls; for (int k = 0; k < g->n; k++) t += g->m[k]->calls; return t; }
static int gerr(b200lk_group *g, int k, int rc) { snprintf(g->err, sizeof g->err, "member %d: %s", k, b200lk_last_error(g->m[k])); return rc; }
int b200lk_group_enqueue_branch_logl(b200lk_group *g, const b200lk_op *const *ops, const int32_t *nops, const int32_t *node_slot,
                                     const int32_t *dad_kind, const int32_t *dad_id, const double *length, int32_t save_log_value) {
    g->calls++;
    for (int k = 0; k < g->n; k++) {
        int rc = b200lk_enqueue_branch_logl(g->m[k], ops[k], nops[k], node_slot[k], dad_kind[k], dad_id[k], length[k], save_log_value);
        if (rc < 0) return gerr(g, k, rc);
    }
    return 0;
}
int b200lk_group_enqueue_branch_derv(b200lk_group *g, const b200lk_op *const *ops, const int32_t *nops, const int32_t *node_slot,
                                     const int32_t *dad_kind, const int32_t *dad_id, const double *length) {
    g->calls++;
    for (int k = 0; k < g->n; k++) {
        int rc = b200lk_enqueue_branch_derv(g->m[k], ops[k], nops[k], node_slot[k], dad_kind[k], dad_id[k], length[k]);
        if (rc < 0) return gerr(g, k, rc);
    }
    return 0;
}
int b200lk_group_enqueue_logl_from_buffer(b200lk_group *g, const double *length) {
    g->calls++;
    for (int k = 0; k < g->n; k++) {
        int rc = b200lk_enqueue_logl_from_buffer(g->m[k], length[k]);
        if (rc < 0) return gerr(g, k, rc);
    }
    return 0;
}
int b200lk_group_finish(b200lk_group *g, double *logl, double *df, double *ddf, int32_t *status) {
    int worst = 0;
    for (int k = 0; k < g->n; k++) {
        int rc = b200lk_finish(g->m[k], logl ? logl + k : NULL, df ? df + k : NULL, ddf ? ddf + k : NULL);
        if (rc < 0) return gerr(g, k, rc);
        if (status) status[k] = rc;
        if (rc > worst) worst = rc;
    }
    return worst;
}
