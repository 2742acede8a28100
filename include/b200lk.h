/*
 * b200lk.h — C ABI of the B200-native phylogenetic likelihood engine (libb200lk.so).
 *
 * Drop-in boundary for IQ-TREE 2's PhyloTree likelihood path.  Every entry point names the reference
 * interface it replaces (paths relative to the IQ-TREE 2.4.0 source tree).  The five pointers-to-member
 * that PhyloTree::setLikelihoodKernel assigns (tree/phylotreesse.cpp:91-199; declared at
 * tree/phylotree.h:900-901, 947-948, 1000-1001, 1342-1343) would be pointed at thin members that marshal
 * into these calls — see INTEGRATION.md for the shim.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are HOST pointers unless a name says `_dev`;
 *   - every call returns 0 on success, >0 for a numerical condition the reference reports with a warning
 *     (B200LK_WARN_*), <0 for an error (message via b200lk_last_error); nothing throws, nothing exits;
 *   - one engine = one PhyloTree's buffers on one GPU for one pattern range. Distinct engines may be used
 *     concurrently from different host threads (PhyloSuperTree calls partition trees in an omp parallel for,
 *     tree/phylosupertreeplen.cpp:380-409); calls on the same engine must be serialised by the caller;
 *   - there is NO CPU fallback: without a CUDA device b200lk_create fails.
 *   - multi-device engines (cfg.device = -N): every entry point takes them; patterns are sharded, the per-evaluation scalars
 *     are summed on the host in device order (the single-process form of the all-reduce of tree/phylokernelnew.h:2375-2379,
 *     2747-2748); per-pattern read-backs gather.  Not separable per shard and refused with B200LK_ERR_STATE: +ASC, robust phylo,
 *     b200lk_enqueue_newton_branch / b200lk_sweep_finish (use b200lk_newton_branch: one reduction per Newton step),
 *     b200lk_set_stream, b200lk_set_pattern_freq_from_posterior.
 *
 * Host-visible layouts are the reference's with SIMD width V=1 (SURVEY §8a "Layout crib"):
 *   partial_lh[p*B + c*n + i], B = n*ncat_mix, c = m*ncat + r;   scale_num NORM_LH [p], SAFE_LH [p*C + c];
 *   the device keeps a warp-interleaved layout internally and converts on read-back.
 */
#ifndef B200LK_H_
#define B200LK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LK_ABI_VERSION 1

/* status codes (>0: numerical warnings the reference prints and repairs; <0: errors) */
#define B200LK_OK 0
#define B200LK_WARN_UNDERFLOW 1      /* non-finite tree logL repaired per pattern (phylokernelnew.h:3107-3132, 3374-3383) */
#define B200LK_WARN_DERV_UNDERFLOW 2 /* non-finite df: df = ddf = 0 returned (phylokernelnew.h:2629-2632) */
#define B200LK_ERR_INVALID (-1)
#define B200LK_ERR_CUDA (-2)
#define B200LK_ERR_NOMEM (-3)
#define B200LK_ERR_STATE (-4)        /* e.g. logl_from_buffer without a theta buffer (ASSERT at phylokernelnew.h:3240) */

#define B200LK_SEQ_GENERIC 0 /* states >= nstates are all "unknown" (model/modelsubst.cpp:197-207) */
#define B200LK_SEQ_DNA 1     /* states 4..17 are IUPAC ambiguity bitmasks state-3 (model/modeldna.cpp:462-475) */
#define B200LK_SEQ_PROTEIN 2 /* states 20,21,22 = B,Z,J (model/modelprotein.cpp:1149-1166) */

/* What PhyloTree::initializeAllPartialLh sizes its buffers from (tree/phylotree.cpp:893-939, 1060-1188). */
typedef struct b200lk_config {
    int32_t nstates;        /* aln->num_states */
    int32_t ncat;           /* site_rate->getNRate() */
    int32_t nmix;           /* model->getNMixtures() */
    int32_t fused_mix_rate; /* model_factory->fused_mix_rate */
    int32_t safe_numeric;   /* 0 = NORM_LH, 1 = SAFE_LH (tree/phylotreesse.cpp:95-96) */
    int32_t seq_type;       /* B200LK_SEQ_* */
    int32_t state_unknown;  /* aln->STATE_UNKNOWN (alignment/alignment.cpp:738-745) */
    int32_t ntaxa;          /* aln->getNSeq() */
    int64_t nptn;           /* patterns owned by this engine (aln->size(), or this GPU's slice of it) */
    int32_t nslots;         /* CLV slots: (ntaxa-2) for LM_PER_NODE + 2 NNI spares (tree/phylotree.cpp:1071-1104) */
    int32_t device;         /* CUDA device ordinal; -N (N >= 2): the first N devices, one engine per device over contiguous pattern
                               slices behind this one handle (see "multi-device engines" below) */
} b200lk_config;

/* One child of a TraversalInfo entry: a leaf (kind 0, id = alignment row) or an internal CLV (kind 1, id = slot). */
typedef struct b200lk_child {
    int32_t kind;
    int32_t id;
    double length; /* child->length */
} b200lk_child;

/* One pending partial update = one TraversalInfo (tree/phylotree.h:320-331): write dad_branch->partial_lh
 * (slot dst_slot) from the two children in neighbour order. echildren / partial_lh_leaves are built on device. */
typedef struct b200lk_op {
    int32_t dst_slot;
    int32_t reserved; /* 0 for a bifurcating node; k > 0: a multifurcating node (node->degree() > 3,
                         tree/phylokernelnew.h:1407-1631) whose third ... (k+2)-th children, in neighbour order, are the next k
                         entries of the array handed to b200lk_set_extra_children */
    b200lk_child left, right;
} b200lk_op;

typedef struct b200lk_engine b200lk_engine;

/* ---- life-cycle: initializeAllPartialLh / deleteAllPartialLh (tree/phylotree.cpp:893, 754) ------------------ */
int b200lk_abi_version(void);
int b200lk_create(const b200lk_config *cfg, b200lk_engine **out);
void b200lk_destroy(b200lk_engine *e);
const char *b200lk_last_error(const b200lk_engine *e); /* e may be NULL: error of the last failed create */
/* bytes of device memory the engine holds (PhyloTree::getMemoryRequired, tree/phylotree.cpp:967) */
int64_t b200lk_device_bytes(const b200lk_engine *e);
/* run on a caller-owned stream (cudaStream_t as void*); NULL restores the engine's own (non-blocking) stream.  NOTE: the legacy
 * default stream's handle IS NULL — a caller that wants the engine ordered with its other work (a collective, a copy) must pass a
 * stream it created; work on the engine's own stream is ordered with nothing but itself */
int b200lk_set_stream(b200lk_engine *e, void *cuda_stream);
/* Reference defects that change results; OFF by default (the engine computes the intended mathematics), ON to
 * reproduce the reference bit for bit in parity tests.
 *   B200LK_QUIRK_DOTTRIPLE_ODD: computeLikelihoodDerv's dotProductTriple counts the last theta element twice when
 *   nstates is odd and nstates*ncat_mix is even (tree/phylokernelnew.h:700-714; e.g. codon 61 x G4, 5 states x 6). */
#define B200LK_QUIRK_DOTTRIPLE_ODD 1
int b200lk_set_reference_quirks(b200lk_engine *e, int32_t mask);
/* schedule for multi-op traversals: 0 = auto, 1 = one levelled launch per tree depth, 2 = pattern-tile walker;
 * 10 + v (v < 30): development switch between kernel variants of the walker schedule (A/B measurements) */
int b200lk_set_schedule(b200lk_engine *e, int schedule);

/* ---- inputs ------------------------------------------------------------------------------------------------- */
/* tip states, sequence-major [ntaxa][nptn] (AlignmentSummary::sequenceMatrix rows, tree/phylotree.cpp:3251-3256;
 * value = Pattern entry aln->at(ptn)[taxon], 0..state_unknown) */
int b200lk_set_tip_states(b200lk_engine *e, const uint8_t *states);
/* ptn_freq / ptn_invar as computePtnFreq / computePtnInvar fill them (tree/phylotreesse.cpp:537-612) */
int b200lk_set_pattern_freq(b200lk_engine *e, const double *ptn_freq);
int b200lk_set_pattern_invar(b200lk_engine *e, const double *ptn_invar);
/* model->getEigenvalues/getEigenvectors/getInverseEigenvectors/getMixtureWeight (model/modelmixture.cpp:1536-1575);
 * recomputes tip_partial_lh on device = computeTipPartialLikelihood (tree/phylotreesse.cpp:243-375) */
int b200lk_set_model(b200lk_engine *e, const double *eval, const double *evec, const double *inv_evec,
                     const double *mix_weight);
/* ---- non-reversible models: the kernels of tree/phylokernelnonrev.h (computeNonrevPartialLikelihood :32, computeNonrevLikelihoodDerv
 * :583, computeNonrevLikelihoodBranch :1059), which setLikelihoodKernel assigns when !model->useRevKernel() — a non-reversible
 * model, or any model under --kernel-nonrev (model/modelsubst.h:73-75).  After this call the engine's CLV slots hold STATE-space
 * vectors and b200lk_update_partials / (enqueue_)branch_logl / (enqueue_)branch_derv / newton_branch run that kernel:
 *   - transition matrices are built on the device as the reference's default technique does (ModelMarkov::computeTransMatrixNonrev,
 *     model/modelmarkov.cpp:477-486): P_m(t) = Re(cevec_m . diag(exp(ceval_m t)) . cinv_evec_m); the derivatives of the evaluation
 *     branch are rate_matrix_m . P and rate_matrix_m^2 . P (:735-747).  ceval: [nmix][n] complex as (re, im) pairs; cevec, cinv_evec:
 *     [nmix][n][n] complex, row-major; rate_matrix [nmix][n][n]; state_freq [nmix][n] (model->getStateFrequency(., m)).  (The
 *     scaling-and-squaring fall-back the reference takes for non-diagonalisable matrices, :465-476, stays with the caller.)
 *   - a child's `kind` is B200LK_CHILD_TIP / _SLOT / _ROOT (the virtual root leaf of a rooted tree, MTree::isRootLeaf,
 *     tree/mtree.h:135: it contributes the state frequencies, tree/phylokernelnew.h:959-964), or-ed with B200LK_CHILD_TRANSPOSED when
 *     the child's branch points towards the root (child->direction == TOWARD_ROOT: transposed matrix, :939-951);
 *   - branch calls: node_slot = dad_branch->partial_lh AFTER the reference's swap (phylokernelnonrev.h:596-603, 1071-1078: the
 *     root-ward end, or the inner end of a pendant branch); dad_kind = B200LK_CHILD_TIP / _SLOT / _ROOT for the other end;
 *   - rooted = PhyloTree::rooted; 0 = an unrooted tree (the rows of the evaluation branch's matrix are weighted by state_freq,
 *     phylokernelnonrev.h:1112-1120);
 *   - not available in this mode (B200LK_ERR_STATE / _INVALID): b200lk_logl_from_buffer (the reference leaves
 *     computeLikelihoodFromBufferPointer NULL), save_log_value = 0, +ASC, multifurcating ops, grouped launches, the device-resident
 *     sweep (b200lk_enqueue_newton_branch; b200lk_newton_branch iterates from the host). */
#define B200LK_CHILD_TIP 0
#define B200LK_CHILD_SLOT 1
#define B200LK_CHILD_ROOT 2
#define B200LK_CHILD_TRANSPOSED 4
int b200lk_set_model_nonrev(b200lk_engine *e, const double *ceval, const double *cevec, const double *cinv_evec,
                            const double *rate_matrix, const double *state_freq, const double *mix_weight, int32_t rooted);

/* Site-specific models (model->isSiteSpecificModel(): ModelSet — `-fs` site-frequency files and the PMSF approximation of the
 * C10-C60 profile mixtures), the SITE_MODEL template branch of tree/phylokernelnew.h (partial :1362-1374, 1653-1683, 1772-1801,
 * 1937-1968; buffer :2129-2140; derivative :2306-2311, 2469-2487; branch :2708-2713, 2833-2848, 2993-3008; from-buffer
 * :3271-3276, 3316-3327; instantiations tree/phylokernelfma.cpp:43-100): pattern p has its OWN eigensystem,
 *   eval [nptn][nstates], evec / inv_evec [nptn][nstates][nstates] row-major — ModelSet::at(p)'s getEigenvalues() /
 *   getEigenvectors() / getInverseEigenvectors() (what ModelSet::computeTipLikelihood and the kernels index by ptn).
 * The engine must have nmix = 1 (ncat_mix = ncat: the reference asserts it, :1362).  Afterwards every likelihood entry point runs
 * the SITE_MODEL arithmetic; the CLV slots hold eigen coordinates of the pattern's own system, as the reference's.  Not available
 * in this mode (B200LK_ERR_STATE / _INVALID): +ASC, robust phylo, multifurcating ops, grouped launches, the non-reversible kernel,
 * the device-resident sweep (b200lk_enqueue_newton_branch; b200lk_newton_branch iterates from the host).  On a multi-device engine
 * every shard takes its pattern slice of the three arrays. */
int b200lk_set_site_model(b200lk_engine *e, const double *eval, const double *evec, const double *inv_evec);

/* Mixed branch lengths (heterotachy, the GHOST models `+H` / `*H`; PhyloTree::isMixlen(), tree/phylotreemixlen.h:89): every branch
 * carries nmixlen = ncat lengths (PhyloNeighborMixlen::lengths, tree/phylonodemixlen.h:60-69).
 *   b200lk_set_mixlen(e, table, nrows): table [nrows][ncat], row k = the lengths of the branch with Neighbor::id k (what
 *     PhyloTree::saveBranchLengths writes at id * getMixlen(), tree/phylotree.cpp:728-731).  From then on EVERY length this API takes
 *     — b200lk_child::length in ops, the `length` of the branch calls — is the ROW of the branch in this table (an integer value),
 *     and rate category c reads table[row][c % ncat] wherever the reference reads child->getLength(cat_id[c]) /
 *     dad_branch->getLength(mycat) (tree/phylokernelnew.h:1017, 1069, 2325, 2348, 2724, 2738, 3286, 3301).  Partial likelihoods,
 *     b200lk_branch_logl and b200lk_logl_from_buffer run on the same kernels as before.  Setting the table again (new lengths
 *     during optimisation) is a small asynchronous copy; CLVs and the theta buffer are the caller's to invalidate, as after any
 *     length change.  nrows = 0 switches the mode off.
 *   b200lk_branch_derv_mixlen: computeLikelihoodDerv of a mixlen tree (:2391-2460, 2556-2568): df [ncat + 1] = the gradient over
 *     the branch's ncat lengths followed by the log-likelihood ("last entry of df now store log-likelihood"), ddf [ncat][ncat] =
 *     the Hessian.  (b200lk_branch_derv returns B200LK_ERR_STATE in this mode, as do the Newton entry points: the reference
 *     optimises mixlen branches with BFGS / EM over these two calls, tree/phylotreemixlen.cpp.)
 *   b200lk_branch_derv_mixlen_cur: computeLikelihoodDervMixlen (:3431-3616) for class cur_mixlen = getCurMixture(): (df, ddf) of
 *     that class's length alone, from its slice of the theta buffer — the M-step of PhyloTreeMixlen::optimizeOneBranch's EM
 *     (the caller has put the class's posterior weights into ptn_freq).
 * Not available with +ASC (the reference asserts, :2449), site-specific or non-reversible models, grouped launches. */
int b200lk_set_mixlen(b200lk_engine *e, const double *table, int32_t nrows);
int b200lk_branch_derv_mixlen(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                              int32_t dad_id, double length, double *df, double *ddf);
int b200lk_branch_derv_mixlen_cur(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                                  int32_t dad_id, double length, int32_t cur_mixlen, double *df, double *ddf);
/* site_rate->getRate(c) / getProp(c) */
int b200lk_set_rates(b200lk_engine *e, const double *cat_rate, const double *cat_prop);

/* +ASC, Lewis' ascertainment-bias correction (ModelFactory::ASC_type = ASC_VARIANT / ASC_INFORMATIVE;
 * tree/phylokernelnew.h:3076-3091, 3198-3221, 2506-2535, 2618-2626, 3342-3356, 3412-3426): the LAST num_unobserved of the
 * engine's nptn patterns are ModelFactory::unobserved_ptns (their ptn_freq is ignored, their ptn_invar is the caller's);
 * nsite = aln->getNSite().  logL, _pattern_lh, df and ddf come back corrected.  asc_type 0 switches it off.
 * asc_type 2 = Holder's correction for missing data (+ASC_MIS, ASC_VARIANT_MISSING; :3175-3197, 2580-2617, 3385-3410): the
 * unobserved block is what Alignment::getUnobservedConstPatterns builds (alignment/alignment.cpp:4158-4180) — nstates groups
 * of num_unobserved/nstates patterns (>= the observed pattern count; the reference pads it to its SIMD width), pattern p of
 * group s = observed pattern p with every non-gap character replaced by state s; the correction is per pattern,
 * log(1 - sum_s L(const pattern s,p)), and nsite is not used.  (ASC_INFORMATIVE_MISSING aborts in the reference itself,
 * alignment.cpp:4198-4201.)  With pattern sharding Lewis' correction is not separable per rank (it needs the global
 * prob_const): use one engine, or replicate the unobserved block and combine on the host; Holder's is per pattern and shards
 * with its patterns. */
int b200lk_set_ascertainment(b200lk_engine *e, int32_t asc_type, int64_t num_unobserved, int64_t nsite);

/* E-step of the reference's EM algorithms, on the device.  PhyloTree::computePatternLhCat(wsl) (tree/phylotree.cpp:1338-1393;
 * wsl as SiteLoglType, utils/tools.h:555: 2 = WSL_RATECAT, 3 = WSL_MIXTURE, 4 = WSL_MIXTURE_RATECAT) takes the per-category
 * pattern likelihoods L[p][k] of the LAST b200lk_branch_logl — for an unfused mixture summed over the mixture classes (per rate
 * category) or over the rate categories (per class) — and the loops of RateFree::optimizeWithEM (model/ratefree.cpp:556-573,
 * add_invar = 1) / ModelMixture::optimizeWithEM (model/modelmixture.cpp:2008-2024, add_invar = 0) turn them into
 *     posterior[p][k] = ptn_freq[p] * L[p][k] / (add_invar * ptn_invar[p] + sum_k L[p][k]),   new_prop[k] = sum_p posterior[p][k].
 * *ncat_out = K (getNumLhCat(wsl)); new_prop: K doubles (NULL: only K is returned); posterior (may be NULL): [nptn][K] on the
 * host.  The posteriors stay on the device: the M-step's per-category trees (ratefree.cpp:617-650: ptn_freq := one column) take
 * them with b200lk_set_pattern_freq_from_posterior(dst, src, k) — same pattern count, same device — without a host copy. */
int b200lk_em_posterior(b200lk_engine *e, int32_t wsl, int32_t add_invar, double *new_prop, double *posterior, int32_t *ncat_out);
int b200lk_set_pattern_freq_from_posterior(b200lk_engine *dst, b200lk_engine *src, int32_t k);

/* The robust-phylo variants of computeLikelihoodBranch (Params::robust_phy_keep < 1, `--robust-phy`: the log-likelihood summed
 * over the best keep*nsite sites; Params::robust_median, `--robust-median`: the median site log-likelihood;
 * tree/phylokernelnew.h:3134-3173): every later b200lk_branch_logl (and its enqueue form) returns that value; _pattern_lh is
 * the uncorrected one, as in the reference; b200lk_logl_from_buffer is not affected (:3232-3428 has no such branch).
 * (1.0, 0) switches it off.  Fails together with +ASC, as the reference does. */
int b200lk_set_robust_phylo(b200lk_engine *e, double keep, int32_t median);

/* ---- the five kernel entry points --------------------------------------------------------------------------- */
/* computePartialLikelihood over a traversal_info list in post-order
 * (computeTraversalInfo<VC>(node, dad, compute_partial_lh=true), tree/phylokernelnew.h:1150-1281, 1289-2039) */
int b200lk_update_partials(b200lk_engine *e, const b200lk_op *ops, int32_t nops);
/* the extra children of the multifurcating ops of the NEXT call that takes an op list, in op order (copied; consumed by
 * that call).  Op lists with multifurcating nodes run on the thread-per-pattern kernel. */
int b200lk_set_extra_children(b200lk_engine *e, const b200lk_child *children, int32_t n);

/* computeLikelihoodBranch(dad_branch, dad, save_log_value) (tree/phylokernelnew.h:2641-3223): flush `ops`, then the
 * branch between CLV slot `node_slot` and the other end (dad_kind 0: leaf `dad_id`; 1: slot `dad_id`).
 * Fills _pattern_lh (+ _pattern_scaling when !save_log_value) and _pattern_lh_cat; returns this engine's
 * sum_p ptn_freq[p]*_pattern_lh[p]. */
int b200lk_branch_logl(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                       int32_t dad_id, double length, int32_t save_log_value, double *logl);

/* computeLikelihoodDerv(dad_branch, dad, &df, &ddf) (tree/phylokernelnew.h:2233-2633): flush `ops`, build
 * theta_all/buffer_scale_all if !theta_computed (computeLikelihoodBuffer, :2048-2231), then one (df, ddf) pass. */
int b200lk_branch_derv(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                       int32_t dad_id, double length, double *df, double *ddf);
/* theta_computed = false (PhyloTree::optimizeOneBranch, tree/phylotree.cpp:2608) */
int b200lk_theta_reset(b200lk_engine *e);

/* computeLikelihoodFromBuffer() (tree/phylokernelnew.h:3232-3428) at current_it->length = length */
int b200lk_logl_from_buffer(b200lk_engine *e, double length, double *logl);

/* Fused Newton-Raphson on the theta buffer = Optimization::minimizeNewton (utils/optimization.cpp:422-499) driving
 * computeFuncDerv (tree/phylotree.cpp:2529-2536) without a host round trip per step. Flushes `ops` and builds theta
 * first when needed. Returns the optimised length in *length_out. */
int b200lk_newton_branch(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                         int32_t dad_id, double min_len, double guess, double max_len, double tolerance,
                         int32_t max_steps, double *length_out, int32_t *steps_out);

/* ---- asynchronous forms (partition models) ----------------------------------------------------------------------
 * PhyloSuperTree evaluates its partition trees inside `#pragma omp parallel for` (tree/phylosupertree.cpp:716-723,
 * tree/phylosupertreeplen.cpp:321-345, 380-409) and sums the scalars.  On the device the partitions are independent engines
 * on independent streams: the enqueue_* calls queue exactly the work of their synchronous namesakes and return without
 * waiting; b200lk_finish waits for the engine's stream and hands out what was queued (logL from enqueue_branch_logl /
 * enqueue_logl_from_buffer, df and ddf from enqueue_branch_derv; a derv call may be queued behind a logL call).  The
 * synchronous entry points above are enqueue + finish.  NULL output pointers are allowed. */
int b200lk_enqueue_branch_logl(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                               int32_t dad_id, double length, int32_t save_log_value);
int b200lk_enqueue_branch_derv(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                               int32_t dad_id, double length);
int b200lk_enqueue_logl_from_buffer(b200lk_engine *e, double length);
int b200lk_finish(b200lk_engine *e, double *logl, double *df, double *ddf);
/* Multi-process hosts (one process per GPU, each engine a pattern slice): queue, behind the calls queued so far, a copy of
 * this engine's (logL, df, ddf) into three doubles of caller-owned DEVICE memory on the engine's stream — the input of a
 * device-side collective (ncclAllReduce on that buffer sums the ranks' scalars: the reduction sites of
 * tree/phylokernelnew.h:2375-2379, 2747-2748) without a host round trip per evaluation.  b200lk_finish still has to be called. */
int b200lk_enqueue_results_dev(b200lk_engine *e, double *dev_out);

/* A whole optimizeAllBranches sweep without a host round trip per branch (tree/phylotree.cpp:2698-2774).  Per branch the
 * caller queues b200lk_enqueue_newton_branch — partial updates, theta buffer, the complete minimizeNewton iteration and the
 * divergence test of optimizeOneBranch (:2616-2625), all on the device — and goes on planning the next branch; the optimised
 * length lands in slot `len_slot` (0 .. 2*ntaxa-1) of a device-resident length table.  Ops queued later in the same sweep
 * refer to it with a NEGATIVE child length: length = -(len_slot) - 1.  b200lk_sweep_finish waits once and returns, per slot,
 * the length, the number of derivative evaluations and the status of the iteration.  B200LK_ERR_STATE from the enqueue call
 * means "use b200lk_newton_branch" (blocks above 512 doubles, +ASC). */
int b200lk_enqueue_newton_branch(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                                 int32_t dad_id, double min_len, double guess, double max_len, double tolerance,
                                 int32_t max_steps, int32_t len_slot);
int b200lk_sweep_finish(b200lk_engine *e, int32_t nslots, double *lengths, int32_t *evaluations, int32_t *status);

/* ---- grouped evaluation: the partition trees of a PhyloSuperTree in O(1) launches ---------------------------------------------
 * PhyloSuperTreePlen evaluates its partition trees from `#pragma omp parallel for` loops over one pooled arena with per-partition
 * offsets (tree/phylosupertreeplen.cpp:282-290, 322-341, 380-409; arena :1624-1720).  A group is that arena's descriptor table:
 * engines on ONE device, evaluated together — every call takes one op list and one evaluation branch PER MEMBER (arrays of
 * length n, member order = creation order) and queues one launch per kernel type for all members on the group's stream;
 * b200lk_group_finish waits once and hands out every member's scalars (NULL arrays allowed; status[k] = member k's warning code).
 * After b200lk_group_create the members' own entry points keep working and are ordered on the group's stream too; calls on a
 * group and its members are serialised by the caller.  A member the grouped kernels do not cover (+ASC, robust phylo, blocks
 * above 512 doubles, state counts other than 4 and 20, multi-device engines are refused) is queued through its own entry points
 * by the same calls: same results, its own launches.  Destroy the group before its members. */
typedef struct b200lk_group b200lk_group;
int b200lk_group_create(b200lk_engine *const *engines, int32_t n, b200lk_group **out);
void b200lk_group_destroy(b200lk_group *g);
const char *b200lk_group_last_error(const b200lk_group *g); /* g may be NULL: error of the last failed create */
int b200lk_group_enqueue_branch_logl(b200lk_group *g, const b200lk_op *const *ops, const int32_t *nops, const int32_t *node_slot,
                                     const int32_t *dad_kind, const int32_t *dad_id, const double *length, int32_t save_log_value);
int b200lk_group_enqueue_branch_derv(b200lk_group *g, const b200lk_op *const *ops, const int32_t *nops, const int32_t *node_slot,
                                     const int32_t *dad_kind, const int32_t *dad_id, const double *length);
int b200lk_group_enqueue_logl_from_buffer(b200lk_group *g, const double *length);
int b200lk_group_finish(b200lk_group *g, double *logl, double *df, double *ddf, int32_t *status);
/* kernel launches issued by the group and its members since creation */
int64_t b200lk_group_launch_count(const b200lk_group *g);

/* ---- read-backs (computePatternLikelihood & friends, tree/phylotree.cpp:1458-1603; parity dumps) -------------- */
int b200lk_get_pattern_lh(b200lk_engine *e, double *out /* [nptn] */);
int b200lk_get_pattern_scaling(b200lk_engine *e, double *out /* [nptn] */);
int b200lk_get_pattern_lh_cat(b200lk_engine *e, double *out /* [nptn][ncat_mix], after transformPatternLhCat */);
int b200lk_get_theta(b200lk_engine *e, double *out /* [nptn][B] */);
int b200lk_get_buffer_scale_all(b200lk_engine *e, double *out /* [nptn] */);
int b200lk_get_scale_num(b200lk_engine *e, int32_t slot, uint16_t *out /* [nptn] or [nptn][ncat_mix] */);
int b200lk_get_partial(b200lk_engine *e, int32_t slot, double *out /* [nptn][B] */);
int b200lk_get_tip_partial_lh(b200lk_engine *e, double *out /* [(state_unknown+1)][nmix][n] */);
/* device pointer of _pattern_lh (length >= nptn doubles) for on-device consumers (UFBoot RELL) */
void *b200lk_pattern_lh_dev(b200lk_engine *e);

/* ---- UFBoot RELL on the device (IQTree::saveCurrentTree, tree/iqtree.cpp:3551-3637) --------------------------------
 * boot_samples: the resampled pattern frequencies of every bootstrap replicate, [nboot][nptn] floats (IQTree::boot_samples,
 * BootValType = float, tree/phylotree.h:54); uploaded once and kept resident (nboot = 0 frees them).
 * b200lk_rell: rell[b] = dotProduct((float)_pattern_lh, boot_samples[b], nptn) (PhyloTree::dotProductSIMD<float>,
 * tree/phylokernel.h:64-70) for all replicates, on the _pattern_lh left by the last computeLikelihoodBranch /
 * computeLikelihoodFromBuffer — 4*nboot bytes come back instead of 8*nptn going out per candidate tree.  Products in float as
 * in the reference, accumulated in double in a fixed order. The accept/replace bookkeeping (boot_logl, boot_counts, the
 * random tie-break) stays with the caller. */
int b200lk_set_boot_samples(b200lk_engine *e, const float *boot_samples, int32_t nboot);
int b200lk_rell(b200lk_engine *e, float *rell /* [nboot] */);
/* The accept / replace bookkeeping of the same loop on the device (tree/iqtree.cpp:3593-3632), for all replicates in one launch:
 *     better = rell > boot_logl[b] + epsilon;   a tie (|rell - boot_logl[b]| < epsilon) wins when uniform[b] <= 1/(boot_counts[b]+1)
 *     winner: boot_counts[b] = tie ? boot_counts[b] + 1 : 1;  boot_logl[b] = max(boot_logl[b], rell);  boot_orig_logl[b] = cur_logl
 * on the _pattern_lh of the last evaluation (the RELL scores are recomputed; they never leave the device).  boot_logl, boot_counts
 * and boot_orig_logl live on the device from b200lk_set_boot_samples on (-DBL_MAX, 0, -DBL_MAX: tree/iqtree.cpp:375-378;
 * b200lk_boot_reset restores that).  uniform: the caller's random numbers in [0,1), one per replicate (the reference draws
 * random_double(rstream) from per-thread streams, so its tie-breaks depend on the thread count; NULL: ties never win).
 * updated[b] = 1: the caller stores the current tree for replicate b (boot_trees[b] = tree_str); nboot BYTES come back per
 * candidate tree instead of 8*nptn going out.  epsilon = Params::ufboot_epsilon (0.5). */
int b200lk_boot_reset(b200lk_engine *e);
int b200lk_rell_update(b200lk_engine *e, double cur_logl, double epsilon, const double *uniform /* [nboot] or NULL */,
                       uint8_t *updated /* [nboot] or NULL */, int32_t *nupdated);
int b200lk_get_boot_state(b200lk_engine *e, double *boot_logl, int32_t *boot_counts, double *boot_orig_logl);

/* ---- alignment -> site patterns (Alignment::addPattern, alignment/alignment.cpp:1092-1133; buildPattern :1725) ------
 * states: [ntaxa][nsite] state codes, sequence-major.  Returns the distinct columns in FIRST-OCCURRENCE order as
 * patterns [ntaxa][nptn] (row stride nptn; capacity ntaxa*nsite), their frequencies ptn_freq [nptn] (capacity nsite) and
 * site_pattern [nsite] — what the reference's hash-map loop builds one site at a time; all-gap columns are kept, as there.
 * Stand-alone (no engine).  B200LK_ERR_STATE: four hash seeds in a row produced a 64-bit collision (never observed). */
int b200lk_compress_patterns(int32_t device, const uint8_t *states, int32_t ntaxa, int64_t nsite, uint8_t *patterns,
                             double *ptn_freq, int32_t *site_pattern, int64_t *nptn);

/* ---- measurement --------------------------------------------------------------------------------------------- */
/* When enabled, every call brackets its kernel groups with CUDA events on the engine stream. */
int b200lk_set_profiling(b200lk_engine *e, int32_t on);
/* milliseconds of the last call: [0] op staging + partial_info + partial-likelihood kernels, [1] root/logl kernels,
 * [2] theta/derv kernels, [3] whole call on the stream, [4] the partial-likelihood kernel launch(es) alone;
 * launches[0..2] = kernel launches in groups [0..2] */
int b200lk_last_timing(b200lk_engine *e, float ms[5], int32_t launches[3]);
/* total kernel launches issued by this engine since creation */
int64_t b200lk_launch_count(const b200lk_engine *e);

#ifdef __cplusplus
}
#endif
#endif /* B200LK_H_ */
