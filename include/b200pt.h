/*
 * b200pt.h — C binding of the host-side PhyloTree mirror (iqtree2_b200/csrc/phylotree_b200.h) so that non-C++
 * hosts (the Python tests, bench.py) can drive the same tree-level API the reference's callers use:
 * PhyloTree::computeLikelihood / computeLikelihoodBranch / computeLikelihoodDerv / computeLikelihoodFromBuffer /
 * computePatternLikelihood / optimizeOneBranch / optimizeAllBranches / clearAllPartialLH
 * (tree/phylotree.h:899-1048, 1136, 1340; tree/phylotree.cpp:668, 1233, 1458, 2594, 2698).
 * Branches are addressed by branch id (0 .. nbranch-1, Newick reading order); end `a`/`b` by node id.
 * Every call returns 0 on success, <0 on error (b200pt_last_error).
 */
#ifndef B200PT_H_
#define B200PT_H_
#include <stdint.h>
#include "b200lk.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200pt_tree b200pt_tree;

b200pt_tree *b200pt_create(int32_t device, int32_t safe_numeric, int32_t force_norm, int32_t schedule, int32_t fused_newton);
void b200pt_destroy(b200pt_tree *t);
const char *b200pt_last_error(const b200pt_tree *t);

/* names: ntaxa NUL-terminated strings in alignment-row order; states [ntaxa][nptn] */
int b200pt_set_alignment(b200pt_tree *t, int32_t ntaxa, const char *const *names, const uint8_t *states, int64_t nptn,
                         const double *ptn_freq, int32_t nstates, int32_t seq_type, int32_t state_unknown);
int b200pt_set_pattern_invar(b200pt_tree *t, const double *ptn_invar);
/* +ASC (Lewis): the last num_unobserved patterns of the alignment are ModelFactory::unobserved_ptns; nsite = aln->getNSite() */
int b200pt_set_ascertainment(b200pt_tree *t, int32_t asc_type, int64_t num_unobserved, int64_t nsite);
/* Params::robust_phy_keep / robust_median (--robust-phy, --robust-median): see b200lk_set_robust_phylo */
int b200pt_set_robust_phylo(b200pt_tree *t, double keep, int32_t median);
int b200pt_set_model(b200pt_tree *t, int32_t ncat, int32_t nmix, int32_t fused_mix_rate, const double *eval, const double *evec,
                     const double *inv_evec, const double *mix_weight, const double *cat_rate, const double *cat_prop);
int b200pt_read_tree(b200pt_tree *t, const char *newick);
int b200pt_get_tree(b200pt_tree *t, char *out, int64_t cap);
int b200pt_num_branches(b200pt_tree *t);
/* node ids of the two ends of every branch and the current lengths */
int b200pt_get_branches(b200pt_tree *t, int32_t *node_a, int32_t *node_b, double *lengths);
int b200pt_set_branch_length(b200pt_tree *t, int32_t branch, double length, int32_t clear_partials);

int b200pt_clear_all_partial_lh(b200pt_tree *t);
int b200pt_compute_likelihood(b200pt_tree *t, double *logl);
int b200pt_compute_likelihood_branch(b200pt_tree *t, int32_t branch, double *logl);
int b200pt_compute_likelihood_derv(b200pt_tree *t, int32_t branch, double *df, double *ddf);
int b200pt_compute_likelihood_from_buffer(b200pt_tree *t, double *logl);
int b200pt_compute_pattern_likelihood(b200pt_tree *t, double *pattern_lh, double *pattern_lh_cat /* may be NULL */);
int b200pt_optimize_one_branch(b200pt_tree *t, int32_t branch, int32_t clear_lh, int32_t max_nr_step);
int b200pt_optimize_all_branches(b200pt_tree *t, int32_t iterations, double tolerance, int32_t max_nr_step, double *logl);
/* PhyloTree::optimizeTreeLengthScaling (tree/phylotree.cpp:2457-2488): one factor for all branch lengths, Brent's method
 * (utils/optimization.cpp:185-373) over clearAllPartialLH + computeLikelihood; *scaling in (the factor the current lengths
 * stand for) and out; used by the M-step of RateFree::optimizeWithEM and by --tree-len-scale.  b200pt_scale_length: MTree::scaleLength
 * + clearAllPartialLH. */
int b200pt_optimize_tree_length_scaling(b200pt_tree *t, double min_scaling, double *scaling, double max_scaling, double gradient_epsilon,
                                        double *logl);
int b200pt_scale_length(b200pt_tree *t, double norm);
/* optimizeAllBranches queues each sweep as a whole (b200lk_enqueue_newton_branch): one synchronisation per sweep */
int b200pt_set_pipelined_sweep(b200pt_tree *t, int32_t on);
/* NNI evaluation on the resident CLVs — PhyloTree::getBestNNIForBran (tree/phylotree.cpp:4069-4322) on inner branch `branch`
 * (node1 = end a, node2 = end b of b200pt_get_branches; reverse != 0: the other way round — which end is node1 decides the
 * order in which the five branches are re-optimised, hence the last digits of the scores) with Params::nni5 = nni5: both alternative topologies are scored on the
 * engine's two spare CLV slots (the reference's nni_partial_lh) after re-optimising the central branch (nni5: and the four
 * adjacent ones, NNI_MAX_NR_STEP Newton steps each); the tree, its lengths and its CLVs are as before on return.
 * Move m swaps the first other neighbour of node1 with the m-th other neighbour of node2.  Outputs: score[2] =
 * NNIMove::newloglh; new_len[2][5] = NNIMove::newLen ([0] central, [1..2] node1's, [3..4] node2's other branches);
 * swap_nodes[2][2] = node ids of the swapped neighbours; outer_nodes[2][4] = node ids at the far ends of the branches of
 * new_len[1..4]; pattern_lh (may be NULL): [2][nptn] pattern log-likelihoods of the two topologies (NNIMove::ptnlh). */
int b200pt_get_best_nni_for_bran(b200pt_tree *t, int32_t branch, int32_t reverse, int32_t nni5, double *score, double *new_len, int32_t *swap_nodes,
                                 int32_t *outer_nodes, double *pattern_lh);
/* PhyloTree::doNNI (tree/phylotree.cpp:3947-4036) of move 0 / 1 of that branch, then changeNNIBrans (:4038-4067) with new_len
 * (5 doubles as returned above; NULL = keep the lengths) */
int b200pt_do_nni(b200pt_tree *t, int32_t branch, int32_t reverse, int32_t move, int32_t nni5, const double *new_len, int32_t clear_lh);
/* number of partial-likelihood updates planned so far (PhyloTree::num_partial_lh_computations) */
int64_t b200pt_num_partial_lh_computations(b200pt_tree *t);


/* -mem (Params::lh_mem_save = LM_MEM_SAVE, tree/phylotree.cpp:999-1021; slot table tree/memslot.cpp): recycle
 * max_lh_slots CLV slots instead of one per internal node.  max_mem_size: 0 = the minimum log2(T)+1 slots, <= 1 = that
 * fraction of the T-2 slots, > 1 = bytes of device memory for the slots.  Drops the current buffers. */
int b200pt_set_mem_save(b200pt_tree *t, int32_t on, double max_mem_size);
int b200pt_max_lh_slots(b200pt_tree *t);
/* Planning without execution: no engine is created; b200pt_plan_branch returns the traversal_info list of
 * computeLikelihoodBranch(branch) (ops in post-order, the slot of the node end, kind/id of the dad end, the branch
 * length) and updates every partial_lh_computed flag and slot as if the list had been executed. */
int b200pt_set_plan_only(b200pt_tree *t, int32_t on);
int b200pt_plan_branch(b200pt_tree *t, int32_t branch, b200lk_op *ops, int32_t cap, int32_t *nops, int32_t *node_slot,
                       int32_t *dad_kind, int32_t *dad_id, double *length);


/* ---- edge-linked partition model: PhyloSuperTreePlen (tree/phylosupertreeplen.cpp:243-412, tree/phylosupertree.cpp:697-726) ----
 * A super tree over `ntaxa` taxa and any number of partition trees (one device engine each, own alignment and model) whose
 * branch lengths are the super tree's times part_rate.  logL = sum_p logL_p, df = sum_p rate_p df_p, ddf = sum_p rate_p^2 ddf_p.
 * A taxon without data in a partition is a row of state_unknown.  All partitions are evaluated concurrently (one CUDA stream
 * per partition, work queued from num_threads host threads; 0 = OpenMP default). */
typedef struct b200st_tree b200st_tree;
b200st_tree *b200st_create(int32_t ntaxa, const char *const *names, int32_t num_threads);
void b200st_destroy(b200st_tree *t);
const char *b200st_last_error(const b200st_tree *t);
/* alignment + model of one partition (arguments as b200pt_set_alignment / b200pt_set_pattern_invar / b200pt_set_model);
 * returns the partition index (>= 0) or -1 */
int b200st_add_partition(b200st_tree *t, int32_t device, int32_t safe_numeric, int32_t schedule, const uint8_t *states, int64_t nptn,
                         const double *ptn_freq, const double *ptn_invar, int32_t nstates, int32_t seq_type, int32_t state_unknown,
                         int32_t ncat, int32_t nmix, int32_t fused_mix_rate, const double *eval, const double *evec,
                         const double *inv_evec, const double *mix_weight, const double *cat_rate, const double *cat_prop,
                         double part_rate);
int b200st_read_tree(b200st_tree *t, const char *newick);
int b200st_get_tree(b200st_tree *t, char *out, int64_t cap);
int b200st_num_partitions(b200st_tree *t);
int b200st_num_branches(b200st_tree *t);
int b200st_get_branches(b200st_tree *t, int32_t *node_a, int32_t *node_b, double *lengths);
int b200st_set_branch_length(b200st_tree *t, int32_t branch, double length, int32_t clear_partials);
int b200st_set_part_rate(b200st_tree *t, int32_t part, double rate);
int b200st_clear_all_partial_lh(b200st_tree *t);
int b200st_compute_likelihood(b200st_tree *t, double *logl);
int b200st_compute_likelihood_branch(b200st_tree *t, int32_t branch, double *logl);
/* derivatives of logL with respect to the super branch length, evaluated at `length` (computeFuncDerv with the signs of logL) */
int b200st_compute_likelihood_derv(b200st_tree *t, int32_t branch, double length, double *df, double *ddf);
int b200st_optimize_one_branch(b200st_tree *t, int32_t branch, int32_t clear_lh, int32_t max_nr_step);
int b200st_optimize_all_branches(b200st_tree *t, int32_t iterations, double tolerance, int32_t max_nr_step, double *logl);
/* NNI evaluation for the linked model (the role of PhyloSuperTreePlen::getBestNNIForBran, tree/phylosupertreeplen.cpp:1100-1500):
 * arguments and outputs as b200pt_get_best_nni_for_bran; the move is the same swap in the super tree and in every partition tree,
 * each partition on its own engine's spare slots, the five branches re-optimised with the linked derivatives.  b200st_do_nni
 * applies move 0 / 1 to the super tree and every partition tree (new_len: super-tree lengths; partitions get them times their rate). */
int b200st_get_best_nni_for_bran(b200st_tree *t, int32_t branch, int32_t reverse, int32_t nni5, double *score, double *new_len,
                                 int32_t *swap_nodes, int32_t *outer_nodes);
int b200st_do_nni(b200st_tree *t, int32_t branch, int32_t reverse, int32_t move, int32_t nni5, const double *new_len, int32_t clear_lh);
/* part_info[p].cur_score of the last evaluation */
int b200st_get_part_scores(b200st_tree *t, double *scores);
int64_t b200st_launch_count(b200st_tree *t);
int64_t b200st_device_bytes(b200st_tree *t);


/* ---- edge-unlinked partition model: PhyloSuperTree (`-Q` / `-sp`; tree/phylosupertree.cpp:653-791) ----
 * Every partition is a b200pt_tree of its own (own taxa, alignment, model, tree = the induced subtree, own branch lengths),
 * configured through the b200pt_* calls and BORROWED here: it must outlive the b200ut_tree.  logL = sum_p logL_p; the partitions
 * are evaluated concurrently, most expensive first (computePartitionOrder), from num_threads host threads (0 = OpenMP default),
 * one device engine and stream each.  b200ut_optimize_all_branches runs every partition's optimizeAllBranches with
 * tolerance / min(npartitions, 10) (:773-791).  pattern_lh (may be NULL in b200ut_compute_likelihood): the partitions'
 * pattern log-likelihoods one after the other (:703-709, 736-763). */
typedef struct b200ut_tree b200ut_tree;
b200ut_tree *b200ut_create(int32_t num_threads);
void b200ut_destroy(b200ut_tree *t);
const char *b200ut_last_error(const b200ut_tree *t);
int b200ut_add_partition(b200ut_tree *t, b200pt_tree *part);
int b200ut_num_partitions(b200ut_tree *t);
int64_t b200ut_num_patterns(b200ut_tree *t);
int b200ut_clear_all_partial_lh(b200ut_tree *t);
int b200ut_compute_likelihood(b200ut_tree *t, double *logl, double *pattern_lh);
int b200ut_compute_pattern_likelihood(b200ut_tree *t, double *pattern_lh, double *logl);
int b200ut_optimize_all_branches(b200ut_tree *t, int32_t iterations, double tolerance, int32_t max_nr_step, double *logl);
int b200ut_get_part_scores(b200ut_tree *t, double *scores);

#ifdef __cplusplus
}
#endif
#endif
