// phylotree_capi.cpp — extern "C" binding of b200::PhyloTree (include/b200pt.h).
#include <algorithm>
#include <cstring>
#include <exception>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200pt.h"
#include "phylotree_b200.h"
#include "supertree_b200.h"

struct b200pt_tree {
    b200::PhyloTree tree;
    std::string err;
};

static thread_local std::string g_err;

#define GUARD(...)                                    \
    try { __VA_ARGS__; return 0; }                    \
    catch (const std::exception &ex) { t->err = ex.what(); return -1; }

extern "C" {

b200pt_tree *b200pt_create(int32_t device, int32_t safe_numeric, int32_t force_norm, int32_t schedule, int32_t fused_newton) {
    auto *t = new b200pt_tree();
    t->tree.params.device = device;
    t->tree.params.safe_numeric = safe_numeric != 0;
    t->tree.params.force_norm = force_norm != 0;
    t->tree.params.schedule = schedule;
    t->tree.params.fused_newton = fused_newton != 0;
    return t;
}
void b200pt_destroy(b200pt_tree *t) { delete t; }
const char *b200pt_last_error(const b200pt_tree *t) { return t ? t->err.c_str() : g_err.c_str(); }

int b200pt_set_alignment(b200pt_tree *t, int32_t ntaxa, const char *const *names, const uint8_t *states, int64_t nptn,
                         const double *ptn_freq, int32_t nstates, int32_t seq_type, int32_t state_unknown) {
    GUARD({
        std::vector<std::string> nm(names, names + ntaxa);
        t->tree.setAlignment(nm, states, nptn, ptn_freq, nstates, seq_type, state_unknown);
    })
}
int b200pt_set_pattern_invar(b200pt_tree *t, const double *ptn_invar) { GUARD(t->tree.setPatternInvar(ptn_invar)) }
int b200pt_set_ascertainment(b200pt_tree *t, int32_t asc_type, int64_t num_unobserved, int64_t nsite) {
    GUARD(t->tree.setAscertainment(asc_type, num_unobserved, nsite))
}
int b200pt_set_model(b200pt_tree *t, int32_t ncat, int32_t nmix, int32_t fused, const double *eval, const double *evec,
                     const double *inv_evec, const double *mix_weight, const double *cat_rate, const double *cat_prop) {
    GUARD(t->tree.setModel(ncat, nmix, fused != 0, eval, evec, inv_evec, mix_weight, cat_rate, cat_prop))
}
int b200pt_read_tree(b200pt_tree *t, const char *newick) { GUARD(t->tree.readTreeString(newick)) }
int b200pt_get_tree(b200pt_tree *t, char *out, int64_t cap) {
    GUARD({
        std::string s = t->tree.getTreeString();
        if ((int64_t)s.size() + 1 > cap) throw std::runtime_error("buffer too small");
        std::memcpy(out, s.c_str(), s.size() + 1);
    })
}
int b200pt_num_branches(b200pt_tree *t) { return t->tree.numBranches(); }
int b200pt_get_branches(b200pt_tree *t, int32_t *node_a, int32_t *node_b, double *lengths) {
    GUARD({
        std::vector<b200::PhyloNode *> a, b;
        t->tree.getBranches(a, b);
        std::vector<double> len;
        t->tree.saveBranchLengths(len);
        for (size_t k = 0; k < a.size(); k++) { node_a[k] = a[k]->id; node_b[k] = b[k]->id; lengths[k] = len[k]; }
    })
}
static void ends(b200pt_tree *t, int branch, b200::PhyloNode *&a, b200::PhyloNode *&b) {
    std::vector<b200::PhyloNode *> na, nb;
    t->tree.getBranches(na, nb);
    if (branch < 0 || branch >= (int)na.size()) throw std::runtime_error("branch id out of range");
    a = na[branch]; b = nb[branch];
}
int b200pt_set_branch_length(b200pt_tree *t, int32_t branch, double length, int32_t clear_partials) {
    GUARD({
        b200::PhyloNode *a, *b;
        ends(t, branch, a, b);
        a->findNeighbor(b)->length = length;
        b->findNeighbor(a)->length = length;
        if (clear_partials) { a->clearReversePartialLh(b); b->clearReversePartialLh(a); t->tree.theta_computed = false; }
    })
}
int b200pt_clear_all_partial_lh(b200pt_tree *t) { GUARD(t->tree.clearAllPartialLH()) }
int b200pt_compute_likelihood(b200pt_tree *t, double *logl) { GUARD(*logl = t->tree.computeLikelihood()) }
int b200pt_compute_likelihood_branch(b200pt_tree *t, int32_t branch, double *logl) {
    GUARD({
        b200::PhyloNode *a, *b;
        ends(t, branch, a, b);
        t->tree.current_it = a->findNeighbor(b);
        t->tree.current_it_back = b->findNeighbor(a);
        *logl = t->tree.computeLikelihoodBranch(t->tree.current_it, a);
    })
}
int b200pt_compute_likelihood_derv(b200pt_tree *t, int32_t branch, double *df, double *ddf) {
    GUARD({
        b200::PhyloNode *a, *b;
        ends(t, branch, a, b);
        b200::PhyloNeighbor *it = a->findNeighbor(b);
        if (t->tree.current_it != it && t->tree.current_it_back != it) t->tree.theta_computed = false;
        t->tree.current_it = it;
        t->tree.current_it_back = b->findNeighbor(a);
        t->tree.computeLikelihoodDerv(it, a, df, ddf);
    })
}
int b200pt_compute_likelihood_from_buffer(b200pt_tree *t, double *logl) { GUARD(*logl = t->tree.computeLikelihoodFromBuffer()) }
int b200pt_compute_pattern_likelihood(b200pt_tree *t, double *pattern_lh, double *pattern_lh_cat) {
    GUARD(t->tree.computePatternLikelihood(pattern_lh, nullptr, pattern_lh_cat))
}
int b200pt_optimize_one_branch(b200pt_tree *t, int32_t branch, int32_t clear_lh, int32_t max_nr_step) {
    GUARD({
        b200::PhyloNode *a, *b;
        ends(t, branch, a, b);
        t->tree.optimizeOneBranch(a, b, clear_lh != 0, max_nr_step);
    })
}
int b200pt_optimize_all_branches(b200pt_tree *t, int32_t iterations, double tolerance, int32_t max_nr_step, double *logl) {
    GUARD(*logl = t->tree.optimizeAllBranches(iterations, tolerance, max_nr_step))
}
// node1 = end a, node2 = end b of the branch as b200pt_get_branches reports them; move m swaps the first other neighbour of
// node1 with the m-th other neighbour of node2 (tree/phylotree.cpp:4168-4183)
static void nniMovesOf(b200pt_tree *t, int branch, int reverse, b200::NNIMove mv[2]) {
    b200::PhyloNode *a, *b;
    ends(t, branch, a, b);
    if (reverse) std::swap(a, b);
    if (a->isLeaf() || b->isLeaf() || a->neighbors.size() != 3 || b->neighbors.size() != 3)
        throw std::runtime_error("NNI needs an inner branch between two degree-3 nodes");
    int k1 = -1, cnt = 0;
    for (int k = 0; k < 3 && k1 < 0; k++) if (a->neighbors[k]->node != b) k1 = k;
    for (int k = 0; k < 3; k++)
        if (b->neighbors[k]->node != a) {
            mv[cnt].node1 = a; mv[cnt].node2 = b;
            mv[cnt].node1Nei_it = k1; mv[cnt].node2Nei_it = k;
            cnt++;
        }
}
int b200pt_get_best_nni_for_bran(b200pt_tree *t, int32_t branch, int32_t reverse, int32_t nni5, double *score, double *new_len, int32_t *swap_nodes,
                                 int32_t *outer_nodes, double *pattern_lh) {
    GUARD({
        b200::NNIMove mv[2];
        nniMovesOf(t, branch, reverse, mv);
        const int64_t nptn = t->tree.getAlnNPattern();
        if (pattern_lh) { mv[0].ptnlh = pattern_lh; mv[1].ptnlh = pattern_lh + nptn; }
        t->tree.params.nni5 = nni5 != 0;
        b200::PhyloNode *node1 = mv[0].node1, *node2 = mv[0].node2;
        t->tree.getBestNNIForBran(node1, node2, mv);
        for (int m = 0; m < 2; m++) {
            score[m] = mv[m].newloglh;
            for (int i = 0; i < 5; i++) new_len[m * 5 + i] = (i == 0 || nni5) ? mv[m].newLen[i] : 0.0;
            b200::PhyloNode *sa = node1->neighbors[mv[m].node1Nei_it]->node, *sc = node2->neighbors[mv[m].node2Nei_it]->node;
            swap_nodes[m * 2] = sa->id; swap_nodes[m * 2 + 1] = sc->id;
            int k = 0;   // the branches of newLen[1..4]: node1's other neighbours, then node2's, in the swapped topology
            for (auto *it : node1->neighbors) if (it->node != node2) outer_nodes[m * 4 + k++] = (it->node == sa ? sc : it->node)->id;
            for (auto *it : node2->neighbors) if (it->node != node1) outer_nodes[m * 4 + k++] = (it->node == sc ? sa : it->node)->id;
        }
    })
}
int b200pt_do_nni(b200pt_tree *t, int32_t branch, int32_t reverse, int32_t move, int32_t nni5, const double *new_len, int32_t clear_lh) {
    GUARD({
        if (move < 0 || move > 1) throw std::runtime_error("move must be 0 or 1");
        b200::NNIMove mv[2];
        nniMovesOf(t, branch, reverse, mv);
        t->tree.doNNI(mv[move], clear_lh != 0);
        if (new_len) {
            t->tree.params.nni5 = nni5 != 0;
            for (int i = 0; i < 5; i++) mv[move].newLen[i] = (i == 0 || nni5) ? new_len[i] : 0.0;
            t->tree.changeNNIBrans(mv[move]);
        }
    })
}
int b200pt_optimize_tree_length_scaling(b200pt_tree *t, double min_scaling, double *scaling, double max_scaling, double gradient_epsilon,
                                        double *logl) {
    GUARD(*logl = t->tree.optimizeTreeLengthScaling(min_scaling, *scaling, max_scaling, gradient_epsilon))
}
int b200pt_scale_length(b200pt_tree *t, double norm) {
    GUARD({
        t->tree.scaleLength(norm);
        t->tree.clearAllPartialLH();
    })
}
int b200pt_set_robust_phylo(b200pt_tree *t, double keep, int32_t median) {
    GUARD({
        t->tree.params.robust_phy_keep = keep;
        t->tree.params.robust_median = median != 0;
        if (t->tree.engine() && b200lk_set_robust_phylo(t->tree.engine(), keep, median) != 0)
            throw std::runtime_error(b200lk_last_error(t->tree.engine()));
    })
}
int b200pt_set_pipelined_sweep(b200pt_tree *t, int32_t on) { t->tree.params.pipelined_sweep = on != 0; return 0; }
int64_t b200pt_num_partial_lh_computations(b200pt_tree *t) { return t->tree.num_partial_lh_computations; }

int b200pt_set_mem_save(b200pt_tree *t, int32_t on, double max_mem_size) {
    GUARD({
        t->tree.deleteAllPartialLh();
        t->tree.params.lh_mem_save = on ? b200::LM_MEM_SAVE : b200::LM_PER_NODE;
        t->tree.params.max_mem_size = max_mem_size;
    })
}
int b200pt_set_plan_only(b200pt_tree *t, int32_t on) {
    GUARD({
        t->tree.deleteAllPartialLh();
        t->tree.params.plan_only = on != 0;
    })
}
int b200pt_max_lh_slots(b200pt_tree *t) {
    try { t->tree.initializeAllPartialLh(); return t->tree.max_lh_slots; }
    catch (const std::exception &ex) { t->err = ex.what(); return -1; }
}
int b200pt_plan_branch(b200pt_tree *t, int32_t branch, b200lk_op *ops, int32_t cap, int32_t *nops, int32_t *node_slot,
                       int32_t *dad_kind, int32_t *dad_id, double *length) {
    GUARD({
        b200::PhyloNode *a, *b;
        ends(t, branch, a, b);
        std::vector<b200lk_op> v;
        int ns, dk, di;
        t->tree.planBranch(a->findNeighbor(b), a, v, ns, dk, di, *length);
        if ((int)v.size() > cap) throw std::runtime_error("op buffer too small");
        std::memcpy(ops, v.data(), v.size() * sizeof(b200lk_op));
        *nops = (int)v.size(); *node_slot = ns; *dad_kind = dk; *dad_id = di;
    })
}


// ---------------------------------------------------------------------------------------------------------------
// PhyloSuperTreePlen
// ---------------------------------------------------------------------------------------------------------------
struct b200st_tree {
    b200::PhyloSuperTreePlen tree;
    std::vector<std::string> names;
    std::string err;
};

b200st_tree *b200st_create(int32_t ntaxa, const char *const *names, int32_t num_threads) {
    auto *t = new b200st_tree();
    try {
        t->names.assign(names, names + ntaxa);
        t->tree.setTaxa(t->names);
        t->tree.num_threads = num_threads;
    } catch (const std::exception &ex) { g_err = ex.what(); delete t; return nullptr; }
    return t;
}
void b200st_destroy(b200st_tree *t) { delete t; }
const char *b200st_last_error(const b200st_tree *t) { return t ? t->err.c_str() : g_err.c_str(); }

int b200st_add_partition(b200st_tree *t, int32_t device, int32_t safe_numeric, int32_t schedule, const uint8_t *states, int64_t nptn,
                         const double *ptn_freq, const double *ptn_invar, int32_t nstates, int32_t seq_type, int32_t state_unknown,
                         int32_t ncat, int32_t nmix, int32_t fused, const double *eval, const double *evec, const double *inv_evec,
                         const double *mix_weight, const double *cat_rate, const double *cat_prop, double part_rate) {
    b200::PhyloTree *p = nullptr;
    try {
        p = new b200::PhyloTree();
        p->params.device = device;
        p->params.safe_numeric = safe_numeric != 0;
        p->params.schedule = schedule;
        p->setAlignment(t->names, states, nptn, ptn_freq, nstates, seq_type, state_unknown);
        if (ptn_invar) p->setPatternInvar(ptn_invar);
        p->setModel(ncat, nmix, fused != 0, eval, evec, inv_evec, mix_weight, cat_rate, cat_prop);
        return t->tree.addPartition(p, part_rate);
    } catch (const std::exception &ex) { t->err = ex.what(); delete p; return -1; }
}
int b200st_read_tree(b200st_tree *t, const char *newick) { GUARD(t->tree.readSuperTree(newick)) }
int b200st_get_tree(b200st_tree *t, char *out, int64_t cap) {
    GUARD({
        std::string s = t->tree.getTreeString();
        if ((int64_t)s.size() + 1 > cap) throw std::runtime_error("buffer too small");
        std::memcpy(out, s.c_str(), s.size() + 1);
    })
}
int b200st_num_partitions(b200st_tree *t) { return t->tree.numPartitions(); }
int b200st_num_branches(b200st_tree *t) { return t->tree.numBranches(); }
int b200st_get_branches(b200st_tree *t, int32_t *node_a, int32_t *node_b, double *lengths) {
    GUARD({
        std::vector<b200::PhyloNode *> a, b;
        t->tree.getBranches(a, b);
        std::vector<double> len;
        t->tree.saveBranchLengths(len);
        for (size_t k = 0; k < a.size(); k++) { node_a[k] = a[k]->id; node_b[k] = b[k]->id; lengths[k] = len[k]; }
    })
}
static void st_ends(b200st_tree *t, int branch, b200::PhyloNode *&a, b200::PhyloNode *&b) {
    std::vector<b200::PhyloNode *> na, nb;
    t->tree.getBranches(na, nb);
    if (branch < 0 || branch >= (int)na.size()) throw std::runtime_error("branch id out of range");
    a = na[branch]; b = nb[branch];
}
int b200st_set_branch_length(b200st_tree *t, int32_t branch, double length, int32_t clear_partials) {
    GUARD({
        b200::PhyloNode *a, *b;
        st_ends(t, branch, a, b);
        a->findNeighbor(b)->length = length;
        b->findNeighbor(a)->length = length;
        for (int p = 0; p < t->tree.numPartitions(); p++) {
            b200::PhyloTree *pt = t->tree.parts[p];
            b200::PhyloNeighbor *n1 = pt->branchNeighbor(branch, b->id), *n2 = pt->branchNeighbor(branch, a->id);
            n1->length = n2->length = length * t->tree.part_rate[p];
            if (clear_partials) { n2->node->clearReversePartialLh(n1->node); n1->node->clearReversePartialLh(n2->node); pt->theta_computed = false; }
        }
    })
}
int b200st_set_part_rate(b200st_tree *t, int32_t part, double rate) { GUARD(t->tree.setPartRate(part, rate)) }
int b200st_clear_all_partial_lh(b200st_tree *t) { GUARD(t->tree.clearAllPartialLH()) }
int b200st_compute_likelihood(b200st_tree *t, double *logl) { GUARD(*logl = t->tree.computeLikelihood()) }
int b200st_compute_likelihood_branch(b200st_tree *t, int32_t branch, double *logl) {
    GUARD({
        b200::PhyloNode *a, *b;
        st_ends(t, branch, a, b);
        *logl = t->tree.computeLikelihoodBranch(a->findNeighbor(b), a);
    })
}
int b200st_compute_likelihood_derv(b200st_tree *t, int32_t branch, double length, double *df, double *ddf) {
    GUARD({
        b200::PhyloNode *a, *b;
        st_ends(t, branch, a, b);
        b200::PhyloNeighbor *it = a->findNeighbor(b);
        if (t->tree.current_it != it)
            for (auto *p : t->tree.parts) p->theta_computed = false;
        t->tree.current_it = it;
        t->tree.current_it_back = b->findNeighbor(a);
        double f, g;
        t->tree.computeFuncDerv(length, f, g);
        *df = -f; *ddf = -g;
    })
}
int b200st_optimize_one_branch(b200st_tree *t, int32_t branch, int32_t clear_lh, int32_t max_nr_step) {
    GUARD({
        b200::PhyloNode *a, *b;
        st_ends(t, branch, a, b);
        t->tree.optimizeOneBranch(a, b, clear_lh != 0, max_nr_step);
    })
}
int b200st_optimize_all_branches(b200st_tree *t, int32_t iterations, double tolerance, int32_t max_nr_step, double *logl) {
    GUARD(*logl = t->tree.optimizeAllBranches(iterations, tolerance, max_nr_step))
}
int b200st_get_part_scores(b200st_tree *t, double *scores) {
    GUARD(for (int p = 0; p < t->tree.numPartitions(); p++) scores[p] = t->tree.cur_score[p])
}
static void superNniMoves(b200st_tree *t, int branch, int reverse, b200::NNIMove mv[2]) {
    std::vector<b200::PhyloNode *> na, nb;
    t->tree.getBranches(na, nb);
    if (branch < 0 || branch >= (int)na.size()) throw std::runtime_error("branch id out of range");
    b200::PhyloNode *a = na[branch], *b = nb[branch];
    if (reverse) std::swap(a, b);
    if (a->isLeaf() || b->isLeaf() || a->neighbors.size() != 3 || b->neighbors.size() != 3)
        throw std::runtime_error("NNI needs an inner branch between two degree-3 nodes");
    b200::PhyloTree::defaultNNIMoves(a, b, mv);
    mv[0].node1 = mv[1].node1 = a;
    mv[0].node2 = mv[1].node2 = b;
}
int b200st_get_best_nni_for_bran(b200st_tree *t, int32_t branch, int32_t reverse, int32_t nni5, double *score, double *new_len,
                                 int32_t *swap_nodes, int32_t *outer_nodes) {
    GUARD({
        b200::NNIMove mv[2];
        superNniMoves(t, branch, reverse, mv);
        t->tree.params.nni5 = nni5 != 0;
        b200::PhyloNode *node1 = mv[0].node1, *node2 = mv[0].node2;
        t->tree.getBestNNIForBran(node1, node2, mv);
        for (int m = 0; m < 2; m++) {
            score[m] = mv[m].newloglh;
            for (int i = 0; i < 5; i++) new_len[m * 5 + i] = (i == 0 || nni5) ? mv[m].newLen[i] : 0.0;
            b200::PhyloNode *sa = node1->neighbors[mv[m].node1Nei_it]->node, *sc = node2->neighbors[mv[m].node2Nei_it]->node;
            swap_nodes[m * 2] = sa->id; swap_nodes[m * 2 + 1] = sc->id;
            int k = 0;
            for (auto *it : node1->neighbors) if (it->node != node2) outer_nodes[m * 4 + k++] = (it->node == sa ? sc : it->node)->id;
            for (auto *it : node2->neighbors) if (it->node != node1) outer_nodes[m * 4 + k++] = (it->node == sc ? sa : it->node)->id;
        }
    })
}
int b200st_do_nni(b200st_tree *t, int32_t branch, int32_t reverse, int32_t move, int32_t nni5, const double *new_len, int32_t clear_lh) {
    GUARD({
        if (move < 0 || move > 1) throw std::runtime_error("move must be 0 or 1");
        b200::NNIMove mv[2];
        superNniMoves(t, branch, reverse, mv);
        t->tree.params.nni5 = nni5 != 0;
        if (new_len) for (int i = 0; i < 5; i++) mv[move].newLen[i] = (i == 0 || nni5) ? new_len[i] : 0.0;
        t->tree.doNNIAll(mv[move], clear_lh != 0, new_len != nullptr);
    })
}
int64_t b200st_launch_count(b200st_tree *t) { return t->tree.launches(); }
int64_t b200st_device_bytes(b200st_tree *t) {
    int64_t n = 0;
    for (auto *p : t->tree.parts) if (p->engine()) n += b200lk_device_bytes(p->engine());
    return n;
}

// ---- edge-unlinked partition model: PhyloSuperTree (tree/phylosupertree.cpp:653-791) ----
struct b200ut_tree {
    b200::PhyloSuperTree tree;
    std::string err;
};
b200ut_tree *b200ut_create(int32_t num_threads) {
    auto *t = new b200ut_tree();
    t->tree.num_threads = num_threads;
    return t;
}
void b200ut_destroy(b200ut_tree *t) { delete t; }
const char *b200ut_last_error(const b200ut_tree *t) { return t ? t->err.c_str() : g_err.c_str(); }
int b200ut_add_partition(b200ut_tree *t, b200pt_tree *part) {
    try { return t->tree.addPartition(part ? &part->tree : nullptr); }
    catch (const std::exception &ex) { t->err = ex.what(); return -1; }
}
int b200ut_num_partitions(b200ut_tree *t) { return t->tree.numPartitions(); }
int64_t b200ut_num_patterns(b200ut_tree *t) { return t->tree.numPatterns(); }
int b200ut_clear_all_partial_lh(b200ut_tree *t) { GUARD(t->tree.clearAllPartialLH()) }
int b200ut_compute_likelihood(b200ut_tree *t, double *logl, double *pattern_lh) { GUARD(*logl = t->tree.computeLikelihood(pattern_lh)) }
int b200ut_compute_pattern_likelihood(b200ut_tree *t, double *pattern_lh, double *logl) { GUARD(t->tree.computePatternLikelihood(pattern_lh, logl)) }
int b200ut_optimize_all_branches(b200ut_tree *t, int32_t iterations, double tolerance, int32_t max_nr_step, double *logl) {
    GUARD(*logl = t->tree.optimizeAllBranches(iterations, tolerance, max_nr_step))
}
int b200ut_get_part_scores(b200ut_tree *t, double *scores) {
    GUARD(for (size_t p = 0; p < t->tree.cur_score.size(); p++) scores[p] = t->tree.cur_score[p])
}

}  // extern "C"
