// supertree_b200.cpp — see supertree_b200.h.  Reference: tree/phylosupertreeplen.cpp, tree/phylosupertree.cpp (IQ-TREE 2.4.0).
#include "supertree_b200.h"

#include <algorithm>
#include <cmath>
#include <exception>
#include <numeric>
#include <stdexcept>
#include <string>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#ifdef _OPENMP
#include <cfloat>
#include <omp.h>
#endif

namespace b200 {

namespace {
// run f(part) for every partition from OpenMP threads (the reference's `omp parallel for schedule(dynamic)` over part_order),
// carrying the first exception out of the parallel region
template <class F>
void for_each_partition(const std::vector<int> &order, int num_threads, F f) {
    std::exception_ptr err;
    const int n = (int)order.size();
#ifdef _OPENMP
    const int nt = num_threads > 0 ? num_threads : std::min(omp_get_max_threads(), 16);
#pragma omp parallel for schedule(dynamic) num_threads(nt) if (n > 1 && nt > 1)
#endif
    for (int j = 0; j < n; j++) {
        try { f(order[j]); }
        catch (...) {
#ifdef _OPENMP
#pragma omp critical(b200_supertree_err)
#endif
            if (!err) err = std::current_exception();
        }
    }
    (void)num_threads;
    if (err) std::rethrow_exception(err);
}
}  // namespace

PhyloSuperTreePlen::PhyloSuperTreePlen() {
    // One stream per partition only overlaps if the streams do not share hardware work queues: the driver's default is 8
    // connections, on which 500 partition streams serialise each other (config 5: 29.7 ms per evaluation against 18.9 ms with
    // 32).  The variable is read when the CUDA context is created, so this helps hosts that have not touched CUDA yet; others
    // export CUDA_DEVICE_MAX_CONNECTIONS=32 themselves.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    params.plan_only = true;        // the super tree owns no device buffers: topology, linked lengths, flags only
    params.fused_newton = false;    // Newton-Raphson runs on the host over the summed derivatives (minimizeNewton)
}

PhyloSuperTreePlen::~PhyloSuperTreePlen() {
    dropGroup();
    for (auto *p : parts) delete p;
}

void PhyloSuperTreePlen::dropGroup() {
    if (group_ && getenv("B200ST_TIMING") && timing_[3] > 0)
        fprintf(stderr, "[b200st] grouped evaluations: %.0f calls; per call: host planning %.3f ms, group enqueue %.3f ms, wait %.3f ms\n", timing_[3],
                1e3 * timing_[0] / timing_[3], 1e3 * timing_[1] / timing_[3], 1e3 * timing_[2] / timing_[3]);
    if (group_) b200lk_group_destroy(group_);
    group_ = nullptr;
    group_engines_.clear();
}

// all partitions on one device -> one group over their engines (created by the partitions' initializeAllPartialLh)
bool PhyloSuperTreePlen::ensureGroup() {
    if (!use_group || parts.empty() || getenv("B200ST_NO_GROUP")) return false;
    // a partition tree re-creates its engine when its alignment / model shape changes: the group follows
    std::vector<b200lk_engine *> engines;
    for (auto *p : parts) {
        p->initializeAllPartialLh();
        if (!p->engine() || p->params.device != parts[0]->params.device || p->params.device < 0) { dropGroup(); return false; }
        engines.push_back(p->engine());
    }
    if (group_ && engines == group_engines_) return true;
    dropGroup();
    if (b200lk_group_create(engines.data(), (int32_t)engines.size(), &group_) != 0)
        throw std::runtime_error(std::string("b200lk_group_create: ") + b200lk_group_last_error(nullptr));
    group_engines_ = engines;
    return true;
}

// what: 1 = branch logL, 2 = branch derivatives, 3 = logL from the theta buffers.  plan(p, EvalPlan&) runs on host threads.
template <class PlanFn>
void PhyloSuperTreePlen::groupEvaluate(int what, PlanFn plan, double *logl, double *df, double *ddf) {
    const size_t np = parts.size();
    std::vector<PhyloTree::EvalPlan> plans(np);
    std::vector<const b200lk_op *> ops(np);
    std::vector<int32_t> nops(np), node_slot(np), dad_kind(np), dad_id(np);
    std::vector<double> length(np);
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    if (what != 3) for_each_partition(part_order, num_threads, [&](int p) { plan(p, plans[p]); });
    for (size_t p = 0; p < np; p++) {
        if (what == 3) { length[p] = parts[p]->current_it->length; continue; }
        ops[p] = plans[p].ops; nops[p] = plans[p].nops; node_slot[p] = plans[p].node_slot; dad_kind[p] = plans[p].dad_kind;
        dad_id[p] = plans[p].dad_id; length[p] = plans[p].length;
    }
    int rc;
    const double t1 = now();
    if (what == 1) rc = b200lk_group_enqueue_branch_logl(group_, ops.data(), nops.data(), node_slot.data(), dad_kind.data(), dad_id.data(), length.data(), 1);
    else if (what == 2) rc = b200lk_group_enqueue_branch_derv(group_, ops.data(), nops.data(), node_slot.data(), dad_kind.data(), dad_id.data(), length.data());
    else rc = b200lk_group_enqueue_logl_from_buffer(group_, length.data());
    const double t2 = now();
    if (rc >= 0) rc = b200lk_group_finish(group_, logl, df, ddf, nullptr);
    if (++timing_calls_ > 4) { timing_[0] += t1 - t0; timing_[1] += t2 - t1; timing_[2] += now() - t2; timing_[3] += 1.0; }   // (skip the calls that allocate)
    if (rc < 0) {
        const std::string msg = std::string("grouped evaluation: ") + b200lk_group_last_error(group_);
        if (what != 3) for (auto *p : parts) p->planFailed();
        throw std::runtime_error(msg);
    }
    if (what != 3) for (auto *p : parts) p->planDone(what == 2);
}

void PhyloSuperTreePlen::setTaxa(const std::vector<std::string> &names) {
    // a one-pattern placeholder alignment: the super tree needs the names (leaf ids = alignment rows), never the data
    std::vector<uint8_t> st(names.size(), 0);
    const double one = 1.0;
    setAlignment(names, st.data(), 1, &one, 4, B200LK_SEQ_GENERIC, 4);
    const double ev[4] = {0, 0, 0, 0}, m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, w = 1.0, r = 1.0, pr = 1.0;
    setModel(1, 1, false, ev, m, m, &w, &r, &pr);
}

int PhyloSuperTreePlen::addPartition(PhyloTree *part, double rate) {
    dropGroup();
    parts.push_back(part);
    part_rate.push_back(rate);
    cur_score.push_back(0.0);
    part_order.clear();
    return (int)parts.size() - 1;
}

void PhyloSuperTreePlen::setPartRate(int part, double rate) {
    part_rate.at(part) = rate;
    parts[part]->clearAllPartialLH();
    mapTrees();
}

void PhyloSuperTreePlen::readSuperTree(const std::string &newick) {
    readTreeString(newick);
    for (auto *p : parts) {
        p->readTreeString(newick);
        if (p->numBranches() != numBranches() || p->nodeNum != nodeNum) throw std::runtime_error("partition tree does not match the super tree");
    }
    mapTrees();
}

void PhyloSuperTreePlen::mapTrees() {
    std::vector<double> len, plen;
    saveBranchLengths(len);
    plen.resize(len.size());
    for (size_t p = 0; p < parts.size(); p++) {
        for (size_t b = 0; b < len.size(); b++) plen[b] = len[b] * part_rate[p];
        parts[p]->restoreBranchLengths(plen);
    }
}

void PhyloSuperTreePlen::restoreBranchLengths(const std::vector<double> &lenvec) {
    PhyloTree::restoreBranchLengths(lenvec);
    mapTrees();
}

void PhyloSuperTreePlen::clearAllPartialLH() {
    PhyloTree::clearAllPartialLH();
    for (auto *p : parts) p->clearAllPartialLH();
    std::fill(cur_score.begin(), cur_score.end(), 0.0);
}

// PhyloSuperTree::computePartitionOrder (tree/phylosupertree.cpp:653-695): most expensive partitions first
void PhyloSuperTreePlen::computePartitionOrder() {
    if (part_order.size() == parts.size()) return;
    part_order.resize(parts.size());
    std::iota(part_order.begin(), part_order.end(), 0);
    std::vector<double> cost(parts.size());
    for (size_t p = 0; p < parts.size(); p++) cost[p] = (double)parts[p]->leafNum * (double)parts[p]->getAlnNPattern() * parts[p]->getNumStates();
    std::stable_sort(part_order.begin(), part_order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
}

// the partition's two directed neighbours of the super tree's current branch: SuperNeighbor::link_neighbors[part]
void PhyloSuperTreePlen::linkBranch(int part, PhyloNeighbor *&nei1_part, PhyloNeighbor *&nei2_part) const {
    // nei1 = current_it (at node1, pointing to node2), nei2 = current_it_back (phylosupertreeplen.cpp:318-320)
    nei1_part = parts[part]->branchNeighbor(current_it->id, current_it->node->id);
    nei2_part = parts[part]->branchNeighbor(current_it->id, current_it_back->node->id);
    if (!nei1_part || !nei2_part) throw std::runtime_error("partition tree is not linked to the super tree's branch");
}

void PhyloSuperTreePlen::setPartLengths(double value) {
    current_it->length = value;
    current_it_back->length = value;
    for (size_t p = 0; p < parts.size(); p++) {
        PhyloNeighbor *n1, *n2;
        linkBranch((int)p, n1, n2);
        // the reference adds lambda * part_rate to the partition branch (:337-338); with linked lengths that is value * rate
        n1->length = n2->length = value * part_rate[p];
        parts[p]->current_it = n1;
        parts[p]->current_it_back = n2;
    }
}


double PhyloSuperTreePlen::computeFunction(double value) {
    computePartitionOrder();
    const bool grouped = ensureGroup();   // (may initialise the partitions' buffers: before the branch is linked)
    setPartLengths(value);
    if (grouped) {
        groupEvaluate(1, [&](int p, PhyloTree::EvalPlan &pl) { parts[p]->planLikelihoodBranch(parts[p]->current_it_back, parts[p]->current_it->node, pl); },
                      cur_score.data(), nullptr, nullptr);
        double lh = 0.0;
        for (size_t p = 0; p < parts.size(); p++) lh += cur_score[p];   // summed in partition order
        return -lh;
    }
    for_each_partition(part_order, num_threads, [&](int p) {
        parts[p]->enqueueLikelihoodBranch(parts[p]->current_it_back, parts[p]->current_it->node);   // (nei2_part, nei1_part->node), :339
    });
    double tree_lh = 0.0;
    for (size_t p = 0; p < parts.size(); p++) {   // summed in partition order: reproducible whatever the thread count
        parts[p]->finish(&cur_score[p], nullptr, nullptr);
        tree_lh += cur_score[p];
    }
    return -tree_lh;
}

double PhyloSuperTreePlen::computeLikelihoodBranch(PhyloNeighbor *dad_branch, PhyloNode *dad, bool) {
    current_it = dad_branch;
    current_it_back = dad_branch->node->findNeighbor(dad);
    return -computeFunction(dad_branch->length);
}

double PhyloSuperTreePlen::computeLikelihood() {
    if (!current_it) {
        PhyloNode *leaf = findFarthestLeaf();
        current_it = leaf->neighbors[0];
        current_it_back = current_it->node->findNeighbor(leaf);
    }
    curScore = computeLikelihoodBranch(current_it, current_it_back->node);
    return curScore;
}

double PhyloSuperTreePlen::computeLikelihoodFromBuffer() {
    double score = 0.0;
    for (double s : cur_score) score += s;
    return score;
}

void PhyloSuperTreePlen::computeFuncDerv(double value, double &df_ret, double &ddf_ret) {
    computePartitionOrder();
    const bool grouped = ensureGroup();
    setPartLengths(value);
    if (grouped) {
        std::vector<double> d1(parts.size(), 0.0), d2(parts.size(), 0.0);
        groupEvaluate(2, [&](int p, PhyloTree::EvalPlan &pl) { parts[p]->planLikelihoodDerv(parts[p]->current_it_back, parts[p]->current_it->node, pl); },
                      nullptr, d1.data(), d2.data());
        double gdf = 0.0, gddf = 0.0;
        for (size_t p = 0; p < parts.size(); p++) {
            gdf += part_rate[p] * d1[p];                     // :398-399
            gddf += part_rate[p] * part_rate[p] * d2[p];
        }
        df_ret = -gdf;
        ddf_ret = -gddf;
        return;
    }
    for_each_partition(part_order, num_threads, [&](int p) {
        parts[p]->enqueueLikelihoodDerv(parts[p]->current_it_back, parts[p]->current_it->node);   // :397
    });
    double df = 0.0, ddf = 0.0;
    for (size_t p = 0; p < parts.size(); p++) {
        double df_aux = 0.0, ddf_aux = 0.0;
        parts[p]->finish(nullptr, &df_aux, &ddf_aux);
        df += part_rate[p] * df_aux;                     // :398-399
        ddf += part_rate[p] * part_rate[p] * ddf_aux;
    }
    df_ret = -df;
    ddf_ret = -ddf;
}

void PhyloSuperTreePlen::optimizeOneBranch(PhyloNode *node1, PhyloNode *node2, bool clearLH, int maxNRStep) {
    current_it = node1->findNeighbor(node2);
    current_it_back = node2->findNeighbor(node1);
    const double current_len = current_it->length;
    for (auto *p : parts) {   // :271-273
        p->theta_computed = false;
        if (p->engine()) b200lk_theta_reset(p->engine());
    }
    PhyloTree::optimizeOneBranch(node1, node2, false, maxNRStep);   // :276
    // minimizeNewton may return the point BEFORE the last one it evaluated; bring the partition branches to the final length
    // (the reference leaves them at the last evaluated point, a drift below the Newton tolerance)
    const bool grouped = ensureGroup();
    setPartLengths(current_it->length);
    computePartitionOrder();
    if (grouped) {
        for (auto *p : parts) if (!p->theta_computed || !p->current_it) throw std::runtime_error("computeLikelihoodFromBuffer: theta_all is not computed");
        groupEvaluate(3, [](int, PhyloTree::EvalPlan &) {}, cur_score.data(), nullptr, nullptr);
    } else {
        for_each_partition(part_order, num_threads, [&](int p) { parts[p]->enqueueLikelihoodFromBuffer(); });   // :283-290
        for (size_t p = 0; p < parts.size(); p++) parts[p]->finish(&cur_score[p], nullptr, nullptr);
    }
    if (clearLH && current_len != current_it->length) {   // :292-302
        for (size_t p = 0; p < parts.size(); p++) {
            PhyloNeighbor *n1, *n2;
            linkBranch((int)p, n1, n2);
            n1->node->clearReversePartialLh(n2->node);
            n2->node->clearReversePartialLh(n1->node);
        }
    }
}

double PhyloSuperTreePlen::optimizeAllBranches(int my_iterations, double tolerance, int maxNRStep) {
    std::fill(cur_score.begin(), cur_score.end(), 0.0);   // :246-248
    return PhyloTree::optimizeAllBranches(my_iterations, tolerance, maxNRStep);
}

int64_t PhyloSuperTreePlen::launches() const {
    if (group_) return b200lk_group_launch_count(group_);
    int64_t n = 0;
    for (auto *p : parts) if (p->engine()) n += b200lk_launch_count(p->engine());
    return n;
}

// ---------------------------------------------------------------------------------------------------------------
// NNI for the linked model
// ---------------------------------------------------------------------------------------------------------------
static int posOfNodeId(const PhyloNode *node, int id) {
    for (size_t k = 0; k < node->neighbors.size(); k++)
        if (node->neighbors[k]->node->id == id) return (int)k;
    throw std::runtime_error("partition tree and super tree differ in topology");
}

NNIMove PhyloSuperTreePlen::getBestNNIForBran(PhyloNode *node1, PhyloNode *node2, NNIMove *nniMoves) {
    initializeAllPartialLh();
    const size_t np = parts.size();
    NNIContext ctx;
    std::vector<NNIContext> pctx(np);
    std::vector<PhyloNode *> p1(np), p2(np);
    NNIMove local[2];
    if (!nniMoves) nniMoves = local;
    if (!nniMoves[0].node1) defaultNNIMoves(node1, node2, nniMoves);
    nniMoves[0].node1 = nniMoves[1].node1 = node1;
    nniMoves[0].node2 = nniMoves[1].node2 = node2;
    nniMoves[0].newloglh = nniMoves[1].newloglh = -DBL_MAX;
    const double backupScore = curScore;
    const std::vector<double> backup_part_score = cur_score;
    // flag of the CLV that `from` holds of the subtree behind `to`, in the super tree and in every partition tree
    auto clearFlag = [&](PhyloNode *from, PhyloNode *to) {
        from->findNeighbor(to)->partial_lh_computed = false;
        for (size_t p = 0; p < np; p++) parts[p]->nodeById(from->id)->findNeighbor(parts[p]->nodeById(to->id))->partial_lh_computed = false;
    };
    auto endAll = [&]() {
        nniEnd(ctx);
        for (size_t p = 0; p < np; p++) parts[p]->nniEnd(pctx[p]);
    };
    try {
        nniBegin(node1, node2, params.nni5, ctx);
        for (size_t p = 0; p < np; p++) {
            p1[p] = parts[p]->nodeById(node1->id);
            p2[p] = parts[p]->nodeById(node2->id);
            parts[p]->nniBegin(p1[p], p2[p], params.nni5, pctx[p]);
        }
        for (int cnt = 0; cnt < 2; cnt++) {
            const int k1 = nniMoves[cnt].node1Nei_it, k2 = nniMoves[cnt].node2Nei_it;
            const int a_id = node1->neighbors[k1]->node->id, c_id = node2->neighbors[k2]->node->id;
            std::vector<int> pk1(np), pk2(np);
            for (size_t p = 0; p < np; p++) { pk1[p] = posOfNodeId(p1[p], a_id); pk2[p] = posOfNodeId(p2[p], c_id); }
            nniSwap(ctx, k1, k2);
            for (size_t p = 0; p < np; p++) parts[p]->nniSwap(pctx[p], pk1[p], pk2[p]);
            for (int step = 0; step < params.nni5_num_eval; step++) {
                clearFlag(node1, node2);
                clearFlag(node2, node1);
                int i = 1;
                if (params.nni5) {
                    for (int k = 0; k < 3; k++) {
                        PhyloNode *other = node1->neighbors[k]->node;
                        if (other == node2) continue;
                        clearFlag(other, node1);
                        optimizeOneBranch(node1, other, false, NNI_MAX_NR_STEP);
                        nniMoves[cnt].newLen[i++] = node1->findNeighbor(other)->length;
                    }
                    clearFlag(node2, node1);
                }
                optimizeOneBranch(node1, node2, false, NNI_MAX_NR_STEP);
                nniMoves[cnt].newLen[0] = node1->findNeighbor(node2)->length;
                if (params.nni5) {
                    for (int k = 0; k < 3; k++) {
                        PhyloNode *other = node2->neighbors[k]->node;
                        if (other == node1) continue;
                        clearFlag(other, node2);
                        optimizeOneBranch(node2, other, false, NNI_MAX_NR_STEP);
                        nniMoves[cnt].newLen[i++] = node2->findNeighbor(other)->length;
                    }
                    clearFlag(node1, node2);
                }
            }
            nniMoves[cnt].newloglh = computeLikelihoodFromBuffer();
            nniSwap(ctx, k1, k2);
            for (size_t p = 0; p < np; p++) parts[p]->nniSwap(pctx[p], pk1[p], pk2[p]);
        }
    } catch (...) {
        endAll();
        throw;
    }
    endAll();
    curScore = backupScore;
    cur_score = backup_part_score;
    for (auto *p : parts) {   // the partitions' theta buffers belong to the last alternative topology
        p->theta_computed = false;
        if (p->engine()) b200lk_theta_reset(p->engine());
    }
    return nniMoves[0].newloglh > nniMoves[1].newloglh ? nniMoves[0] : nniMoves[1];
}

void PhyloSuperTreePlen::doNNIAll(const NNIMove &move, bool clearLH, bool set_lengths) {
    const int a_id = move.node1->neighbors[move.node1Nei_it]->node->id, c_id = move.node2->neighbors[move.node2Nei_it]->node->id;
    for (size_t p = 0; p < parts.size(); p++) {
        NNIMove pm = move;
        pm.node1 = parts[p]->nodeById(move.node1->id);
        pm.node2 = parts[p]->nodeById(move.node2->id);
        pm.node1Nei_it = posOfNodeId(pm.node1, a_id);
        pm.node2Nei_it = posOfNodeId(pm.node2, c_id);
        parts[p]->doNNI(pm, clearLH);
        if (set_lengths) {
            for (int i = 0; i < 5; i++) pm.newLen[i] = move.newLen[i] * part_rate[p];
            parts[p]->params.nni5 = params.nni5;
            parts[p]->changeNNIBrans(pm);
        }
    }
    doNNI(move, clearLH);
    if (set_lengths) changeNNIBrans(move);
}

// ---------------------------------------------------------------------------------------------------------------
// PhyloSuperTree: edge-unlinked partitions (tree/phylosupertree.cpp:653-791)
// ---------------------------------------------------------------------------------------------------------------
PhyloSuperTree::PhyloSuperTree() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

int PhyloSuperTree::addPartition(PhyloTree *part) {
    if (!part) throw std::runtime_error("addPartition: null partition tree");
    parts.push_back(part);
    cur_score.push_back(0.0);
    part_order.clear();
    return (int)parts.size() - 1;
}

void PhyloSuperTree::computePartitionOrder() {
    if (part_order.size() == parts.size()) return;
    part_order.resize(parts.size());
    std::iota(part_order.begin(), part_order.end(), 0);
    std::vector<double> cost(parts.size());
    for (size_t p = 0; p < parts.size(); p++) cost[p] = (double)parts[p]->leafNum * (double)parts[p]->getAlnNPattern() * parts[p]->getNumStates();
    std::stable_sort(part_order.begin(), part_order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
}

int64_t PhyloSuperTree::numPatterns() const {
    int64_t n = 0;
    for (auto *p : parts) n += p->getAlnNPattern();
    return n;
}

double PhyloSuperTree::computeLikelihood(double *pattern_lh) {
    computePartitionOrder();
    std::vector<int64_t> off(parts.size() + 1, 0);
    for (size_t p = 0; p < parts.size(); p++) off[p + 1] = off[p] + parts[p]->getAlnNPattern();
    for_each_partition(part_order, num_threads, [&](int p) { cur_score[p] = parts[p]->computeLikelihood(pattern_lh ? pattern_lh + off[p] : nullptr); });
    double tree_lh = 0.0;
    for (double x : cur_score) tree_lh += x;   // in partition order: the same sum for any number of threads
    return tree_lh;
}

void PhyloSuperTree::computePatternLikelihood(double *pattern_lh, double *cur_logl) {
    // cur_logl is an input in the reference (a sanity check against sum_p freq * pattern_lh, :748-761); here the partitions'
    // scores of this very evaluation are handed back instead
    int64_t off = 0;
    double sum = 0.0;
    for (size_t p = 0; p < parts.size(); p++) {
        cur_score[p] = parts[p]->computeLikelihood(pattern_lh + off);
        sum += cur_score[p];
        off += parts[p]->getAlnNPattern();
    }
    if (cur_logl) *cur_logl = sum;
}

double PhyloSuperTree::optimizeAllBranches(int my_iterations, double tolerance, int maxNRStep) {
    computePartitionOrder();
    const double tol = tolerance / std::min<int>((int)parts.size(), 10);
    for_each_partition(part_order, num_threads, [&](int p) { cur_score[p] = parts[p]->optimizeAllBranches(my_iterations, tol, maxNRStep); });
    double tree_lh = 0.0;
    for (double x : cur_score) tree_lh += x;
    return tree_lh;
}

void PhyloSuperTree::clearAllPartialLH() {
    for (auto *p : parts) p->clearAllPartialLH();
}

}  // namespace b200
