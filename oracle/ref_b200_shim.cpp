// ref_b200_shim.cpp — TEST INFRASTRUCTURE: the reference-side binding of INTEGRATION.md, compiled for real.
//
// A PhyloTree of the UNMODIFIED reference (IQ-TREE 2.4.0, built by oracle/build_ref.sh from /root/reference) whose five
// kernel pointers (tree/phylotree.h:900-901, 947-948, 1000-1001, 1342-1346) point at members that marshal into the product's
// C ABI (include/b200lk.h, iqtree2_b200/libb200lk.so).  Everything above the kernels — computeTraversalInfo, reorientPartialLh,
// optimizeOneBranch / minimizeNewton, optimizeAllBranches, getBestNNIForBran, optimizeTreeLengthScaling, computePatternLhCat — is
// the reference's own code; only the arithmetic runs on the GPU.  PhyloNeighbor::partial_lh stays what the reference makes it:
// a pointer into central_partial_lh / nni_partial_lh; here it is only ever converted into a slot number.
//
// The reference's class cannot be given new members without editing its header, so the shim members live in a derived class
// and are stored in the base's pointers-to-member by static_cast (valid: every object they are called on is a PhyloTreeB200).
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "alignment/alignment.h"
#include "model/modelfactory.h"
#include "model/modelsubst.h"
#include "model/rateheterogeneity.h"
#include "tree/phylotree.h"
#include "tree/phylonodemixlen.h"
#include "utils/tools.h"

#include "b200lk.h"

extern PhyloTree *(*iqref_tree_factory)();   // ref_harness.cpp: how iqref_create makes its tree
extern int iqref_next_mixlen_classes;         // ref_harness.cpp: > 0 = the tree being created has mixed branch lengths
extern void (*iqref_derv_mixlen_cur_hook)(PhyloTree *, PhyloNeighbor *, PhyloNode *, int, double *, double *);
// ref_harness.cpp: the complex eigensystem and rate matrices of the harness's non-reversible model (in a full IQ-TREE build these
// are ModelMarkov's members ceval / cevec / cinv_evec / rate_matrix, model/modelmarkov.h); false for a reversible model
extern bool iqref_nonrev_arrays(ModelSubst *model, const double **ceval, const double **cevec, const double **cinv, const double **qmat);

namespace {

class PhyloTreeB200 : public PhyloTree {
public:
    b200lk_engine *b200 = nullptr;
    std::vector<b200lk_op> b200_ops;
    std::vector<b200lk_child> b200_extra;
    long long kernel_calls = 0;
    bool fused_newton = false;   // optimizeOneBranch through b200lk_newton_branch (the optional hook of INTEGRATION.md)
    // mixed branch lengths: in a full build this class derives from PhyloTreeMixlen; here it answers the three questions the kernels
    // ask (tree/phylotreemixlen.h:89-98, 218) and makes PhyloNodeMixlen nodes (tree/phylotreemixlen.cpp:95-101)
    int mixlen_classes = 0, cur_mixture = 0;
    bool isMixlen() override { return mixlen_classes > 0; }
    int getMixlen() override { return mixlen_classes > 0 ? mixlen_classes : 1; }
    int getCurMixture() override { return cur_mixture; }
    Node *newNode(int node_id = -1, const char *node_name = NULL) override {
        return mixlen_classes ? (Node *)new PhyloNodeMixlen(node_id, node_name) : PhyloTree::newNode(node_id, node_name);
    }
    Node *newNode(int node_id, int node_name) override {
        return mixlen_classes ? (Node *)new PhyloNodeMixlen(node_id, node_name) : PhyloTree::newNode(node_id, node_name);
    }

    ~PhyloTreeB200() override { if (b200) b200lk_destroy(b200); }

    void check(int rc, const char *what) {
        if (rc < 0) outError(std::string(what) + ": " + b200lk_last_error(b200));
    }

    size_t numCat() { return model_factory->fused_mix_rate ? site_rate->getNRate() : site_rate->getNRate() * model->getNMixtures(); }

    // slot = (partial_lh - central_partial_lh) / block; nni_partial_lh -> the two spare slots T-2, T-1
    int slotOf(PhyloNeighbor *nei) {
        const size_t block = getPartialLhSize();
        double *p = nei->partial_lh;
        ASSERT(p);
        if (p >= nni_partial_lh && p < nni_partial_lh + 2 * block) return (int)(leafNum - 2 + (p - nni_partial_lh) / block);
        return (int)((p - central_partial_lh) / block);
    }
    b200lk_child childOf(PhyloNeighbor *nei) {
        b200lk_child c;
        c.kind = nei->node->isLeaf() ? 0 : 1;
        c.id = nei->node->isLeaf() ? nei->node->id : slotOf(nei);
        c.length = isMixlen() ? (double)nei->id : nei->length;   // mixlen: the row of the branch in the table of uploadMixlen()
        if (!model->useRevKernel()) {   // computePartialInfo's non-reversible branch (tree/phylokernelnew.h:932-1008)
            if (isRootLeaf(nei->node)) { c.kind = B200LK_CHILD_ROOT; c.id = 0; }
            if (nei->direction == TOWARD_ROOT) c.kind |= B200LK_CHILD_TRANSPOSED;
        }
        return c;
    }
    int dadKind(PhyloNode *dad) { return isRootLeaf(dad) && !model->useRevKernel() ? B200LK_CHILD_ROOT : (dad->isLeaf() ? 0 : 1); }
    int dadId(PhyloNode *dad, PhyloNeighbor *node_branch) { return isRootLeaf(dad) && !model->useRevKernel() ? 0 : (dad->isLeaf() ? dad->id : slotOf(node_branch)); }

    void uploadModel() {
        const int n = aln->num_states, nmix = model->getNMixtures();
        std::vector<double> w(nmix, 1.0);
        if (model->isMixture()) for (int m = 0; m < nmix; m++) w[m] = model->getMixtureWeight(m);
        check(b200lk_set_model(b200, model->getEigenvalues(), model->getEigenvectors(), model->getInverseEigenvectors(), w.data()), "set_model");
        const int ncat = site_rate->getNRate();
        std::vector<double> rate(ncat), prop(ncat);
        for (int c = 0; c < ncat; c++) { rate[c] = site_rate->getRate(c); prop[c] = site_rate->getProp(c); }
        check(b200lk_set_rates(b200, rate.data(), prop.data()), "set_rates");
        if (!model->useRevKernel()) {   // the kernels of tree/phylokernelnonrev.h
            const double *ceval, *cevec, *cinv, *qmat;
            if (!iqref_nonrev_arrays(model, &ceval, &cevec, &cinv, &qmat)) outError("non-reversible kernel: the model has no complex eigensystem");
            std::vector<double> freq((size_t)nmix * n);
            for (int m = 0; m < nmix; m++) model->getStateFrequency(&freq[(size_t)m * n], m);
            check(b200lk_set_model_nonrev(b200, ceval, cevec, cinv, qmat, freq.data(), w.data(), rooted ? 1 : 0), "set_model_nonrev");
        }
        if (model->isSiteSpecificModel()) {   // ModelSet: one eigensystem per pattern, V-interleaved where the kernels read them
            const size_t nobs = aln->size(), V = vector_size, ns = (size_t)n, nn = ns * ns;   // (tree/phylotreesse.cpp:271-281)
            const double *se = model->getEigenvalues(), *su = model->getEigenvectors(), *si = model->getInverseEigenvectors();
            std::vector<double> ev(nobs * ns), U(nobs * nn), Ui(nobs * nn);
            for (size_t p = 0; p < nobs; p++) {
                const size_t base = p / V, v = p % V;
                for (size_t i = 0; i < ns; i++) ev[p * ns + i] = se[base * ns * V + i * V + v];
                for (size_t k = 0; k < nn; k++) { U[p * nn + k] = su[base * nn * V + k * V + v]; Ui[p * nn + k] = si[base * nn * V + k * V + v]; }
            }
            check(b200lk_set_site_model(b200, ev.data(), U.data(), Ui.data()), "set_site_model");
        }
        // ptn_invar: observed patterns, then the unobserved constant patterns of +ASC, which the reference keeps behind the
        // padded observed block (tree/phylotreesse.cpp:585-611)
        const size_t nobs = aln->size(), U = model_factory->unobserved_ptns.size(), max_orig = get_safe_upper_limit(nobs);
        std::vector<double> inv(nobs + U);
        for (size_t p = 0; p < nobs; p++) inv[p] = ptn_invar[p];
        for (size_t u = 0; u < U; u++) inv[nobs + u] = ptn_invar[max_orig + u];
        check(b200lk_set_pattern_invar(b200, inv.data()), "set_pattern_invar");
        (void)n;
    }

    // mixed branch lengths: row Neighbor::id of the table = that branch's lengths (what saveBranchLengths writes, tree/phylotree.cpp:728-731)
    void collectMixlen(Node *node, Node *dad, std::vector<double> &tab) {
        FOR_NEIGHBOR_IT(node, dad, it) {
            const int id = (*it)->id;
            ASSERT(id >= 0 && id < (int)branchNum);
            for (int c = 0; c < mixlen_classes; c++) tab[(size_t)id * mixlen_classes + c] = (*it)->getLength(c);
            collectMixlen((*it)->node, node, tab);
        }
    }
    void uploadMixlen() {
        if (!isMixlen()) return;
        std::vector<double> tab((size_t)branchNum * mixlen_classes, 0.0);
        collectMixlen(root, nullptr, tab);
        check(b200lk_set_mixlen(b200, tab.data(), (int)branchNum), "set_mixlen");
    }

    // after initializeAllPartialLh: create the engine, hand over the inputs, take over the kernel pointers
    void attach() {
        ASSERT(params->lh_mem_save == LM_PER_NODE);
        const int nseq = (int)aln->getNSeq();   // (a rooted tree has one leaf more: the root leaf, without a sequence)
        computeTipPartialLikelihood();   // also ptn_freq / ptn_invar (tree/phylotreesse.cpp:243-375, 537-612)
        b200lk_config cfg{};
        cfg.nstates = aln->num_states; cfg.ncat = site_rate->getNRate(); cfg.nmix = model->getNMixtures();
        cfg.fused_mix_rate = model_factory->fused_mix_rate; cfg.safe_numeric = safe_numeric;
        cfg.seq_type = aln->seq_type == SEQ_DNA ? 1 : aln->seq_type == SEQ_PROTEIN ? 2 : 0;
        const size_t U = model_factory->unobserved_ptns.size();   // +ASC: appended as the engine's last patterns
        cfg.state_unknown = aln->STATE_UNKNOWN; cfg.ntaxa = nseq; cfg.nptn = (int64_t)(aln->size() + U);
        cfg.nslots = (leafNum - 2) + 2; cfg.device = 0;
        // Params::num_threads' counterpart for devices: B200LK_NDEV=N puts the tree on the first N GPUs (cfg.device = -N: one
        // engine per device over pattern slices behind the same handle, include/b200lk.h)
        if (const char *nd = getenv("B200LK_NDEV")) if (atoi(nd) > 1 && U == 0) cfg.device = -atoi(nd);
        if (b200lk_create(&cfg, &b200) != 0) outError(std::string("b200lk_create: ") + b200lk_last_error(nullptr));
        const size_t nobs = aln->size(), nptn = nobs + U;
        std::vector<uint8_t> states((size_t)nseq * nptn);
        for (int t = 0; t < nseq; t++) {
            for (size_t p = 0; p < nobs; p++) states[(size_t)t * nptn + p] = (uint8_t)aln->at(p)[t];
            for (size_t u = 0; u < U; u++) states[(size_t)t * nptn + nobs + u] = (uint8_t)model_factory->unobserved_ptns[u][t];
        }
        check(b200lk_set_tip_states(b200, states.data()), "set_tip_states");
        std::vector<double> freq(nptn, 0.0);
        for (size_t p = 0; p < nobs; p++) freq[p] = ptn_freq[p];
        check(b200lk_set_pattern_freq(b200, freq.data()), "set_pattern_freq");
        uploadModel();
        if (U) {
            const ASCType asc = model_factory->getASC();
            check(b200lk_set_ascertainment(b200, (asc == ASC_VARIANT_MISSING || asc == ASC_INFORMATIVE_MISSING) ? 2 : 1, (int64_t)U,
                                           (int64_t)aln->getNSite()), "set_ascertainment");
        }
        // setLikelihoodKernelB200 of INTEGRATION.md
        computeLikelihoodBranchPointer = static_cast<ComputeLikelihoodBranchType>(&PhyloTreeB200::computeLikelihoodBranchB200);
        computeLikelihoodDervPointer = static_cast<ComputeLikelihoodDervType>(&PhyloTreeB200::computeLikelihoodDervB200);
        computeLikelihoodFromBufferPointer = static_cast<ComputeLikelihoodFromBufferType>(&PhyloTreeB200::computeLikelihoodFromBufferB200);
        if (!model->useRevKernel()) computeLikelihoodFromBufferPointer = NULL;   // as setLikelihoodKernel leaves it for the non-reversible kernel
        computePartialLikelihoodPointer = static_cast<ComputePartialLikelihoodType>(&PhyloTreeB200::computePartialLikelihoodB200);
        computeLikelihoodDervMixlenPointer = isMixlen() ? static_cast<ComputeLikelihoodDervMixlenType>(&PhyloTreeB200::computeLikelihoodDervMixlenB200) : NULL;
    }

    // traversal_info (filled by the reference's own computeTraversalInfo, tree/phylotree.cpp:5918-6028) -> b200lk_op[]
    void marshalTraversalInfo() {
        b200_ops.clear();
        b200_extra.clear();
        for (auto &info : traversal_info) {
            PhyloNode *node = (PhyloNode *)info.dad_branch->node;
            b200lk_op op{};
            op.dst_slot = slotOf(info.dad_branch);
            int k = 0;
            FOR_NEIGHBOR_IT(node, info.dad, it) {
                b200lk_child c = childOf((PhyloNeighbor *)*it);
                if (k == 0) op.left = c;
                else if (k == 1) op.right = c;
                else { b200_extra.push_back(c); op.reserved++; }
                k++;
            }
            b200_ops.push_back(op);
        }
        traversal_info.clear();
        if (!b200_extra.empty()) check(b200lk_set_extra_children(b200, b200_extra.data(), (int)b200_extra.size()), "set_extra_children");
    }

    // the head of every kernel of tree/phylokernelnew.h: orient the branch, list the stale partials
    void planBranch(PhyloNeighbor *&dad_branch, PhyloNode *&dad, PhyloNeighbor *&node_branch) {
        PhyloNode *node = (PhyloNode *)dad_branch->node;
        node_branch = (PhyloNeighbor *)node->findNeighbor(dad);
        const bool swap = model->useRevKernel() ? node->isLeaf()   // :2654-2661
                                                : (node->isLeaf() || (dad_branch->direction == AWAYFROM_ROOT && !isRootLeaf(dad)));   // phylokernelnonrev.h:596-603
        if (swap) {
            std::swap(dad, node);
            std::swap(dad_branch, node_branch);
        }
        if ((tip_partial_lh_computed & 1) == 0) {   // the model changed (clearAllPartialLH): :1151-1152
            computeTipPartialLikelihood();
            uploadModel();
        }
        uploadMixlen();   // (a production binding re-sends the table only after a length changed)
        traversal_info.clear();
        double *buffer = buffer_partial_lh;   // echildren / partial_lh_leaves are carved from it but never touched here
        computeTraversalInfo(dad_branch, dad, buffer);
        computeTraversalInfo(node_branch, node, buffer);
        marshalTraversalInfo();
        kernel_calls++;
    }

    void copyPatternLh() {
        std::vector<double> plh(aln->size() + model_factory->unobserved_ptns.size());
        check(b200lk_get_pattern_lh(b200, plh.data()), "get_pattern_lh");
        memcpy(_pattern_lh, plh.data(), sizeof(double) * aln->size());
    }

    double computeLikelihoodBranchB200(PhyloNeighbor *dad_branch, PhyloNode *dad, bool save_log_value) {
        PhyloNeighbor *node_branch;
        const double length = isMixlen() ? (double)dad_branch->id : dad_branch->length;
        planBranch(dad_branch, dad, node_branch);
        double logl = 0.0;
        int rc = b200lk_branch_logl(b200, b200_ops.data(), (int)b200_ops.size(), slotOf(dad_branch), dadKind(dad),
                                    dadId(dad, node_branch), length, save_log_value, &logl);
        if (rc == B200LK_WARN_UNDERFLOW) outWarning("Numerical underflow for lh-branch");
        check(rc, "branch_logl");
        // callers read _pattern_lh / _pattern_lh_cat on the host (computePatternLikelihood, computePatternLhCat): tests are small,
        // so they are always brought back here; a production shim does it on demand
        copyPatternLh();
        if (_pattern_lh_cat && model_factory->unobserved_ptns.empty()) {
            // transformPatternLhCat (tree/phylotree.cpp:1313-1336) expects the kernel's V-interleaved layout
            const size_t nptn = aln->size(), C = numCat(), V = vector_size;
            std::vector<double> cat(nptn * C);
            check(b200lk_get_pattern_lh_cat(b200, cat.data()), "get_pattern_lh_cat");
            for (size_t p = 0; p < nptn; p++)
                for (size_t c = 0; c < C; c++) _pattern_lh_cat[(p / V) * C * V + c * V + p % V] = cat[p * C + c];
        }
        return logl;
    }

    void computeLikelihoodDervB200(PhyloNeighbor *dad_branch, PhyloNode *dad, double *df, double *ddf) {
        PhyloNeighbor *node_branch;
        const double length = isMixlen() ? (double)dad_branch->id : dad_branch->length;
        planBranch(dad_branch, dad, node_branch);
        if (!b200_ops.empty()) theta_computed = false;
        if (!theta_computed) check(b200lk_theta_reset(b200), "theta_reset");   // :2361-2389: the buffer is rebuilt
        if (isMixlen()) {   // df [K + 1] = gradient, logL; ddf [K][K] = Hessian (:2391-2460, 2556-2568)
            check(b200lk_branch_derv_mixlen(b200, b200_ops.data(), (int)b200_ops.size(), slotOf(dad_branch), dadKind(dad),
                                            dadId(dad, node_branch), length, df, ddf), "branch_derv_mixlen");
            theta_computed = true;
            return;
        }
        int rc = b200lk_branch_derv(b200, b200_ops.data(), (int)b200_ops.size(), slotOf(dad_branch), dadKind(dad),
                                    dadId(dad, node_branch), length, df, ddf);
        check(rc, "branch_derv");
        theta_computed = true;
    }

    // computeLikelihoodDervMixlen (:3431-3616): the current class's length alone
    void computeLikelihoodDervMixlenB200(PhyloNeighbor *dad_branch, PhyloNode *dad, double &df, double &ddf) {
        PhyloNeighbor *node_branch;
        const double length = (double)dad_branch->id;
        planBranch(dad_branch, dad, node_branch);
        if (!b200_ops.empty()) theta_computed = false;
        if (!theta_computed) check(b200lk_theta_reset(b200), "theta_reset");
        check(b200lk_branch_derv_mixlen_cur(b200, b200_ops.data(), (int)b200_ops.size(), slotOf(dad_branch), dadKind(dad),
                                            dadId(dad, node_branch), length, getCurMixture(), &df, &ddf), "branch_derv_mixlen_cur");
        theta_computed = true;
    }

    double computeLikelihoodFromBufferB200() {
        double logl = 0.0;
        if (isMixlen()) uploadMixlen();
        check(b200lk_logl_from_buffer(b200, isMixlen() ? (double)current_it->id : current_it->length, &logl), "logl_from_buffer");
        copyPatternLh();
        return logl;
    }

    // Optional hook: PhyloTree::optimizeOneBranch (tree/phylotree.cpp:2594-2646) with the minimizeNewton loop replaced by ONE C call
    // (the whole Newton-Raphson iteration runs on the device on the resident theta buffer); everything else as in the reference.
    void optimizeOneBranch(PhyloNode *node1, PhyloNode *node2, bool clearLH = true, int maxNRStep = 100) override {
        if (!b200 || !fused_newton || !optimize_by_newton) { PhyloTree::optimizeOneBranch(node1, node2, clearLH, maxNRStep); return; }
        if (rooted && (node1 == root || node2 == root)) return;
        current_it = (PhyloNeighbor *)node1->findNeighbor(node2);
        current_it_back = (PhyloNeighbor *)node2->findNeighbor(node1);
        const double current_len = current_it->length;
        theta_computed = false;
        check(b200lk_theta_reset(b200), "theta_reset");
        PhyloNeighbor *dad_branch = current_it, *node_branch;
        PhyloNode *dad = (PhyloNode *)current_it_back->node;
        planBranch(dad_branch, dad, node_branch);
        double optx = current_len;
        int steps = 0;
        check(b200lk_newton_branch(b200, b200_ops.data(), (int)b200_ops.size(), slotOf(dad_branch), dadKind(dad),
                                   dadId(dad, node_branch), params->min_branch_length, current_len,
                                   params->max_branch_length, params->min_branch_length, maxNRStep, &optx, &steps), "newton_branch");
        theta_computed = true;
        if (optx > params->max_branch_length * 0.95) {   // Newton-Raphson diverged: keep the better of the two (:2616-2625)
            current_it->length = current_it_back->length = optx;
            const double opt_lh = computeLikelihoodFromBuffer();
            current_it->length = current_it_back->length = current_len;
            if (computeLikelihoodFromBuffer() > opt_lh) optx = current_len;
        }
        current_it->length = optx;
        current_it_back->length = optx;
        if (clearLH && current_len != optx) {
            node1->clearReversePartialLh(node2);
            node2->clearReversePartialLh(node1);
        }
    }

    // only reached through computeTraversalInfo<VC>(..., compute_partial_lh = true) of callers outside the kernels: one op,
    // once (the pattern range and the packet are the CPU kernel's work split)
    void computePartialLikelihoodB200(TraversalInfo &info, size_t, size_t, int packet_id) {
        if (packet_id > 0) return;
        PhyloNode *node = (PhyloNode *)info.dad_branch->node;
        b200lk_op op{};
        op.dst_slot = slotOf(info.dad_branch);
        int k = 0;
        FOR_NEIGHBOR_IT(node, info.dad, it) {
            if (k == 0) op.left = childOf((PhyloNeighbor *)*it);
            else if (k == 1) op.right = childOf((PhyloNeighbor *)*it);
            k++;
        }
        ASSERT(k == 2);
        check(b200lk_update_partials(b200, &op, 1), "update_partials");
    }
};

PhyloTree *make_b200_tree() {
    PhyloTreeB200 *t = new PhyloTreeB200();
    t->mixlen_classes = iqref_next_mixlen_classes;
    return t;
}

void derv_mixlen_cur(PhyloTree *tree, PhyloNeighbor *nb, PhyloNode *dad, int cur, double *df, double *ddf) {
    PhyloTreeB200 *t = dynamic_cast<PhyloTreeB200 *>(tree);
    if (!t || !t->computeLikelihoodDervMixlenPointer) { *df = *ddf = 0.0 / 0.0; return; }
    t->cur_mixture = cur;
    (t->*(t->computeLikelihoodDervMixlenPointer))(nb, dad, *df, *ddf);   // PhyloTreeMixlen::computeFuncDerv's call
}

}  // namespace

extern "C" {

// trees created by iqref_create from now on are PhyloTreeB200 objects (still running the CPU kernel until attached)
void iqref_b200_enable(int on) {
    iqref_tree_factory = on ? make_b200_tree : nullptr;
    iqref_derv_mixlen_cur_hook = on ? derv_mixlen_cur : iqref_derv_mixlen_cur_hook;
}

// switch the tree of this handle to the B200 kernel; returns 0, or -1 if it was not created as a PhyloTreeB200
int iqref_b200_attach(PhyloTree *tree) {
    PhyloTreeB200 *t = dynamic_cast<PhyloTreeB200 *>(tree);
    if (!t) return -1;
    if (!t->b200) t->attach();
    return 0;
}

int iqref_b200_set_fused_newton(PhyloTree *tree, int on) {
    PhyloTreeB200 *t = dynamic_cast<PhyloTreeB200 *>(tree);
    if (!t) return -1;
    t->fused_newton = on != 0;
    return 0;
}

long long iqref_b200_kernel_calls(PhyloTree *tree) {
    PhyloTreeB200 *t = dynamic_cast<PhyloTreeB200 *>(tree);
    return t ? t->kernel_calls : -1;
}

long long iqref_b200_launches(PhyloTree *tree) {
    PhyloTreeB200 *t = dynamic_cast<PhyloTreeB200 *>(tree);
    return t && t->b200 ? b200lk_launch_count(t->b200) : -1;
}

}
