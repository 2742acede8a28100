// phylotree_b200.h — host-side mirror of the reference's PhyloTree likelihood interface, driving the C ABI
// (include/b200lk.h).  Same member names, argument meaning and invalidation semantics as tree/phylotree.h:899-1048,
// tree/phylonode.h:28-185 so that the shim in INTEGRATION.md is a field-for-field mapping:
//
//   reference                                            here
//   PhyloNeighbor::partial_lh / scale_num (views)   ->   PhyloNeighbor::slot (device CLV slot handle, -1 = none)
//   PhyloNeighbor::partial_lh_computed              ->   same flag, same clearing rules
//   PhyloTree::traversal_info (vector<TraversalInfo>) ->  std::vector<b200lk_op> handed to the engine per call
//   central_partial_lh / nni_partial_lh              ->   engine slots [0, T-2) + 2 NNI spares, resident in HBM
//   theta_all / theta_computed                       ->   engine theta buffer + the same flag
//
// The tree, branch lengths and flags live on the host; CLVs never leave the GPU (b200lk_get_partial serves the few
// readers outside the kernels).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "b200lk.h"

namespace b200 {

const double TOL_BRANCH_LEN = 0.000001;   // tree/phylotree.h:62
const double TOL_LIKELIHOOD = 0.001;      // tree/phylotree.h:63

struct PhyloNode;

// directed branch dad -> node: carries the partial likelihood of the subtree rooted at `node` seen from dad
struct PhyloNeighbor {
    PhyloNode *node = nullptr;
    double length = 0.0;
    int id = -1;                      // branch id
    int slot = -1;                    // device CLV slot (the reference's partial_lh pointer)
    bool partial_lh_computed = false;
    int size = 0;                     // subtree size (PhyloNeighbor::size), 0 = unknown
};

struct PhyloNode {
    int id = -1;                      // leaves: alignment row (tree/phylokernelnew.h:1691-1711); internal: >= ntaxa
    std::string name;
    std::vector<PhyloNeighbor *> neighbors;
    double height = 0.0;
    bool isLeaf() const { return neighbors.size() <= 1; }
    PhyloNeighbor *findNeighbor(const PhyloNode *other) const;
    void clearReversePartialLh(PhyloNode *dad);   // tree/phylonode.cpp:22-34
};

enum LhMemSave { LM_PER_NODE = 0, LM_MEM_SAVE = 1 };   // utils/tools.h (LikelihoodMemSave; -mem option)

struct Params {
    LhMemSave lh_mem_save = LM_PER_NODE;   // -mem: LM_MEM_SAVE recycles max_lh_slots CLV slots (tree/phylotree.cpp:999-1021)
    double max_mem_size = 0.0;             // LM_MEM_SAVE: 0 = the minimum log2(T)+1 slots, <= 1 = fraction of the T-2 slots,
                                           // > 1 = bytes of device memory for CLV slots (incl. the 2 NNI spares)
    double min_branch_length = 1e-6;   // utils/tools.cpp:1235 (reset to 0.1/nsites-style values by the caller if wanted)
    double max_branch_length = 10.0;   // utils/tools.cpp:1238
    int device = 0;
    bool safe_numeric = false;         // -safe; forced for ntaxa >= 2000 or nstates not in {4,20} (phylotreesse.cpp:95-96)
    bool force_norm = false;           // test hook: keep NORM_LH where the reference would switch to SAFE_LH
    int schedule = 0;
    bool fused_newton = true;          // optimizeOneBranch through b200lk_newton_branch (one C call per branch)
    bool pipelined_sweep = false;      // optimizeAllBranches queues the whole sweep (b200lk_enqueue_newton_branch): branch lengths
                                       // stay on the device between branches, one synchronisation per sweep
    bool plan_only = false;            // no engine: traversals are planned (flags, slots, op lists) but not executed
    double robust_phy_keep = 1.0;      // --robust-phy: computeLikelihoodBranch sums the best keep*nsite sites (utils/tools.cpp:4379)
    bool robust_median = false;        // --robust-median: ... returns the median site log-likelihood (:4392)
    bool nni5 = true;                  // utils/tools.cpp:1155: NNI evaluation re-optimises the five branches around the move
    int nni5_num_eval = 1;             // utils/tools.cpp:1156
};

const int NNI_MAX_NR_STEP = 10;        // utils/pllnni.cpp:42

// tree/phylotree.h:169-191.  The reference keeps NeighborVec iterators to the two swapped neighbours; here their positions
// in node1->neighbors / node2->neighbors.
struct NNIMove {
    PhyloNode *node1 = nullptr, *node2 = nullptr;
    int node1Nei_it = -1, node2Nei_it = -1;
    double newloglh = 0.0;
    double newLen[5] = {0.0, 0.0, 0.0, 0.0, 0.0};   // [0] central branch; nni5: [1..2] node1's other branches, [3..4] node2's
    double *ptnlh = nullptr;                        // if set: pattern log-likelihoods of the swapped topology (nptn doubles)
    bool operator<(const NNIMove &rhs) const { return newloglh > rhs.newloglh; }
};

// tree/memslot.h, tree/memslot.cpp: the slot table of the memory-saving mode.  Same operations, same eviction rule
// (the unlocked slot whose neighbour has the smallest subtree); a slot is a device CLV slot index instead of a pointer pair.
struct MemSlot {
    int status = 0;
    PhyloNeighbor *nei = nullptr;
};
class MemSlotVector : public std::vector<MemSlot> {
public:
    bool enabled = false;
    void init(int num_slot);                                   // memslot.cpp:28-40
    void reset();                                              // :42-51
    bool lock(PhyloNeighbor *nei);                             // :92-104
    void unlock(PhyloNeighbor *nei);                           // :106-117
    bool locked(PhyloNeighbor *nei);                           // :119-135
    int allocate(PhyloNeighbor *nei);                          // :137-176
    void update(PhyloNeighbor *nei);                           // :178-192
    void takeover(PhyloNeighbor *nei, PhyloNeighbor *taken);   // :211-229
    int findNei(PhyloNeighbor *nei) const;                     // :54-59
private:
    void addNei(PhyloNeighbor *nei, int id);                   // :61-67
    std::map<PhyloNeighbor *, int> nei_id_map;
    int free_count = 0;
};

class PhyloTree {
public:
    PhyloTree();
    virtual ~PhyloTree();
    PhyloTree(const PhyloTree &) = delete;

    // ---- inputs (what aln / model / site_rate supply in the reference) ----------------------------------------
    // taxa names in alignment-row order + pattern-compressed tip states [ntaxa][nptn] + ptn_freq
    void setAlignment(const std::vector<std::string> &names, const uint8_t *states, int64_t nptn, const double *ptn_freq,
                      int nstates, int seq_type, int state_unknown);
    void setPatternInvar(const double *ptn_invar);
    void setAscertainment(int asc_type, int64_t num_unobserved, int64_t nsite);   // ModelFactory::ASC_type / unobserved_ptns (+ASC, Lewis)
    void setModel(int ncat, int nmix, bool fused_mix_rate, const double *eval, const double *evec, const double *inv_evec,
                  const double *mix_weight, const double *cat_rate, const double *cat_prop);
    void readTreeString(const std::string &newick);   // MTree::readTree; leaf ids := alignment rows
    std::string getTreeString(int precision = 10) const;

    // ---- buffers ------------------------------------------------------------------------------------------------
    void initializeAllPartialLh();    // tree/phylotree.cpp:893-939, 1060-1188 (LM_PER_NODE slot assignment)
    void deleteAllPartialLh();
    virtual void clearAllPartialLH();   // tree/phylotree.cpp:668-677

    // ---- the likelihood API (tree/phylotree.h:899-1048, 1136, 1340) ----------------------------------------------
    double computeLikelihood(double *pattern_lh = nullptr, bool save_log_value = true);   // tree/phylotree.cpp:1233-1274
    virtual double computeLikelihoodBranch(PhyloNeighbor *dad_branch, PhyloNode *dad, bool save_log_value = true);
    void computeLikelihoodDerv(PhyloNeighbor *dad_branch, PhyloNode *dad, double *df, double *ddf);
    virtual double computeLikelihoodFromBuffer();
    // asynchronous forms for partition models (b200lk_enqueue_* / b200lk_finish): same planning and flag updates as the
    // synchronous members, the device work is only queued on the engine's stream; finish() waits and returns the scalars
    // the same three steps for a GROUPED evaluation (b200lk_group_*: all partition trees of a super tree in O(1) launches): plan*
    // fills traversal_info and says what to evaluate; the caller hands the plans of all members to ONE b200lk_group_enqueue_* call
    // and then reports the outcome with planDone / planFailed (the roll-back of checkPlanned)
    struct EvalPlan { const b200lk_op *ops; int32_t nops, node_slot, dad_kind, dad_id; double length; };
    void planLikelihoodBranch(PhyloNeighbor *dad_branch, PhyloNode *dad, EvalPlan &out);
    void planLikelihoodDerv(PhyloNeighbor *dad_branch, PhyloNode *dad, EvalPlan &out);
    void planDone(bool derv);
    void planFailed();
    void enqueueLikelihoodBranch(PhyloNeighbor *dad_branch, PhyloNode *dad, bool save_log_value = true);
    void enqueueLikelihoodDerv(PhyloNeighbor *dad_branch, PhyloNode *dad);
    void enqueueLikelihoodFromBuffer();
    void finish(double *logl, double *df, double *ddf);
    void computePartialLikelihood(PhyloNeighbor *dad_branch, PhyloNode *dad);   // flush the stale partials below one directed branch
    void planBranch(PhyloNeighbor *dad_branch, PhyloNode *dad, std::vector<b200lk_op> &ops, int &node_slot, int &dad_kind,
                    int &dad_id, double &length);
    void computePatternLikelihood(double *pattern_lh, double *cur_logl = nullptr, double *pattern_lh_cat = nullptr);
    void computePatternLhCat(double *pattern_lh_cat);

    // ---- branch-length optimisation (tree/phylotree.cpp:2514-2774, utils/optimization.cpp:422-499) ---------------
    virtual double computeFunction(double value);
    virtual void computeFuncDerv(double value, double &df, double &ddf);
    double minimizeNewton(double x1, double xguess, double x2, double xacc, double &d2l, int maxNRStep = 100);
    virtual void optimizeOneBranch(PhyloNode *node1, PhyloNode *node2, bool clearLH = true, int maxNRStep = 100);
    virtual double optimizeAllBranches(int my_iterations = 100, double tolerance = TOL_LIKELIHOOD, int maxNRStep = 100);

    // ---- tree-length scaling (tree/phylotree.cpp:2455-2528; Brent's method utils/optimization.cpp:185-373) --------------
    void scaleLength(double norm);                                                                  // tree/mtree.cpp:1835-1846
    double minimizeOneDimen(double xmin, double xguess, double xmax, double tolerance, double *fx, double *f2x);
    double optimizeTreeLengthScaling(double min_scaling, double &scaling, double max_scaling, double gradient_epsilon);

    // ---- NNI evaluation on the resident CLVs (tree/phylotree.cpp:3975-4322) ------------------------------------------
    // Evaluates the two NNIs around the inner branch (node1, node2) on the engine's two spare slots (nni_partial_lh): the
    // CLVs, flags and lengths of the current tree are untouched on return.  nniMoves: optional array of two, filled.
    virtual NNIMove getBestNNIForBran(PhyloNode *node1, PhyloNode *node2, NNIMove *nniMoves = nullptr);
    // the pieces of it, for drivers that evaluate the same move in several trees at once (PhyloSuperTreePlen): nniBegin replaces the
    // six (nni5) or two neighbours by copies on the spare slots, nniSwap does / undoes the swap of the neighbours at positions k1 of
    // node1 and k2 of node2 (re-orienting the two central CLVs first), nniEnd undoes a pending swap and puts the saved neighbours back
    struct NNIContext {
        PhyloNode *node1 = nullptr, *node2 = nullptr;
        int IT_NUM = 0;
        struct Saved { PhyloNode *owner; int pos; PhyloNeighbor *nei; } saved[6];
        PhyloNeighbor copies[6];
        PhyloNeighbor *node12_it = nullptr, *node21_it = nullptr;
        int swapped_k1 = -1, swapped_k2 = -1;
    };
    void nniBegin(PhyloNode *node1, PhyloNode *node2, bool nni5, NNIContext &ctx);
    void nniSwap(NNIContext &ctx, int k1, int k2);
    void nniEnd(NNIContext &ctx);
    PhyloNode *nodeById(int id) const;
    static void defaultNNIMoves(PhyloNode *node1, PhyloNode *node2, NNIMove *nniMoves);   // :4168-4183
    void doNNI(const NNIMove &move, bool clearLH = true);   // :3947-4036
    void changeNNIBrans(const NNIMove &move);               // :4038-4067

    // ---- tree access ----------------------------------------------------------------------------------------------
    PhyloNode *findNodeName(const std::string &name) const;
    PhyloNode *findFarthestLeaf(PhyloNode *node = nullptr, PhyloNode *dad = nullptr);   // tree/mtree.cpp:3040-3060
    void getPreOrderBranches(std::vector<PhyloNode *> &nodes, std::vector<PhyloNode *> &nodes2, PhyloNode *node,
                             PhyloNode *dad = nullptr);                               // tree/mtree.cpp:3062-3082
    void getBranches(std::vector<PhyloNode *> &nodes, std::vector<PhyloNode *> &nodes2) const;   // by branch id
    void saveBranchLengths(std::vector<double> &lenvec) const;
    virtual void restoreBranchLengths(const std::vector<double> &lenvec);
    // the directed neighbour of branch `branch_id` that points TO the node with id `to_node_id` (nullptr if neither end has it)
    PhyloNeighbor *branchNeighbor(int branch_id, int to_node_id) const;
    int numBranches() const { return (int)branches_.size(); }
    int64_t getAlnNPattern() const { return nptn_; }
    int getNumStates() const { return nstates_; }

    // diagnostics for parity tests
    int64_t num_partial_lh_computations = 0;   // tree/phylotree.h:2127
    b200lk_engine *engine() const { return engine_; }
    const std::vector<b200lk_op> &lastTraversal() const { return traversal_info; }
    const std::vector<b200lk_child> &lastPlanExtraChildren() const { return last_plan_extra_; }

    Params params;
    PhyloNode *root = nullptr;
    PhyloNeighbor *current_it = nullptr, *current_it_back = nullptr;
    bool theta_computed = false;
    double curScore = 0.0;
    double current_scaling = 1.0;     // tree/phylotree.h: the scaling factor while optimizeTreeLengthScaling runs
    bool is_opt_scaling = false;
    int leafNum = 0, nodeNum = 0;
    int max_lh_slots = 0;             // tree/phylotree.h (set by initializeAllPartialLh; T-2 for LM_PER_NODE)
    MemSlotVector mem_slots;

protected:
    bool computeTraversalInfo(PhyloNeighbor *dad_branch, PhyloNode *dad);   // tree/phylotree.cpp:5918-6028
    void computeTraversalInfo(PhyloNode *node, PhyloNode *dad);             // tree/phylokernelnew.h:1150-1281
    void reorientPartialLh(PhyloNeighbor *dad_branch, PhyloNode *dad);      // tree/phylotree.cpp:5898-5912
    int computeSize(PhyloNeighbor *nei, PhyloNode *dad);
    void check(int rc, const char *what);
    void checkPlanned(int rc, const char *what);   // check() + roll the host state back when a planned traversal was refused
    void freeTree();
    double brent_opt(double ax, double bx, double cx, double tol, double *foptx, double *f2optx, double fax, double fbx, double fcx);
    void branchEnds(PhyloNeighbor *&dad_branch, PhyloNode *&dad, int &node_slot, int &dad_kind, int &dad_id);

    std::vector<PhyloNode *> nodes_;
    std::vector<PhyloNeighbor *> neighbors_;
    std::vector<std::pair<PhyloNeighbor *, PhyloNeighbor *>> branches_;   // by branch id: (a->b, b->a)
    std::vector<std::string> taxa_;
    std::vector<uint8_t> states_;
    std::vector<double> ptn_freq_, ptn_invar_;
    int64_t nptn_ = 0;
    int nstates_ = 0, seq_type_ = 0, state_unknown_ = 0;
    int ncat_ = 0, nmix_ = 1;
    bool fused_ = false;
    std::vector<double> eval_, evec_, inv_evec_, mix_weight_, cat_rate_, cat_prop_;
    b200lk_engine *engine_ = nullptr;
    bool initialized_ = false;
    std::vector<b200lk_op> traversal_info;
    std::vector<b200lk_child> traversal_extra;   // third and later children of multifurcating nodes, in op order
    std::vector<b200lk_child> last_plan_extra_;
    int asc_type_ = 0;
    int64_t asc_nunobs_ = 0, asc_nsite_ = 0;
    void pushExtraChildren();
    bool sweepPipelined(const std::vector<PhyloNode *> &nodes, const std::vector<PhyloNode *> &nodes2, int maxNRStep);
    std::vector<char> dev_len_valid_;   // by branch id: the current length lives in the engine's device length table
    std::vector<double> pattern_lh_;
    mutable std::vector<PhyloNode *> node_by_id_;
};

}  // namespace b200
