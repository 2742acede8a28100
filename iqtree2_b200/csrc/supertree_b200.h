// supertree_b200.h — host-side mirror of the reference's edge-linked partition model, PhyloSuperTreePlen
// (tree/phylosupertreeplen.h, tree/phylosupertreeplen.cpp:243-412; PhyloSuperTree::computeLikelihood,
// tree/phylosupertree.cpp:697-726), over per-partition b200::PhyloTree objects = per-partition device engines.
//
// Like the reference's class it IS a PhyloTree: the super tree carries the topology and the linked branch lengths and inherits
// optimizeOneBranch / optimizeAllBranches / minimizeNewton unchanged; what it overrides is where the numbers come from —
// computeFunction / computeFuncDerv / computeLikelihoodBranch / computeLikelihoodFromBuffer sum over the partition trees:
//     logL = sum_p logL_p,     df = sum_p rate_p df_p,     ddf = sum_p rate_p^2 ddf_p        (phylosupertreeplen.cpp:307-412)
// with every partition branch at length super_length * part_rate.  Where the reference fans the partitions out over OpenMP
// threads that each block in their partition's kernel, here every partition's work is QUEUED on its own CUDA stream
// (b200lk_enqueue_*), from OpenMP threads as well, and collected with one b200lk_finish per partition: the GPU sees hundreds
// of small independent traversals at once instead of one at a time.
//
// Scope of this mirror: every partition tree has the super tree's topology over the full taxon set; a taxon without data in a
// partition is a row of STATE_UNKNOWN there (mathematically the likelihood of the reference's induced subtree, whose merged
// branches have the summed lengths: an all-gap leaf contributes the factor 1).  SuperNeighbor::link_neighbors therefore is the
// identity on branch ids, and `if (nei1_part && nei2_part)` is always true.
#pragma once
#include <vector>

#include "phylotree_b200.h"

namespace b200 {

class PhyloSuperTreePlen : public PhyloTree {
public:
    PhyloSuperTreePlen();
    ~PhyloSuperTreePlen() override;

    // partition tree with alignment and model already set (same taxon names, same order as the super tree); takes ownership
    int addPartition(PhyloTree *part, double part_rate);
    // the super tree's taxa (alignment-row order); must precede readSuperTree
    void setTaxa(const std::vector<std::string> &names);
    // reads the topology into the super tree and every partition tree, then mapTrees()
    void readSuperTree(const std::string &newick);
    // partition branch lengths := super branch lengths * part_rate (PhyloSuperTreePlen::mapTrees, phylosupertreeplen.cpp:119-160)
    void mapTrees();
    void setPartRate(int part, double rate);

    double computeLikelihood();                                                                  // phylosupertree.cpp:697-726
    double computeLikelihoodBranch(PhyloNeighbor *dad_branch, PhyloNode *dad, bool save_log_value = true) override;   // plen.cpp:355-359
    double computeLikelihoodFromBuffer() override;                                               // plen.cpp:361-370
    double computeFunction(double value) override;                                               // plen.cpp:307-353
    void computeFuncDerv(double value, double &df, double &ddf) override;                        // plen.cpp:372-412
    void optimizeOneBranch(PhyloNode *node1, PhyloNode *node2, bool clearLH = true, int maxNRStep = 100) override;   // plen.cpp:253-305
    double optimizeAllBranches(int my_iterations = 100, double tolerance = TOL_LIKELIHOOD, int maxNRStep = 100) override;   // plen.cpp:243-251
    void clearAllPartialLH() override;
    void restoreBranchLengths(const std::vector<double> &lenvec) override;
    // NNI evaluation for the linked model (the role of PhyloSuperTreePlen::getBestNNIForBran, phylosupertreeplen.cpp:1100-1500;
    // here every partition tree has the full topology, so the move is the same swap in the super tree and in every partition
    // tree, each on its own engine's spare slots; the five branches are re-optimised with the linked derivatives)
    NNIMove getBestNNIForBran(PhyloNode *node1, PhyloNode *node2, NNIMove *nniMoves = nullptr) override;
    // doNNI (+ changeNNIBrans when set_lengths) in the super tree and in every partition tree
    void doNNIAll(const NNIMove &move, bool clearLH, bool set_lengths);

    int numPartitions() const { return (int)parts.size(); }
    std::vector<PhyloTree *> parts;
    std::vector<double> part_rate;
    std::vector<double> cur_score;   // part_info[part].cur_score
    int num_threads = 0;             // host threads that queue partition work (0 = OpenMP default)
    int64_t launches() const;        // kernel launches of all partition engines
    // grouped launches (b200lk_group_*): one launch per kernel type for all partitions instead of one stream and ~5 launches per
    // partition.  On by default when all partitions sit on one device; B200ST_NO_GROUP=1 (or use_group = false before the first
    // evaluation) keeps the per-partition streams.
    bool use_group = true;

private:
    b200lk_group *group_ = nullptr;
    std::vector<b200lk_engine *> group_engines_;   // the members the group was built over
    int timing_calls_ = 0;
    double timing_[4] = {0, 0, 0, 0};              // B200ST_TIMING=1: seconds in planning / enqueue / wait, calls
    bool ensureGroup();
    void dropGroup();
    template <class PlanFn> void groupEvaluate(int what, PlanFn plan, double *logl, double *df, double *ddf);
    void linkBranch(int part, PhyloNeighbor *&nei1_part, PhyloNeighbor *&nei2_part) const;
    void setPartLengths(double value);
    std::vector<int> part_order;     // PhyloSuperTree::computePartitionOrder: decreasing cost
    void computePartitionOrder();
};

// The edge-unlinked partition model (PhyloSuperTree proper; `-Q` / `-sp`: every partition tree has its own topology-induced
// subtree and its own branch lengths; tree/phylosupertree.cpp:653-791).  The partition trees are independent PhyloTree objects
// (borrowed: the caller configures and owns them); the class is the driver: logL = sum_p logL_p with the partitions evaluated
// concurrently, most expensive first (computePartitionOrder :653-695), one device engine and stream each, and
// optimizeAllBranches = every partition's own optimizeAllBranches with tolerance / min(ntrees, 10) (:773-791).
// computeBranchLengths (the super tree's lengths as averages over the partitions, :787) is display logic of the caller.
class PhyloSuperTree {
public:
    PhyloSuperTree();
    int addPartition(PhyloTree *part);
    double computeLikelihood(double *pattern_lh = nullptr);                                                  // :697-726
    void computePatternLikelihood(double *pattern_lh, double *cur_logl = nullptr);                           // :736-763 (concatenated)
    double optimizeAllBranches(int my_iterations = 100, double tolerance = TOL_LIKELIHOOD, int maxNRStep = 100);   // :773-791
    void clearAllPartialLH();
    int numPartitions() const { return (int)parts.size(); }
    int64_t numPatterns() const;
    std::vector<PhyloTree *> parts;
    std::vector<double> cur_score;   // part_info[part].cur_score
    int num_threads = 0;

private:
    std::vector<int> part_order;
    void computePartitionOrder();
};

}  // namespace b200
