// ref_harness.cpp — TEST INFRASTRUCTURE.  Drives the UNMODIFIED reference likelihood kernel
// (IQ-TREE 2.4.0, tree/phylokernelnew.h instantiated by tree/phylokernelfma.cpp for Vec4d + FMA, dispatched
// through PhyloTree's five pointers-to-member, tree/phylotreesse.cpp:91-236) behind a small C interface, so that
//   (a) oracle/lk_oracle.c (the scalar restatement) and the CUDA engine can be pinned against the real thing, and
//   (b) bench.py can time the reference's own AVX2/FMA OpenMP path on the host cores ("cpu_baseline.kind": "reference").
//
// Nothing here is reference source: this file only #includes the reference's headers where they lie under
// /root/reference and calls its public members.  It is compiled by oracle/build_ref.sh together with the reference
// TUs (compiled in place, objects and the .so go to oracle/_ref/, git-ignored).  Only tests/, smoke() and bench.py's
// CPU arms may load the result.
//
// What the harness supplies itself (and why that is legitimate for a *kernel* oracle):
//   * HModel / HRate: ModelSubst / RateHeterogeneity subclasses that return a caller-provided eigensystem
//     (eval, evec, inv_evec per mixture class), mixture weights, state frequencies, category rates/proportions and
//     p_invar.  The C ABI of the product takes the eigensystem as an input as well (the default
//     eigensolver is Eigen3, absent here, and eigenvector sign/order is solver dependent), so parity is defined for
//     identical eigensystems.  computeTipLikelihood's 0/1 indicator vectors follow model/modelsubst.cpp:197,
//     model/modeldna.cpp:462, model/modelprotein.cpp:1149 (those TUs need ModelMarkov -> Eigen3).
//   * link-time stubs for PhyloTree members defined in TUs that are not compiled (SSE/AVX kernel TUs).
// Everything downstream — Alignment/Pattern, MTree::readTree, initializeAllPartialLh, computeTipPartialLikelihood,
// computePtnFreq/Invar, computeTraversalInfo, computePartialInfo, the partial/branch/derv/from-buffer kernels,
// optimizeOneBranch/optimizeAllBranches + minimizeNewton — is the reference's compiled code.
#include <cstdint>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

#include "alignment/alignment.h"
#include "model/modelfactory.h"
#include "model/modelsubst.h"
#include "model/rateheterogeneity.h"
#include "tree/phylotree.h"
#include "tree/phylonodemixlen.h"
#include "utils/starttree.h"
#include "utils/timeutil.h"
#include "utils/tools.h"

#ifdef _OPENMP
#include <omp.h>
#endif
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

namespace {

// the kernels load eval/evec/inv_evec with aligned SIMD loads (the reference allocates them with aligned_alloc_double)
struct AVec {
    double *p = nullptr;
    size_t n = 0;
    ~AVec() { free(p); }
    void assign(const double *a, const double *b) {
        free(p);
        n = (size_t)(b - a);
        if (posix_memalign((void **)&p, 64, sizeof(double) * (n + 8)) != 0) abort();
        memcpy(p, a, sizeof(double) * n);
    }
    double *data() const { return p; }
    double &operator[](size_t i) const { return p[i]; }
};

struct HModel : public ModelSubst {
    int nmix, seqtype;   // seqtype: 0 generic, 1 DNA, 2 protein (include/b200lk.h B200LK_SEQ_*)
    int state_unknown;
    AVec eval, evec, inv_evec;
    std::vector<double> weight, freq;
    // non-reversible mode (the kernels of tree/phylokernelnonrev.h): the complex eigensystem ModelMarkov would hold (ceval, cevec,
    // cinv_evec; [nmix][n](re,im) / [nmix][n][n](re,im) row-major) and the rate matrices, supplied by the caller
    bool nonrev = false;
    std::vector<double> ceval, cevec, cinv, qmat;
    // site-specific mode (ModelSet: `-fs` site-frequency files / PMSF; the SITE_MODEL template branch of tree/phylokernelnew.h): one
    // eigensystem per pattern, interleaved by the kernel's SIMD width V as ModelSet lays them out — eval[(p/V) n V + i V + p%V],
    // evec / inv_evec[(p/V) n^2 V + k V + p%V] (read that way at tree/phylotreesse.cpp:271-281, phylokernelnew.h:1660-1662)
    bool site = false;
    AVec s_eval, s_evec, s_inv;
    bool isSiteSpecificModel() override { return site; }
    void make_site(const std::vector<double> &ev, const std::vector<double> &U, const std::vector<double> &Ui, size_t nptn, size_t V) {
        const size_t n = num_states, nn = n * n, maxp = (nptn + V - 1) / V * V;
        std::vector<double> a(maxp * n), b(maxp * nn), c(maxp * nn);
        for (size_t p = 0; p < maxp; p++) {
            const size_t q = p < nptn ? p : nptn - 1, base = p / V, v = p % V;   // padding patterns: the last pattern's system
            for (size_t i = 0; i < n; i++) a[base * n * V + i * V + v] = ev[q * n + i];
            for (size_t k = 0; k < nn; k++) { b[base * nn * V + k * V + v] = U[q * nn + k]; c[base * nn * V + k * V + v] = Ui[q * nn + k]; }
        }
        s_eval.assign(a.data(), a.data() + a.size()); s_evec.assign(b.data(), b.data() + b.size()); s_inv.assign(c.data(), c.data() + c.size());
        site = true;
    }
    HModel(int n, int nmix_, int seqtype_, int su) : ModelSubst(n), nmix(nmix_), seqtype(seqtype_), state_unknown(su) {
        name = "HARNESS";
        full_name = "harness-supplied eigensystem";
    }
    bool isReversible() override { return !nonrev; }
    // ModelMarkov::computeTransMatrixNonrev, MET_EIGEN3LIB_DECOMPOSITION (model/modelmarkov.cpp:477-486; ModelMarkov itself needs
    // Eigen3 and is not compiled): P = Re(cevec . diag(exp(ceval t)) . cinv_evec)
    void computeTransMatrix(double time, double *trans_matrix, int mixture = 0, int selected_row = -1) override {
        const int n = num_states;
        const double *ev = &ceval[(size_t)mixture * n * 2], *V = &cevec[(size_t)mixture * n * n * 2], *W = &cinv[(size_t)mixture * n * n * 2];
        std::vector<double> er(n), ei(n);
        for (int k = 0; k < n; k++) {
            const double a = exp(ev[2 * k] * time);
            er[k] = a * cos(ev[2 * k + 1] * time);
            ei[k] = a * sin(ev[2 * k + 1] * time);
        }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                double sum = 0.0;
                for (int k = 0; k < n; k++) {
                    const double vr = V[(size_t)(i * n + k) * 2], vi = V[(size_t)(i * n + k) * 2 + 1];
                    const double ar = vr * er[k] - vi * ei[k], ai = vr * ei[k] + vi * er[k];
                    sum += ar * W[(size_t)(k * n + j) * 2] - ai * W[(size_t)(k * n + j) * 2 + 1];
                }
                trans_matrix[i * n + j] = sum;
            }
    }
    // ModelMarkov::computeTransDerv, non-reversible (model/modelmarkov.cpp:735-747): Q.P and Q.Q.P
    void computeTransDerv(double time, double *trans_matrix, double *trans_derv1, double *trans_derv2, int mixture = 0) override {
        const int n = num_states;
        const double *Q = &qmat[(size_t)mixture * n * n];
        computeTransMatrix(time, trans_matrix, mixture);
        for (int pass = 0; pass < 2; pass++) {
            const double *src = pass ? trans_derv1 : trans_matrix;
            double *dst = pass ? trans_derv2 : trans_derv1;
            for (int i = 0; i < n; i++)
                for (int j = 0; j < n; j++) {
                    double sum = 0.0;
                    for (int k = 0; k < n; k++) sum += Q[i * n + k] * src[k * n + j];
                    dst[i * n + j] = sum;
                }
        }
    }
    bool isMixture() override { return nmix > 1; }
    int getNMixtures() override { return nmix; }
    double getMixtureWeight(int cat) override { return weight[cat]; }
    double *getEigenvalues() const override { return site ? s_eval.data() : eval.data(); }
    double *getEigenvectors() const override { return site ? s_evec.data() : evec.data(); }
    double *getInverseEigenvectors() const override { return site ? s_inv.data() : inv_evec.data(); }
    void getStateFrequency(double *f, int mixture = 0) override {
        // mixture -1: weighted average over classes (ModelMixture::getStateFrequency, model/modelmixture.cpp)
        if (mixture >= 0) { memcpy(f, &freq[(size_t)mixture * num_states], sizeof(double) * num_states); return; }
        for (int i = 0; i < num_states; i++) f[i] = 0.0;
        for (int m = 0; m < nmix; m++)
            for (int i = 0; i < num_states; i++) f[i] += weight[m] * freq[(size_t)m * num_states + i];
    }
    void computeTipLikelihood(PML::StateType state, double *lk) override {
        const int n = num_states;
        for (int i = 0; i < n; i++) lk[i] = 0.0;
        if ((int)state < n) { lk[state] = 1.0; return; }
        if ((int)state == state_unknown) { for (int i = 0; i < n; i++) lk[i] = 1.0; return; }
        if (seqtype == 1) {
            int cstate = (int)state - n + 1;
            for (int i = 0; i < n; i++) if ((cstate >> i) & 1) lk[i] = 1.0;
        } else if (seqtype == 2) {
            static const int ambi_aa[] = {4 + 8, 32 + 64, 512 + 1024};
            int cstate = (int)state - n;
            for (int i = 0; i < 11; i++) if (ambi_aa[cstate] & (1 << i)) lk[i] = 1.0;
        } else {
            for (int i = 0; i < n; i++) lk[i] = 1.0;
        }
    }
    void startCheckpoint() override {}
    void saveCheckpoint() override {}
    void restoreCheckpoint() override {}
};

struct HRate : public RateHeterogeneity {
    std::vector<double> rate, prop;
    double pinv = 0.0;
    int getNRate() override { return (int)rate.size(); }
    int getNDiscreteRate() override { return (int)rate.size(); }
    double getRate(int c) override { return rate[c]; }
    double getProp(int c) override { return prop[c]; }
    double getPInvar() override { return pinv; }
    void startCheckpoint() override {}
    void saveCheckpoint() override {}
    void restoreCheckpoint() override {}
};

struct Handle {
    Alignment *aln = nullptr;
    PhyloTree *tree = nullptr;
    HModel *model = nullptr;
    HRate *rate = nullptr;
    ModelFactory *fac = nullptr;
    int n = 0, C = 0, B = 0, ntaxa = 0;
    size_t nptn = 0;
    std::vector<PhyloNode *> leaf;   // by alignment row
    std::string err;
};

bool g_params_ready = false;

// IQREF_DEBUG=1: print a symbolised backtrace on SIGSEGV (no debugger in the image)
void segv_backtrace(int sig) {
    void *frames[64];
    int nf = backtrace(frames, 64);
    backtrace_symbols_fd(frames, nf, 2);
    _exit(128 + sig);
}

void init_params(int nthreads) {
    Params &p = Params::getInstance();
    if (!g_params_ready) {
        if (getenv("IQREF_DEBUG")) signal(SIGSEGV, segv_backtrace);
        char a0[] = "iqtree2", a1[] = "-s", a2[] = "harness.phy", a3[] = "-quiet";
        char *argv[] = {a0, a1, a2, a3, nullptr};
        parseArg(4, argv, p);
        g_params_ready = true;
    }
    p.num_threads = nthreads;
#ifdef _OPENMP
    omp_set_num_threads(nthreads);
#endif
}

PhyloNeighbor *nei_towards(PhyloNode *from, PhyloNode *to) { return (PhyloNeighbor *)from->findNeighbor(to); }

void collect_leaves(Node *node, Node *dad, std::vector<uint8_t> &mask) {
    if (node->isLeaf()) { if (node->id >= 0 && node->id < (int)mask.size()) mask[node->id] = 1; return; }   // (the root leaf of a rooted tree has id ntaxa)
    FOR_NEIGHBOR_IT(node, dad, it) collect_leaves((*it)->node, node, mask);
}

struct Directed { PhyloNeighbor *nei; PhyloNode *node, *dad; };   // nei = dad->findNeighbor(node): CLV of node's subtree seen from dad

void collect_directed(PhyloNode *node, PhyloNode *dad, std::vector<Directed> &out) {
    FOR_NEIGHBOR_IT(node, dad, it) {
        PhyloNode *child = (PhyloNode *)(*it)->node;
        // both directions of edge (node, child)
        out.push_back({(PhyloNeighbor *)(*it), child, node});
        out.push_back({nei_towards(child, node), node, child});
        collect_directed(child, node, out);
    }
}

// a PhyloTree with mixed branch lengths (heterotachy): what PhyloTreeMixlen (tree/phylotreemixlen.h:89-98, 218; newNode
// tree/phylotreemixlen.cpp:95-101) answers to the likelihood kernels — isMixlen(), getMixlen(), getCurMixture() — over
// PhyloNodeMixlen / PhyloNeighborMixlen nodes; none of its optimisers (they need ModelFactoryMixlen, RateHeterotachy, BFGS ...)
struct MixlenTree : public PhyloTree {
    int K = 1, cur = 0;
    explicit MixlenTree(int k) : PhyloTree(), K(k) {}
    bool isMixlen() override { return true; }
    int getMixlen() override { return K; }
    int getCurMixture() override { return cur; }
    Node *newNode(int node_id = -1, const char *node_name = NULL) override { return new PhyloNodeMixlen(node_id, node_name); }
    Node *newNode(int node_id, int node_name) override { return new PhyloNodeMixlen(node_id, node_name); }
    // PhyloTreeMixlen::computeFuncDerv's call (tree/phylotreemixlen.cpp:736)
    void dervCur(PhyloNeighbor *nb, PhyloNode *dad, double &df, double &ddf) { (this->*computeLikelihoodDervMixlenPointer)(nb, dad, df, ddf); }
};

void set_mixlen_lengths(Node *node, Node *dad, const std::vector<double> &factor) {
    FOR_NEIGHBOR_IT(node, dad, it) {
        DoubleVector v;
        for (double f : factor) v.push_back((*it)->length * f);
        (*it)->setLength(v);
        (*it)->node->findNeighbor(node)->setLength(v);
        set_mixlen_lengths((*it)->node, node, factor);
    }
}

}  // namespace

// the shim's view of the non-reversible arrays (oracle/ref_b200_shim.cpp)
bool iqref_nonrev_arrays(ModelSubst *model, const double **ceval, const double **cevec, const double **cinv, const double **qmat) {
    HModel *m = dynamic_cast<HModel *>(model);
    if (!m || !m->nonrev) return false;
    *ceval = m->ceval.data(); *cevec = m->cevec.data(); *cinv = m->cinv.data(); *qmat = m->qmat.data();
    return true;
}

// ---- link-time stubs for members whose TUs (phylokernelsse.cpp, phylotreeavx.cpp) are not compiled -------------
void PhyloTree::setParsimonyKernelSSE() {}   // parsimony is not on the likelihood path
void PhyloTree::setParsimonyKernelAVX() {}
void PhyloTree::setDotProductSSE() {}
void PhyloTree::setDotProductAVX() {}
void PhyloTree::setLikelihoodKernelSSE() { outError("harness: SSE kernel TU not linked"); }
void PhyloTree::setLikelihoodKernelAVX() { outError("harness: AVX kernel TU not linked"); }

// parseArg (utils/tools.cpp:1015) asks the distance-tree factory for the name of its default builder (utils/starttree.cpp pulls
// in every NJ/BIONJ implementation; none is on the likelihood path)
const std::string &StartTree::Factory::getNameOfDefaultTreeBuilder() {
    static const std::string name = "NONE";
    return name;
}

// how iqref_create makes its tree: nullptr = `new PhyloTree()`; oracle/ref_b200_shim.cpp installs a factory for its derived class
PhyloTree *(*iqref_tree_factory)() = nullptr;
int iqref_next_mixlen_classes = 0;   // > 0 while iqref_create builds a tree with mixed branch lengths (read by the shim's factory)

void (*iqref_derv_mixlen_cur_hook)(PhyloTree *, PhyloNeighbor *, PhyloNode *, int, double *, double *) = nullptr;

int NNI_MAX_NR_STEP = 10;   // utils/pllnni.cpp:42 (PLL glue, not compiled); read by getBestNNIForBran

extern "C" {

struct iqref_desc {
    int32_t nstates, ncat, nmix, fused_mix_rate, safe_numeric, seq_type, state_unknown, ntaxa;
    int64_t nptn;
    int32_t nthreads, reserved;
    const uint8_t *states;      // [ntaxa][nptn]
    const int32_t *ptn_freq;    // [nptn]
    const char *newick;         // leaf names are alignment row numbers (PhyloTree::readTreeString)
    const double *eval, *evec, *inv_evec, *mix_weight, *state_freq, *cat_rate, *cat_prop;
    double pinvar;
};

void iqref_destroy(void *hv);

// +ASC for the NEXT iqref_create: ModelFactory::ASC_type and ModelFactory::unobserved_ptns (what ModelFactory's constructor
// fills from Alignment::getUnobservedConstPatterns, model/modelfactory.cpp) — unobs is [nunobs][ntaxa] state codes
static int g_asc_type = 0;
static std::vector<std::vector<uint8_t>> g_unobs;
void iqref_set_asc_next(int asc_type, const uint8_t *unobs, int nunobs, int ntaxa) {
    g_asc_type = asc_type;
    g_unobs.clear();
    for (int u = 0; u < nunobs; u++) g_unobs.emplace_back(unobs + (size_t)u * ntaxa, unobs + (size_t)(u + 1) * ntaxa);
}

// non-reversible model for the NEXT iqref_create: HModel::nonrev and its arrays; the Newick handed to that create has a bifurcating
// root, so MTree::readTree makes the tree rooted and attaches the root leaf (tree/mtree.cpp:742-754) — or, with rooted = 0, a
// trifurcating one (the non-reversible kernel on an unrooted tree, as --kernel-nonrev runs it)
static bool g_nonrev = false;
static std::vector<double> g_ceval, g_cevec, g_cinv, g_qmat;
void iqref_set_nonrev_next(const double *ceval, const double *cevec, const double *cinv, const double *qmat, int n, int nmix) {
    g_nonrev = true;
    g_ceval.assign(ceval, ceval + (size_t)2 * nmix * n);
    g_cevec.assign(cevec, cevec + (size_t)2 * nmix * n * n);
    g_cinv.assign(cinv, cinv + (size_t)2 * nmix * n * n);
    g_qmat.assign(qmat, qmat + (size_t)nmix * n * n);
}

// site-specific model for the NEXT iqref_create: per-pattern eigensystems, plain [p][n] / [p][n][n] (HModel::make_site interleaves them)
static bool g_site = false;
static std::vector<double> g_s_eval, g_s_evec, g_s_inv;
void iqref_set_site_next(const double *eval, const double *evec, const double *inv_evec, int64_t nptn, int n) {
    g_site = true;
    g_s_eval.assign(eval, eval + (size_t)nptn * n);
    g_s_evec.assign(evec, evec + (size_t)nptn * n * n);
    g_s_inv.assign(inv_evec, inv_evec + (size_t)nptn * n * n);
}

// mixed branch lengths for the NEXT iqref_create: the tree is a MixlenTree with k = ncat classes, every branch gets
// lengths[c] = length * factor[c] (the callers build the same table from the same tree)
static std::vector<double> g_mixlen_factor;
void iqref_set_mixlen_next(const double *factor, int k) { g_mixlen_factor.assign(factor, factor + k); }

void *iqref_create(const iqref_desc *d) {
    init_params(d->nthreads > 0 ? d->nthreads : 1);
    Params &params = Params::getInstance();
    params.lk_safe_scaling = d->safe_numeric != 0;
    // what runPhyloAnalysis does after loading the alignment (main/phyloanalysis.cpp:4424-4430): parseArg leaves
    // min_branch_length at 0; it is also minimizeNewton's tolerance (tree/phylotree.cpp:2612)
    {
        double nsites = 0;
        for (int64_t p = 0; p < d->nptn; p++) nsites += d->ptn_freq ? d->ptn_freq[p] : 1;
        params.min_branch_length = nsites >= 100000 ? 0.1 / nsites : 1e-6;
    }
    Handle *h = new Handle();
    const int n = d->nstates;
    h->n = n; h->ntaxa = d->ntaxa; h->nptn = (size_t)d->nptn;
    h->C = d->fused_mix_rate ? d->ncat : d->ncat * d->nmix;
    h->B = n * h->C;

    Alignment *aln = h->aln = new Alignment();
    aln->num_states = n;
    aln->seq_type = d->seq_type == 1 ? SEQ_DNA : d->seq_type == 2 ? SEQ_PROTEIN : SEQ_MORPH;
    aln->STATE_UNKNOWN = (StateType)d->state_unknown;
    for (int t = 0; t < d->ntaxa; t++) aln->seq_names.push_back("T" + std::to_string(t));
    // one site per pattern occurrence is not needed by the kernel: site_pattern maps site -> pattern; we register
    // each pattern once at its own site and carry the multiplicity in Pattern::frequency (alignment/pattern.h)
    aln->site_pattern.resize(d->nptn, -1);
    aln->reserve(d->nptn);
    for (int64_t p = 0; p < d->nptn; p++) {
        Pattern pat;
        pat.resize(d->ntaxa);
        for (int t = 0; t < d->ntaxa; t++) pat[t] = d->states[(size_t)t * d->nptn + p];
        pat.frequency = d->ptn_freq ? d->ptn_freq[p] : 1;
        aln->computeConst(pat);
        aln->push_back(pat);             // no pattern merging: caller's pattern order is kept
        aln->site_pattern[p] = (int)p;
    }

    HModel *m = h->model = new HModel(n, d->nmix, d->seq_type, d->state_unknown);
    m->eval.assign(d->eval, d->eval + (size_t)d->nmix * n);
    m->evec.assign(d->evec, d->evec + (size_t)d->nmix * n * n);
    m->inv_evec.assign(d->inv_evec, d->inv_evec + (size_t)d->nmix * n * n);
    m->weight.assign(d->mix_weight, d->mix_weight + d->nmix);
    m->freq.assign(d->state_freq, d->state_freq + (size_t)d->nmix * n);
    if (g_nonrev) { m->nonrev = true; m->ceval = g_ceval; m->cevec = g_cevec; m->cinv = g_cinv; m->qmat = g_qmat; g_nonrev = false; }
    HRate *r = h->rate = new HRate();
    r->rate.assign(d->cat_rate, d->cat_rate + d->ncat);
    r->prop.assign(d->cat_prop, d->cat_prop + d->ncat);
    r->pinv = d->pinvar;
    ModelFactory *fac = h->fac = new ModelFactory();
    fac->model = m;
    fac->site_rate = r;
    fac->fused_mix_rate = d->fused_mix_rate != 0;
    fac->ASC_type = (ASCType)g_asc_type;
    for (auto &u : g_unobs) {
        Pattern pt;
        pt.assign(u.begin(), u.end());
        fac->unobserved_ptns.push_back(pt);
    }
    g_asc_type = 0;
    g_unobs.clear();

    const std::vector<double> mixlen_factor = g_mixlen_factor;
    g_mixlen_factor.clear();
    iqref_next_mixlen_classes = (int)mixlen_factor.size();
    PhyloTree *tree = h->tree = iqref_tree_factory ? iqref_tree_factory() : !mixlen_factor.empty() ? new MixlenTree((int)mixlen_factor.size()) : new PhyloTree();
    iqref_next_mixlen_classes = 0;
    tree->params = &params;
    tree->aln = aln;                       // readTreeString -> assignLeafNames needs it only for named leaves
    tree->readTreeString(std::string(d->newick));
    if (!mixlen_factor.empty()) set_mixlen_lengths(tree->root, nullptr, mixlen_factor);
    tree->setAlignment(aln);
    tree->setRootNode(params.root);
    tree->sse = LK_AVX_FMA;
    tree->num_threads = d->nthreads > 0 ? d->nthreads : 1;
    if (g_site) m->site = true;            // setLikelihoodKernel picks the SITE_MODEL instantiations from isSiteSpecificModel()
    tree->setModelFactory(fac);            // -> setLikelihoodKernel(sse) + setNumThreads
    if (g_site) { m->make_site(g_s_eval, g_s_evec, g_s_inv, (size_t)d->nptn, tree->vector_size); g_site = false; }
    r->setTree(tree);
    tree->initializeAllPartialLh();
    h->leaf.assign(d->ntaxa + (tree->rooted ? 1 : 0), nullptr);   // rooted: the root leaf has id ntaxa (PhyloTree::setAlignment)
    NodeVector taxa;
    tree->getTaxa(taxa);
    for (Node *x : taxa) if (x->id >= 0 && x->id < (int)h->leaf.size()) h->leaf[x->id] = (PhyloNode *)x;
    if (tree->rooted) h->leaf[d->ntaxa] = (PhyloNode *)tree->root;
    for (auto *x : h->leaf) if (!x) { h->err = "leaf missing"; }
    return h;
}

void iqref_destroy(void *hv) {
    Handle *h = (Handle *)hv;
    if (!h) return;
    // the tree does not own model/rate here (model_factory is ours); detach before deletion
    if (h->tree) { h->tree->model_factory = nullptr; h->tree->model = nullptr; h->tree->site_rate = nullptr; delete h->tree; }
    delete h->fac; delete h->model; delete h->rate; delete h->aln;
    delete h;
}

void *iqref_tree_ptr(void *hv) { return ((Handle *)hv)->tree; }

int iqref_info(void *hv, int32_t out[6]) {
    Handle *h = (Handle *)hv;
    out[0] = (int)h->tree->vector_size;
    out[1] = h->tree->safe_numeric ? 1 : 0;
    out[2] = h->tree->num_threads;
    out[3] = h->tree->num_packets;
    out[4] = (int)h->tree->sse;
    out[5] = (int)h->tree->num_partial_lh_computations;
    return 0;
}

void iqref_set_rates(void *hv, const double *rate, const double *prop, double pinvar) {
    Handle *h = (Handle *)hv;
    h->rate->rate.assign(rate, rate + h->rate->rate.size());
    h->rate->prop.assign(prop, prop + h->rate->prop.size());
    h->rate->pinv = pinvar;
    h->tree->clearAllPartialLH();
}

void iqref_clear_all_partial_lh(void *hv) { ((Handle *)hv)->tree->clearAllPartialLH(); }

// computeLikelihoodBranch on the pendant branch of alignment row `leaf` (dad = the leaf, dad_branch = leaf -> inner)
double iqref_branch_logl(void *hv, int leaf, double *pattern_lh, double *pattern_lh_cat) {
    Handle *h = (Handle *)hv;
    PhyloNode *L = h->leaf[leaf];
    PhyloNeighbor *nb = (PhyloNeighbor *)L->neighbors[0];
    h->tree->current_it = nb;
    h->tree->current_it_back = nei_towards((PhyloNode *)nb->node, L);
    double ll = h->tree->computeLikelihoodBranch(nb, L);
    if (pattern_lh) memcpy(pattern_lh, h->tree->_pattern_lh, sizeof(double) * h->nptn);
    if (pattern_lh_cat) {
        // _pattern_lh_cat is V-interleaved until transformPatternLhCat (tree/phylotree.cpp:1313); de-interleave here
        const size_t V = h->tree->vector_size, C = h->C;
        for (size_t p = 0; p < h->nptn; p++)
            for (size_t c = 0; c < C; c++) pattern_lh_cat[p * C + c] = h->tree->_pattern_lh_cat[(p / V) * C * V + c * V + p % V];
    }
    return ll;
}

// PhyloTree::computeLikelihood (tree/phylotree.cpp:1233): the tree picks the branch itself
double iqref_compute_likelihood(void *hv, double *pattern_lh) { return ((Handle *)hv)->tree->computeLikelihood(pattern_lh); }

void iqref_branch_derv(void *hv, int leaf, double *df, double *ddf) {
    Handle *h = (Handle *)hv;
    PhyloNode *L = h->leaf[leaf];
    PhyloNeighbor *nb = (PhyloNeighbor *)L->neighbors[0];
    h->tree->current_it = nb;
    h->tree->current_it_back = nei_towards((PhyloNode *)nb->node, L);
    h->tree->theta_computed = false;
    h->tree->computeLikelihoodDerv(nb, L, df, ddf);
}

// computeLikelihoodDerv of a mixlen tree (tree/phylokernelnew.h:2391-2460, 2556-2568): df [K + 1] (gradient, then the
// log-likelihood), ddf [K][K]
void iqref_branch_derv_mixlen(void *hv, int leaf, double *df, double *ddf) {
    Handle *h = (Handle *)hv;
    PhyloNode *L = h->leaf[leaf];
    PhyloNeighbor *nb = (PhyloNeighbor *)L->neighbors[0];
    h->tree->current_it = nb;
    h->tree->current_it_back = nei_towards((PhyloNode *)nb->node, L);
    h->tree->theta_computed = false;
    h->tree->computeLikelihoodDerv(nb, L, df, ddf);
}

// computeLikelihoodDervMixlen (:3431-3616) for class `cur`
void iqref_branch_derv_mixlen_cur(void *hv, int leaf, int cur, double *df, double *ddf) {
    Handle *h = (Handle *)hv;
    PhyloNode *L = h->leaf[leaf];
    PhyloNeighbor *nb = (PhyloNeighbor *)L->neighbors[0];
    h->tree->current_it = nb;
    h->tree->current_it_back = nei_towards((PhyloNode *)nb->node, L);
    h->tree->theta_computed = false;
    if (MixlenTree *t = dynamic_cast<MixlenTree *>(h->tree)) { t->cur = cur; t->dervCur(nb, L, *df, *ddf); }
    else if (iqref_derv_mixlen_cur_hook) iqref_derv_mixlen_cur_hook(h->tree, nb, L, cur, df, ddf);   // the shim's tree class
    else *df = *ddf = 0.0 / 0.0;
}

double iqref_logl_from_buffer(void *hv) { return ((Handle *)hv)->tree->computeLikelihoodFromBuffer(); }

// one full evaluation of SURVEY §8(d): clearAllPartialLH -> computeLikelihoodBranch -> computeLikelihoodDerv;
// returns seconds (getRealTime, utils/timeutil.h:138), best of `reps`
double iqref_time_eval(void *hv, int leaf, int reps, double *logl, double *df, double *ddf, double *mean_sec) {
    Handle *h = (Handle *)hv;
    double best = 1e300, sum = 0.0;
    for (int r = 0; r < reps; r++) {
        double t0 = getRealTime();
        h->tree->clearAllPartialLH();
        *logl = iqref_branch_logl(hv, leaf, nullptr, nullptr);
        iqref_branch_derv(hv, leaf, df, ddf);
        double t = getRealTime() - t0;
        sum += t;
        if (t < best) best = t;
    }
    if (mean_sec) *mean_sec = sum / reps;
    return best;
}

// every directed CLV that is currently computed: count, then fetch k-th as (leaf mask of the subtree, partial_lh
// de-interleaved to [p][B], scale_num [p] (NORM) or [p][C] (SAFE))
static void all_directed(Handle *h, std::vector<Directed> &v) {
    PhyloNode *root = (PhyloNode *)h->tree->root;
    collect_directed(root, nullptr, v);
}

int iqref_num_directed(void *hv) {
    Handle *h = (Handle *)hv;
    std::vector<Directed> v;
    all_directed(h, v);
    return (int)v.size();
}

// returns 1 if the k-th directed CLV is an internal node's and is computed (outputs filled), else 0
int iqref_get_directed(void *hv, int k, uint8_t *leafmask, double *partial, uint16_t *scale) {
    Handle *h = (Handle *)hv;
    std::vector<Directed> v;
    all_directed(h, v);
    const Directed &d = v[k];
    if (d.node->isLeaf() || !d.nei->partial_lh || !(d.nei->partial_lh_computed & 1)) return 0;
    std::vector<uint8_t> mask(h->ntaxa, 0);
    collect_leaves(d.node, d.dad, mask);
    memcpy(leafmask, mask.data(), h->ntaxa);
    const size_t V = h->tree->vector_size, B = h->B, C = h->C;
    if (partial)
        for (size_t p = 0; p < h->nptn; p++)
            for (size_t k2 = 0; k2 < B; k2++) partial[p * B + k2] = d.nei->partial_lh[(p / V) * B * V + k2 * V + p % V];
    if (scale) {
        if (h->tree->safe_numeric) memcpy(scale, d.nei->scale_num, sizeof(uint16_t) * h->nptn * C);
        else memcpy(scale, d.nei->scale_num, sizeof(uint16_t) * h->nptn);
    }
    return 1;
}

int iqref_get_theta(void *hv, double *theta, double *buffer_scale_all) {
    Handle *h = (Handle *)hv;
    const size_t V = h->tree->vector_size, B = h->B;
    for (size_t p = 0; p < h->nptn; p++)
        for (size_t k = 0; k < B; k++) theta[p * B + k] = h->tree->theta_all[(p / V) * B * V + k * V + p % V];
    if (buffer_scale_all) memcpy(buffer_scale_all, h->tree->buffer_scale_all, sizeof(double) * h->nptn);
    return 0;
}

int iqref_get_tip_partial_lh(void *hv, double *out) {
    Handle *h = (Handle *)hv;
    h->tree->computeTipPartialLikelihood();
    memcpy(out, h->tree->tip_partial_lh, sizeof(double) * (h->model->state_unknown + 1) * h->model->nmix * h->n);
    return 0;
}

int iqref_get_ptn_invar(void *hv, double *out) {
    Handle *h = (Handle *)hv;
    h->tree->computeTipPartialLikelihood();
    memcpy(out, h->tree->ptn_invar, sizeof(double) * h->nptn);
    return 0;
}

// PhyloTree::optimizeAllBranches (tree/phylotree.cpp:2698); returns the tree log-likelihood
double iqref_optimize_all_branches(void *hv, int iterations, double tolerance, int max_nr_step) {
    Handle *h = (Handle *)hv;
    return h->tree->optimizeAllBranches(iterations, tolerance, max_nr_step);
}

// PhyloTree::optimizeOneBranch (tree/phylotree.cpp:2594) on the pendant branch of alignment row `leaf`; returns the new length
double iqref_optimize_one_branch(void *hv, int leaf, int max_nr_step) {
    Handle *h = (Handle *)hv;
    PhyloNode *L = h->leaf[leaf];
    PhyloNode *inner = (PhyloNode *)L->neighbors[0]->node;
    h->tree->optimizeOneBranch(L, inner, true, max_nr_step);
    return L->neighbors[0]->length;
}

// the branch order optimizeAllBranches walks (computeBestTraversal, tree/phylotree.cpp:2689): for each branch the leaf
// mask of the side not containing taxon 0
int iqref_best_traversal(void *hv, uint8_t *masks) {
    Handle *h = (Handle *)hv;
    NodeVector n1, n2;
    h->tree->computeBestTraversal(n1, n2);
    for (size_t b = 0; b < n1.size(); b++) {
        std::vector<uint8_t> mask(h->ntaxa, 0);
        collect_leaves(n2[b], n1[b], mask);
        if (mask[0]) for (auto &x : mask) x = !x;
        memcpy(masks + b * h->ntaxa, mask.data(), h->ntaxa);
    }
    return (int)n1.size();
}

// branch lengths keyed by split: for each branch, mask[ntaxa] of the side NOT containing taxon 0, and its length
int iqref_get_branches(void *hv, uint8_t *masks /* [nbranch][ntaxa] */, double *lengths) {
    Handle *h = (Handle *)hv;
    NodeVector n1, n2;
    h->tree->getBranches(n1, n2);
    for (size_t b = 0; b < n1.size(); b++) {
        std::vector<uint8_t> mask(h->ntaxa, 0);
        collect_leaves(n2[b], n1[b], mask);
        if (mask[0]) for (auto &x : mask) x = !x;
        memcpy(masks + b * h->ntaxa, mask.data(), h->ntaxa);
        lengths[b] = n1[b]->findNeighbor(n2[b])->length;
    }
    return (int)n1.size();
}

// PhyloTree::getBestNNIForBran (tree/phylotree.cpp:4069) on the inner branch whose side without taxon 0 has leaf set `mask`,
// with Params::nni5 = nni5.  For each of the two moves: score[m] = NNIMove::newloglh, new_len[m][5] = newLen[0..4] (only [0]
// without nni5), swap_masks[m][2][ntaxa] = leaf sets of the two swapped subtrees (node1's, node2's), outer_masks[m][4][ntaxa] =
// leaf sets of the subtrees on the branches newLen[1..4] belong to (node1's neighbours, then node2's, after the swap).
// Returns 0, or -1 when no inner branch has that split.
int iqref_best_nni(void *hv, const uint8_t *mask, int nni5, double *score, double *new_len, uint8_t *swap_masks, uint8_t *outer_masks) {
    Handle *h = (Handle *)hv;
    NodeVector n1, n2;
    h->tree->getBranches(n1, n2);
    PhyloNode *node1 = nullptr, *node2 = nullptr;
    for (size_t b = 0; b < n1.size(); b++) {
        if (n1[b]->isLeaf() || n2[b]->isLeaf()) continue;
        std::vector<uint8_t> m(h->ntaxa, 0);
        collect_leaves(n2[b], n1[b], m);
        if (m[0]) for (auto &x : m) x = !x;
        if (memcmp(m.data(), mask, h->ntaxa) == 0) { node1 = (PhyloNode *)n1[b]; node2 = (PhyloNode *)n2[b]; break; }
    }
    if (!node1) return -1;
    Params::getInstance().nni5 = nni5 != 0;
    NNIMove moves[2];
    moves[0].ptnlh = moves[1].ptnlh = nullptr;
    moves[0].node1 = nullptr;
    h->tree->getBestNNIForBran(node1, node2, moves);
    node1 = moves[0].node1; node2 = moves[0].node2;
    std::vector<Node *> o1, o2;
    FOR_NEIGHBOR_IT(node1, node2, it) o1.push_back((*it)->node);
    FOR_NEIGHBOR_IT(node2, node1, it) o2.push_back((*it)->node);
    auto put = [&](uint8_t *dst, Node *sub, Node *from) {
        std::vector<uint8_t> m(h->ntaxa, 0);
        collect_leaves(sub, from, m);
        memcpy(dst, m.data(), h->ntaxa);
    };
    for (int m = 0; m < 2; m++) {
        score[m] = moves[m].newloglh;
        for (int i = 0; i < 5; i++) new_len[m * 5 + i] = (i == 0 || nni5) && !moves[m].newLen[i].empty() ? moves[m].newLen[i][0] : 0.0;
        Node *a = (*moves[m].node1Nei_it)->node, *c = (*moves[m].node2Nei_it)->node;
        put(swap_masks + (size_t)(m * 2 + 0) * h->ntaxa, a, node1);
        put(swap_masks + (size_t)(m * 2 + 1) * h->ntaxa, c, node2);
        // neighbour lists after the swap: a's position in node1 holds c, c's position in node2 holds a
        int k = 0;
        for (Node *x : o1) put(outer_masks + (size_t)(m * 4 + k++) * h->ntaxa, x == a ? c : x, x == a ? node2 : node1);
        for (Node *x : o2) put(outer_masks + (size_t)(m * 4 + k++) * h->ntaxa, x == c ? a : x, x == c ? node1 : node2);
    }
    return 0;
}

// entries of traversal_info left by the last computeLikelihoodBranch / computeLikelihoodDerv call (= partial updates it ran;
// the new kernel does not maintain num_partial_lh_computations)
int iqref_last_traversal_size(void *hv) { return (int)((Handle *)hv)->tree->traversal_info.size(); }

// PhyloTree::optimizeTreeLengthScaling (tree/phylotree.cpp:2457-2488): returns the log-likelihood, *scaling in/out
double iqref_optimize_tree_length_scaling(void *hv, double min_scaling, double *scaling, double max_scaling, double gradient_epsilon) {
    return ((Handle *)hv)->tree->optimizeTreeLengthScaling(min_scaling, *scaling, max_scaling, gradient_epsilon);
}

// PhyloTree::computePatternLhCat(wsl) (tree/phylotree.cpp:1338-1393) on the pendant branch of alignment row `leaf`: copies
// _pattern_lh_cat, nptn x getNumLhCat(wsl), and returns getNumLhCat(wsl)
int iqref_pattern_lh_cat_wsl(void *hv, int leaf, int wsl, double *out) {
    Handle *h = (Handle *)hv;
    PhyloNode *L = h->leaf[leaf];
    PhyloNeighbor *nb = (PhyloNeighbor *)L->neighbors[0];
    h->tree->current_it = nb;
    h->tree->current_it_back = nei_towards((PhyloNode *)nb->node, L);
    h->tree->computePatternLhCat((SiteLoglType)wsl);
    const int K = h->tree->getNumLhCat((SiteLoglType)wsl);
    if (out) memcpy(out, h->tree->_pattern_lh_cat, sizeof(double) * h->nptn * K);
    return K;
}

// --robust-phy keep / --robust-median (utils/tools.cpp:4374-4392): Params::robust_phy_keep / robust_median for trees created
// AFTERWARDS (initializeAllPartialLh allocates _site_lh when they are set, tree/phylotree.cpp:908-909); (1.0, 0) = off
void iqref_set_robust(double keep, int median) {
    init_params(Params::getInstance().num_threads > 0 ? Params::getInstance().num_threads : 1);
    Params::getInstance().robust_phy_keep = keep;
    Params::getInstance().robust_median = median != 0;
}

// -mem: Params::lh_mem_save / max_mem_size as parseArg sets them (utils/tools.cpp, "-mem"); applies to trees created
// AFTERWARDS (getMemoryRequired / MemSlotVector::init read the global Params)
void iqref_set_mem_save(int on, double max_mem_size) {
    init_params(Params::getInstance().num_threads > 0 ? Params::getInstance().num_threads : 1);
    Params::getInstance().lh_mem_save = on ? LM_MEM_SAVE : LM_PER_NODE;
    Params::getInstance().max_mem_size = max_mem_size;
}
int iqref_max_lh_slots(void *hv) { return (int)((Handle *)hv)->tree->max_lh_slots; }

// UFBoot RELL as IQTree::saveCurrentTree computes it (tree/iqtree.cpp:3564-3611): the tree's own `dotProduct` member
// (PhyloTree::dotProductSIMD<float, Vec8f>, tree/phylokernel.h:64-70, bound by setDotProductFMA) over aligned float arrays padded
// with zeros to the SIMD width.  boot: [nboot][nptn] floats.
int iqref_rell(void *hv, const double *pattern_lh, const float *boot, int nptn, int nboot, float *out) {
    PhyloTree *t = ((Handle *)hv)->tree;
    if (!t->dotProduct) return -1;
    const int maxnptn = (nptn + 15) / 16 * 16;
    BootValType *lh = aligned_alloc<BootValType>(maxnptn), *bs = aligned_alloc<BootValType>(maxnptn);
    memset(lh, 0, maxnptn * sizeof(BootValType));
    for (int i = 0; i < nptn; i++) lh[i] = (float)pattern_lh[i];
    for (int b = 0; b < nboot; b++) {
        memset(bs, 0, maxnptn * sizeof(BootValType));
        memcpy(bs, boot + (size_t)b * nptn, sizeof(float) * nptn);
        out[b] = (t->*(t->dotProduct))(lh, bs, nptn);
    }
    aligned_free(lh); aligned_free(bs);
    return 0;
}

int iqref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
