// b200lk.cu — engine (device buffers, op staging, launches) and the C ABI declared in include/b200lk.h.
//
// One engine = the likelihood buffers of one PhyloTree on one GPU for one pattern range
// (central_partial_lh / central_scale_num / theta_all / buffer_scale_all / _pattern_lh / _pattern_lh_cat /
// ptn_freq / ptn_invar / tip_partial_lh, tree/phylotree.cpp:893-1188), resident in HBM across calls.
// No CPU fallback: every compute entry point launches CUDA kernels or fails.
#include "b200lk.h"
#include "b200lk_kernels.cuh"
#include "b200lk_walk20ss.cuh"
#include "b200lk_walk20v2.cuh"
#include "b200lk_walk4.cuh"
#include "b200lk_walkgen.cuh"
#include "b200lk_group.cuh"
#include "b200lk_nonrev.cuh"
#include "b200lk_site.cuh"
#include "b200lk_mixlen.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace b200lk;

struct b200lk_engine {
    b200lk_config cfg{};
    int C = 0, B = 0, denom = 1, S = 1;
    long long P = 0, Ppad = 0;
    long long clv_width = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    Dev d{};
    // owned device buffers
    uint8_t *d_tips = nullptr;
    double *d_freq = nullptr, *d_invar = nullptr;
    double *d_model = nullptr;   // eval | evec | inv_evec | mix_weight | cat_rate | cat_prop | tip_lh | val(3B)
    double *d_partial = nullptr;
    uint16_t *d_scale = nullptr;
    double *d_theta = nullptr, *d_buffer_scale = nullptr, *d_pattern_lh = nullptr, *d_pattern_scaling = nullptr,
           *d_pattern_lh_cat = nullptr;
    double *d_opinfo = nullptr;
    size_t opinfo_cap = 0;   // doubles
    DevOp *d_ops = nullptr;
    ChildJob *d_jobs = nullptr;
    size_t ops_cap = 0;
    void *h_stage = nullptr;   // pinned staging: a ring of stage_ring segments of stage_seg bytes (ops + jobs of one call each)
    size_t h_stage_cap = 0, stage_seg = 0;
    int stage_ring = 32, stage_next = 0, stage_used = 0;
    double *d_red = nullptr;   // per-block partial sums
    double *d_result = nullptr;
    double *h_result = nullptr;   // pinned + mapped, 8 doubles: the final reductions write the scalars straight into it
    unsigned *d_ticket = nullptr;  // last-block ticket of the fused final reductions
    unsigned long long *d_counter = nullptr;
    double *d_scratch = nullptr;   // read-back conversion buffer
    size_t scratch_cap = 0;
    std::vector<double> h_freq;
    std::vector<double> h_evec, h_inv_evec;   // host copies: constant-bank operands of the 4-state walker
    int64_t device_bytes = 0;
    bool theta_computed = false;
    int schedule = 0;
    int num_sms = 148;
    int info_smem_optin = 0;
    bool walk20_attr_set = false;
    unsigned walk20v2_attr = 0;
    bool walkgen_attr_set = false;
    int walkgen_ctas_per_sm = 1;   // bit per (NW, CS) configuration of k_walk_n20v2 whose attributes are set
    // EM E-step (b200lk_em_posterior): posteriors [K][Ppad], per-block partial sums, the K sums
    double *d_post = nullptr, *d_em_red = nullptr, *d_em_out = nullptr;
    int post_K = 0;
    bool have_lh_cat = false;   // _pattern_lh_cat holds the categories of a branch_logl call
    // robust phylo (b200lk_set_robust_phylo): keys / values / running site count of the sorted patterns + cub's scratch
    double robust_keep = 1.0;
    int robust_median = 0;
    double *d_robust = nullptr;
    void *d_robust_tmp = nullptr;
    size_t robust_tmp_bytes = 0;
    int kernel_variant = 0;   // b200lk_set_schedule(e, 10 + v): development switch between kernel variants of one schedule
    // state-space CLVs (20-state engines, b200lk_walk20ss.cuh): two scratch slots behind the caller's nslots hold
    // partial_lh = inv_evec . t of the CLVs at the evaluation branch; slot_ver counts the writes of every slot
    int nslots_total = 0;
    double *d_inv_frag20 = nullptr;
    std::vector<uint64_t> slot_ver;
    int eig_src[2] = {-1, -1};
    uint64_t eig_ver[2] = {0, 0};
    int eig_next = 0;
    int walk20_ctas_per_sm = 1;
    // profiling
    bool profiling = false;
    bool partials_timed = false;
    cudaEvent_t ev[6]{};
    float last_ms[5] = {0, 0, 0, 0, 0};
    int last_launch[3] = {0, 0, 0};
    int64_t launches = 0;
    float *d_boot = nullptr, *d_rell = nullptr;   // UFBoot: resampled pattern frequencies [nboot][boot_stride], RELL scores
    double *d_boot_state = nullptr;               // boot_logl | boot_orig_logl | uniform (3 x nboot doubles), then boot_counts, nupdated, updated
    int nboot = 0;
    long long boot_stride = 0;
    size_t boot_bytes = 0;
    int newton_grid = 0;   // co-resident CTAs of k_newton_coop
    int64_t asc_tmp_bytes = 0;
    double *d_lens = nullptr;      // device-resident branch lengths of an asynchronous sweep (slots)
    double *h_sweep = nullptr;     // mapped host table [nlen_slots][3]: (length, evaluations, status) per slot
    double *d_sweep = nullptr;     // its device alias
    int nlen_slots = 0;
    std::vector<b200lk_child> extra_children;   // b200lk_set_extra_children: consumed by the next call that takes ops
    MultiChild *d_mchild = nullptr;
    bool has_multi = false;
    int pending = 0, pending_marks = 0;   // results queued by b200lk_enqueue_*, collected by b200lk_finish
    int last_finish_status = 0;
    struct b200lk_group *group = nullptr;   // the group this engine is a member of (b200lk_group.inc)
    int group_index = -1;
    bool last_derv_nonfinite = false;
    // multi-device engine (b200lk_multi.inc): one shard engine per device over pattern slices [shard_lo[k], shard_lo[k+1])
    std::vector<b200lk_engine *> shards;
    std::vector<long long> shard_lo;
    // non-reversible models (b200lk_nonrev.inc): ceval | cevec | cinv_evec | rate matrices | state frequencies | state indicators;
    // the evaluation branch's matrices and leaf tables; the per-child jobs of the info kernel
    bool nonrev = false, nr_rooted = false;
    double *d_nr_model = nullptr, *d_nr_branch = nullptr;
    NrChildJob *d_nr_jobs = nullptr;
    size_t nr_jobs_cap = 0;
    // site-specific models (b200lk_site.inc): eval | evec | inv_evec of every pattern, warp-interleaved
    bool site = false;
    double *d_site = nullptr;
    unsigned site_attr = 0, xsm_attr = 0;   // cudaFuncSetAttribute done on this engine's device (k_site_walk20, k_walk_n4c4 variant 26)
    // mixed branch lengths (b200lk_mixlen.inc): the [rows][ncat] table, the gradient / Hessian pass's per-block sums
    double *d_mixlen = nullptr, *d_mixlen_red = nullptr;
    size_t mixlen_cap = 0;
    int mixlen_rows = 0;
    std::string err;
};

static thread_local std::string g_create_err;

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t err__ = (call);                                                                   \
        if (err__ != cudaSuccess) {                                                                   \
            e->err = std::string(#call) + ": " + cudaGetErrorString(err__);                           \
            return B200LK_ERR_CUDA;                                                                   \
        }                                                                                             \
    } while (0)

// every stream synchronisation goes through here: the pinned staging ring is free again afterwards
static cudaError_t sync_stream(b200lk_engine *e) {
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    e->stage_used = 0;
    return ce;
}

// Host staging for op lists travels through a RING of pinned segments, so that several calls can be queued without waiting
// for the previous call's host-to-device copy (the asynchronous branch sweep); the ring wraps only after a synchronisation.
static void *stage_acquire(b200lk_engine *e) {
    if (e->stage_used == e->stage_ring) sync_stream(e);
    void *p = (char *)e->h_stage + (size_t)(e->stage_next % e->stage_ring) * e->stage_seg;
    e->stage_next++;
    e->stage_used++;
    return p;
}

static int fail(b200lk_engine *e, int code, const std::string &msg) {
    e->err = msg;
    return code;
}

template <typename T>
static int dev_alloc(b200lk_engine *e, T **ptr, size_t count) {
    cudaError_t err = cudaMalloc((void **)ptr, std::max<size_t>(count, 1) * sizeof(T));
    if (err != cudaSuccess) {
        e->err = std::string("cudaMalloc(") + std::to_string(count * sizeof(T)) + " bytes): " + cudaGetErrorString(err);
        return err == cudaErrorMemoryAllocation ? B200LK_ERR_NOMEM : B200LK_ERR_CUDA;
    }
    e->device_bytes += (int64_t)(count * sizeof(T));
    return 0;
}

#include "b200lk_multi.inc"

extern "C" int b200lk_abi_version(void) { return B200LK_ABI_VERSION; }

extern "C" const char *b200lk_last_error(const b200lk_engine *e) { return e ? e->err.c_str() : g_create_err.c_str(); }

extern "C" int64_t b200lk_device_bytes(const b200lk_engine *e) {
    if (!e->shards.empty()) { int64_t t = 0; for (auto *x : e->shards) t += x->device_bytes; return t; } return e->device_bytes; }
extern "C" int64_t b200lk_launch_count(const b200lk_engine *e) {
    if (!e->shards.empty()) { int64_t t = 0; for (auto *x : e->shards) t += x->launches; return t; } return e->launches; }

static void group_detach(b200lk_engine *e);
extern "C" void b200lk_destroy(b200lk_engine *e) {
    if (e && !e->shards.empty()) { mg_destroy(e); return; }
    if (!e) return;
    if (e->group) group_detach(e);
    cudaSetDevice(e->cfg.device);
    if (e->stream) sync_stream(e);
    cudaFree(e->d_tips); cudaFree(e->d_freq); cudaFree(e->d_invar); cudaFree(e->d_model); cudaFree(e->d_partial);
    cudaFree(e->d_scale); cudaFree(e->d_theta); cudaFree(e->d_buffer_scale); cudaFree(e->d_pattern_lh);
    cudaFree(e->d_pattern_scaling); cudaFree(e->d_pattern_lh_cat); cudaFree(e->d_opinfo); cudaFree(e->d_ops);
    cudaFree(e->d.asc_tmp); cudaFree(e->d_robust); cudaFree(e->d_robust_tmp);
    cudaFree(e->d_post); cudaFree(e->d_em_red); cudaFree(e->d_em_out);
    cudaFree(e->d_jobs); cudaFree(e->d_red); cudaFree(e->d_ticket); cudaFree(e->d_counter); cudaFree(e->d_scratch);
    cudaFree(e->d_nr_model); cudaFree(e->d_nr_branch); cudaFree(e->d_nr_jobs); cudaFree(e->d_site); cudaFree(e->d_mixlen); cudaFree(e->d_mixlen_red);
    cudaFree(e->d_boot); cudaFree(e->d_rell); cudaFree(e->d_boot_state); cudaFree(e->d_mchild); cudaFree(e->d_lens); cudaFree(e->d_inv_frag20);
    if (e->h_sweep) cudaFreeHost(e->h_sweep);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    if (e->h_result) cudaFreeHost(e->h_result);
    for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
}

extern "C" int b200lk_create(const b200lk_config *cfg, b200lk_engine **out) {
    if (cfg && cfg->device < 0) { *out = nullptr; if (cfg->nstates < 1 || cfg->ncat < 1 || cfg->nmix < 1 || cfg->nptn < 1) { g_create_err = "b200lk_create: invalid configuration"; return B200LK_ERR_INVALID; } return mg_create(cfg, out); }
    *out = nullptr;
    if (!cfg || cfg->nstates < 1 || cfg->ncat < 1 || cfg->nmix < 1 || cfg->ntaxa < 2 || cfg->nptn < 1 || cfg->nslots < 1 ||
        cfg->state_unknown < cfg->nstates || cfg->state_unknown > 255) {
        g_create_err = "b200lk_create: invalid configuration";
        return B200LK_ERR_INVALID;
    }
    if (cfg->nstates > 64) {
        g_create_err = "b200lk_create: nstates > 64 not supported by this build";
        return B200LK_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
        g_create_err = std::string("b200lk_create: no usable CUDA device (") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "count=0 or bad ordinal") +
                       "); this engine has no CPU fallback";
        return B200LK_ERR_CUDA;
    }
    b200lk_engine *e = new b200lk_engine();
    e->cfg = *cfg;
    auto bail = [&](int rc) { g_create_err = e->err; b200lk_destroy(e); return rc; };
    if (cudaSetDevice(cfg->device) != cudaSuccess) { e->err = "cudaSetDevice failed"; return bail(B200LK_ERR_CUDA); }
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, cfg->device);
    e->num_sms = prop.multiProcessorCount;
    const int n = cfg->nstates;
    e->C = cfg->fused_mix_rate ? cfg->ncat : cfg->ncat * cfg->nmix;   // phylokernelnew.h:1322-1323
    e->denom = cfg->fused_mix_rate ? 1 : cfg->ncat;                    // :1325
    e->B = n * e->C;
    e->S = cfg->safe_numeric ? e->C : 1;
    e->P = cfg->nptn;
    e->Ppad = (cfg->nptn + kTile - 1) / kTile * kTile;
    int rc;
#define TRY(x) if ((rc = (x)) != 0) return bail(rc)
    if (cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking) != cudaSuccess) { e->err = "cudaStreamCreate failed"; return bail(B200LK_ERR_CUDA); }
    e->stream = e->own_stream;
    for (auto &ev : e->ev) cudaEventCreate(&ev);
    const long long Ppad = e->Ppad;
    // 20-state engines keep state-space CLVs in the fragment-major layout (b200lk_walk20ss.cuh)
    const bool f20 = n == 20 && e->C % cfg->nmix == 0;
    // general state counts above 8 (codon models ...): the same with n padded to a multiple of 8 (b200lk_walkgen.cuh)
    const int n8 = (n + 7) / 8 * 8;
    const bool f8 = n > 8 && n <= 64 && n != 20 && gen_smem_doubles(n8, cfg->state_unknown) * sizeof(double) + 4096 <= 227 * 1024;
    const bool state_space = f20 || f8;
    e->clv_width = f8 ? (long long)e->C * n8 : (long long)e->B;   // doubles per pattern in a CLV slot
    e->nslots_total = cfg->nslots + (state_space ? 2 : 0);
    e->slot_ver.assign(e->nslots_total, 0);
    TRY(dev_alloc(e, &e->d_tips, (size_t)cfg->ntaxa * Ppad));
    TRY(dev_alloc(e, &e->d_freq, Ppad));
    TRY(dev_alloc(e, &e->d_invar, Ppad));
    const size_t n_eval = (size_t)cfg->nmix * n, n_mat = (size_t)cfg->nmix * n * n;
    const size_t n_tip = (size_t)(cfg->state_unknown + 1) * cfg->nmix * n;
    const size_t model_doubles = n_eval + 2 * n_mat + cfg->nmix + 2 * cfg->ncat + n_tip + 3 * (size_t)e->B;
    TRY(dev_alloc(e, &e->d_model, model_doubles));
    TRY(dev_alloc(e, &e->d_partial, (size_t)e->nslots_total * Ppad * e->clv_width));
    if (f8) cudaMemsetAsync(e->d_partial, 0, sizeof(double) * (size_t)e->nslots_total * Ppad * e->clv_width, e->stream);   // padded values stay exact zeros
    TRY(dev_alloc(e, &e->d_scale, (size_t)e->nslots_total * Ppad * e->S));
    if (f20) TRY(dev_alloc(e, &e->d_inv_frag20, (size_t)cfg->nmix * 480));
    TRY(dev_alloc(e, &e->d_theta, (size_t)Ppad * e->B));
    TRY(dev_alloc(e, &e->d_buffer_scale, Ppad));
    TRY(dev_alloc(e, &e->d_pattern_lh, Ppad));
    TRY(dev_alloc(e, &e->d_pattern_scaling, Ppad));
    TRY(dev_alloc(e, &e->d_pattern_lh_cat, (size_t)Ppad * e->C));
    TRY(dev_alloc(e, &e->d_red, (size_t)(Ppad / kTile) * 5));   // per-block partial sums: up to 5 values (+ASC derivative pass); k_newton_coop: two buffers of 2
    TRY(dev_alloc(e, &e->d_ticket, 4));
    TRY(dev_alloc(e, &e->d_counter, 8));
#undef TRY
    if (cudaHostAlloc((void **)&e->h_result, 16 * sizeof(double), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&e->d_result, e->h_result, 0) != cudaSuccess) { e->err = "cudaHostAlloc(mapped) failed"; return bail(B200LK_ERR_CUDA); }
    cudaMemsetAsync(e->d_ticket, 0, 4 * sizeof(unsigned), e->stream);
    e->nlen_slots = 2 * cfg->ntaxa;
    if (dev_alloc(e, &e->d_lens, (size_t)e->nlen_slots) != 0) return bail(B200LK_ERR_NOMEM);
    if (cudaHostAlloc((void **)&e->h_sweep, sizeof(double) * 3 * e->nlen_slots, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&e->d_sweep, e->h_sweep, 0) != cudaSuccess) { e->err = "cudaHostAlloc(mapped) failed"; return bail(B200LK_ERR_CUDA); }
    cudaMemsetAsync(e->d_tips, cfg->state_unknown, (size_t)cfg->ntaxa * Ppad, e->stream);
    cudaMemsetAsync(e->d_freq, 0, sizeof(double) * Ppad, e->stream);
    cudaMemsetAsync(e->d_invar, 0, sizeof(double) * Ppad, e->stream);
    cudaMemsetAsync(e->d_scale, 0, sizeof(uint16_t) * (size_t)e->nslots_total * Ppad * e->S, e->stream);
    cudaMemsetAsync(e->d_model, 0, sizeof(double) * model_doubles, e->stream);
    e->h_freq.assign(cfg->nptn, 0.0);

    Dev &d = e->d;
    d.n = n; d.C = e->C; d.ncat = cfg->ncat; d.nmix = cfg->nmix; d.denom = e->denom; d.B = e->B; d.S = e->S;
    d.su = cfg->state_unknown; d.P = e->P; d.Ppad = Ppad;
    d.asc = 0; d.Pobs = e->P; d.asc_nsite = 0; d.asc_step = 0; d.asc_tmp = nullptr;
    d.clv_f20 = f20 ? 1 : 0;   // 20-state CLVs are fragment-major (b200lk_kernels.cuh: clv20_off)
    d.clv_state = state_space ? 1 : 0;
    d.clv_f8 = f8 ? 1 : 0; d.n8 = n8;
    d.inv_frag20 = e->d_inv_frag20;
    d.ss_lo = 0u; d.ss_hi = 0xffffffffu;   // set_model narrows the band from the eigenvector norms
    d.slot_stride = Ppad * e->clv_width; d.scale_stride = Ppad * e->S;
    d.partial = e->d_partial; d.scale = e->d_scale; d.tips = e->d_tips; d.freq = e->d_freq; d.invar = e->d_invar;
    double *m = e->d_model;
    d.eval = m; m += n_eval;
    d.evec = m; m += n_mat;
    d.inv_evec = m; m += n_mat;
    d.mix_weight = m; m += cfg->nmix;
    d.cat_rate = m; m += cfg->ncat;
    d.cat_prop = m; m += cfg->ncat;
    d.tip_lh = m; m += n_tip;
    d.val = m;
    d.opinfo = nullptr;
    d.dev_len = e->d_lens;
    d.mchild = nullptr;
    d.theta = e->d_theta; d.buffer_scale = e->d_buffer_scale; d.pattern_lh = e->d_pattern_lh;
    d.pattern_scaling = e->d_pattern_scaling; d.pattern_lh_cat = e->d_pattern_lh_cat;

    // k_partial_info keeps exp(eval*r*t) for the whole block in shared memory
    size_t info_smem = sizeof(double) * e->B;
    if (info_smem > 48 * 1024) {
        if (info_smem > 200 * 1024) { e->err = "ncat_mix*nstates too large for k_partial_info"; return bail(B200LK_ERR_INVALID); }
        cudaFuncSetAttribute(k_partial_info, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)info_smem);
    }
    if (sync_stream(e) != cudaSuccess) { e->err = "initialisation failed"; return bail(B200LK_ERR_CUDA); }
    *out = e;
    return 0;
}

extern "C" int b200lk_set_stream(b200lk_engine *e, void *cuda_stream) {
    if (!e->shards.empty()) MG_UNSUPPORTED(e, "set_stream");
    sync_stream(e);
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return 0;
}

extern "C" int b200lk_set_schedule(b200lk_engine *e, int schedule) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_set_schedule(e->shards[k], schedule); });
    if (schedule >= 10 && schedule < 40) {   // development switch: kernel variant of the walker schedule (A/B measurements)
        e->kernel_variant = schedule - 10;
        e->walk20_attr_set = false;
        return 0;
    }
    if (schedule < 0 || schedule > 2) return fail(e, B200LK_ERR_INVALID, "schedule must be 0, 1 or 2");
    e->schedule = schedule;
    return 0;
}

extern "C" int b200lk_set_ascertainment(b200lk_engine *e, int32_t asc_type, int64_t num_unobserved, int64_t nsite) {
    if (!e->shards.empty()) { if (asc_type != 0) MG_UNSUPPORTED(e, "+ASC"); return mg_each(e, [&](int k) { return b200lk_set_ascertainment(e->shards[k], 0, 0, nsite); }); }
    if (asc_type != 0 && e->site) return fail(e, B200LK_ERR_STATE, "+ASC is not available with site-specific models");
    if (asc_type < 0 || asc_type > 2) return fail(e, B200LK_ERR_INVALID, "asc_type: 0 = none, 1 = Lewis (ASC_VARIANT / ASC_INFORMATIVE), 2 = Holder (ASC_VARIANT_MISSING)");
    if (num_unobserved < 0 || num_unobserved >= e->P) return fail(e, B200LK_ERR_INVALID, "num_unobserved out of range");
    if (asc_type && (e->robust_keep < 1.0 || e->robust_median)) return fail(e, B200LK_ERR_INVALID, "+ASC not supported with robust phylo");
    if (asc_type == 2 && (num_unobserved % e->cfg.nstates != 0 || num_unobserved / e->cfg.nstates < e->P - num_unobserved))
        return fail(e, B200LK_ERR_INVALID, "+ASC_MIS: the unobserved block must be nstates groups of at least (observed patterns) constant patterns");
    cudaSetDevice(e->cfg.device);
    CK(sync_stream(e));
    if (e->d.asc_tmp) { cudaFree(e->d.asc_tmp); e->d.asc_tmp = nullptr; e->device_bytes -= e->asc_tmp_bytes; e->asc_tmp_bytes = 0; }
    if (asc_type == 2) {
        CK(cudaMalloc(&e->d.asc_tmp, sizeof(double) * 3 * (size_t)num_unobserved));
        e->asc_tmp_bytes = (int64_t)(sizeof(double) * 3 * (size_t)num_unobserved);
        e->device_bytes += e->asc_tmp_bytes;
    }
    e->d.asc = asc_type;
    e->d.Pobs = asc_type ? e->P - num_unobserved : e->P;
    e->d.asc_step = asc_type == 2 ? num_unobserved / e->cfg.nstates : 0;
    e->d.asc_nsite = nsite;
    e->theta_computed = false;
    return 0;
}

// --robust-phy / --robust-median (Params::robust_phy_keep, robust_median; tree/phylokernelnew.h:3134-3173)
extern "C" int b200lk_set_robust_phylo(b200lk_engine *e, double keep, int32_t median) {
    if (!e->shards.empty()) { if (keep < 1.0 || median) MG_UNSUPPORTED(e, "robust phylo"); return 0; }
    if (!(keep > 0.0 && keep <= 1.0) || (keep < 1.0 && median)) return fail(e, B200LK_ERR_INVALID, "robust phylo: 0 < keep <= 1, and keep < 1 excludes median");
    const bool on = keep < 1.0 || median;
    if (on && e->d.asc) return fail(e, B200LK_ERR_INVALID, "+ASC not supported with robust phylo");   // :3136-3138
    cudaSetDevice(e->cfg.device);
    if (on && !e->d_robust) {
        size_t b1 = 0, b2 = 0;
        double *nul = nullptr;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, nul, nul, nul, nul, (int)e->P, 0, 64, e->stream);
        cub::DeviceScan::InclusiveSum(nullptr, b2, nul, nul, (int)e->P, e->stream);
        e->robust_tmp_bytes = std::max(b1, b2);
        CK(cudaMalloc(&e->d_robust, sizeof(double) * 3 * (size_t)e->P));
        if (cudaMalloc(&e->d_robust_tmp, e->robust_tmp_bytes) != cudaSuccess) {
            cudaFree(e->d_robust); e->d_robust = nullptr;
            return fail(e, B200LK_ERR_NOMEM, "set_robust_phylo: cudaMalloc failed");
        }
        e->device_bytes += (int64_t)(sizeof(double) * 3 * (size_t)e->P + e->robust_tmp_bytes);
    }
    e->robust_keep = keep;
    e->robust_median = median != 0;
    return 0;
}

extern "C" int b200lk_set_extra_children(b200lk_engine *e, const b200lk_child *children, int32_t n) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_set_extra_children(e->shards[k], children, n); });
    if (n < 0 || (n > 0 && !children)) return fail(e, B200LK_ERR_INVALID, "set_extra_children: invalid arguments");
    e->extra_children.assign(children, children + n);
    return 0;
}

extern "C" int b200lk_set_reference_quirks(b200lk_engine *e, int32_t mask) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_set_reference_quirks(e->shards[k], mask); });
    if (mask & ~B200LK_QUIRK_DOTTRIPLE_ODD) return fail(e, B200LK_ERR_INVALID, "unknown quirk bit");
    e->d.quirks = mask;
    return 0;
}

extern "C" int b200lk_set_profiling(b200lk_engine *e, int32_t on) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_set_profiling(e->shards[k], on); }); e->profiling = on != 0; return 0; }

extern "C" int b200lk_last_timing(b200lk_engine *e, float ms[5], int32_t launches[3]) {
    if (!e->shards.empty()) { for (int i = 0; i < 5; i++) ms[i] = 0.f; for (int i = 0; i < 3; i++) launches[i] = 0; for (auto *x : e->shards) { for (int i = 0; i < 5; i++) ms[i] = std::max(ms[i], x->last_ms[i]); for (int i = 0; i < 3; i++) launches[i] += x->last_launch[i]; } return 0; }
    for (int i = 0; i < 5; i++) ms[i] = e->last_ms[i];
    for (int i = 0; i < 3; i++) launches[i] = e->last_launch[i];
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// inputs
// ---------------------------------------------------------------------------------------------------------------
extern "C" int b200lk_set_tip_states(b200lk_engine *e, const uint8_t *states) {
    if (!e->shards.empty()) return mg_set_tip_states(e, states);
    cudaSetDevice(e->cfg.device);
    const size_t total = (size_t)e->cfg.ntaxa * e->P;
    for (size_t k = 0; k < total; k++)
        if (states[k] > e->cfg.state_unknown) return fail(e, B200LK_ERR_INVALID, "tip state > STATE_UNKNOWN");
    CK(cudaMemcpy2DAsync(e->d_tips, e->Ppad, states, e->P, e->P, e->cfg.ntaxa, cudaMemcpyHostToDevice, e->stream));
    CK(sync_stream(e));
    return 0;
}

extern "C" int b200lk_set_pattern_freq(b200lk_engine *e, const double *f) {
    if (!e->shards.empty()) return mg_scatter(e, f, 1, [](b200lk_engine *x, const double *p) { return b200lk_set_pattern_freq(x, p); });
    cudaSetDevice(e->cfg.device);
    e->h_freq.assign(f, f + e->P);
    CK(cudaMemcpyAsync(e->d_freq, f, sizeof(double) * e->P, cudaMemcpyHostToDevice, e->stream));
    CK(sync_stream(e));
    return 0;
}

extern "C" int b200lk_set_pattern_invar(b200lk_engine *e, const double *f) {
    if (!e->shards.empty()) return mg_scatter(e, f, 1, [](b200lk_engine *x, const double *p) { return b200lk_set_pattern_invar(x, p); });
    cudaSetDevice(e->cfg.device);
    CK(cudaMemcpyAsync(e->d_invar, f, sizeof(double) * e->P, cudaMemcpyHostToDevice, e->stream));
    CK(sync_stream(e));
    e->theta_computed = false;
    return 0;
}

extern "C" int b200lk_set_model(b200lk_engine *e, const double *eval, const double *evec, const double *inv_evec,
                                const double *mix_weight) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_set_model(e->shards[k], eval, evec, inv_evec, mix_weight); });
    cudaSetDevice(e->cfg.device);
    const int n = e->cfg.nstates, nmix = e->cfg.nmix;
    CK(cudaMemcpyAsync((void *)e->d.eval, eval, sizeof(double) * nmix * n, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync((void *)e->d.evec, evec, sizeof(double) * nmix * n * n, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync((void *)e->d.inv_evec, inv_evec, sizeof(double) * nmix * n * n, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync((void *)e->d.mix_weight, mix_weight, sizeof(double) * nmix, cudaMemcpyHostToDevice, e->stream));
    e->h_evec.assign(evec, evec + (size_t)nmix * n * n);
    e->h_inv_evec.assign(inv_evec, inv_evec + (size_t)nmix * n * n);
    const int total = (e->cfg.state_unknown + 1) * nmix * n;
    k_tip_partial_lh<<<(total + 127) / 128, 128, 0, e->stream>>>(e->d, (double *)e->d.tip_lh, e->cfg.seq_type);
    e->launches++;
    if (e->d.clv_state) {
        // max|t| / ||evec||_inf <= max|inv_evec . t| <= ||inv_evec||_inf . max|t|  (t = evec . W): outside the band
        // [2^-256 / ||inv_evec||, 2^-256 . ||evec||) the reference's underflow test is decided by max|t| alone; one more
        // binade on either side absorbs the rounding of the norms and of t
        double r_evec = 0.0, r_inv = 0.0;
        for (int m = 0; m < nmix; m++)
            for (int r = 0; r < n; r++) {
                double a = 0.0, b = 0.0;
                for (int k = 0; k < n; k++) { a += std::fabs(evec[((size_t)m * n + r) * n + k]); b += std::fabs(inv_evec[((size_t)m * n + r) * n + k]); }
                r_evec = std::max(r_evec, a); r_inv = std::max(r_inv, b);
            }
        if (!(r_evec > 0.0) || !(r_inv > 0.0) || !std::isfinite(r_evec) || !std::isfinite(r_inv)) return fail(e, B200LK_ERR_INVALID, "set_model: degenerate eigenvectors");
        const int e_hi = (int)std::ceil(std::log2(r_evec)) + 1, e_lo = (int)std::ceil(std::log2(r_inv)) + 1;
        e->d.ss_hi = (unsigned)std::min(2046, std::max(1, 1023 - 256 + e_hi)) << 20;
        e->d.ss_lo = (unsigned)std::min(2046, std::max(0, 1023 - 256 - e_lo)) << 20;
        if (e->d.clv_f20) { k_inv_frag20<<<nmix, 128, 0, e->stream>>>(e->d, e->d_inv_frag20); e->launches++; }
        e->eig_src[0] = e->eig_src[1] = -1;
    }
    CK(cudaGetLastError());
    // pageable host sources: make sure the copies have consumed them before returning
    CK(sync_stream(e));
    e->theta_computed = false;
    return 0;
}

extern "C" int b200lk_set_rates(b200lk_engine *e, const double *cat_rate, const double *cat_prop) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_set_rates(e->shards[k], cat_rate, cat_prop); });
    cudaSetDevice(e->cfg.device);
    CK(cudaMemcpyAsync((void *)e->d.cat_rate, cat_rate, sizeof(double) * e->cfg.ncat, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync((void *)e->d.cat_prop, cat_prop, sizeof(double) * e->cfg.ncat, cudaMemcpyHostToDevice, e->stream));
    CK(sync_stream(e));
    e->theta_computed = false;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// op staging + partial-likelihood launches
// ---------------------------------------------------------------------------------------------------------------
static int ensure_op_capacity(b200lk_engine *e, size_t nops) {
    const size_t e_sz = (size_t)e->B * e->cfg.nstates;
    const size_t leaf_sz = (size_t)(e->cfg.state_unknown + 1) * e->B;
    const size_t need_info = nops * 2 * (e_sz + leaf_sz);
    if (need_info > e->opinfo_cap) {
        CK(sync_stream(e));
        if (e->d_opinfo) { cudaFree(e->d_opinfo); e->device_bytes -= (int64_t)(e->opinfo_cap * sizeof(double)); }
        e->d_opinfo = nullptr; e->opinfo_cap = 0;
        size_t cap = std::max(need_info, (size_t)4 * 2 * (e_sz + leaf_sz));
        int rc = dev_alloc(e, &e->d_opinfo, cap);
        if (rc) return rc;
        e->opinfo_cap = cap;
        e->d.opinfo = e->d_opinfo;
    }
    if (nops > e->ops_cap) {
        CK(sync_stream(e));
        cudaFree(e->d_ops); cudaFree(e->d_jobs); cudaFree(e->d_mchild);
        if (e->h_stage) cudaFreeHost(e->h_stage);
        e->d_ops = nullptr; e->d_jobs = nullptr; e->h_stage = nullptr; e->d_mchild = nullptr;
        e->device_bytes -= (int64_t)(e->ops_cap * (sizeof(DevOp) + 2 * sizeof(ChildJob) + 2 * sizeof(MultiChild)));
        e->ops_cap = 0;   // set again only when all four allocations below have succeeded
        size_t cap = std::max<size_t>(nops, 16);
        int rc = dev_alloc(e, &e->d_ops, cap);
        if (rc) return rc;
        rc = dev_alloc(e, &e->d_jobs, cap * 2);
        if (rc) return rc;
        rc = dev_alloc(e, &e->d_mchild, cap * 2);
        if (rc) return rc;
        e->d.mchild = e->d_mchild;
        e->stage_seg = (cap * (sizeof(DevOp) + 2 * sizeof(ChildJob) + 2 * sizeof(MultiChild)) + 255) / 256 * 256;
        e->h_stage_cap = e->stage_seg * e->stage_ring;
        CK(cudaMallocHost(&e->h_stage, e->h_stage_cap));
        e->stage_next = e->stage_used = 0;
        e->ops_cap = cap;
    }
    return 0;
}

template <bool SAFE>
static void launch_partials(b200lk_engine *e, int op_begin, int op_end, int walk, int grid) {
    const int n = e->cfg.nstates, C = e->C;
    cudaMemsetAsync(e->d_counter, 0, sizeof(unsigned long long), e->stream);
    if (n == 4 && C == 4 && e->cfg.nmix == 1 && e->cfg.state_unknown < kMaxLeafRows4 && !e->has_multi)
        k_partials_n4c4<SAFE><<<grid, kTile, 0, e->stream>>>(e->d, e->d_ops, op_begin, op_end, walk, e->d_counter);
    else if (n == 4)
        k_partials_generic<4, SAFE><<<grid, kTile, 0, e->stream>>>(e->d, e->d_ops, op_begin, op_end, walk, e->d_counter);
    else if (n == 20)
        k_partials_generic<20, SAFE><<<grid, kTile, 0, e->stream>>>(e->d, e->d_ops, op_begin, op_end, walk, e->d_counter);
    else if (n == 2)
        k_partials_generic<2, SAFE><<<grid, kTile, 0, e->stream>>>(e->d, e->d_ops, op_begin, op_end, walk, e->d_counter);
    else
        k_partials_generic<0, SAFE><<<grid, kTile, 0, e->stream>>>(e->d, e->d_ops, op_begin, op_end, walk, e->d_counter);
    e->launches++;
    e->last_launch[0]++;
}


// a child length is a number >= 0, or -(k)-1 = slot k of the device length table (b200lk_enqueue_newton_branch)
static bool length_ok(const b200lk_engine *e, double len) {
    if (e->mixlen_rows) return len >= 0.0 && len < (double)e->mixlen_rows && len == std::floor(len);   // a row of the mixlen table
    if (len >= 0.0) return true;
    const double k = -len - 1.0;
    return k >= 0.0 && k < (double)e->nlen_slots && k == std::floor(k);
}

static int validate_op(b200lk_engine *e, const b200lk_op &op) {
    if (op.dst_slot < 0 || op.dst_slot >= e->cfg.nslots) return fail(e, B200LK_ERR_INVALID, "op: dst_slot out of range");
    const b200lk_child ch[2] = {op.left, op.right};
    for (auto &c : ch) {
        if (c.kind == 0) { if (c.id < 0 || c.id >= e->cfg.ntaxa) return fail(e, B200LK_ERR_INVALID, "op: taxon id out of range"); }
        else if (c.kind == 1) { if (c.id < 0 || c.id >= e->cfg.nslots || c.id == op.dst_slot) return fail(e, B200LK_ERR_INVALID, "op: child slot invalid"); }
        else return fail(e, B200LK_ERR_INVALID, "op: child kind must be 0 or 1");
        if (!length_ok(e, c.length)) return fail(e, B200LK_ERR_INVALID, "op: child length must be >= 0 or -(len_slot)-1 of the device length table");
    }
    return 0;
}

// the specialised 4-state x 4-category walker (k_walk_n4c4) serves the walker schedule of every single-class DNA model
static bool use_walk4(const b200lk_engine *e) {
    return e->schedule != 1 && e->cfg.nstates == 4 && e->C == 4 && e->cfg.nmix == 1 && e->cfg.ncat == 4 &&
           e->cfg.state_unknown < kMaxLeafRows4 && e->h_evec.size() == 16;
}

// ... and k_walk_n4<C> (b200lk_walk4.cuh) every other single-class 4-state model: no rate heterogeneity, +G8, +R2 .. +R10, ...
struct Walk4Cfg { int C; const void *f[2]; };
#define W4CFG(C) { C, { (const void *)k_walk_n4<C, false>, (const void *)k_walk_n4<C, true> } }
static const Walk4Cfg walk4_cfgs[] = { W4CFG(1), W4CFG(2), W4CFG(3), W4CFG(4), W4CFG(5), W4CFG(6), W4CFG(7), W4CFG(8), W4CFG(9), W4CFG(10), W4CFG(12), W4CFG(16) };
static const Walk4Cfg *walk4n_cfg(const b200lk_engine *e) {
    if (e->schedule == 1 || e->cfg.nstates != 4 || e->cfg.nmix != 1 || e->cfg.state_unknown >= kMaxLeafRows4 || e->h_evec.size() != 16) return nullptr;
    if (e->C == 4 && e->kernel_variant != 5) return nullptr;   // four categories: k_walk_n4c4 (variant 5 = the templated kernel, A/B)
    for (const Walk4Cfg &c : walk4_cfgs) if (c.C == e->C) return &c;
    return nullptr;
}

// b200lk_op[] -> WalkJob4[]: child A is the CLV that stays in registers (chained from the previous op when it is its parent)
static int fill_jobs4(b200lk_engine *e, const b200lk_op *ops, int32_t nops, WalkJob4 *hj) {
    int prev_dst = -1;
    for (int k = 0; k < nops; k++) {
        const b200lk_op &op = ops[k];
        int rc = validate_op(e, op);
        if (rc) return rc;
        if (op.reserved != 0) return fail(e, B200LK_ERR_INVALID, "multifurcating ops are not served by the walkers");
        const b200lk_child *A = &op.left, *B = &op.right;
        WalkJob4 &j = hj[k];
        if (A->kind == 0 && B->kind == 0) { j.ncase = 0; j.loadA = 0; }
        else if (A->kind == 1 && B->kind == 1) {
            j.ncase = 2;
            if (B->id == prev_dst) std::swap(A, B);   // chain: the CLV produced by the previous op stays in registers
            j.loadA = A->id != prev_dst;
        } else {
            j.ncase = 1;
            if (A->kind == 0) std::swap(A, B);        // A = the internal child
            j.loadA = A->id != prev_dst;
        }
        j.lenA = A->length; j.lenB = B->length;
        j.aid = A->id; j.bid = B->id;
        j.dst = op.dst_slot; j.pad = 0;
        prev_dst = op.dst_slot;
    }
    return 0;
}

static int enqueue_walk4(b200lk_engine *e, const b200lk_op *ops, int32_t nops, const Walk4Cfg *wcfg) {
    WalkJob4 *hj = (WalkJob4 *)stage_acquire(e);
    int rcj = fill_jobs4(e, ops, nops, hj);
    if (rcj) return rcj;
    WalkJob4 *dj = (WalkJob4 *)e->d_jobs;
    CK(cudaMemcpyAsync(dj, hj, sizeof(WalkJob4) * nops, cudaMemcpyHostToDevice, e->stream));
    Model4 M;
    for (int k = 0; k < 16; k++) { M.evec[k] = e->h_evec[k]; M.inv[k] = e->h_inv_evec[k]; }
    const long long ntiles = e->Ppad / kTile;
    if (wcfg) {   // any number of categories: k_walk_n4<C>
        k_walk_info_n4<<<nops, 128, sizeof(double) * 8 * e->C, e->stream>>>(e->d, dj, e->d_opinfo, e->d_counter);
        const int per_sm = e->C <= 4 ? 4 : e->C <= 8 ? 3 : 2;
        const int grid = (int)std::min<long long>(ntiles, (long long)e->num_sms * per_sm);
        if (e->profiling) cudaEventRecord(e->ev[4], e->stream);
        void *args[] = {(void *)&e->d, (void *)&M, (void *)&e->d_opinfo, (void *)&nops, (void *)&e->d_counter};
        CK(cudaLaunchKernel(wcfg->f[e->cfg.safe_numeric ? 1 : 0], dim3(grid), dim3(kTile), args, 0, e->stream));
        if (e->profiling) cudaEventRecord(e->ev[5], e->stream);
        e->partials_timed = e->profiling;
        e->launches += 2;
        e->last_launch[0] += 2;
        CK(cudaGetLastError());
        e->theta_computed = false;
        return 0;
    }
    k_walk_info_n4c4<<<nops, 64, 0, e->stream>>>(e->d, dj, e->d_opinfo, e->d_counter);
    const int grid = (int)std::min<long long>(ntiles, (long long)e->num_sms * 4);   // 4 resident CTAs per SM (128 regs/thread)
    if (e->profiling) cudaEventRecord(e->ev[4], e->stream);
    {   // kernel_variant (A/B): 0 = per-category direct stores, 1 = TMA bulk stores from shared memory, 2 = one burst per op
        const bool safe = e->cfg.safe_numeric != 0;
        static const void *const kernels[22] = {(const void *)k_walk_n4c4<false, 1>, (const void *)k_walk_n4c4<true, 1>,
                                                (const void *)k_walk_n4c4<false, 2>, (const void *)k_walk_n4c4<true, 2>,
                                                (const void *)k_walk_n4c4<false, 0>, (const void *)k_walk_n4c4<true, 0>,
                                                // variants 20..23 (A/B): L2 prefetches none / tip rows / sibling CLVs / both
                                                (const void *)k_walk_n4c4<false, 1, 0, 0>, (const void *)k_walk_n4c4<true, 1, 0, 0>,
                                                (const void *)k_walk_n4c4<false, 1, 0, 1>, (const void *)k_walk_n4c4<true, 1, 0, 1>,
                                                (const void *)k_walk_n4c4<false, 1, 0, 2>, (const void *)k_walk_n4c4<true, 1, 0, 2>,
                                                (const void *)k_walk_n4c4<false, 1, 0, 3>, (const void *)k_walk_n4c4<true, 1, 0, 3>,
                                                // 24, 25: tip states through shared memory (cp.async), with / without the L2 prefetch of tip rows
                                                (const void *)k_walk_n4c4<false, 1, 0, 5>, (const void *)k_walk_n4c4<true, 1, 0, 5>,
                                                (const void *)k_walk_n4c4<false, 1, 0, 4>, (const void *)k_walk_n4c4<true, 1, 0, 4>,
                                                // 26: + the next op's sibling CLV through shared memory (cp.async one op ahead)
                                                (const void *)k_walk_n4c4<false, 1, 0, 12>, (const void *)k_walk_n4c4<true, 1, 0, 12>,
                                                // 27: the default without the L1 prefetches of the next op's CLV operands
                                                (const void *)k_walk_n4c4<false, 1, 0, 20>, (const void *)k_walk_n4c4<true, 1, 0, 20>};
        int v = (e->kernel_variant >= 0 && e->kernel_variant <= 2) ? e->kernel_variant : (e->kernel_variant >= 20 && e->kernel_variant <= 27) ? e->kernel_variant - 17 : 0;
#ifdef B200LK_DIAG
        static const void *const dbg_kernels[3] = {(const void *)k_walk_n4c4<false, 2, 1>, (const void *)k_walk_n4c4<false, 2, 2>,
                                                   (const void *)k_walk_n4c4<false, 2, 3>};   // variants 6..8: no stores / no loads / neither
        if (e->kernel_variant >= 6 && e->kernel_variant <= 8) {
            void *dargs[] = {(void *)&e->d, (void *)&M, (void *)&e->d_opinfo, (void *)&nops, (void *)&e->d_counter};
            CK(cudaLaunchKernel(dbg_kernels[e->kernel_variant - 6], dim3(grid), dim3(kTile), dargs, 0, e->stream));
            v = -1;
        }
#endif
        void *args[] = {(void *)&e->d, (void *)&M, (void *)&e->d_opinfo, (void *)&nops, (void *)&e->d_counter};
        const size_t xsm = v == 9 ? sizeof(double) * 2 * 16 * kTile : 0;   // variant 26: two sibling-CLV buffers in dynamic shared memory
        if (xsm) {
            if (!(e->xsm_attr & (1u << (safe ? 1 : 0)))) { CK(cudaFuncSetAttribute(kernels[v * 2 + (safe ? 1 : 0)], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsm)); e->xsm_attr |= 1u << (safe ? 1 : 0); }
        }
        if (v >= 0) CK(cudaLaunchKernel(kernels[v * 2 + (safe ? 1 : 0)], dim3(grid), dim3(kTile), args, xsm, e->stream));
    }
    if (e->profiling) cudaEventRecord(e->ev[5], e->stream);
    e->partials_timed = e->profiling;
    e->launches += 2;
    e->last_launch[0] += 2;
    CK(cudaGetLastError());
    e->theta_computed = false;
    return 0;
}

// the 20-state FP64-MMA walker serves the walker schedule of every protein model: any number of categories, mixtures,
// fused classes
static int walk20_cpb(const b200lk_engine *e) { return e->C / e->cfg.nmix; }   // categories per class block
static size_t walk20_stage_doubles(const b200lk_engine *e) {
    return ss_stage_doubles(walk20_cpb(e), e->cfg.state_unknown);
}
static bool use_walk20(const b200lk_engine *e) {
    return e->schedule != 1 && e->cfg.nstates == 20 && e->C % e->cfg.nmix == 0 && e->h_evec.size() == (size_t)400 * e->cfg.nmix &&
           2 * walk20_stage_doubles(e) * sizeof(double) <= 200 * 1024;
}

// every op rewrites its destination slot: the eigen-space copies of state-space CLVs (eigen_slot) are keyed on these counters
static void note_writes(b200lk_engine *e, const b200lk_op *ops, int32_t nops) {
    if (!e->d.clv_state) return;
    for (int k = 0; k < nops; k++) e->slot_ver[ops[k].dst_slot]++;
}

// state-space engines: the slot that holds partial_lh = inv_evec . t of CLV `slot` (one of two scratch slots behind the caller's;
// converted on the stream when the cached copy is stale).  Eigen-space engines: the slot itself.
// the scratch slot for `slot`; *convert = whether its content has to be (re)computed (the caller launches the conversion)
static int eigen_slot_plan(b200lk_engine *e, int slot, int keep, bool *convert) {
    *convert = false;
    if (!e->d.clv_state) return slot;
    for (int k = 0; k < 2; k++)
        if (e->eig_src[k] == slot && e->eig_ver[k] == e->slot_ver[slot]) return e->cfg.nslots + k;
    int k = e->eig_next;
    if (e->cfg.nslots + k == keep) k ^= 1;   // do not evict the other end of the same branch
    e->eig_next = k ^ 1;
    e->eig_src[k] = slot;
    e->eig_ver[k] = e->slot_ver[slot];
    *convert = true;
    return e->cfg.nslots + k;
}
static int eigen_slot(b200lk_engine *e, int slot, int keep = -1) {
    bool convert;
    const int dst = eigen_slot_plan(e, slot, keep, &convert);
    if (convert) {
        if (e->d.clv_f8) {
            const long long total = e->Ppad * e->C;
            k_clv_gen_to_eigen<<<(unsigned)((total + 127) / 128), 128, 0, e->stream>>>(e->d, slot, dst);
        } else {
            const long long nblocks = e->Ppad / 32;
            const int grid = (int)std::min<long long>((nblocks + 3) / 4, (long long)e->num_sms * 8);
            k_clv20_to_eigen<<<grid, 128, 0, e->stream>>>(e->d, slot, dst);
        }
        e->launches++;
    }
    return dst;
}

struct Walk20Cfg { int nw; const void *f[4]; };
// k_walk_n20ss configurations {warps per CTA, kernels [MIX*2 + SAFE]}; index = kernel_variant (b200lk_set_schedule(e, 10 + v));
// and the category-split form [SAFE]
#define W20SS(NW, MINB)                                                                                                        \
    { NW, { (const void *)k_walk_n20ss<false, false, NW, MINB, false>, (const void *)k_walk_n20ss<true, false, NW, MINB, false>, \
            (const void *)k_walk_n20ss<false, true, NW, MINB, false>, (const void *)k_walk_n20ss<true, true, NW, MINB, false> } }
static const Walk20Cfg walk20ss_cfgs[] = {
    W20SS(4, 3),   // 0 (default): 4 warps per CTA, 3 CTAs per SM (168 registers)
    W20SS(4, 2),   // 1: 8 warps per SM, no register cap pressure
    W20SS(6, 2),   // 2: 12 warps per SM in 2 CTAs: a staging block serves 6 pattern blocks
    W20SS(8, 1),   // 3: 8 warps per SM in one CTA
    W20SS(4, 3),   // 4: as 0 without the category-split form for small problems (A/B)
};
constexpr int kNumWalk20ssCfgs = (int)(sizeof(walk20ss_cfgs) / sizeof(walk20ss_cfgs[0]));
static const void *const walk20ss_csp[2] = {(const void *)k_walk_n20ss<false, false, 4, 3, true>, (const void *)k_walk_n20ss<true, false, 4, 3, true>};

// k_walk_n20v2 [staging depth 3 / 2][NW index][CS index][MIX*2 + SAFE]; MIX only with CS = 1
#define W20V2(NW, CS, SD) { (const void *)k_walk_n20v2<false, false, NW, CS, SD>, (const void *)k_walk_n20v2<true, false, NW, CS, SD>, \
                            CS == 1 ? (const void *)k_walk_n20v2<false, true, NW, 1, SD> : nullptr, CS == 1 ? (const void *)k_walk_n20v2<true, true, NW, 1, SD> : nullptr }
static const int walk20v2_nw[3] = {12, 8, 4};
static const int walk20v2_cs[3] = {1, 2, 4};
static const void *const walk20v2_f[2][3][3][4] = {
    {{W20V2(12, 1, 3), W20V2(12, 2, 3), W20V2(12, 4, 3)}, {W20V2(8, 1, 3), W20V2(8, 2, 3), W20V2(8, 4, 3)}, {W20V2(4, 1, 3), W20V2(4, 2, 3), W20V2(4, 4, 3)}},
    {{W20V2(12, 1, 2), W20V2(12, 2, 2), W20V2(12, 4, 2)}, {W20V2(8, 1, 2), W20V2(8, 2, 2), W20V2(8, 4, 2)}, {W20V2(4, 1, 2), W20V2(4, 2, 2), W20V2(4, 4, 2)}}};
constexpr size_t kV2SmemBudget = 227 * 1024 - 7 * 1024;   // dynamic shared memory a v2 CTA may ask for (static: mbarriers, NORM_LH maxima <= 6.2 KB)

// (NW index, staging depth) that fits: prefer 12 warps, then the deeper staging ring
static bool walk20v2_pick(int cpb, int su, int &wi, int &sd, size_t &smem) {
    for (wi = 0; wi < 3; wi++)
        for (sd = 3; sd >= 2; sd--) {
            smem = v2_smem_doubles(cpb, su, walk20v2_nw[wi], sd) * sizeof(double);
            if (smem <= kV2SmemBudget) return true;
        }
    return false;
}

// the second-generation walker (b200lk_walk20v2.cuh): one CTA per SM, TMA rings; returns 1 when this engine / op list cannot use it
static int enqueue_walk20v2(b200lk_engine *e, int32_t nops, const WalkJob20 *dj) {
    const int nblk = e->cfg.nmix, cpb = walk20_cpb(e);
    const bool mix = nblk > 1, safe = e->cfg.safe_numeric != 0;
    int wi, sd;
    size_t smem;
    if (!walk20v2_pick(cpb, e->cfg.state_unknown, wi, sd, smem)) return 1;
    if (e->kernel_variant == 1 && wi < 1) { wi = 1; sd = 3; smem = v2_smem_doubles(cpb, e->cfg.state_unknown, 8, 3) * sizeof(double); }   // A/B and parity coverage: 8 / 4 warps per CTA
    if (e->kernel_variant == 2) { wi = 2; sd = 2; smem = v2_smem_doubles(cpb, e->cfg.state_unknown, 4, 2) * sizeof(double); }                 // ... and the 2-deep staging ring
    if (smem > kV2SmemBudget) return 1;
    const int nw = walk20v2_nw[wi];
    const long long nblocks = e->Ppad / 32, slots = e->num_sms;
    // small problems: split the categories of a block over 2 or 4 warps so that every warp of the machine has work
    int ci = 0;
    if (!mix) {
        while (ci < 2 && walk20v2_cs[ci + 1] <= cpb && nblocks * walk20v2_cs[ci + 1] <= slots * nw) ci++;
        if (e->kernel_variant == 3 && cpb >= 2) ci = 1;   // forced category split (parity coverage), or none
        if (e->kernel_variant == 4) ci = cpb >= 4 ? 2 : 0;
    }
    const int csplit = walk20v2_cs[ci], nb = nw / csplit;
    const void *f = walk20v2_f[3 - sd][wi][ci][(mix ? 2 : 0) + (safe ? 1 : 0)];
    if (!f) return 1;
    if (!(e->walk20v2_attr & (1u << ((3 - sd) * 9 + wi * 3 + ci)))) {
        for (int k = 0; k < 4; k++) {
            const void *fk = walk20v2_f[3 - sd][wi][ci][k];
            if (!fk) continue;
            CK(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaFuncSetAttribute(fk, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        e->walk20v2_attr |= 1u << ((3 - sd) * 9 + wi * 3 + ci);
    }
    const long long waves = (nblocks + nb * slots - 1) / (nb * slots);
    int ntiles = (int)std::min<long long>(waves * slots, nblocks);
    const int grid = (int)std::min<long long>(ntiles, slots);
    k_walk_info_n20ss<<<nops * nblk, 256, sizeof(double) * 2 * cpb * 20, e->stream>>>(e->d, dj, e->d_opinfo, e->d_counter);
    if (e->profiling) cudaEventRecord(e->ev[4], e->stream);
    void *args[] = {(void *)&e->d, (void *)&e->d_opinfo, (void *)&nops, (void *)&e->d_counter, (void *)&ntiles};
    CK(cudaLaunchKernel(f, dim3(grid), dim3(nw * 32), args, smem, e->stream));
    if (e->profiling) cudaEventRecord(e->ev[5], e->stream);
    e->partials_timed = e->profiling;
    e->launches += 2;
    e->last_launch[0] += 2;
    CK(cudaGetLastError());
    e->theta_computed = false;
    return 0;
}

// b200lk_op[] -> WalkJob20[]: a single tip child is A
static int fill_jobs20(b200lk_engine *e, const b200lk_op *ops, int32_t nops, WalkJob20 *hj) {
    for (int k = 0; k < nops; k++) {
        const b200lk_op &op = ops[k];
        int rc = validate_op(e, op);
        if (rc) return rc;
        if (op.reserved != 0) return fail(e, B200LK_ERR_INVALID, "multifurcating ops are not served by the walkers");
        const b200lk_child *A = &op.left, *B = &op.right;
        WalkJob20 &j = hj[k];
        if (A->kind == 0 && B->kind == 0) j.ncase = 0;
        else if (A->kind == 1 && B->kind == 1) j.ncase = 2;
        else { j.ncase = 1; if (A->kind == 1) std::swap(A, B); }   // A = the tip
        j.lenA = A->length; j.lenB = B->length;
        j.aid = A->id; j.bid = B->id;
        j.dst = op.dst_slot;
    }
    return 0;
}

// k_walk_gen [NT8 - 2][SAFE]: 9..16, 17..24 (not 20), ... 57..64 states
#define WGEN(NT8) { (const void *)k_walk_gen<false, NT8>, (const void *)k_walk_gen<true, NT8> }
static const void *const walkgen_f[7][2] = { WGEN(2), WGEN(3), WGEN(4), WGEN(5), WGEN(6), WGEN(7), WGEN(8) };
static bool use_walkgen(const b200lk_engine *e) { return e->schedule != 1 && e->d.clv_f8 && e->h_evec.size() == (size_t)e->cfg.nmix * e->cfg.nstates * e->cfg.nstates; }

static int enqueue_walkgen(b200lk_engine *e, const b200lk_op *ops, int32_t nops) {
    const int n8 = e->d.n8, C = e->C;
    const size_t stage_sz = gen_stage_doubles(n8, e->cfg.state_unknown);   // one staging block per (op, category)
    if (stage_sz * nops * C > e->opinfo_cap) {
        CK(sync_stream(e));
        if (e->d_opinfo) { cudaFree(e->d_opinfo); e->device_bytes -= (int64_t)(e->opinfo_cap * sizeof(double)); }
        e->d_opinfo = nullptr; e->opinfo_cap = 0;
        int rc = dev_alloc(e, &e->d_opinfo, stage_sz * nops * C);
        if (rc) return rc;
        e->opinfo_cap = stage_sz * nops * C;
        e->d.opinfo = e->d_opinfo;
    }
    WalkJob20 *hj = (WalkJob20 *)stage_acquire(e);
    int rcj = fill_jobs20(e, ops, nops, hj);
    if (rcj) return rcj;
    WalkJob20 *dj = (WalkJob20 *)e->d_jobs;
    CK(cudaMemcpyAsync(dj, hj, sizeof(WalkJob20) * nops, cudaMemcpyHostToDevice, e->stream));
    note_writes(e, ops, nops);
    const void *f = walkgen_f[n8 / 8 - 2][e->cfg.safe_numeric ? 1 : 0];
    const size_t smem = gen_smem_doubles(n8, e->cfg.state_unknown) * sizeof(double);
    if (!e->walkgen_attr_set) {
        for (int k = 0; k < 2; k++) {
            CK(cudaFuncSetAttribute(walkgen_f[n8 / 8 - 2][k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaFuncSetAttribute(walkgen_f[n8 / 8 - 2][k], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, kGenNW * 32, smem));
        e->walkgen_ctas_per_sm = std::max(per_sm, 1);
        e->walkgen_attr_set = true;
    }
    k_walk_info_gen<<<nops * C, 256, sizeof(double) * 2 * e->cfg.nstates, e->stream>>>(e->d, dj, e->d_opinfo, e->d_counter);
    const long long nblocks = e->Ppad / 32, slots = (long long)e->num_sms * e->walkgen_ctas_per_sm;
    const long long waves = (nblocks + kGenNW * slots - 1) / (kGenNW * slots);
    int ntiles = (int)std::min<long long>(waves * slots, nblocks);
    const int grid = (int)std::min<long long>(ntiles, slots);
    if (e->profiling) cudaEventRecord(e->ev[4], e->stream);
    void *args[] = {(void *)&e->d, (void *)&e->d_opinfo, (void *)&nops, (void *)&e->d_counter, (void *)&ntiles};
    CK(cudaLaunchKernel(f, dim3(grid), dim3(kGenNW * 32), args, smem, e->stream));
    if (e->profiling) cudaEventRecord(e->ev[5], e->stream);
    e->partials_timed = e->profiling;
    e->launches += 2;
    e->last_launch[0] += 2;
    CK(cudaGetLastError());
    e->theta_computed = false;
    return 0;
}

static int enqueue_walk20(b200lk_engine *e, const b200lk_op *ops, int32_t nops) {
    const int nblk = e->cfg.nmix, cpb = walk20_cpb(e);
    const bool mix = nblk > 1;
    const size_t stage_sz = walk20_stage_doubles(e);   // one staging block per (op, class block)
    if (stage_sz * nops * nblk > e->opinfo_cap) {
        CK(sync_stream(e));
        if (e->d_opinfo) { cudaFree(e->d_opinfo); e->device_bytes -= (int64_t)(e->opinfo_cap * sizeof(double)); }
        e->d_opinfo = nullptr; e->opinfo_cap = 0;
        int rc = dev_alloc(e, &e->d_opinfo, stage_sz * nops * nblk);
        if (rc) return rc;
        e->opinfo_cap = stage_sz * nops * nblk;
        e->d.opinfo = e->d_opinfo;
    }
    static_assert(sizeof(WalkJob20) <= sizeof(DevOp) + 2 * sizeof(ChildJob), "pinned staging buffer too small for WalkJob20");
    WalkJob20 *hj = (WalkJob20 *)stage_acquire(e);
    int rcj = fill_jobs20(e, ops, nops, hj);
    if (rcj) return rcj;
    WalkJob20 *dj = (WalkJob20 *)e->d_jobs;
    CK(cudaMemcpyAsync(dj, hj, sizeof(WalkJob20) * nops, cudaMemcpyHostToDevice, e->stream));
    if (e->kernel_variant < 5) {   // default: the second-generation walker; variants 5.. = k_walk_n20ss configurations (A/B)
        note_writes(e, ops, nops);
        const int r = enqueue_walk20v2(e, nops, dj);
        if (r <= 0) return r;
        for (int k = 0; k < nops; k++) e->slot_ver[ops[k].dst_slot]--;   // not launched: the first-generation kernel below counts the writes
    }
    k_walk_info_n20ss<<<nops * nblk, 256, sizeof(double) * 2 * cpb * 20, e->stream>>>(e->d, dj, e->d_opinfo, e->d_counter);
    note_writes(e, ops, nops);
    const Walk20Cfg &cfg = walk20ss_cfgs[(e->kernel_variant >= 5 && e->kernel_variant - 5 < kNumWalk20ssCfgs) ? e->kernel_variant - 5 : 0];
    const void *const *csp_f = walk20ss_csp;
    const int tile_ptn = cfg.nw * 32;
    const size_t smem = 2 * stage_sz * sizeof(double);   // the walker double-buffers whole staging blocks
    if (!e->walk20_attr_set) {
        for (const void *f : cfg.f) {
            CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // the staging tables must not be what limits residency: ask for the largest shared-memory carve-out
            cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        for (int k = 0; k < 2; k++) {
            CK(cudaFuncSetAttribute(csp_f[k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaFuncSetAttribute(csp_f[k], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        }
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cfg.f[0], tile_ptn, smem));
        e->walk20_ctas_per_sm = std::max(per_sm, 1);
        e->walk20_attr_set = true;
    }
    // balanced tiles (see the kernel): whole waves of the resident CTAs, each tile floor/ceil(nblocks/ntiles) <= nw warp blocks
    const long long nblocks = e->Ppad / 32, slots = (long long)e->num_sms * e->walk20_ctas_per_sm;
    const long long waves = (nblocks + cfg.nw * slots - 1) / (cfg.nw * slots);
    int ntiles = (int)std::min<long long>(waves * slots, nblocks);
    if (e->kernel_variant == 4) ntiles = (int)((nblocks + cfg.nw - 1) / cfg.nw);   // A/B: fixed full tiles
    // small problems: one 32-pattern block per CTA, its categories split over the CTA's warps (k_walk_n20ss<.., CSP>): a
    // traversal's latency drops by min(C, 4); worth it only while every block still gets its own resident CTA
    const bool csp = !mix && e->C >= 2 && e->kernel_variant <= 5 && cfg.nw == 4 && nblocks <= slots;
    if (csp) ntiles = (int)nblocks;
    const int grid = (int)std::min<long long>(ntiles, slots);
    if (e->profiling) cudaEventRecord(e->ev[4], e->stream);
    {
        const bool safe = e->cfg.safe_numeric != 0;
        void *args[] = {(void *)&e->d, (void *)&e->d_opinfo, (void *)&nops, (void *)&e->d_counter, (void *)&ntiles};
        CK(cudaLaunchKernel(csp ? csp_f[safe ? 1 : 0] : cfg.f[(mix ? 2 : 0) + (safe ? 1 : 0)], dim3(grid), dim3(tile_ptn), args, smem, e->stream));
    }
    if (e->profiling) cudaEventRecord(e->ev[5], e->stream);
    e->partials_timed = e->profiling;
    e->launches += 2;
    e->last_launch[0] += 2;
    CK(cudaGetLastError());
    e->theta_computed = false;
    return 0;
}

#include "b200lk_nonrev.inc"
#include "b200lk_site.inc"

// Enqueue (no sync) everything computePartialLikelihood needs for the op list.
static int enqueue_partials(b200lk_engine *e, const b200lk_op *ops, int32_t nops) {
    if (nops <= 0) return 0;
    if (e->nonrev) return nonrev_enqueue_partials(e, ops, nops);
    if (e->site) return site_enqueue_partials(e, ops, nops);
    // multifurcating ops (reserved = number of children beyond left/right, taken in op order from b200lk_set_extra_children)
    int nextra = 0;
    for (int k = 0; k < nops; k++) {
        if (ops[k].reserved < 0) return fail(e, B200LK_ERR_INVALID, "op: negative number of extra children");
        nextra += ops[k].reserved;
    }
    if (nextra > (int)e->extra_children.size()) return fail(e, B200LK_ERR_INVALID, "multifurcating ops: not enough extra children set (b200lk_set_extra_children)");
    const std::vector<b200lk_child> extra(e->extra_children.begin(), e->extra_children.begin() + nextra);
    e->extra_children.clear();
    int rc = ensure_op_capacity(e, (size_t)nops + (nextra + 1) / 2);
    if (rc) return rc;
    if (nextra == 0 && use_walk4(e) && e->kernel_variant != 5) return enqueue_walk4(e, ops, nops, nullptr);
    if (nextra == 0 && walk4n_cfg(e)) return enqueue_walk4(e, ops, nops, walk4n_cfg(e));
    if (nextra == 0 && use_walk20(e)) return enqueue_walk20(e, ops, nops);
    if (nextra == 0 && use_walkgen(e)) return enqueue_walkgen(e, ops, nops);
    const size_t e_sz = (size_t)e->B * e->cfg.nstates;
    const size_t leaf_sz = (size_t)(e->cfg.state_unknown + 1) * e->B;
    char *seg = (char *)stage_acquire(e);
    DevOp *hops = (DevOp *)seg;
    ChildJob *hjobs = (ChildJob *)(seg + e->ops_cap * sizeof(DevOp));
    MultiChild *hmc = (MultiChild *)((char *)hjobs + e->ops_cap * 2 * sizeof(ChildJob));
    int xcur = 0, nmc = 0;   // cursor into `extra`, MultiChild records written
    // level_of_slot: level of the latest writer; last_touch: highest level of any op that has read or written the slot.
    // With recycled slots (LM_MEM_SAVE: one slot serves several nodes of one traversal) an op must also run after
    // every earlier reader/writer of its destination (WAR / WAW), not only after its producers.
    std::vector<int> level_of_slot(e->cfg.nslots, -1), last_touch(e->cfg.nslots, -1);
    std::vector<int> level(nops, 0);
    int max_level = 0;
    bool ordered_levels = true;
    for (int k = 0; k < nops; k++) {
        const b200lk_op &op = ops[k];
        if (op.dst_slot < 0 || op.dst_slot >= e->cfg.nslots) return fail(e, B200LK_ERR_INVALID, "op: dst_slot out of range");
        b200lk_child ch[2] = {op.left, op.right};
        for (auto &c : ch) {
            if (c.kind == 0) { if (c.id < 0 || c.id >= e->cfg.ntaxa) return fail(e, B200LK_ERR_INVALID, "op: taxon id out of range"); }
            else if (c.kind == 1) { if (c.id < 0 || c.id >= e->cfg.nslots || c.id == op.dst_slot) return fail(e, B200LK_ERR_INVALID, "op: child slot invalid"); }
            else return fail(e, B200LK_ERR_INVALID, "op: child kind must be 0 or 1");
            if (!length_ok(e, c.length)) return fail(e, B200LK_ERR_INVALID, "op: child length must be >= 0 or -(len_slot)-1 of the device length table");
        }
        const size_t base = (size_t)k * 2 * (e_sz + leaf_sz);
        int64_t eoff[2] = {(int64_t)base, (int64_t)(base + e_sz)};
        int64_t loff[2] = {(int64_t)(base + 2 * e_sz), (int64_t)(base + 2 * e_sz + leaf_sz)};
        for (int s = 0; s < 2; s++) {
            hjobs[2 * k + s].length = ch[s].length;
            hjobs[2 * k + s].e_off = eoff[s];
            hjobs[2 * k + s].leaf_off = ch[s].kind == 0 ? loff[s] : -1;
        }
        int a = 0, b = 1;   // swap so that a single leaf child is `left` (phylokernelnew.h:1395-1405)
        if (ch[0].kind != 0 && ch[1].kind == 0) { a = 1; b = 0; }
        DevOp &d = hops[k];
        d.dst = op.dst_slot;
        d.nchild = 2; d.pad_ = 0; d.mc_off = 0;
        d.ncase = (ch[a].kind == 0 && ch[b].kind == 0) ? 0 : (ch[a].kind == 0 ? 1 : 2);
        std::vector<b200lk_child> more;
        if (op.reserved > 0) {   // multifurcating node: all children in neighbour order (phylokernelnew.h:1407-1631)
            d.ncase = 3;
            d.nchild = 2 + op.reserved;
            d.mc_off = nmc;
            for (int s = 0; s < 2; s++) hmc[nmc++] = MultiChild{ch[s].kind, ch[s].id, eoff[s], ch[s].kind == 0 ? loff[s] : -1};
            for (int x = 0; x < op.reserved; x++) {
                const b200lk_child &c = extra[xcur];
                if (c.kind == 0) { if (c.id < 0 || c.id >= e->cfg.ntaxa) return fail(e, B200LK_ERR_INVALID, "op: taxon id out of range"); }
                else if (c.kind == 1) { if (c.id < 0 || c.id >= e->cfg.nslots || c.id == op.dst_slot) return fail(e, B200LK_ERR_INVALID, "op: child slot invalid"); }
                else return fail(e, B200LK_ERR_INVALID, "op: child kind must be 0 or 1");
                if (!length_ok(e, c.length)) return fail(e, B200LK_ERR_INVALID, "op: child length must be >= 0 or -(len_slot)-1 of the device length table");
                const int64_t eo = (int64_t)((size_t)nops * 2 * (e_sz + leaf_sz) + (size_t)xcur * (e_sz + leaf_sz));
                ChildJob &j = hjobs[2 * nops + xcur];
                j.length = c.length; j.e_off = eo; j.leaf_off = c.kind == 0 ? eo + (int64_t)e_sz : -1;
                hmc[nmc++] = MultiChild{c.kind, c.id, j.e_off, j.leaf_off};
                more.push_back(c);
                xcur++;
            }
        }
        d.lid = ch[a].id; d.rid = ch[b].id;
        d.eL = eoff[a]; d.eR = eoff[b];
        d.leafL = ch[a].kind == 0 ? loff[a] : -1;
        d.leafR = ch[b].kind == 0 ? loff[b] : -1;
        int lv = 0;
        for (auto &c : ch) if (c.kind == 1 && level_of_slot[c.id] >= 0) lv = std::max(lv, level_of_slot[c.id] + 1);
        for (auto &c : more) if (c.kind == 1 && level_of_slot[c.id] >= 0) lv = std::max(lv, level_of_slot[c.id] + 1);
        lv = std::max(lv, last_touch[op.dst_slot] + 1);
        for (auto &c : ch) if (c.kind == 1) last_touch[c.id] = std::max(last_touch[c.id], lv);
        for (auto &c : more) if (c.kind == 1) last_touch[c.id] = std::max(last_touch[c.id], lv);
        level[k] = lv;
        level_of_slot[op.dst_slot] = lv;
        last_touch[op.dst_slot] = lv;
        max_level = std::max(max_level, lv);
        if (k > 0 && level[k] < level[k - 1]) ordered_levels = false;
    }
    int schedule = e->schedule;
    if (schedule == 0) schedule = 2;
    // the levelled schedule needs the ops grouped by depth: stable-sort a copy of the staged ops by level
    std::vector<int> level_start;
    if (schedule == 1) {
        if (!ordered_levels) {
            std::vector<int> idx(nops);
            for (int k = 0; k < nops; k++) idx[k] = k;
            std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return level[x] < level[y]; });
            std::vector<DevOp> tmp(hops, hops + nops);
            std::vector<int> lv2(nops);
            for (int k = 0; k < nops; k++) { hops[k] = tmp[idx[k]]; lv2[k] = level[idx[k]]; }
            level = lv2;
        }
        for (int k = 0; k < nops; k++) if (k == 0 || level[k] != level[k - 1]) level_start.push_back(k);
        level_start.push_back(nops);
    }
    note_writes(e, ops, nops);
    CK(cudaMemcpyAsync(e->d_ops, hops, sizeof(DevOp) * nops, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->d_jobs, hjobs, sizeof(ChildJob) * (2 * nops + nextra), cudaMemcpyHostToDevice, e->stream));
    if (nmc) CK(cudaMemcpyAsync(e->d_mchild, hmc, sizeof(MultiChild) * nmc, cudaMemcpyHostToDevice, e->stream));
    e->has_multi = nmc > 0;
    k_partial_info<<<2 * nops + nextra, 256, sizeof(double) * e->B, e->stream>>>(e->d, e->d_jobs);
    e->launches++;
    e->last_launch[0]++;
    const long long ntiles = e->Ppad / kTile;
    const int resident = e->num_sms * 4;
    if (e->profiling) cudaEventRecord(e->ev[4], e->stream);
    if (schedule == 2) {
        int grid = (int)std::min<long long>(ntiles, resident);
        if (e->cfg.safe_numeric) launch_partials<true>(e, 0, nops, 1, grid);
        else launch_partials<false>(e, 0, nops, 1, grid);
    } else {
        for (size_t l = 0; l + 1 < level_start.size(); l++) {
            long long items = ntiles * (level_start[l + 1] - level_start[l]);
            int grid = (int)std::min<long long>(items, resident);
            if (e->cfg.safe_numeric) launch_partials<true>(e, level_start[l], level_start[l + 1], 0, grid);
            else launch_partials<false>(e, level_start[l], level_start[l + 1], 0, grid);
        }
    }
    if (e->profiling) cudaEventRecord(e->ev[5], e->stream);
    e->partials_timed = e->profiling;
    CK(cudaGetLastError());
    e->theta_computed = false;   // any partial update invalidates the theta buffer
    return 0;
}

static void prof_begin(b200lk_engine *e) {
    e->last_launch[0] = e->last_launch[1] = e->last_launch[2] = 0;
    e->partials_timed = false;
    if (e->profiling) cudaEventRecord(e->ev[0], e->stream);
}
static void prof_mark(b200lk_engine *e, int i) {
    if (e->profiling) cudaEventRecord(e->ev[i], e->stream);
}
static void prof_end(b200lk_engine *e, int last) {
    if (!e->profiling) return;
    cudaEventSynchronize(e->ev[last]);
    for (int i = 0; i < 3; i++) e->last_ms[i] = 0.f;
    for (int i = 1; i <= last; i++) cudaEventElapsedTime(&e->last_ms[i - 1], e->ev[i - 1], e->ev[i]);
    cudaEventElapsedTime(&e->last_ms[3], e->ev[0], e->ev[last]);
    e->last_ms[4] = 0.f;
    if (e->partials_timed) cudaEventElapsedTime(&e->last_ms[4], e->ev[4], e->ev[5]);
}

extern "C" int b200lk_update_partials(b200lk_engine *e, const b200lk_op *ops, int32_t nops) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_update_partials(e->shards[k], ops, nops); });
    cudaSetDevice(e->cfg.device);
    prof_begin(e);
    int rc = enqueue_partials(e, ops, nops);
    if (rc) return rc;
    prof_mark(e, 1);
    CK(sync_stream(e));
    prof_end(e, 1);
    return 0;
}

static int check_branch(b200lk_engine *e, int node_slot, int dad_kind, int dad_id) {
    if (node_slot < 0 || node_slot >= e->cfg.nslots) return fail(e, B200LK_ERR_INVALID, "node_slot out of range");
    if (dad_kind == 0) { if (dad_id < 0 || dad_id >= e->cfg.ntaxa) return fail(e, B200LK_ERR_INVALID, "dad taxon out of range"); }
    else if (dad_kind == 1) { if (dad_id < 0 || dad_id >= e->cfg.nslots || dad_id == node_slot) return fail(e, B200LK_ERR_INVALID, "dad slot invalid"); }
    else if (!(e->nonrev && dad_kind == kNrKindRoot)) return fail(e, B200LK_ERR_INVALID, "dad_kind must be 0 or 1 (2 = the root leaf with the non-reversible kernel)");
    return 0;
}

// NaN repair of the reference (phylokernelnew.h:3115-3132, 3374-3383): rare, done on the host copy.
static int repair_pattern_lh(b200lk_engine *e, bool save_log, double *logl) {
    std::vector<double> lh(e->P), sc;
    CK(cudaMemcpy(lh.data(), e->d_pattern_lh, sizeof(double) * e->P, cudaMemcpyDeviceToHost));
    if (!save_log) {
        sc.resize(e->P);
        CK(cudaMemcpy(sc.data(), e->d_pattern_scaling, sizeof(double) * e->P, cudaMemcpyDeviceToHost));
    }
    double tree_lh = 0.0;
    for (long long p = 0; p < e->P; p++) {
        // the reference tests and sums _pattern_lh[ptn] as it stands (:3119-3130) — with save_log_value = false that is the raw
        // likelihood, not its logarithm; reproduced to the letter
        double ptn_lh = lh[p];
        if (!std::isfinite(ptn_lh)) {
            if (save_log) lh[p] = kLogScalingThreshold * 4;
            else { lh[p] = 1.0; sc[p] = kLogScalingThreshold * 4; }
            ptn_lh = kLogScalingThreshold * 4;
        }
        tree_lh += ptn_lh * e->h_freq[p];
    }
    CK(cudaMemcpy(e->d_pattern_lh, lh.data(), sizeof(double) * e->P, cudaMemcpyHostToDevice));
    if (!save_log) CK(cudaMemcpy(e->d_pattern_scaling, sc.data(), sizeof(double) * e->P, cudaMemcpyHostToDevice));
    *logl = tree_lh;
    return B200LK_WARN_UNDERFLOW;
}

// +ASC epilogue of the two log-likelihood entry points: raw (tree_lh, prob_const) in d_result[8..9] -> d_result[0], _pattern_lh
static void asc_finish_logl(b200lk_engine *e, int nblocks) {
    if (e->d.asc == 2) k_asc_holder_logl<<<nblocks, kTile, 0, e->stream>>>(e->d, e->d_result + 8, e->d_red, e->d_ticket, e->d_result);
    else k_asc_finish_logl<<<std::min(nblocks, 296), 256, 0, e->stream>>>(e->d, e->d_result + 8, e->d_result);
}

// pending results of the enqueue_* entry points (b200lk_finish collects them)
enum { PEND_LOGL = 1, PEND_DERV = 2, PEND_SAVE_LOG = 4 };

extern "C" int b200lk_enqueue_branch_logl(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                                          int32_t dad_id, double length, int32_t save_log_value) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_enqueue_branch_logl(e->shards[k], ops, nops, node_slot, dad_kind, dad_id, length, save_log_value); });
    cudaSetDevice(e->cfg.device);
    int rc = check_branch(e, node_slot, dad_kind, dad_id);
    if (rc) return rc;
    if (e->mixlen_rows && !length_ok(e, length)) return fail(e, B200LK_ERR_INVALID, "length must be a row of the mixlen table");
    prof_begin(e);
    rc = enqueue_partials(e, ops, nops);
    if (rc) return rc;
    prof_mark(e, 1);
    if (e->nonrev) {   // computeNonrevLikelihoodBranch (tree/phylokernelnonrev.h:1062-1528); save_log_value = false has no form there
        if (!save_log_value) return fail(e, B200LK_ERR_STATE, "non-reversible kernel: save_log_value = false is not available");
        rc = nonrev_enqueue_root(e, node_slot, dad_kind, dad_id, length);
        if (rc) return rc;
        prof_mark(e, 2);
        CK(cudaGetLastError());
        e->have_lh_cat = true;
        e->pending = (e->pending & ~PEND_SAVE_LOG) | PEND_LOGL | PEND_SAVE_LOG;
        e->pending_marks = 2;
        return 0;
    }
    if (e->site) {   // SITE_MODEL: the pattern's own eigenvalues (tree/phylokernelnew.h:2833-2848, 2993-3008)
        if (e->robust_keep < 1.0 || e->robust_median) return fail(e, B200LK_ERR_STATE, "site-specific models: robust phylo is not available");
        site_enqueue_root(e, node_slot, dad_kind, dad_id, length, save_log_value);
        prof_mark(e, 2);
        CK(cudaGetLastError());
        e->have_lh_cat = true;
        e->pending = (e->pending & ~PEND_SAVE_LOG) | PEND_LOGL | (save_log_value ? PEND_SAVE_LOG : 0);
        e->pending_marks = 2;
        return 0;
    }
    const int nblocks = (int)(e->Ppad / kTile);
    const int fuse = e->B <= kFuseValsMaxB;
    const size_t vsm = fuse ? sizeof(double) * e->B : 0;
    if (!fuse) { k_vals<<<1, 256, 0, e->stream>>>(e->d, length, 0); e->launches++; e->last_launch[1]++; }
    if (e->d.clv_state) {   // the root stage reads partial_lh = inv_evec . t
        node_slot = eigen_slot(e, node_slot);
        if (dad_kind == 1) dad_id = eigen_slot(e, dad_id, node_slot);
    }
    if (e->d.asc) {   // +ASC: (tree_lh, prob_const) into slots 8..9, then the correction
        if (e->cfg.safe_numeric) k_root<true, true><<<nblocks, kTile, vsm, e->stream>>>(e->d, node_slot, dad_kind, dad_id, save_log_value, e->d_red, e->d_ticket, e->d_result + 8, length, fuse);
        else k_root<false, true><<<nblocks, kTile, vsm, e->stream>>>(e->d, node_slot, dad_kind, dad_id, save_log_value, e->d_red, e->d_ticket, e->d_result + 8, length, fuse);
        asc_finish_logl(e, nblocks);
        e->launches += 1;
        e->last_launch[1] += 1;
    } else if (e->cfg.safe_numeric) k_root<true><<<nblocks, kTile, vsm, e->stream>>>(e->d, node_slot, dad_kind, dad_id, save_log_value, e->d_red, e->d_ticket, e->d_result, length, fuse);
    else k_root<false><<<nblocks, kTile, vsm, e->stream>>>(e->d, node_slot, dad_kind, dad_id, save_log_value, e->d_red, e->d_ticket, e->d_result, length, fuse);
    e->launches += 1;
    e->last_launch[1] += 1;
    if (e->robust_keep < 1.0 || e->robust_median) {
        // the site log-likelihoods in ascending order without expanding them: patterns sorted by _pattern_lh, the running sum of
        // their frequencies gives every pattern its range of site ranks; k_robust_phylo weighs each pattern by the part of its
        // range that is kept (or that holds the median).  Library sort and scan on a per-call side path, as in b200lk_patterns.cu.
        double *keys = e->d_robust, *vals = keys + e->P, *cum = vals + e->P;
        size_t bytes = e->robust_tmp_bytes;
        cub::DeviceRadixSort::SortPairs(e->d_robust_tmp, bytes, e->d_pattern_lh, keys, e->d_freq, vals, (int)e->P, 0, 64, e->stream);
        bytes = e->robust_tmp_bytes;
        cub::DeviceScan::InclusiveSum(e->d_robust_tmp, bytes, vals, cum, (int)e->P, e->stream);
        k_robust_phylo<<<nblocks, kTile, 0, e->stream>>>(e->P, keys, vals, cum, e->robust_keep, e->robust_median, e->d_red, e->d_ticket, e->d_result);
        e->launches += 3;
        e->last_launch[1] += 3;
    }
    prof_mark(e, 2);
    CK(cudaGetLastError());
    e->have_lh_cat = true;
    e->pending = (e->pending & ~PEND_SAVE_LOG) | PEND_LOGL | (save_log_value ? PEND_SAVE_LOG : 0);
    e->pending_marks = 2;
    return 0;
}

extern "C" int b200lk_enqueue_results_dev(b200lk_engine *e, double *dev_out) {
    if (!e->shards.empty()) MG_UNSUPPORTED(e, "enqueue_results_dev");
    if (!dev_out) return fail(e, B200LK_ERR_INVALID, "enqueue_results_dev: null pointer");
    cudaSetDevice(e->cfg.device);
    // the final reductions write the scalars through the mapped alias of h_result; stream order makes them visible to this copy
    CK(cudaMemcpyAsync(dev_out, e->h_result, 3 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    return 0;
}

static int finish_collect(b200lk_engine *e, double *logl, double *df, double *ddf);
extern "C" int b200lk_finish(b200lk_engine *e, double *logl, double *df, double *ddf) {
    if (!e->shards.empty()) return mg_finish(e, logl, df, ddf);
    cudaSetDevice(e->cfg.device);
    CK(sync_stream(e));
    return finish_collect(e, logl, df, ddf);
}
// the stream has been synchronised: hand out what was queued (also used by b200lk_group_finish after ONE synchronisation)
static int finish_collect(b200lk_engine *e, double *logl, double *df, double *ddf) {
    e->stage_used = 0;
    if (e->pending_marks) prof_end(e, e->pending_marks);
    e->pending_marks = 0;
    const int pending = e->pending;
    e->pending = 0;
    e->last_derv_nonfinite = false;
    int status = 0;
    if (pending & PEND_DERV) {
        double f = e->h_result[1], g = e->h_result[2];
        e->last_derv_nonfinite = !std::isfinite(f);
        if (!std::isfinite(f)) {   // phylokernelnew.h:2629-2632 (the NORM_LH outError at :2577 is the caller's policy)
            f = g = 0.0;
            status = B200LK_WARN_DERV_UNDERFLOW;
        }
        if (df) *df = f;
        if (ddf) *ddf = g;
    }
    if (pending & PEND_LOGL) {
        double l = e->h_result[0];
        if (!std::isfinite(l)) {
            int rc = repair_pattern_lh(e, (pending & PEND_SAVE_LOG) != 0, &l);
            if (rc < 0) return rc;
            status = rc;
        }
        if (logl) *logl = l;
    }
    e->last_finish_status = status;
    return status;
}

extern "C" int b200lk_branch_logl(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                                  int32_t dad_id, double length, int32_t save_log_value, double *logl) {
    int rc = b200lk_enqueue_branch_logl(e, ops, nops, node_slot, dad_kind, dad_id, length, save_log_value);
    if (rc) return rc;
    return b200lk_finish(e, logl, nullptr, nullptr);
}

static int enqueue_theta(b200lk_engine *e, int node_slot, int dad_kind, int dad_id) {
    if (e->site) { site_enqueue_theta(e, node_slot, dad_kind, dad_id); return 0; }
    const int nblocks = (int)(e->Ppad / kTile);
    if (e->d.clv_state) {
        node_slot = eigen_slot(e, node_slot);
        if (dad_kind == 1) dad_id = eigen_slot(e, dad_id, node_slot);
    }
    if (e->cfg.safe_numeric) k_theta<true><<<nblocks, kTile, 0, e->stream>>>(e->d, node_slot, dad_kind, dad_id);
    else k_theta<false><<<nblocks, kTile, 0, e->stream>>>(e->d, node_slot, dad_kind, dad_id);
    e->launches++;
    e->last_launch[2]++;
    e->theta_computed = true;
    return 0;
}

static void enqueue_derv(b200lk_engine *e, double length) {
    if (e->site) { site_enqueue_derv(e, length); return; }
    const int nblocks = (int)(e->Ppad / kTile);
    const int fuse = e->B <= kFuseValsMaxB;
    const size_t vsm = fuse ? sizeof(double) * 3 * e->B : 0;
    if (!fuse) { k_vals<<<1, 256, 0, e->stream>>>(e->d, length, 1); e->launches++; e->last_launch[2]++; }
    if (e->d.asc) {   // +ASC: (df, ddf, prob_const, df_const, ddf_const) into slots 10..14, then the correction
        k_derv<0, true><<<nblocks, kTile, vsm, e->stream>>>(e->d, e->d_red, e->d_ticket, e->d_result + 10, length, fuse);
        if (e->d.asc == 2) k_asc_holder_derv<<<nblocks, kTile, 0, e->stream>>>(e->d, e->d_result + 10, e->d_red, e->d_ticket, e->d_result + 1);
        else k_asc_finish_derv<<<1, 1, 0, e->stream>>>(e->d, e->d_result + 10, e->d_result + 1);
        e->launches += 1;
        e->last_launch[2] += 1;
    } else if (e->B == 16) k_derv<16><<<nblocks, kTile, vsm, e->stream>>>(e->d, e->d_red, e->d_ticket, e->d_result + 1, length, fuse);
    else if (e->B == 80) k_derv<80><<<nblocks, kTile, vsm, e->stream>>>(e->d, e->d_red, e->d_ticket, e->d_result + 1, length, fuse);
    else k_derv<0><<<nblocks, kTile, vsm, e->stream>>>(e->d, e->d_red, e->d_ticket, e->d_result + 1, length, fuse);
    e->launches += 1;
    e->last_launch[2] += 1;
}

extern "C" int b200lk_theta_reset(b200lk_engine *e) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_theta_reset(e->shards[k]); }); e->theta_computed = false; return 0; }

extern "C" int b200lk_enqueue_branch_derv(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                                          int32_t dad_id, double length) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_enqueue_branch_derv(e->shards[k], ops, nops, node_slot, dad_kind, dad_id, length); });
    cudaSetDevice(e->cfg.device);
    int rc = check_branch(e, node_slot, dad_kind, dad_id);
    if (rc) return rc;
    if (e->mixlen_rows) return fail(e, B200LK_ERR_STATE, "branch_derv: with mixed branch lengths computeLikelihoodDerv returns a gradient and a Hessian (b200lk_branch_derv_mixlen)");
    const bool chained = (e->pending & PEND_LOGL) != 0;   // queued behind a logL call: keep its profiling marks
    if (!chained) prof_begin(e);
    if (nops > 0) e->theta_computed = false;
    rc = enqueue_partials(e, ops, nops);
    if (rc) return rc;
    if (!chained) { prof_mark(e, 1); prof_mark(e, 2); }
    if (e->nonrev) {   // computeNonrevLikelihoodDerv (tree/phylokernelnonrev.h:586-1056): both ends' CLVs, no theta buffer
        rc = nonrev_enqueue_derv(e, node_slot, dad_kind, dad_id, length);
        if (rc) return rc;
    } else {
    if (!e->theta_computed) enqueue_theta(e, node_slot, dad_kind, dad_id);
    enqueue_derv(e, length);
    }
    prof_mark(e, 3);
    CK(cudaGetLastError());
    e->pending |= PEND_DERV;
    e->pending_marks = 3;
    return 0;
}

extern "C" int b200lk_branch_derv(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                                  int32_t dad_id, double length, double *df, double *ddf) {
    int rc = b200lk_enqueue_branch_derv(e, ops, nops, node_slot, dad_kind, dad_id, length);
    if (rc) return rc;
    return b200lk_finish(e, nullptr, df, ddf);
}

extern "C" int b200lk_enqueue_logl_from_buffer(b200lk_engine *e, double length) {
    if (!e->shards.empty()) return mg_each(e, [&](int k) { return b200lk_enqueue_logl_from_buffer(e->shards[k], length); });
    cudaSetDevice(e->cfg.device);
    if (e->nonrev) return fail(e, B200LK_ERR_STATE, "logl_from_buffer: the non-reversible kernel has no theta buffer (the reference leaves computeLikelihoodFromBufferPointer NULL)");
    if (!e->theta_computed) return fail(e, B200LK_ERR_STATE, "logl_from_buffer: theta buffer not computed");
    if (e->mixlen_rows && !length_ok(e, length)) return fail(e, B200LK_ERR_INVALID, "length must be a row of the mixlen table");
    prof_begin(e);
    prof_mark(e, 1);
    if (e->site) {
        site_enqueue_from_buffer(e, length);
        prof_mark(e, 2);
        CK(cudaGetLastError());
        e->pending = (e->pending & PEND_DERV) | PEND_LOGL | PEND_SAVE_LOG;
        e->pending_marks = 2;
        return 0;
    }
    const int nblocks = (int)(e->Ppad / kTile);
    const int fuse = e->B <= kFuseValsMaxB;
    const size_t vsm = fuse ? sizeof(double) * e->B : 0;
    if (!fuse) { k_vals<<<1, 256, 0, e->stream>>>(e->d, length, 0); e->launches++; e->last_launch[1]++; }
    if (e->d.asc) {
        k_from_buffer<0, true><<<nblocks, kTile, vsm, e->stream>>>(e->d, e->d_red, e->d_ticket, e->d_result + 8, length, fuse);
        asc_finish_logl(e, nblocks);
        e->launches += 1;
        e->last_launch[1] += 1;
    } else if (e->B == 16) k_from_buffer<16><<<nblocks, kTile, vsm, e->stream>>>(e->d, e->d_red, e->d_ticket, e->d_result, length, fuse);
    else if (e->B == 80) k_from_buffer<80><<<nblocks, kTile, vsm, e->stream>>>(e->d, e->d_red, e->d_ticket, e->d_result, length, fuse);
    else k_from_buffer<0><<<nblocks, kTile, vsm, e->stream>>>(e->d, e->d_red, e->d_ticket, e->d_result, length, fuse);
    e->launches += 1;
    e->last_launch[1] += 1;
    prof_mark(e, 2);
    CK(cudaGetLastError());
    e->pending = (e->pending & PEND_DERV) | PEND_LOGL | PEND_SAVE_LOG;   // a queued derivative result stays collectable
    e->pending_marks = 2;
    return 0;
}

extern "C" int b200lk_logl_from_buffer(b200lk_engine *e, double length, double *logl) {
    int rc = b200lk_enqueue_logl_from_buffer(e, length);
    if (rc) return rc;
    return b200lk_finish(e, logl, nullptr, nullptr);
}

static bool newton_coop_ok(const b200lk_engine *e) { return e->B <= kFuseValsMaxB && e->kernel_variant != 9 && !e->d.asc && !e->site && !e->mixlen_rows; }

static int launch_newton_coop(b200lk_engine *e, double x1, double xguess, double x2, double xacc, int maxNRStep, double *io,
                              double *len_out, int check_divergence) {
    const void *kern = e->B == 16 ? (const void *)k_newton_coop<16> : e->B == 80 ? (const void *)k_newton_coop<80> : (const void *)k_newton_coop<0>;
    const size_t vsm = sizeof(double) * 3 * e->B;
    if (e->newton_grid == 0) {
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTile, vsm));
        e->newton_grid = std::max(1, per_sm) * e->num_sms;
    }
    const int grid = std::min((int)(e->Ppad / kTile), e->newton_grid);   // d_red holds 5 doubles per tile: room for 2 x 2 x grid
    void *args[] = {(void *)&e->d, (void *)&e->d_red, (void *)&io, (void *)&x1, (void *)&xguess, (void *)&x2, (void *)&xacc,
                    (void *)&maxNRStep, (void *)&len_out, (void *)&check_divergence};
    CK(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kTile), args, vsm, e->stream));
    e->launches++;
    return 0;
}

// The asynchronous form of b200lk_newton_branch for a whole optimizeAllBranches sweep: nothing waits.  The optimised length
// goes to slot `len_slot` of the engine's device length table, where later ops of the same sweep pick it up (a child length
// of -(len_slot)-1 in a b200lk_op), and — with the number of derivative evaluations and the status — to a host-mapped table
// that b200lk_sweep_finish hands out after one synchronisation.  The Newton-Raphson divergence test of optimizeOneBranch
// (tree/phylotree.cpp:2616-2625) runs on the device too.  Returns B200LK_ERR_STATE when this engine's configuration has no
// device-resident Newton iteration (block > 512 doubles, +ASC): the caller then uses the synchronous call.
extern "C" int b200lk_enqueue_newton_branch(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                                            int32_t dad_id, double x1, double xguess, double x2, double xacc, int32_t maxNRStep,
                                            int32_t len_slot) {
    if (!e->shards.empty()) MG_UNSUPPORTED(e, "enqueue_newton_branch");
    cudaSetDevice(e->cfg.device);
    if (e->nonrev || !newton_coop_ok(e)) return fail(e, B200LK_ERR_STATE, "enqueue_newton_branch: no device-resident Newton iteration for this configuration");
    if (len_slot < 0 || len_slot >= e->nlen_slots) return fail(e, B200LK_ERR_INVALID, "len_slot out of range");
    int rc = check_branch(e, node_slot, dad_kind, dad_id);
    if (rc) return rc;
    e->theta_computed = false;
    rc = enqueue_partials(e, ops, nops);
    if (rc) return rc;
    enqueue_theta(e, node_slot, dad_kind, dad_id);
    return launch_newton_coop(e, x1, xguess, x2, xacc, maxNRStep, e->d_sweep + 3 * (size_t)len_slot, e->d_lens + len_slot, 1);
}

extern "C" int b200lk_sweep_finish(b200lk_engine *e, int32_t nslots, double *lengths, int32_t *evaluations, int32_t *status) {
    if (!e->shards.empty()) MG_UNSUPPORTED(e, "sweep_finish");
    cudaSetDevice(e->cfg.device);
    if (nslots < 0 || nslots > e->nlen_slots) return fail(e, B200LK_ERR_INVALID, "nslots out of range");
    CK(sync_stream(e));
    for (int k = 0; k < nslots; k++) {
        if (lengths) lengths[k] = e->h_sweep[3 * k];
        if (evaluations) evaluations[k] = (int)e->h_sweep[3 * k + 1];
        if (status) status[k] = e->h_sweep[3 * k + 2] != 0.0 ? B200LK_WARN_DERV_UNDERFLOW : 0;
    }
    return 0;
}

// The host-driven form of the iteration (engines without a device-resident Newton loop; multi-device engines, where every step
// sums the shards' derivatives): Optimization::minimizeNewton, utils/optimization.cpp:422-499, on f = -df, f' = -ddf.
// func_derv(x, f, df) returns < 0 on error.
template <typename F>
static int host_newton(F func_derv, double x1, double xguess, double x2, double xacc, int maxNRStep, double *length_out) {
    int rc;
    double df, dx, f, temp, xh, xl, rts, rts_old;
    rts = xguess;
    if (rts < x1) rts = x1;
    if (rts > x2) rts = x2;
    auto done = [&](double x) { *length_out = x; return 0; };
    if ((rc = func_derv(rts, f, df)) < 0) return rc;
    if (!std::isfinite(f) || !std::isfinite(df)) { *length_out = rts; return B200LK_WARN_DERV_UNDERFLOW; }
    if (df >= 0.0 && std::fabs(f) < xacc) return done(rts);
    if (f < 0.0) { xl = rts; xh = x2; } else { xh = rts; xl = x1; }
    for (int j = 1; j <= maxNRStep; j++) {
        rts_old = rts;
        if ((df <= 0.0) || (((rts - xh) * df - f) * ((rts - xl) * df - f) >= 0.0)) {
            dx = 0.5 * (xh - xl);
            rts = xl + dx;
            if (xl == rts) return done(rts);
        } else {
            dx = f / df;
            temp = rts;
            rts -= dx;
            if (temp == rts) return done(rts);
        }
        if (std::fabs(dx) < xacc || (j == maxNRStep)) return done(rts_old);
        if ((rc = func_derv(rts, f, df)) < 0) return rc;
        if (!std::isfinite(f) || !std::isfinite(df)) { *length_out = rts; return B200LK_WARN_DERV_UNDERFLOW; }
        if (df > 0.0 && std::fabs(f) < xacc) return done(rts);
        if (f < 0.0) xl = rts;
        else if (f > 0.0) xh = rts;
    }
    return done(rts);
}

// Optimization::minimizeNewton (utils/optimization.cpp:422-499) with f = -df, f' = -ddf (computeFuncDerv,
// tree/phylotree.cpp:2529-2536), evaluated on the resident theta buffer.
extern "C" int b200lk_newton_branch(b200lk_engine *e, const b200lk_op *ops, int32_t nops, int32_t node_slot, int32_t dad_kind,
                                    int32_t dad_id, double x1, double xguess, double x2, double xacc, int32_t maxNRStep,
                                    double *length_out, int32_t *steps_out) {
    if (!e->shards.empty()) {   // every Newton step sums the shards' (df, ddf): the shards walk the same iterates
        int nsteps = 0;
        bool nonfinite = false;
        auto fd = [&](double x, double &f, double &df) -> int {
            double a = 0.0, b = 0.0;
            const int r = b200lk_branch_derv(e, nsteps == 0 ? ops : nullptr, nsteps == 0 ? nops : 0, node_slot, dad_kind, dad_id, x, &a, &b);
            if (r < 0) return r;
            if (r == B200LK_WARN_DERV_UNDERFLOW) { nonfinite = true; a = b = std::nan(""); }
            f = -a; df = -b;
            nsteps++;
            return 0;
        };
        if (nops > 0) b200lk_theta_reset(e);
        int r = host_newton(fd, x1, xguess, x2, xacc, maxNRStep, length_out);
        if (steps_out) *steps_out = nsteps;
        (void)nonfinite;
        return r;
    }
    cudaSetDevice(e->cfg.device);
    int rc = check_branch(e, node_slot, dad_kind, dad_id);
    if (rc) return rc;
    if (nops > 0) e->theta_computed = false;
    rc = enqueue_partials(e, ops, nops);
    if (rc) return rc;
    if (e->nonrev) {   // host-driven iteration over computeNonrevLikelihoodDerv: one matrices kernel + one pattern pass per step
        int nsteps = 0;
        auto fd = [&](double x, double &f, double &df) -> int {
            int r = nonrev_enqueue_derv(e, node_slot, dad_kind, dad_id, x);
            if (r) return r;
            cudaError_t ce = sync_stream(e);
            if (ce != cudaSuccess) { e->err = cudaGetErrorString(ce); return B200LK_ERR_CUDA; }
            f = -e->h_result[1];
            df = -e->h_result[2];
            nsteps++;
            return 0;
        };
        rc = host_newton(fd, x1, xguess, x2, xacc, maxNRStep, length_out);
        if (steps_out) *steps_out = nsteps;
        return rc;
    }
    if (!e->theta_computed) enqueue_theta(e, node_slot, dad_kind, dad_id);
    if (newton_coop_ok(e)) {
        // the whole iteration in one cooperative launch (k_newton_coop); variant 9 keeps the host-driven loop below for A/B
        double *io = e->d_result + 4;   // mapped host memory; slots 4..6 = (length, evaluations, status)
        rc = launch_newton_coop(e, x1, xguess, x2, xacc, maxNRStep, io, nullptr, 0);
        if (rc) return rc;
        CK(sync_stream(e));
        *length_out = e->h_result[4];
        if (steps_out) *steps_out = (int)e->h_result[5];
        return e->h_result[6] != 0.0 ? B200LK_WARN_DERV_UNDERFLOW : 0;
    }
    int nsteps = 0;
    auto func_derv = [&](double x, double &f, double &df) -> int {
        enqueue_derv(e, x);
        cudaError_t ce = sync_stream(e);
        if (ce != cudaSuccess) { e->err = cudaGetErrorString(ce); return B200LK_ERR_CUDA; }
        f = -e->h_result[1];
        df = -e->h_result[2];
        nsteps++;
        return 0;
    };
    rc = host_newton(func_derv, x1, xguess, x2, xacc, maxNRStep, length_out);
    if (steps_out) *steps_out = nsteps;
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// read-backs
// ---------------------------------------------------------------------------------------------------------------
static int ensure_scratch(b200lk_engine *e, size_t bytes) {
    if (bytes <= e->scratch_cap) return 0;
    if (e->d_scratch) { cudaFree(e->d_scratch); e->device_bytes -= (int64_t)e->scratch_cap; }
    e->d_scratch = nullptr; e->scratch_cap = 0;
    int rc = dev_alloc(e, (char **)&e->d_scratch, bytes);
    if (rc) return rc;
    e->scratch_cap = bytes;
    return 0;
}

static int read_interleaved_f64(b200lk_engine *e, const double *src, int W, double *out) {
    cudaSetDevice(e->cfg.device);
    const size_t total = (size_t)e->P * W;
    int rc = ensure_scratch(e, total * sizeof(double));
    if (rc) return rc;
    k_deinterleave_f64<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(src, e->d_scratch, e->P, W);
    e->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, e->d_scratch, total * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(sync_stream(e));
    return 0;
}

static int read_plain(b200lk_engine *e, const void *src, size_t bytes, void *out) {
    cudaSetDevice(e->cfg.device);
    CK(cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(sync_stream(e));
    return 0;
}

extern "C" int b200lk_get_pattern_lh(b200lk_engine *e, double *out) {
    if (!e->shards.empty()) return mg_gather(e, out, 1, [](b200lk_engine *x, double *p) { return b200lk_get_pattern_lh(x, p); }); return read_plain(e, e->d_pattern_lh, sizeof(double) * e->P, out); }
extern "C" int b200lk_get_pattern_scaling(b200lk_engine *e, double *out) {
    if (!e->shards.empty()) return mg_gather(e, out, 1, [](b200lk_engine *x, double *p) { return b200lk_get_pattern_scaling(x, p); }); return read_plain(e, e->d_pattern_scaling, sizeof(double) * e->P, out); }
extern "C" int b200lk_get_buffer_scale_all(b200lk_engine *e, double *out) {
    if (!e->shards.empty()) return mg_gather(e, out, 1, [](b200lk_engine *x, double *p) { return b200lk_get_buffer_scale_all(x, p); }); return read_plain(e, e->d_buffer_scale, sizeof(double) * e->P, out); }
extern "C" int b200lk_get_pattern_lh_cat(b200lk_engine *e, double *out) {
    if (!e->shards.empty()) return mg_gather(e, out, (size_t)e->C, [](b200lk_engine *x, double *p) { return b200lk_get_pattern_lh_cat(x, p); }); return read_interleaved_f64(e, e->d_pattern_lh_cat, e->C, out); }
// PhyloTree::computePatternLhCat(wsl) + the E-step loop of RateFree / ModelMixture::optimizeWithEM, on the device
extern "C" int b200lk_em_posterior(b200lk_engine *e, int32_t wsl, int32_t add_invar, double *new_prop, double *posterior, int32_t *ncat_out) {
    if (!e->shards.empty()) return mg_em_posterior(e, wsl, add_invar, new_prop, posterior, ncat_out);
    if (wsl < 2 || wsl > 4) return fail(e, B200LK_ERR_INVALID, "wsl: 2 = WSL_RATECAT, 3 = WSL_MIXTURE, 4 = WSL_MIXTURE_RATECAT");
    if (!e->have_lh_cat) return fail(e, B200LK_ERR_STATE, "em_posterior: needs the per-category likelihoods of a b200lk_branch_logl call");
    const bool marginal = e->cfg.nmix > 1 && !e->cfg.fused_mix_rate && (wsl == 2 || wsl == 3);
    const int mode = !marginal ? 0 : wsl == 2 ? 1 : 2;
    const int K = mode == 0 ? e->C : mode == 1 ? e->cfg.ncat : e->cfg.nmix;
    if (ncat_out) *ncat_out = K;
    if (!new_prop) return 0;
    cudaSetDevice(e->cfg.device);
    const int nblocks = (int)(e->Ppad / kTile);
    if (!e->d_post || !e->d_em_red || !e->d_em_out) {   // all three or none
        const int64_t before = e->device_bytes;
        cudaFree(e->d_post); cudaFree(e->d_em_red); cudaFree(e->d_em_out);
        e->d_post = e->d_em_red = e->d_em_out = nullptr;
        int rc = dev_alloc(e, &e->d_post, (size_t)e->C * e->Ppad);
        if (!rc) rc = dev_alloc(e, &e->d_em_red, (size_t)e->C * nblocks);
        if (!rc) rc = dev_alloc(e, &e->d_em_out, (size_t)e->C);
        if (rc) {
            cudaFree(e->d_post); cudaFree(e->d_em_red); cudaFree(e->d_em_out);
            e->d_post = e->d_em_red = e->d_em_out = nullptr;
            e->device_bytes = before;
            return rc;
        }
    }
    const size_t smem = sizeof(double) * (kTile / 32) * K;
    if (smem > 200 * 1024) return fail(e, B200LK_ERR_INVALID, "em_posterior: too many categories");
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_em_posterior, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_em_posterior<<<nblocks, kTile, smem, e->stream>>>(e->d, mode, K, add_invar != 0, e->d_post, e->d_em_red);
    k_em_reduce<<<(K + 127) / 128, 128, 0, e->stream>>>(nblocks, K, e->d_em_red, e->d_em_out);
    e->launches += 2;
    CK(cudaGetLastError());
    e->post_K = K;
    CK(cudaMemcpyAsync(new_prop, e->d_em_out, sizeof(double) * K, cudaMemcpyDeviceToHost, e->stream));
    if (posterior) {   // [nptn][K] for callers that want the matrix on the host; the M-step does not need it
        std::vector<double> col((size_t)e->Ppad);
        for (int k = 0; k < K; k++) {
            CK(cudaMemcpyAsync(col.data(), e->d_post + (size_t)k * e->Ppad, sizeof(double) * e->P, cudaMemcpyDeviceToHost, e->stream));
            CK(sync_stream(e));
            for (int64_t p = 0; p < e->P; p++) posterior[(size_t)p * K + k] = col[(size_t)p];
        }
    }
    CK(sync_stream(e));
    return 0;
}

extern "C" int b200lk_set_pattern_freq_from_posterior(b200lk_engine *dst, b200lk_engine *src, int32_t k) {
    if (!dst->shards.empty() || !src->shards.empty()) { dst->err = "set_pattern_freq_from_posterior: not available on a multi-device engine"; return B200LK_ERR_STATE; }
    b200lk_engine *e = dst;   // errors are reported on the destination
    if (!src->d_post || k < 0 || k >= src->post_K) return fail(dst, B200LK_ERR_STATE, "no posteriors for that category: call b200lk_em_posterior on the source engine first");
    if (dst->P != src->P || dst->cfg.device != src->cfg.device) return fail(dst, B200LK_ERR_INVALID, "engines differ in pattern count or device");
    cudaSetDevice(dst->cfg.device);
    CK(sync_stream(src));
    k_copy_column<<<(unsigned)((dst->P + 255) / 256), 256, 0, dst->stream>>>(dst->P, src->d_post + (size_t)k * src->Ppad, dst->d_freq);
    dst->launches++;
    CK(cudaGetLastError());
    // the host copy of ptn_freq (NaN repair, robust phylo) follows
    CK(cudaMemcpyAsync(dst->h_freq.data(), dst->d_freq, sizeof(double) * dst->P, cudaMemcpyDeviceToHost, dst->stream));
    CK(sync_stream(dst));
    dst->theta_computed = false;
    return 0;
}

extern "C" int b200lk_get_theta(b200lk_engine *e, double *out) {
    if (!e->shards.empty()) return mg_gather(e, out, (size_t)e->B, [](b200lk_engine *x, double *p) { return b200lk_get_theta(x, p); }); return read_interleaved_f64(e, e->d_theta, e->B, out); }
extern "C" int b200lk_get_partial(b200lk_engine *e, int32_t slot, double *out) {
    if (!e->shards.empty()) return mg_gather(e, out, (size_t)e->B, [&](b200lk_engine *x, double *p) { return b200lk_get_partial(x, slot, p); });
    if (slot < 0 || slot >= e->cfg.nslots) return fail(e, B200LK_ERR_INVALID, "slot out of range");
    cudaSetDevice(e->cfg.device);
    const double *src = e->d_partial + (size_t)eigen_slot(e, slot) * e->d.slot_stride;
    if (!e->d.clv_f20 && !e->d.clv_f8) return read_interleaved_f64(e, src, e->B, out);
    const size_t total = (size_t)e->P * e->B;
    int rc = ensure_scratch(e, total * sizeof(double));
    if (rc) return rc;
    k_deinterleave_clv<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(e->d, src, e->d_scratch);
    e->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, e->d_scratch, total * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CK(sync_stream(e));
    return 0;
}
extern "C" int b200lk_get_scale_num(b200lk_engine *e, int32_t slot, uint16_t *out) {
    if (!e->shards.empty()) return mg_gather(e, out, (size_t)e->S, [&](b200lk_engine *x, uint16_t *p) { return b200lk_get_scale_num(x, slot, p); });
    if (slot < 0 || slot >= e->cfg.nslots) return fail(e, B200LK_ERR_INVALID, "slot out of range");
    const uint16_t *src = e->d_scale + (size_t)slot * e->d.scale_stride;
    if (!e->cfg.safe_numeric) return read_plain(e, src, sizeof(uint16_t) * e->P, out);
    cudaSetDevice(e->cfg.device);
    const size_t total = (size_t)e->P * e->C;
    int rc = ensure_scratch(e, total * sizeof(uint16_t));
    if (rc) return rc;
    k_deinterleave_u16<<<(unsigned)((total + 255) / 256), 256, 0, e->stream>>>(src, (uint16_t *)e->d_scratch, e->P, e->C);
    e->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, e->d_scratch, total * sizeof(uint16_t), cudaMemcpyDeviceToHost, e->stream));
    CK(sync_stream(e));
    return 0;
}
extern "C" int b200lk_get_tip_partial_lh(b200lk_engine *e, double *out) {
    if (!e->shards.empty()) return b200lk_get_tip_partial_lh(e->shards[0], out);
    return read_plain(e, e->d.tip_lh, sizeof(double) * (e->cfg.state_unknown + 1) * e->cfg.nmix * e->cfg.nstates, out);
}
extern "C" void *b200lk_pattern_lh_dev(b200lk_engine *e) {
    if (!e->shards.empty()) return nullptr; return e->d_pattern_lh; }

// ---------------------------------------------------------------------------------------------------------------
// UFBoot RELL on the device
// ---------------------------------------------------------------------------------------------------------------
extern "C" int b200lk_set_boot_samples(b200lk_engine *e, const float *boot_samples, int32_t nboot) {
    if (!e->shards.empty()) return mg_set_boot_samples(e, boot_samples, nboot);
    cudaSetDevice(e->cfg.device);
    if (nboot < 0) return fail(e, B200LK_ERR_INVALID, "nboot < 0");
    CK(sync_stream(e));
    if (e->d_boot) { cudaFree(e->d_boot); e->device_bytes -= (int64_t)e->boot_bytes; e->d_boot = nullptr; }
    if (e->d_rell) { cudaFree(e->d_rell); e->d_rell = nullptr; }
    if (e->d_boot_state) { cudaFree(e->d_boot_state); e->d_boot_state = nullptr; }
    e->nboot = 0; e->boot_bytes = 0;
    if (nboot == 0) return 0;
    e->boot_stride = (e->P + 3) / 4 * 4;   // rows 16-byte aligned
    const size_t n = (size_t)nboot * e->boot_stride;
    int rc = dev_alloc(e, &e->d_boot, n);
    if (rc) return rc;
    e->boot_bytes = n * sizeof(float);
    rc = dev_alloc(e, &e->d_rell, (size_t)nboot);
    if (rc) return rc;
    CK(cudaMemsetAsync(e->d_boot, 0, n * sizeof(float), e->stream));
    CK(cudaMemcpy2DAsync(e->d_boot, e->boot_stride * sizeof(float), boot_samples, e->P * sizeof(float), e->P * sizeof(float), nboot,
                         cudaMemcpyHostToDevice, e->stream));
    // bookkeeping state of b200lk_rell_update: boot_logl, boot_orig_logl = -DBL_MAX, boot_counts = 0 (tree/iqtree.cpp:375-378)
    rc = dev_alloc(e, &e->d_boot_state, (size_t)3 * nboot + (size_t)(nboot + 1) / 2 + 1 + (size_t)(nboot + 7) / 8 + 1);   // 3 double arrays | counts | nupdated | updated
    if (rc) return rc;
    e->nboot = nboot;
    {
        double *bl = e->d_boot_state, *bo = bl + nboot;
        int *bc = (int *)(bl + 3 * (size_t)nboot);
        k_boot_reset<<<(nboot + 127) / 128, 128, 0, e->stream>>>(nboot, bl, bc, bo);
        e->launches++;
    }
    CK(sync_stream(e));
    return 0;
}

// device pointers of the bookkeeping arrays inside d_boot_state
struct BootState { double *logl, *orig, *uniform; int *counts, *nupdated; unsigned char *updated; };
static BootState boot_state(b200lk_engine *e) {
    BootState b;
    b.logl = e->d_boot_state; b.orig = b.logl + e->nboot; b.uniform = b.orig + e->nboot;
    b.counts = (int *)(b.uniform + e->nboot); b.nupdated = b.counts + e->nboot;
    b.updated = (unsigned char *)(b.nupdated + 2);
    return b;
}

extern "C" int b200lk_boot_reset(b200lk_engine *e) {
    if (!e->shards.empty()) MG_UNSUPPORTED(e, "boot_reset");
    cudaSetDevice(e->cfg.device);
    if (e->nboot <= 0) return fail(e, B200LK_ERR_STATE, "boot_reset: no bootstrap samples set");
    BootState b = boot_state(e);
    k_boot_reset<<<(e->nboot + 127) / 128, 128, 0, e->stream>>>(e->nboot, b.logl, b.counts, b.orig);
    e->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b200lk_rell_update(b200lk_engine *e, double cur_logl, double epsilon, const double *uniform, uint8_t *updated, int32_t *nupdated) {
    if (!e->shards.empty()) MG_UNSUPPORTED(e, "rell_update");
    cudaSetDevice(e->cfg.device);
    if (e->nboot <= 0) return fail(e, B200LK_ERR_STATE, "rell_update: no bootstrap samples set");
    BootState b = boot_state(e);
    if (uniform) CK(cudaMemcpyAsync(b.uniform, uniform, sizeof(double) * e->nboot, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemsetAsync(b.nupdated, 0, sizeof(int), e->stream));
    k_rell<<<e->nboot, 256, 0, e->stream>>>(e->d_pattern_lh, e->d_boot, e->P, e->boot_stride, e->d_rell);
    k_rell_update<<<(e->nboot + 127) / 128, 128, 0, e->stream>>>(e->nboot, e->d_rell, cur_logl, epsilon, uniform ? b.uniform : nullptr, b.logl, b.counts, b.orig,
                                                                 b.updated, b.nupdated);
    e->launches += 2;
    CK(cudaGetLastError());
    int n = 0;
    CK(cudaMemcpyAsync(&n, b.nupdated, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    if (updated) CK(cudaMemcpyAsync(updated, b.updated, (size_t)e->nboot, cudaMemcpyDeviceToHost, e->stream));
    CK(sync_stream(e));
    if (nupdated) *nupdated = n;
    return 0;
}

extern "C" int b200lk_get_boot_state(b200lk_engine *e, double *boot_logl, int32_t *boot_counts, double *boot_orig_logl) {
    if (!e->shards.empty()) MG_UNSUPPORTED(e, "get_boot_state");
    cudaSetDevice(e->cfg.device);
    if (e->nboot <= 0) return fail(e, B200LK_ERR_STATE, "get_boot_state: no bootstrap samples set");
    BootState b = boot_state(e);
    if (boot_logl) CK(cudaMemcpyAsync(boot_logl, b.logl, sizeof(double) * e->nboot, cudaMemcpyDeviceToHost, e->stream));
    if (boot_counts) CK(cudaMemcpyAsync(boot_counts, b.counts, sizeof(int) * e->nboot, cudaMemcpyDeviceToHost, e->stream));
    if (boot_orig_logl) CK(cudaMemcpyAsync(boot_orig_logl, b.orig, sizeof(double) * e->nboot, cudaMemcpyDeviceToHost, e->stream));
    CK(sync_stream(e));
    return 0;
}

extern "C" int b200lk_rell(b200lk_engine *e, float *rell) {
    if (!e->shards.empty()) return mg_rell(e, rell);
    cudaSetDevice(e->cfg.device);
    if (e->nboot <= 0) return fail(e, B200LK_ERR_STATE, "rell: no bootstrap samples set");
    k_rell<<<e->nboot, 256, 0, e->stream>>>(e->d_pattern_lh, e->d_boot, e->P, e->boot_stride, e->d_rell);
    e->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rell, e->d_rell, sizeof(float) * e->nboot, cudaMemcpyDeviceToHost, e->stream));
    CK(sync_stream(e));
    return 0;
}

#include "b200lk_mixlen.inc"
#include "b200lk_group.inc"
