#!/usr/bin/env python3
"""bench.py — full-tree logL+derivative evaluations/s (BASELINE.json metric) through the C ABI.

One STEP = one EVAL of SURVEY §8(d):  invalidate every CLV (clearAllPartialLH) -> computeLikelihoodBranch at a
fixed branch (full post-order traversal + root reduction) -> computeLikelihoodDerv at that branch (theta build +
one (df, ddf) pass).  Default workload = BASELINE.json configs[1]: synthetic DNA 1000 taxa x 1,000,000 patterns,
GTR+G4 (NORM_LH), which fits one B200 (127.7 GB of CLVs).

  python bench.py --gpus N --steps K --warmup W               # this framework (CUDA engine; no CPU fallback)
  python bench.py --impl reference --gpus N --steps K ...     # the reference's CPU algorithm on the host cores

N > 1 (torchrun): the pattern axis is sharded across ranks (each GPU owns a contiguous pattern slice of every CLV,
SURVEY §8e); the only exchange is an all-reduce of the (logL, df, ddf) scalars.  Total work is fixed => "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from iqtree2_b200 import models, plan, treesim  # noqa: E402

WORKLOADS = {
    # name: (ntaxa, npatterns, model factory, description)   -- BASELINE.json configs[1..3]
    "dna_1000x1M_gtr_g4": (1000, 1_000_000, lambda: models.make_gtr_gamma(), "synthetic DNA 1000 taxa x 1M patterns, GTR+G4"),
    "aa_200x100k_lg_g4": (200, 100_000, lambda: models.make_lg_gamma(), "synthetic protein 200 taxa x 100k patterns, LG+G4"),
    "aa_100x50k_lg_c60_g4": (100, 50_000, lambda: models.make_lg_c60_gamma(), "synthetic protein 100 taxa x 50k patterns, LG+C60+G4"),
    # not a BASELINE config: the general-n FP64-MMA walker.  61 sense codons, random reversible exchangeabilities,
    # Gamma(4); SAFE_LH as the reference forces for every n outside {4, 20} (tree/phylotreesse.cpp:95-96)
    # not a BASELINE config: the non-reversible kernels (tree/phylokernelnonrev.h) on a ROOTED tree — a general 12-parameter DNA
    # rate matrix (UNREST-like) + Gamma(4); ntaxa real taxa + the virtual root leaf
    "dna_500x200k_unrest_g4": (500, 200_000, lambda: models.make_nonreversible(4, ncat=4, seed=11, state_unknown=models.DNA_STATE_UNKNOWN),
                               "synthetic DNA 500 taxa x 200k patterns, rooted, non-reversible 12-parameter model + G4"),
    # not a BASELINE config: site-specific models (ModelSet / PMSF, the SITE_MODEL branch of tree/phylokernelnew.h): one frequency
    # profile, hence one eigensystem, per pattern + Gamma(4)
    "aa_100x20k_pmsf_g4": (100, 20_000, lambda: models.make_site_specific_fast(20, 20_000, ncat=4, seed=9),
                           "synthetic protein 100 taxa x 20k patterns, site-specific frequency profiles (PMSF-like) + G4"),
    "codon_100x20k_rev61_g4": (100, 20_000, lambda: models.make_random_reversible(61, ncat=4, seed=7), "synthetic codon 100 taxa x 20k patterns, 61-state reversible + G4"),
}
SAFE_WORKLOADS = {"codon_100x20k_rev61_g4"}


SUPER_WORKLOAD = "super_500part_100taxa"   # BASELINE.json configs[4]
DEFAULT_WORKLOAD = "dna_1000x1M_gtr_g4"    # BASELINE.json configs[1]: the headline line
SECONDARY_WORKLOAD = "aa_200x100k_lg_g4"   # BASELINE.json configs[2] (the FP64-MMA path): measured in the same run, reported under "secondary"


def make_super_inputs(ntaxa=100, nparts=500, seed=5):
    """500 partitions, lengths ~ LogNormal(median 1000), 80 % DNA GTR+G4 / 20 % protein LG+G4,
    per-partition rate ~ Gamma(2, 0.5), 10 % of the taxa without data in each partition (rows of STATE_UNKNOWN)."""
    rng = np.random.default_rng(seed)
    tree = treesim.yule_tree(ntaxa, seed=1)
    parts = []
    for p in range(nparts):
        aa = rng.random() < 0.2
        if aa:
            model = models.make_lg_gamma(alpha=0.5 + rng.random())
        else:
            model = models.make_gtr_gamma(rates6=tuple(0.5 + 3.0 * rng.random(6)), pi=tuple(rng.dirichlet([8, 8, 8, 8])), alpha=0.3 + rng.random())
        rate = float(rng.gamma(2.0, 0.5)) + 0.02
        nptn = int(np.clip(rng.lognormal(np.log(1000.0), 0.6), 100, 8000))
        scaled = treesim.Tree(tree.ntaxa, tree.edges, tree.lengths * rate, tree.names)
        pat = treesim.evolve_patterns_fast(scaled, model.nstates, nptn, seed=1000 + p)
        absent = rng.random(ntaxa) < 0.10
        pat[absent, :] = model.state_unknown
        parts.append(dict(model=model, rate=rate, pat=np.ascontiguousarray(pat), freq=np.ones(nptn), aa=bool(aa)))
    return tree, parts


def super_algorithmic(parts, ntaxa):
    """algorithmic bytes / flops of one evaluation summed over the partitions (SURVEY §8d per-pattern figures, all T-2 updates)"""
    by = fl = 0.0
    for p in parts:
        n, C = p["model"].nstates, p["model"].ncat_mix
        B = n * C
        nptn = p["pat"].shape[1]
        by += nptn * (2.0 * (ntaxa - 2) * (8 * B + 2) + ntaxa + 16 + 8 + 2 * 8 * B + 16)
        fl += nptn * C * (4.0 * n * n + n) * (ntaxa - 2)   # a mid estimate of the node mix (tip-inner everywhere)
    return by, fl


def run_super(args, rank, world, local_rank):
    """BASELINE config 5 through the partition-model API (include/b200pt.h b200st_*): one step = clearAllPartialLH on every
    partition tree -> computeLikelihoodBranch (sum of the partition logLs) -> computeFuncDerv at that branch.  Partitions are
    dealt round-robin (in cost order) to the ranks; the only exchange is the all-reduce of (logL, df, ddf)."""
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # one hardware queue per partition stream as far as the driver allows
    import torch
    import torch.distributed as dist
    from iqtree2_b200.supertree import PhyloSuperTree
    ntaxa, nparts = args.taxa or 100, args.parts or 500
    tree, parts = make_super_inputs(ntaxa, nparts)
    from iqtree2_b200.sharded import deal_partitions
    cost = [p["pat"].shape[1] * p["model"].nstates for p in parts]
    order = np.argsort(-np.asarray(cost, dtype=np.float64), kind="stable")
    mine = deal_partitions(cost, world, rank)
    if args.impl == "reference":
        if rank != 0:
            return 0
        from concurrent.futures import ThreadPoolExecutor
        from oracle import ref_engine
        nsample = min(args.cpu_sample or 48, nparts)
        sample = [int(i) for i in order[:: max(1, nparts // nsample)][:nsample]]
        threads = host_threads()
        refs = []
        for i in sample:
            p = parts[i]
            scaled = treesim.Tree(tree.ntaxa, tree.edges, tree.lengths * p["rate"], tree.names)
            refs.append(ref_engine.RefTree(scaled, p["model"], p["pat"], p["freq"], nthreads=1))
        def one(r):
            return r.time_eval(0, 1)
        with ThreadPoolExecutor(threads) as ex:
            for _ in range(max(1, args.warmup)):
                list(ex.map(one, refs))
            t0 = time.perf_counter()
            for _ in range(args.steps):
                out = list(ex.map(one, refs))
            dt = (time.perf_counter() - t0) / args.steps
        frac = sum(parts[i]["pat"].shape[1] * parts[i]["model"].nstates ** 2 for i in sample) / sum(p["pat"].shape[1] * p["model"].nstates ** 2 for p in parts)
        evals = frac / dt
        line = {"impl": "reference", "metric": "full_tree_logl_deriv_evals_per_sec", "value": evals, "unit": "evals/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / evals, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": SUPER_WORKLOAD, "ntaxa": ntaxa, "npartitions": nparts},
                "cpu_baseline": {"value": evals, "unit": "evals/s", "cores": threads, "kind": "reference",
                                 "what": "IQ-TREE 2.4.0 kernel per partition tree (Vec4d AVX+FMA), partitions spread over the host threads as PhyloSuperTree's omp parallel for does",
                                 "sample": f"{len(sample)} of {nparts} partitions (every {max(1, nparts // nsample)}th in cost order), whole evaluations; throughput scaled by their share {frac:.4f} of the n^2-weighted pattern count"},
                "e2e": {"value": evals, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        for r in refs:
            r.close()
        print(json.dumps(line))
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the likelihood engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # host threads that plan the partitions' traversals: set explicitly (torchrun exports OMP_NUM_THREADS=1, which would plan
    # 250 partitions on one thread per rank and hide the GPUs' scaling behind the host)
    nthreads = args.host_threads or max(1, min(16, host_threads() // max(world, 1)))
    st = PhyloSuperTree(tree.names, num_threads=nthreads)
    for i in mine:
        p = parts[i]
        st.add_partition(p["pat"], p["freq"], p["model"], part_rate=p["rate"], device=local_rank)
    st.read_tree(tree.to_newick(fmt="%.17g"))
    br = 0
    _, _, ln = st.branches()
    scal = torch.zeros(3, dtype=torch.float64, device="cuda")

    def one_eval():
        st.clear_all_partial_lh()
        ll = st.compute_likelihood_branch(br)
        df, ddf = st.compute_likelihood_derv(br, float(ln[br]))
        if world > 1:
            scal.copy_(torch.tensor([ll, df, ddf], dtype=torch.float64))
            dist.all_reduce(scal)
            ll, df, ddf = scal.tolist()
        return ll, df, ddf

    for _ in range(args.warmup):
        res = one_eval()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = st.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = one_eval()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = st.launch_count - l0
    if world > 1:
        t = torch.tensor([wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    K = args.steps
    value = K / wall
    by, fl = super_algorithmic(parts, ntaxa)
    peak, peak_src = load_peaks()
    nbytes_h2d = sum((ntaxa - 2) * 40 + 8 * (2 * p["model"].nstates ** 2 + 32) for p in (parts[i] for i in mine))
    line = {
        "metric": "full_tree_logl_deriv_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": wall * 1e3 / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic: Yule-Harding tree (seed 1); per partition tree-evolved columns at the partition's rate, every column its own pattern; "
                "10 % of the taxa per partition without data (rows of STATE_UNKNOWN)",
        "config": {"workload": SUPER_WORKLOAD, "ntaxa": ntaxa, "npartitions": nparts, "partitions_per_gpu": len(mine),
                   "npatterns_total": int(sum(p["pat"].shape[1] for p in parts)), "protein_partitions": int(sum(p["aa"] for p in parts)),
                   "model": "edge-linked proportional partition model (PhyloSuperTreePlen), GTR+G4 / LG+G4 per partition",
                   "host_threads": nthreads, "clv_bytes_resident": st.device_bytes, "algorithmic_GB_per_eval": by / 1e9,
                   "l2": "each evaluation rewrites every CLV of every partition (working set > L2)",
                   "parallelism": f"partition-sharded x{world}" if world > 1 else "single GPU", "timing": "wall clock around K steps, device idle on both sides (the step is host-driven: traversals planned per partition, two grouped calls, one wait each)"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": int(nbytes_h2d), "d2h_bytes_per_step": 24 * len(mine),
                "ms_per_step": wall * 1e3 / K, "what": "the b200st_* calls a PhyloSuperTree host makes: per partition the op list with branch lengths goes up and (logL, df, ddf) come back every step; identical to `value` (there is no resident-input variant of a host-driven partition sweep)"},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_group_walk_n4<4> + k_group_walk_n20v2 (grouped launches: one walker launch per kernel family for all partitions)",
                     "achieved": by / (wall / K) / 1e9, "peak": peak, "unit": "GB/s", "frac": by / (wall / K) / 1e9 / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_step": by,
                     "note": "the figure is the whole step (host planning + all kernels of all partitions), not one launch"},
        "result": {"logl": res[0], "df": res[1], "ddf": res[2]},
    }
    if rank == 0:
        print(json.dumps(line))
    st.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def host_threads():
    """threads the CPU arm may use: the process's CPU affinity, NOT omp_get_max_threads() (torchrun exports OMP_NUM_THREADS=1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def host_memory_available():
    """bytes of host memory a CPU sample may plan with: min(what the OS reports available, what the cgroup still allows)"""
    avail = None
    try:
        import psutil
        avail = int(psutil.virtual_memory().available)
    except Exception:
        pass
    try:
        lim = open("/sys/fs/cgroup/memory.max").read().strip()
        if lim != "max":
            cur = int(open("/sys/fs/cgroup/memory.current").read().strip())
            room = int(lim) - cur
            avail = room if avail is None else min(avail, room)
    except Exception:
        pass
    return avail if avail is not None else 16 << 30


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.device_index = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 2.0:   # nvidia-smi needs a moment before its first line
                time.sleep(0.02)
            self.lines.clear()                                  # keep only samples taken inside the timed region
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def load_traffic(workload, overridden, my_nptn):
    """roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel from the committed
    ncu capture of this workload at full size (profiles/traffic.json), scaled by this rank's share of the patterns; None
    for workloads / overrides without a capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if overridden or not os.path.exists(p):
        return None, None
    t = json.load(open(p)).get(workload)
    if not t:
        return None, None
    return (t["dram_bytes_read"] + t["dram_bytes_write"]) * (my_nptn / t["npatterns"]), t["source"]


def load_fp64_peak():
    p = os.path.join(ROOT, "profiles", "fp64_peaks_b200.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["dmma_m8n8k4_tflops"]), "measured (profiles/fp64_peaks_b200.json: mma.sync m8n8k4.f64, scripts/fp64_mix.cu)"
    return 40.0, "fallback (B200 datasheet FP64 40 TFLOP/s)"


def algorithmic_flops(ops, nptn, n, C):
    """per pattern C*[(6n^2+n) N_ii + (4n^2+n) N_ti + (2n^2+n) N_tt] for the partial-likelihood kernels."""
    n_ii, n_ti, n_tt = plan.count_node_cases(ops)
    return float(nptn) * C * ((6 * n * n + n) * n_ii + (4 * n * n + n) * n_ti + (2 * n * n + n) * n_tt)


def algorithmic_bytes(ops, nptn, B, S, ntaxa, n_root_slots=2):
    """per pattern 2(T-2)(8B+2S) [each internal CLV+scale written once, read once] + T [tip codes]
    + 16 [freq, invar] + 8 [_pattern_lh] + 2*8B + 16 [theta write + read, freq, invar].
    Returns (bytes per eval, bytes of the partial-likelihood kernels only), computed from the actual op list."""
    clv = 8 * B + 2 * S
    partial = 0
    for op in ops:
        partial += clv  # write dst
        for side in ("left", "right"):
            partial += clv if op[side]["kind"] == plan.KIND_SLOT else 1
    # the CLVs at the evaluation branch are read once by the root stage: one slot when the other end is a leaf
    root_reads = clv * n_root_slots
    per_ptn_eval = partial + root_reads + 16 + 8 + 2 * 8 * B + 16
    return per_ptn_eval * nptn, partial * nptn


def make_inputs(ntaxa, nptn, model, seed_tree=1, seed_seq=2, block=262_144):
    tree = treesim.yule_tree(ntaxa, seed=seed_tree)
    nb = min(nptn, block)
    base = treesim.evolve_patterns_fast(tree, model.nstates, nb, seed=seed_seq)
    reps = (nptn + nb - 1) // nb
    tips = np.tile(base, (1, reps))[:, :nptn] if reps > 1 else base
    tips = np.ascontiguousarray(tips)
    freq = np.ones(nptn)
    data = (f"synthetic: Yule-Harding tree (seed {seed_tree}), Exp(0.1) branch lengths; {nb} tree-evolved columns (seed {seed_seq})"
            + (f" tiled x{reps} to {nptn} patterns" if reps > 1 else "") + "; every column counted as its own pattern (freq 1)")
    return tree, tips, freq, data


def run_cpu_reference(tree, model, tips, freq, ops, a, b, L, sample_ptn, steps, warmup, safe=False):
    """The reference's CPU implementation of one eval on a bounded pattern sample, all host threads.

    Preferred: oracle/_ref/libiqref.so = IQ-TREE 2.4.0's own AVX+FMA Vec4d OpenMP kernel (its TUs compiled in place by
    oracle/build_ref.sh; the prebuilt .so travels with the snapshot)  -> kind "reference".
    Otherwise the scalar C restatement oracle/lk_oracle.c (OpenMP over pattern packets) -> kind "port"."""
    from oracle import ref_engine
    sample_ptn = int(min(sample_ptn, tips.shape[1]))
    if ref_engine.available():
        threads = host_threads()   # set explicitly: the harness calls omp_set_num_threads(), which overrides OMP_NUM_THREADS
        r = ref_engine.RefTree(tree, models.site_slice(model, 0, sample_ptn), np.ascontiguousarray(tips[:, :sample_ptn]), freq[:sample_ptn], safe_numeric=safe, nthreads=threads)
        if warmup:
            r.time_eval(0, warmup)
        t = r.time_eval(0, steps)
        info = r.info()
        r.close()
        return {"sec_per_sample_eval": t["mean_sec"], "best_sec": t["best_sec"], "sample_ptn": sample_ptn, "cores": info["num_threads"],
                "logl": t["logl"], "df": t["df"], "ddf": t["ddf"], "kind": "reference",
                "what": "IQ-TREE 2.4.0 phylokernelnew.h, Vec4d AVX+FMA, OpenMP (compiled from the reference's own TUs)"}
    from oracle.oracle_engine import OracleEngine
    o = OracleEngine(models.site_slice(model, 0, sample_ptn), tree.ntaxa, sample_ptn, safe_numeric=safe)
    o.set_tip_states(np.ascontiguousarray(tips[:, :sample_ptn]))
    o.set_pattern_freq(freq[:sample_ptn])
    o.set_pattern_invar(np.zeros(sample_ptn))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        o.theta_reset()
        ll = o.branch_logl(ops, a, b, L)
        df, ddf = o.branch_derv(ops[:0], a, b, L)
        times.append(time.perf_counter() - t0)
    t = times[warmup:]
    return {"sec_per_sample_eval": float(np.mean(t)), "best_sec": float(np.min(t)), "sample_ptn": sample_ptn,
            "cores": OracleEngine.num_threads(), "logl": ll, "df": df, "ddf": ddf, "kind": "port",
            "what": "scalar C restatement of phylokernelnew.h (oracle/lk_oracle.c), OpenMP"}


def shared_config(workload_name, ntaxa, nptn, model, desc, nops, safe=False):
    """the `config` object: identical for this framework's arm and for the reference arm of the same workload"""
    return {"workload": workload_name, "ntaxa": ntaxa, "npatterns": nptn, "model": desc, "kernel": "SAFE_LH" if safe else "NORM_LH", "nstates": model.nstates,
            "ncat_mix": model.ncat_mix, "partial_updates_per_eval": nops,
            "l2": "inputs (CLVs) far larger than L2; no flush needed"}


def run_reference_arm(args, workload, ntaxa, nptn, model, desc, workload_name, steps=None, warmup=None):
    """`--impl reference`: the reference's own CPU kernel (oracle/_ref/libiqref.so) on every host thread the process may use.
    A step evaluates a SAMPLE of the workload's patterns (patterns are independent): the whole workload when the host
    memory holds the reference's (ntaxa-2) x nptn x B doubles of CLVs, otherwise as many patterns as fit in half of the available
    memory; `ms_per_step` is the measured time of that step, `value` scales it to the whole pattern count, `sample_fraction`
    says by how much."""
    steps = steps or args.steps
    warmup = max(1, args.warmup if warmup is None else warmup)
    B = model.nstates * model.ncat_mix
    bytes_per_ptn = (ntaxa + 2) * (8 * B + 2) + 4 * ntaxa + 8 * (4 * B + 8)   # CLVs + scale, tip codes (u32 patterns), theta & co
    mem_cap = int(0.5 * host_memory_available() / bytes_per_ptn)
    sample = nptn if not args.cpu_sample else min(args.cpu_sample, nptn)
    sample = max(256, min(sample, mem_cap))
    # time cap: a short calibration run, then at most ~150 s for warm-up + timed steps (never below a quarter of the patterns
    # when memory allows that)
    nonrev = getattr(model, "nonrev", None) is not None
    tree, tips, freq, data = make_inputs(ntaxa + (1 if nonrev else 0), sample, model)
    if nonrev:
        tips = np.ascontiguousarray(tips[:ntaxa])
    ops, a, b, L = (plan.plan_branch_nonrev if nonrev else plan.plan_branch)(tree, plan.pendant_branch_of(tree, 0))
    if not args.cpu_sample:
        cal_n = min(sample, 8192)
        cal = run_cpu_reference(tree, model, tips, freq, ops, a, b, L, cal_n, 1, 1, safe=workload in SAFE_WORKLOADS)
        per_ptn = cal["best_sec"] / cal_n
        by_time = int(150.0 / ((steps + warmup) * per_ptn))
        sample = min(sample, max(by_time, min(nptn // 4, mem_cap)))
        sample = max(256, sample // 8 * 8)
    r = run_cpu_reference(tree, model, tips, freq, ops, a, b, L, sample, steps, warmup, safe=workload in SAFE_WORKLOADS)
    frac = sample / nptn
    evals = frac / r["sec_per_sample_eval"]
    return {"impl": "reference", "metric": "full_tree_logl_deriv_evals_per_sec", "value": evals, "unit": "evals/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["sec_per_sample_eval"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": data,
            "config": shared_config(workload_name, ntaxa, nptn, model, desc, len(ops), safe=workload in SAFE_WORKLOADS),
            "sample_fraction": frac,
            "cpu_baseline": {"value": evals, "unit": "evals/s", "cores": r["cores"], "kind": r["kind"], "what": r["what"],
                             "sample": f"{sample} of {nptn} patterns per step (fraction {frac:.4f}), all {len(ops)} partial updates + root + derivative; "
                                       f"ms_per_step is that step as measured, value = fraction / seconds per step",
                             "sample_fraction": frac},
            "result": {"logl_sample": r["logl"], "df_sample": r["df"], "ddf_sample": r["ddf"]},
            "e2e": {"value": evals, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def cpu_sample_size(nops, B, nstates, ntaxa, budget_flop, nptn):
    """patterns in the bounded CPU sample: ~budget_flop of kernel work per eval, capped so that the reference's
    per-node CLVs (LM_PER_NODE: (ntaxa-2) x sample x B doubles) stay under a quarter of the free host memory"""
    sample = max(256, int(budget_flop / (nops * B * nstates * 6)))
    try:
        import psutil
        cap = int(psutil.virtual_memory().available * 0.25 / (max(ntaxa - 2, 1) * B * 8))
        sample = min(sample, max(256, cap))
    except Exception:
        sample = min(sample, 50_000)
    return min(sample, nptn) // 8 * 8 or min(sample, nptn)


def dram_model_bytes(ops, nptn, B, S, chained):
    """DRAM-side model of one walker launch: every CLV (+ scale counts) written once, internal children that are not consumed
    from registers / the same tile's previous op read once, one byte per tip child.  An upper bound of the traffic at the DRAM
    pins (sibling reads that hit L2 never reach them) — `traffic` is the measured figure."""
    clv = 8 * B + 2 * S
    total, prev = 0, -1
    for op in ops:
        total += clv
        kids = [op["left"], op["right"]]
        skipped = False
        for k in kids:
            if k["kind"] == plan.KIND_SLOT:
                if chained and not skipped and int(k["id"]) == prev:
                    skipped = True      # register chaining (4-state walker): the previous op's CLV never leaves the thread
                else:
                    total += clv
            else:
                total += 1
        prev = int(op["dst_slot"])
    return float(total) * nptn


def run_tree_workload(args, rank, world, local_rank, workload, ntaxa, nptn, model, desc, workload_name, steps, warmup, cpu_baseline=True):
    """One single-tree workload (BASELINE configs[1..3]) on this rank's pattern shard; returns the JSON line (complete on rank 0)."""
    import torch
    import torch.distributed as dist
    from iqtree2_b200.engine import LikelihoodEngine
    safe = workload in SAFE_WORKLOADS
    B, S = model.nstates * model.ncat_mix, (model.ncat_mix if safe else 1)
    overridden = bool(args.taxa or args.patterns or args.ncat)
    # pattern shard of this rank: contiguous, multiple of the 128-pattern tile except for the last rank
    per = (nptn + world - 1) // world
    per = (per + 127) // 128 * 128
    lo, hi = min(rank * per, nptn), min((rank + 1) * per, nptn)
    my_nptn = hi - lo
    nonrev = getattr(model, "nonrev", None) is not None
    tree, tips_all, freq_all, data = make_inputs(ntaxa + (1 if nonrev else 0), nptn, model)   # rooted: the last leaf is the root leaf
    if nonrev:
        tips_all = np.ascontiguousarray(tips_all[:ntaxa])
        data += "; rooted at an extra (virtual) leaf"
    tips = np.ascontiguousarray(tips_all[:, lo:hi])
    freq = freq_all[lo:hi]
    model = models.site_slice(model, lo, hi)   # site-specific models: this rank's per-pattern eigensystems
    br = plan.pendant_branch_of(tree, 0)
    ops, a, b, L = plan.plan_branch_nonrev(tree, br) if nonrev else plan.plan_branch(tree, br)
    ops_acc = ops.copy()   # byte / flop accounting: child kinds without the non-reversible flags (the root leaf counts as a tip)
    for side in ("left", "right"):
        k = ops_acc[side]["kind"] & 3
        ops_acc[side]["kind"] = np.where(k == plan.KIND_ROOT, plan.KIND_TIP, k)

    # CLV slots: one per internal node (LM_PER_NODE) when that fits the device; otherwise the reference's -mem mode
    # (LM_MEM_SAVE, tree/phylotree.cpp:999-1021): the host mirror's planner recycles a bounded number of slots
    nslots, mem_mode = (ntaxa - 1 + 2 if nonrev else None), "LM_PER_NODE"
    slot_bytes = ((my_nptn + 127) // 128 * 128) * (8 * B + 2 * S)
    dev_total = torch.cuda.get_device_properties(local_rank).total_memory
    mem_frac = args.mem
    if mem_frac < 0 and (ntaxa + 4) * slot_bytes > 0.85 * dev_total:
        mem_frac = min(1.0, (0.70 * dev_total / slot_bytes - 4) / (ntaxa - 2))
    if mem_frac >= 0:
        ops, a, b, L, nslots = plan_mem_save(tree, model, ntaxa, mem_frac)
        mem_mode = f"LM_MEM_SAVE ({nslots - 2} of {ntaxa - 2} CLV slots recycled by the host planner)"

    eng = LikelihoodEngine(model, ntaxa, my_nptn, nslots=nslots, safe_numeric=safe, device=local_rank, schedule=args.schedule)
    # An explicit stream, made torch's current one: the engine queues on it (a null handle — torch's legacy default stream — would
    # mean "the engine's own stream" to b200lk_set_stream) and torch's NCCL collectives order themselves behind the CURRENT stream, so
    # the all-reduce of the scalars waits for the evaluation that produces them
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    if args.variant:
        eng.set_schedule(10 + args.variant)
    eng.set_stream(stream.cuda_stream)
    eng.set_tip_states(tips)
    eng.set_pattern_freq(freq)
    eng.set_pattern_invar(np.zeros(my_nptn))
    eng.set_profiling(True)

    scal = torch.zeros(3, dtype=torch.float64, device="cuda")
    pinned_model = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in
                    (model.eval, model.evec, model.inv_evec, model.mix_weight, model.cat_rate, model.cat_prop)]
    h2d = int(ops.nbytes) + int(sum(t.numel() * 8 for t in pinned_model))
    d2h = 3 * 8

    # what an optimisation step re-sends: the class-level eigensystem and the rates; the per-pattern eigensystems of a site-specific
    # model are fixed for the whole analysis (the PMSF profiles come from a guide tree) and stay resident like the alignment
    site = getattr(model, "site", None) is not None
    import dataclasses
    model_step = dataclasses.replace(model, site=None) if site else model

    def one_eval(upload_model, L=L):
        """clearAllPartialLH -> computeLikelihoodBranch -> computeLikelihoodDerv, host buffers in, scalars out."""
        if upload_model:
            eng.set_model(model_step)  # eigensystem + rates re-sent as a ModelFinder / BFGS step would
        # both calls are queued on the stream and waited for ONCE (the asynchronous forms of the C ABI); with several ranks the
        # scalars go from the engine's result slot to a device buffer, through ncclAllReduce and back in the same stream order:
        # no host round trip between the kernels and the collective
        eng.theta_reset()
        eng.enqueue_branch_logl(ops, a, b, L)
        eng.enqueue_branch_derv(ops[:0], a, b, L)
        if world > 1:
            eng.enqueue_results_dev(scal.data_ptr())
            dist.all_reduce(scal)
            ll, df, ddf = scal.tolist()          # the one synchronisation of the evaluation
            local = eng.finish()
        else:
            local = eng.finish()
            ll, df, ddf = local
        t1 = eng.last_timing()
        return (ll, df, ddf), t1, None, local

    for _ in range(warmup):
        res, _, _, _ = one_eval(True)

    def timed(upload_model):
        sampler = ClockSampler(local_rank)
        sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        t0 = time.perf_counter()
        dev_ms = partial_ms = 0.0
        launches0 = eng.launch_count
        for _ in range(steps):
            r, t1, t2, local = one_eval(upload_model)
            dev_ms += t1["total_ms"]
            partial_ms += t1["partials_kernel_ms"]
        ev1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        out = {"event_ms": ev0.elapsed_time(ev1), "wall_s": wall, "device_ms": dev_ms, "partials_kernel_ms": partial_ms,
               "launches": eng.launch_count - launches0, "result": r, "local": local, "clocks": clocks}
        if world > 1:
            t = torch.tensor([out["event_ms"], out["device_ms"], out["partials_kernel_ms"], wall * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["event_ms"], out["device_ms"], out["partials_kernel_ms"], w = t.tolist()
            out["wall_s"] = w / 1e3
        return out

    e2e_run = timed(True)     # host buffers in every step, H2D/D2H inside the timed region
    res_run = timed(False)    # inputs resident; device-time of the kernels only

    K = steps
    value = K / (res_run["event_ms"] * 1e-3)     # CUDA events around the K timed steps (inputs resident), max over ranks
    e2e_value = K / (e2e_run["event_ms"] * 1e-3)
    bytes_eval, bytes_partials = algorithmic_bytes(ops_acc, nptn, B, S, ntaxa, n_root_slots=1 + (b[0] == plan.KIND_SLOT))
    my_bytes_partials = bytes_partials * (my_nptn / nptn)
    peak, peak_src = load_peaks()
    ker_s = res_run["partials_kernel_ms"] * 1e-3 / K
    alg_gbps = my_bytes_partials / ker_s / 1e9
    n_ii, n_ti, n_tt = plan.count_node_cases(ops_acc)
    walk4 = args.schedule != 1 and model.nstates == 4 and model.nmix == 1 and not nonrev
    walk20 = args.schedule != 1 and model.nstates == 20
    if site:
        walk4 = walk20 = False
        kernel_name = ("k_site_walk20 (32-pattern group per CTA, per-pattern eigensystems resident in shared memory, FP64 MMA m8n8k4 over categories x children)"
                       if model.nstates == 20 and args.variant != 5 else
                       f"k_site_partials<{model.nstates}> (thread per pattern, per-pattern eigensystems streamed from HBM / L2)")
    elif nonrev:
        kernel_name = (f"k_nonrev_walk4<{model.ncat_mix}> (pattern-tile walker with register chaining, state-space CLVs, device-built transition matrices)"
                       if model.nstates == 4 and 1 <= model.ncat_mix <= 4 and args.variant != 5 else
                       f"k_nonrev_partials<{model.nstates}> (thread per pattern, state-space CLVs, device-built transition matrices)")
    elif walk20:
        kernel_name = "k_walk_n20v2 (pattern-tile walker, FP64 MMA m8n8k4 on state-space CLVs, cp.async rings)"
    elif walk4:
        kernel_name = ("k_walk_n4c4 (pattern-tile walker, 4 states x 4 categories)" if model.ncat_mix == 4 else
                       f"k_walk_n4<{model.ncat_mix}> (pattern-tile walker, 4 states x {model.ncat_mix} categories)")
    elif args.schedule != 1 and 8 < model.nstates <= 64:
        kernel_name = f"k_walk_gen<{(model.nstates + 7) // 8}> (pattern-tile walker, FP64 MMA m8n8k4, {model.nstates} states padded to {(model.nstates + 7) // 8 * 8})"
    else:
        kernel_name = "k_partials_generic"
    nonrev_walk = nonrev and model.nstates == 4 and 1 <= model.ncat_mix <= 4 and args.variant != 5
    model_bytes = dram_model_bytes(ops_acc, my_nptn, B, S, chained=walk4 or nonrev_walk)
    if site and not (model.nstates == 20 and args.variant != 5):
        # thread-per-pattern kernel: every op reads the pattern's (n + 2 n^2) doubles once (re-reads across categories are L1 / L2 hits at best)
        model_bytes += len(ops) * my_nptn * 8 * (model.nstates + 2 * model.nstates ** 2)
    elif site:   # k_site_walk20: the eigensystems are read once per evaluation
        model_bytes += my_nptn * 8 * (model.nstates + 2 * model.nstates ** 2)
    traffic, traffic_src = load_traffic(workload, overridden or bool(args.variant or args.schedule), my_nptn)
    # roofline.frac is PHYSICAL: bytes at the DRAM pins (the committed ncu capture of this workload, else the DRAM-side model)
    # per launch / launch time / measured copy peak.  The SURVEY 8(d) algorithmic figure (every CLV written once AND read once)
    # is kept beside it: it exceeds the pins' traffic because chained / L2-resident CLVs are never re-read.
    phys_bytes = traffic if traffic else model_bytes
    roofline = {"bound": "hbm", "kernel": kernel_name, "achieved": phys_bytes / ker_s / 1e9, "peak": peak, "unit": "GB/s",
                "frac": phys_bytes / ker_s / 1e9 / peak, "frac_basis": "measured DRAM traffic (ncu)" if traffic else "DRAM-side model (upper bound of the traffic)",
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launch_ms": ker_s * 1e3,
                "dram_model_bytes_per_launch": model_bytes, "dram_model_GBps": model_bytes / ker_s / 1e9,
                "algorithmic_bytes_per_launch": my_bytes_partials, "algorithmic_equivalent_GBps": alg_gbps, "partial_lh_GBps": alg_gbps}
    if model.nstates >= 20 and not site:
        # 20-state updates: the FP64-MMA pipe and HBM are both within reach of binding (~5 flop/B at a ridge of ~5.7).
        # `tensor` = DMMA flops the kernel EXECUTES (state-space form: 30 / 15 / 0 DMMAs of 512 flops per 8 patterns x category
        # for inner-inner / tip-inner / tip-tip nodes) over the measured DMMA peak; the reference-form flops (SURVEY 8d:
        # C(6n^2+n) ...) over the same time are the `algorithmic_equivalent`.
        fpeak, fsrc = load_fp64_peak()
        flops_alg = algorithmic_flops(ops, my_nptn, model.nstates, model.ncat_mix)
        steps8 = (my_nptn + 7) // 8 * model.ncat_mix
        nt8 = (model.nstates + 7) // 8
        per_child = 15 if model.nstates == 20 else nt8 * 2 * nt8     # general n: NT8 state tiles x n8/4 K-steps per mat-vec
        flops_exec = 512.0 * steps8 * per_child * (2 * n_ii + n_ti)
        roofline["tensor"] = {"executed_TFLOPs": flops_exec / ker_s / 1e12, "peak": fpeak, "unit": "TFLOP/s", "frac": flops_exec / ker_s / 1e12 / fpeak,
                              "peak_source": fsrc, "executed_flops_per_launch": flops_exec,
                              "algorithmic_flops_per_launch": flops_alg, "algorithmic_equivalent_TFLOPs": flops_alg / ker_s / 1e12,
                              "algorithmic_equivalent_frac": flops_alg / ker_s / 1e12 / fpeak}
        if roofline["tensor"]["frac"] > roofline["frac"]:
            roofline["bound"] = "tensor"
    if site and model.nstates == 20 and args.variant != 5:
        # k_site_walk20 shares the FP64 pipe between its DMMAs (30 per pattern, node and 4-category chunk) and the exponentials
        # (160 per pattern, node and chunk, ~45 FP64 instructions each): the DMMA flops alone over the measured DMMA peak
        fpeak, fsrc = load_fp64_peak()
        chunks = (model.ncat_mix + 3) // 4
        flops_exec = 512.0 * 30 * chunks * my_nptn * len(ops)
        roofline["tensor"] = {"executed_TFLOPs": flops_exec / ker_s / 1e12, "peak": fpeak, "unit": "TFLOP/s", "frac": flops_exec / ker_s / 1e12 / fpeak,
                              "peak_source": fsrc, "executed_flops_per_launch": flops_exec,
                              "note": "DMMA flops only; the exp() calls of the same launch cost about as many FP64-pipe cycles again"}
        if roofline["tensor"]["frac"] > roofline["frac"]:
            roofline["bound"] = "tensor"
    # ---- N > 1: the sharded result checks itself -------------------------------------------------------------------
    check = None
    if world > 1:
        parts = torch.zeros(world * 3, dtype=torch.float64, device="cuda")
        parts[rank * 3:(rank + 1) * 3] = torch.tensor(res_run["local"], dtype=torch.float64)
        dist.all_reduce(parts)
        per_rank = parts.view(world, 3).cpu().numpy()
        tot = per_rank.sum(axis=0)
        ok_sum = bool(abs(tot[0] - res_run["result"][0]) <= 1e-9 * abs(tot[0]))
        # every timed step evaluates the same thing, so a collective that read the scalars of the PREVIOUS step would pass the test
        # above: one more evaluation at another length of the evaluation branch, reduced and summed the same two ways
        (ll2, df2, ddf2), _, _, local2 = one_eval(False, L * 1.37)
        parts2 = torch.zeros(world * 3, dtype=torch.float64, device="cuda")
        parts2[rank * 3:(rank + 1) * 3] = torch.tensor(local2, dtype=torch.float64)
        dist.all_reduce(parts2)
        tot2 = parts2.view(world, 3).cpu().numpy().sum(axis=0)
        ok_fresh = bool(abs(tot2[0] - ll2) <= 1e-9 * abs(tot2[0]) and abs(tot2[1] - df2) <= 1e-7 * max(1.0, abs(tot2[1])) and abs(ll2 - res_run["result"][0]) > 1e-6)
        ok_sum = ok_sum and ok_fresh
        one_eval(False)   # back to the workload's evaluation (the per-pattern values below are its)
        # an independent engine on rank 0 recomputes the first patterns of the LAST rank's shard
        sl = int(min(2048, nptn - min((world - 1) * per, nptn)))
        lh_last = torch.zeros(max(sl, 1), dtype=torch.float64, device="cuda")
        if rank == world - 1 and sl > 0:
            lh_last.copy_(torch.from_numpy(eng.get_pattern_lh()[:sl].copy()))
        dist.broadcast(lh_last, src=world - 1)
        slice_diff = None
        if rank == 0 and sl > 0:
            lo2 = min((world - 1) * per, nptn)
            e2 = LikelihoodEngine(model, ntaxa, sl, nslots=nslots, device=local_rank)
            e2.set_tip_states(np.ascontiguousarray(tips_all[:, lo2:lo2 + sl])); e2.set_pattern_freq(freq_all[lo2:lo2 + sl]); e2.set_pattern_invar(np.zeros(sl))
            e2.branch_logl(ops, a, b, L)
            ref = e2.get_pattern_lh()
            slice_diff = float(np.max(np.abs(ref - lh_last.cpu().numpy()) / np.abs(ref)))
            e2.close()
        check = {"per_rank_logl": [float(x) for x in per_rank[:, 0]], "sum_of_ranks_equals_allreduce": ok_sum,
                 "fresh_evaluation_at_another_length": {"allreduced_logl": float(ll2), "sum_of_ranks": float(tot2[0]), "ok": ok_fresh},
                 "slice_recomputed_on_rank0": {"rank": world - 1, "patterns": sl, "max_rel_diff": slice_diff}}
        if rank == 0:
            assert ok_sum, ("all-reduced logL differs from the sum of the per-rank values", tot[0], res_run["result"][0])
            assert slice_diff is None or slice_diff < 1e-11, ("sharded per-pattern logL differs from an independent engine", slice_diff)
    exp = None if overridden else load_expected(workload)
    if exp is not None:
        d = abs(res_run["result"][0] - exp)
        check = dict(check or {}, expected_logl=exp, expected_abs_diff=d)
        if rank == 0:
            assert d <= 1e-9 * abs(exp), ("logL differs from the recorded value of this workload", res_run["result"][0], exp)
    line = {
        "metric": "full_tree_logl_deriv_evals_per_sec", "value": value, "unit": "evals/s", "n_gpus": world, "steps": K,
        "warmup": warmup, "ms_per_step": res_run["event_ms"] / K, "kernel_ms_per_step": res_run["device_ms"] / K,
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": data,
        "config": shared_config(workload_name, ntaxa, nptn, model, desc, len(ops), safe=safe),
        "detail": {"patterns_per_gpu": my_nptn, "node_cases": {"inner_inner": n_ii, "tip_inner": n_ti, "tip_tip": n_tt},
                   "schedule": {0: "auto(walker)", 1: "levelled", 2: "walker"}.get(args.schedule, f"variant{args.schedule}"),
                   "clv_bytes_resident": eng.device_bytes, "clv_slots": mem_mode,
                   "parallelism": f"pattern-sharded x{world}" if world > 1 else "single GPU", "algorithmic_GB_per_eval": bytes_eval / 1e9},
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_run["event_ms"] / K, "wall_ms_per_step": e2e_run["wall_s"] * 1e3 / K,
                "what": "set_model+set_rates (eigensystem, rates), op list with branch lengths from host memory, "
                        "branch_logl + branch_derv through the C ABI, (logL, df, ddf) read back each step"},
        "gpu_launches": int(res_run["launches"]),
        "clocks": res_run["clocks"],
        "roofline": roofline,
        "result": {"logl": res_run["result"][0], "df": res_run["result"][1], "ddf": res_run["result"][2]},
    }
    if check:
        line["check"] = check
    eng.close()
    del eng
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and cpu_baseline:
        sample = args.cpu_sample or cpu_sample_size(len(ops), B, model.nstates, ntaxa, 2.0e10, my_nptn)
        r = run_cpu_reference(tree, model, tips, freq, ops, a, b, L, sample, 5, 1, safe=safe)
        cpu_evals = (sample / nptn) / r["sec_per_sample_eval"]
        line["cpu_baseline"] = {"value": cpu_evals, "unit": "evals/s", "cores": r["cores"], "kind": r["kind"], "what": r["what"],
                                "sample": f"first {sample} of {nptn} patterns, all {len(ops)} partial updates + root + derivative, "
                                          f"5 timed evals after 1 warm-up; throughput scaled by {sample}/{nptn}",
                                "sample_fraction": sample / nptn, "sec_per_sample_eval": r["sec_per_sample_eval"]}
    return line


def plan_mem_save(tree, model, ntaxa, mem_frac):
    """the op list of the headline evaluation under the reference's -mem mode, planned by the host mirror (plan-only tree)"""
    from iqtree2_b200.phylotree import PhyloTree
    pt = PhyloTree()
    pt.set_plan_only(True)
    pt.set_alignment(tree.names, np.zeros((ntaxa, 1), dtype=np.uint8), np.ones(1), model)   # planning needs no patterns
    pt.set_model(model)
    pt.read_tree(tree.to_newick(fmt="%.17g"))   # the exact branch lengths
    pt.set_mem_save(True, mem_frac)
    nslots = pt.max_lh_slots + 2
    pa, pb, _ = pt.branches()
    pbr = [k for k in range(len(pa)) if min(pa[k], pb[k]) == 0][0]   # the pendant branch of alignment row 0
    ops, a, b, L = pt.plan_branch(pbr)
    assert len(ops) == ntaxa - 2
    pt.close()
    return ops, a, b, L, nslots


def load_expected(workload):
    """logL of the full workload recorded from a single-GPU run (profiles/expected_logl.json): every N must reproduce it"""
    p = os.path.join(ROOT, "profiles", "expected_logl.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(workload)



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS) + [SUPER_WORKLOAD])
    ap.add_argument("--parts", type=int, default=0, help="override the number of partitions of the partition workload (debug)")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads that queue partition work (0 = OpenMP default)")
    ap.add_argument("--taxa", type=int, default=0, help="override taxa (debug)")
    ap.add_argument("--patterns", type=int, default=0, help="override patterns (debug)")
    ap.add_argument("--ncat", type=int, default=0, help="override the number of Gamma categories (debug; single-class workloads)")
    ap.add_argument("--schedule", type=int, default=0, help="0 auto, 1 levelled, 2 walker")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant of the walker schedule (development A/B switch)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="patterns in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="default workload only: skip the LG+G4 `secondary` block")
    ap.add_argument("--mem", type=float, default=-1.0,
                    help="-mem of the reference (LM_MEM_SAVE): fraction of the T-2 CLV slots to keep (0 = log2(T)+1 slots); "
                         "default: only when one slot per internal node does not fit the device")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.workload == SUPER_WORKLOAD:
        return run_super(args, rank, world, local_rank)
    ntaxa, nptn, mk, desc = WORKLOADS[args.workload]
    if args.taxa:
        ntaxa = args.taxa
    if args.patterns:
        nptn = args.patterns
    model = mk()
    if args.ncat:
        model = {"dna_1000x1M_gtr_g4": models.make_gtr_gamma, "aa_200x100k_lg_g4": models.make_lg_gamma}[args.workload](ncat=args.ncat)
    B, S = model.nstates * model.ncat_mix, 1
    workload_name = args.workload if not (args.taxa or args.patterns or args.ncat) else f"{args.workload}[override {ntaxa}x{nptn} ncat={model.ncat_mix}]"

    # ---------------------------------------------------------------------------------------------------------
    # reference arm: the reference's own CPU algorithm on the host cores (rank 0 only)
    # ---------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        line = run_reference_arm(args, args.workload, ntaxa, nptn, model, desc, workload_name)
        if args.workload == DEFAULT_WORKLOAD and not (args.taxa or args.patterns or args.ncat) and not args.no_secondary:
            ntaxa2, nptn2, mk2, desc2 = WORKLOADS[SECONDARY_WORKLOAD]
            line["secondary"] = run_reference_arm(args, SECONDARY_WORKLOAD, ntaxa2, nptn2, mk2(), desc2, SECONDARY_WORKLOAD, steps=min(args.steps, 5), warmup=1)
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------------------------------------------------
    # this framework
    # ---------------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the likelihood engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = run_tree_workload(args, rank, world, local_rank, args.workload, ntaxa, nptn, model, desc, workload_name, args.steps, args.warmup,
                             cpu_baseline=not args.no_cpu_baseline)
    if args.workload == DEFAULT_WORKLOAD and not (args.taxa or args.patterns or args.ncat or args.variant or args.schedule) and not args.no_secondary:
        # the FP64-MMA path (BASELINE configs[2], named in north_star's target) measured in the same run: ~0.1 s of GPU time
        ntaxa2, nptn2, mk2, desc2 = WORKLOADS[SECONDARY_WORKLOAD]
        sec = run_tree_workload(args, rank, world, local_rank, SECONDARY_WORKLOAD, ntaxa2, nptn2, mk2(), desc2, SECONDARY_WORKLOAD,
                                args.steps, args.warmup, cpu_baseline=not args.no_cpu_baseline)
        line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
