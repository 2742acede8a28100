#!/bin/bash
# AddressSanitizer over the product's C++ host mirrors (phylotree_b200.cpp, supertree_b200.cpp, phylotree_capi.cpp) linked against
# the CPU oracle through oracle/b200lk_over_oracle.c (test infrastructure): the tree-level tests — NNI evaluation on the spare
# slots with its neighbour copies and swaps, doNNI, -mem slot recycling, tree-length scaling, both partition drivers — run on the
# instrumented build.  No GPU needed.  Output: "asan run complete" or the first report.
# SAN=undefined bash scripts/asan_host_mirror.sh runs the same under UndefinedBehaviorSanitizer.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
B="$ROOT/oracle/_build/asan"; mkdir -p "$B"
SAN="${SAN:-address}"
F="-O1 -g -fsanitize=$SAN -fno-omit-frame-pointer -fopenmp -fPIC"
[ "$SAN" = undefined ] && F="$F -fno-sanitize-recover=undefined"
RT=$([ "$SAN" = undefined ] && echo libubsan.so || echo libasan.so)
gcc $F -march=x86-64-v3 -mfma -ffp-contract=off -c "$ROOT/oracle/lk_oracle.c" -o "$B/lk_oracle.o"
gcc $F -c "$ROOT/oracle/b200lk_over_oracle.c" -o "$B/adapter.o"
g++ -std=c++17 $F -I"$ROOT/include" -I"$ROOT/iqtree2_b200/csrc" -shared -o "$B/libb200pt_over_oracle.so" \
    "$ROOT"/iqtree2_b200/csrc/{phylotree_b200,supertree_b200,phylotree_capi}.cpp "$B/lk_oracle.o" "$B/adapter.o" -lm
cat > "$B/run.py" <<PY
import ctypes, os, sys
ROOT = "$ROOT"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
lib = ctypes.CDLL(os.path.join(ROOT, "oracle/_build/asan/libb200pt_over_oracle.so"))
import iqtree2_b200.phylotree as pt, iqtree2_b200.supertree as st
pt.load_library = lambda: lib
st.load_library = lambda: lib
import test_gpu_nni as n, test_gpu_phylotree as p, test_gpu_supertree as s, test_gpu_supertree_nni as sn, test_gpu_supertree_unlinked as u
n.test_nni_against_reference_golden(True)
n.test_nni_evaluation_on_resident_clvs("dna", False, False)
n.test_do_nni_applies_the_evaluated_move(True)
p.test_config1_branch_length_optimisation(True)
p.test_mem_save_slot_recycling_on_gpu("dna", 0, 0.0)
p.test_tree_length_scaling("dna", 1.0)
s.test_supertree_branch_length_optimisation()
s.test_supertree_incremental_after_one_branch()
sn.test_single_partition_equals_plain_tree(True)
sn.test_partitioned_nni_scores_are_the_likelihoods_of_the_moved_trees()
u.test_unlinked_partitions(2)
print("asan run complete")
PY
LD_PRELOAD=$(gcc -print-file-name=$RT) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
    OMP_NUM_THREADS=2 python "$B/run.py"
