#!/usr/bin/env bash
# build_ref.sh — TEST INFRASTRUCTURE.  Compiles the reference's OWN likelihood path (IQ-TREE 2.4.0) from the sources
# where they lie under /root/reference into oracle/_ref/libiqref.so, together with oracle/ref_harness.cpp.
# No reference source is copied; objects and the library go to oracle/_ref/ (git-ignored, travels with gpurun).
#
# Not the reference's build system: its CMakeLists hard-requires Eigen3 and Boost (CMakeLists.txt:84-96), absent
# here.  The likelihood path itself needs neither: only alignment/alignment.cpp includes <Eigen/LU> (for the
# symmetry tests at :5236-5320, never reached from the likelihood path); it is compiled against a declaration-only
# overlay (oracle/ref_stubs/Eigen/LU).  Flags follow the reference's kernelfma target (CMakeLists.txt:586, 622, 844):
# -O3 -mavx -mfma -D__SSE3 -D__AVX -DIQ_TREE, OpenMP on.
# Members of classes whose TUs are not compiled (tree search, parsimony, PLL, ModelMarkov, ...) are never called by
# the harness; they are bound to a trap (ref_unresolved.s, generated) so that the library loads with RTLD_NOW.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
R="${IQREF_SRC:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
[ -d "$R/tree" ] || { echo "build_ref.sh: $R not present; keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OBJ" "$OUT/inc"

cat > "$OUT/inc/iqtree_config.h" <<'EOF'
#define iqtree_VERSION_MAJOR 2
#define iqtree_VERSION_MINOR 4
#define iqtree_VERSION_PATCH "0"
#define HAVE_GETTIMEOFDAY
#define HAVE_GETRUSAGE
#define HAVE_STRNDUP
#define HAVE_STRTOK_R
EOF

TUS=(
  tree/phylokernelfma.cpp tree/phylotreesse.cpp tree/phylotree.cpp tree/mtree.cpp tree/node.cpp tree/phylonode.cpp tree/phylonodemixlen.cpp
  tree/memslot.cpp tree/constrainttree.cpp pda/split.cpp pda/hashsplitset.cpp
  alignment/alignment.cpp alignment/pattern.cpp alignment/alignmentsummary.cpp phylo-yaml/statespace.cpp
  model/modelsubst.cpp model/rateheterogeneity.cpp model/modelfactory.cpp
  utils/tools.cpp utils/optimization.cpp utils/checkpoint.cpp utils/progress.cpp utils/timeutil.cpp utils/operatingsystem.cpp
)
CXXFLAGS=(-std=c++17 -O3 -mavx -mfma -D__SSE3 -D__AVX -DIQ_TREE -fopenmp -fPIC -w -fno-access-control
          -I"$OUT/inc" -I"$HERE/ref_stubs" -I"$R" -I"$R/yaml-cpp/include" -I"$R/zlib-1.2.7" -I"$R/lbfgsb")

compile_one() {
  src="$1"; obj="$OBJ/$(echo "$src" | tr '/' '_' | sed 's/\.cpp$/.o/')"
  if [ ! -f "$obj" ] || [ "$R/$src" -nt "$obj" ]; then
    g++ "${CXXFLAGS[@]}" -c "$R/$src" -o "$obj"
  fi
}
export -f compile_one
export R OBJ
export CXXFLAGS_STR="${CXXFLAGS[*]}"
pids=()
for s in "${TUS[@]}"; do
  [ -f "$R/$s" ] || continue
  compile_one "$s" &
  pids+=($!)
done
g++ "${CXXFLAGS[@]}" -c "$HERE/ref_harness.cpp" -o "$OBJ/ref_harness.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done

# first link: see what stays unresolved, bind those names to a trap, link again
g++ -shared -fopenmp -o "$OUT/libiqref_trial.so" "$OBJ"/*.o -lz 2>/dev/null || g++ -shared -fopenmp -o "$OUT/libiqref_trial.so" "$OBJ"/*.o
ldd -r "$OUT/libiqref_trial.so" 2>&1 | sed -n 's/^undefined symbol: \([^ \t]*\).*/\1/p' | sort -u > "$OUT/unresolved.txt"
{
  echo '.text'
  i=0
  while read -r sym; do
    [ -n "$sym" ] || continue
    printf '.weak %s\n%s:\n  leaq .Lname%d(%%rip), %%rdi\n  jmp iqref_unresolved_trap@PLT\n' "$sym" "$sym" "$i"
    i=$((i+1))
  done < "$OUT/unresolved.txt"
  echo '.section .rodata'
  i=0
  while read -r sym; do
    [ -n "$sym" ] || continue
    printf '.Lname%d: .string "%s"\n' "$i" "$sym"
    i=$((i+1))
  done < "$OUT/unresolved.txt"
  echo '.section .note.GNU-stack,"",@progbits'
} > "$OUT/ref_unresolved.s"
cat > "$OUT/ref_trap.c" <<'EOF'
#include <stdio.h>
#include <stdlib.h>
void iqref_unresolved_trap(const char *name) {
    fprintf(stderr, "libiqref: %s was called, but its translation unit is not compiled into the harness\n", name);
    abort();
}
EOF
gcc -fPIC -c "$OUT/ref_trap.c" -o "$OUT/stub_trap.o"
gcc -fPIC -c "$OUT/ref_unresolved.s" -o "$OUT/stub_unresolved.o"
g++ -shared -fopenmp -o "$OUT/libiqref.so" "$OBJ"/*.o "$OUT/stub_trap.o" "$OUT/stub_unresolved.o" $(ldd "$OUT/libiqref_trial.so" | grep -q libz && echo -lz)
rm -f "$OUT/libiqref_trial.so"
left=$(ldd -r "$OUT/libiqref.so" 2>&1 | grep -c '^undefined symbol' || true)
echo "build_ref.sh: $OUT/libiqref.so built ($(wc -l < "$OUT/unresolved.txt") names bound to the trap, $left still undefined)"

# The reference-side binding of INTEGRATION.md, compiled for real: the same objects + ref_b200_shim.cpp (a PhyloTree whose
# kernel pointers marshal into the product's C ABI), linked against iqtree2_b200/libb200lk.so when that has been built.
B200LIB="$HERE/../iqtree2_b200/libb200lk.so"
if [ -f "$B200LIB" ]; then
  g++ "${CXXFLAGS[@]}" -I"$HERE/../include" -c "$HERE/ref_b200_shim.cpp" -o "$OUT/ref_b200_shim.o"
  g++ -shared -fopenmp -o "$OUT/libiqref_b200.so" "$OBJ"/*.o "$OUT/ref_b200_shim.o" "$OUT/stub_trap.o" "$OUT/stub_unresolved.o" \
      -L"$HERE/../iqtree2_b200" -lb200lk -Wl,-rpath,'$ORIGIN/../../iqtree2_b200'
  left=$(ldd -r "$OUT/libiqref_b200.so" 2>&1 | grep -c '^undefined symbol' || true)
  echo "build_ref.sh: $OUT/libiqref_b200.so built ($left undefined)"
else
  echo "build_ref.sh: iqtree2_b200/libb200lk.so not built yet; skipping libiqref_b200.so"
fi

# The same binding over the CPU oracle (oracle/b200lk_over_oracle.c implements the C ABI it calls): the binding's own logic — slot
# numbers from partial_lh pointers, marshalling of traversal_info, theta_computed, +ASC columns — is then covered by the CPU tests.
if [ -f "$OUT/ref_b200_shim.o" ]; then
  gcc -O3 -march=x86-64-v3 -mfma -ffp-contract=off -fopenmp -fPIC -c "$HERE/lk_oracle.c" -o "$OUT/lk_oracle.o"
  gcc -O2 -fopenmp -fPIC -c "$HERE/b200lk_over_oracle.c" -o "$OUT/b200lk_over_oracle.o"
  g++ -shared -fopenmp -o "$OUT/libiqref_over_oracle.so" "$OBJ"/*.o "$OUT/ref_b200_shim.o" "$OUT/lk_oracle.o" "$OUT/b200lk_over_oracle.o" \
      "$OUT/stub_trap.o" "$OUT/stub_unresolved.o" -lm
  left=$(ldd -r "$OUT/libiqref_over_oracle.so" 2>&1 | grep -c '^undefined symbol' || true)
  echo "build_ref.sh: $OUT/libiqref_over_oracle.so built ($left undefined)"
fi
