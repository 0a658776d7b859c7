#!/bin/bash
# TEST INFRASTRUCTURE. Builds the reference's OPEN driver (gblanco92/GamBa: source/{gamba,io,f4,params,stats,debug,getRSS}.cpp
# and the headers they include, compiled from where they lie under $GAMBA_REFERENCE) against gamba-b200's C ABI through the
# two adapter headers of integration/reference_adapter/ (the reference's own linalg_v3.hpp / linalg.hpp are withheld):
#   oracle/_ref/gamba_ref        linked with gamba_b200/libgamba_cuda.so      (runs on the GPU box: tests/test_gpu_ref_adapter.py)
#   oracle/_ref/gamba_ref_cpu    linked with oracle/_ref/libcabi_oracle.so, the CPU oracle behind the same C ABI
#                                (checks the adapters without a GPU: tests/test_ref_adapter_cpu.py)
# Follows SURVEY.md Appendix B. Nothing of the reference is copied or patched: every reference source is compiled where it
# lies; outputs are binaries under oracle/_ref/ (git-ignored), object files go to a temporary directory.
set -eu
cd "$(dirname "$0")/.."
REF=${GAMBA_REFERENCE:-/root/reference}
[ -f "$REF/source/f4.hpp" ] || { echo "build_ref: no reference sources at $REF (nothing to do)"; exit 0; }
OUT=oracle/_ref
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
CXX=${CXX:-g++}
# include order: the adapters, the shims (config.hpp, kernel/find.hpp, gmpxx.h), then the reference's own directory
FLAGS="-std=c++23 -O2 -march=x86-64-v3 -w -include oracle/ref_shims/prefix.hpp -Iintegration/reference_adapter -Ioracle/ref_shims -Iinclude -I$REF/source"
OBJS=""
for f in gamba io f4 params stats debug getRSS; do
  $CXX $FLAGS -c "$REF/source/$f.cpp" -o "$TMP/$f.o" &
  OBJS="$OBJS $TMP/$f.o"
done
$CXX $FLAGS -c oracle/ref_shims/ref_main.cpp -o "$TMP/ref_main.o" &
$CXX -std=c++20 -O2 -march=x86-64-v3 -fPIC -shared -pthread -Iinclude -o "$OUT/libcabi_oracle.so" oracle/ref_shims/cabi_oracle_shim.cpp oracle/oracle_engine.cpp oracle/oracle_checker.cpp &
wait
$CXX -o "$OUT/gamba_ref_cpu" $OBJS "$TMP/ref_main.o" -L"$OUT" -lcabi_oracle -pthread -Wl,-rpath,'$ORIGIN'
if [ -f gamba_b200/libgamba_cuda.so ]; then
  $CXX -o "$OUT/gamba_ref" $OBJS "$TMP/ref_main.o" -Lgamba_b200 -lgamba_cuda -pthread -Wl,-rpath,'$ORIGIN/../../gamba_b200'
fi
ls -la "$OUT"
