#!/usr/bin/env bash
# Builds the UNMODIFIED reference (oracle A) from the sources where they lie under
# $MEDGP_REFERENCE (default /root/reference) into oracle/_ref/ (git-ignored, travels with
# gpurun).  Nothing is copied out of the reference tree; only binaries are produced.
#   oracle/_ref/main_one_train.o, main_one_test.o  -- the reference executables
#   oracle/_ref/ref_eval                           -- one-evaluation driver (oracle/ref/ref_eval.cpp)
#   oracle/_ref/ref_scg                            -- reference SCG on an analytic objective
# BLAS/LAPACK: the LP64 OpenBLAS bundled in the scipy wheel via the mkl.h shim (NOT Intel MKL;
# g++ not icpc) -- stated wherever a number from these binaries is reported.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MEDGP_REFERENCE:-/root/reference}"
OUT="$HERE/../_ref"
SRC="$REF/medgpc/src"
if [ ! -d "$SRC" ]; then echo "reference not present at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; fi
SITE="$(python -c 'import scipy, os; print(os.path.dirname(os.path.dirname(scipy.__file__)))')"
BLASDIR="$SITE/scipy.libs"
BLAS="$(basename "$(ls "$BLASDIR"/libscipy_openblas*.so | head -1)")"
RJ="$SITE/tilelang/3rdparty/composable_kernel/include"
mkdir -p "$OUT/obj"
CXX="${MEDGP_CXX:-/usr/bin/g++}"   # the env CXX in this image (/opt/gcc) lacks libgomp.spec
FLAGS="-std=c++11 -O2 -fopenmp -fpermissive -w -I$HERE -I$SRC -I$RJ"
objs=()
for f in "$SRC"/*/*.cpp; do
  o="$OUT/obj/$(basename "${f%.cpp}").o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then $CXX $FLAGS -c "$f" -o "$o" & fi
  objs+=("$o")
done
wait
LINK="-L$BLASDIR -l:$BLAS -Wl,-rpath,$BLASDIR"
$CXX $FLAGS -o "$OUT/main_one_train.o" "${objs[@]}" "$SRC/main_one_train.cpp" $LINK &
$CXX $FLAGS -o "$OUT/main_one_test.o" "${objs[@]}" "$SRC/main_one_test.cpp" $LINK &
$CXX $FLAGS -o "$OUT/ref_eval" "${objs[@]}" "$HERE/ref_eval.cpp" $LINK &
if [ -f "$HERE/ref_scg.cpp" ]; then $CXX $FLAGS -o "$OUT/ref_scg" "${objs[@]}" "$HERE/ref_scg.cpp" $LINK & fi
wait
echo "built: $(ls "$OUT" | tr '\n' ' ')"
