#!/usr/bin/env bash
# Builds the UNMODIFIED reference (oracle A) from the sources where they lie under
# $MEDGP_REFERENCE (default /root/reference) into oracle/_ref/ (git-ignored, travels with
# gpurun).  Nothing is copied out of the reference tree; only binaries are produced.
#   oracle/_ref/main_one_train.o, main_one_test.o  -- the reference executables
#   oracle/_ref/ref_eval                           -- one-evaluation driver (oracle/ref/ref_eval.cpp)
#   oracle/_ref/ref_scg                            -- reference SCG on an analytic objective
#   oracle/_ref/main_one_train_cuda.o, main_one_test_cuda.o, ref_eval_cuda -- the SAME unmodified sources with the
#       reference-side binding c_inference_cuda (oracle/ref/c_inference_cuda.{h,cpp}) taking the
#       place of c_inference_prior, linked against medgp_b200/libmedgp_cuda.so: the drop-in
#       boundary proven by compilation
#   oracle/_ref/main_one_{train,test}_orb.o, ref_eval_orb  -- the same with the oracle behind the C ABI
#       (oracle/_build/liborb.a), so the binding itself is testable without a GPU
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
# ---- the drop-in boundary: reference sources + c_inference_cuda, no source modified.
# -include puts the binding's header first; its trailing macro (MEDGP_REPLACE_C_INFERENCE_PRIOR)
# turns the `c_inference_prior inffunc(thread_num)` of run_model_LMC_SM into a c_inference_cuda.
ROOT="$HERE/../.."
BIND="-I$ROOT/include -include $HERE/c_inference_cuda.h -DMEDGP_REPLACE_C_INFERENCE_PRIOR"
$CXX $FLAGS -I"$ROOT/include" -c "$HERE/c_inference_cuda.cpp" -o "$OUT/obj/c_inference_cuda.o"
CUDALIB="$ROOT/medgp_b200"
if [ -f "$CUDALIB/libmedgp_cuda.so" ]; then
  CUDALINK="-L$CUDALIB -lmedgp_cuda -Wl,-rpath,\$ORIGIN/../../medgp_b200 -Wl,-rpath,/usr/local/cuda/lib64 -Wl,-rpath-link,/usr/local/cuda/lib64"
  $CXX $FLAGS $BIND -o "$OUT/main_one_train_cuda.o" "${objs[@]}" "$OUT/obj/c_inference_cuda.o" "$SRC/main_one_train.cpp" $LINK $CUDALINK &
  $CXX $FLAGS $BIND -o "$OUT/main_one_test_cuda.o" "${objs[@]}" "$OUT/obj/c_inference_cuda.o" "$SRC/main_one_test.cpp" $LINK $CUDALINK &
  $CXX $FLAGS $BIND -o "$OUT/ref_eval_cuda" "${objs[@]}" "$OUT/obj/c_inference_cuda.o" "$HERE/ref_eval.cpp" $LINK $CUDALINK &
fi
ORB="$HERE/../_build/liborb.a"
if [ -f "$ORB" ]; then
  $CXX $FLAGS $BIND -o "$OUT/main_one_train_orb.o" "${objs[@]}" "$OUT/obj/c_inference_cuda.o" "$SRC/main_one_train.cpp" "$ORB" $LINK &
  $CXX $FLAGS $BIND -o "$OUT/main_one_test_orb.o" "${objs[@]}" "$OUT/obj/c_inference_cuda.o" "$SRC/main_one_test.cpp" "$ORB" $LINK &
  $CXX $FLAGS $BIND -o "$OUT/ref_eval_orb" "${objs[@]}" "$OUT/obj/c_inference_cuda.o" "$HERE/ref_eval.cpp" "$ORB" $LINK &
fi
wait
echo "built: $(ls "$OUT" | tr '\n' ' ')"
