#!/bin/sh
# Build libstacb.so (sm_100a) in-tree; one translation unit per kernel variant, compiled in parallel.
# -fmad=false: fused multiply-adds are explicit in the source (canonical arithmetic).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -I ../../include ${STACB_NVCC_EXTRA}"
mkdir -p _obj
rm -f _obj/*.o
VARIANTS=$(sed -n 's/^#define STACB_VARIANTS(X)//p' stacb_variants.h | sed 's/X(\([0-9]*\), *\([0-9]*\), *\([0-9]*\), *\([0-9]*\), *\([0-9]*\))/\1_\2_\3_\4_\5/g')
pids=""
for v in $VARIANTS; do
  set -- $(echo $v | tr '_' ' ')
  $NVCC $FLAGS -DV_CPL=$1 -DV_NB=$2 -DV_NBF=$3 -DV_SPL=$4 -DV_JM=$5 -c stacb_variant.cu -o _obj/variant_$v.o > _obj/variant_$v.log 2>&1 &
  pids="$pids $!"
done
FAST=$(sed -n 's/^#define STACB_FAST_VARIANTS(X)//p' stacb_variants.h | sed 's/X(\([0-9]*\), *\([0-9]*\), *\([0-9]*\))/\1_\2_\3/g')
for v in $FAST; do
  set -- $(echo $v | tr '_' ' ')
  $NVCC $FLAGS -Xptxas -v -DV_FJM=$1 -DV_FRT=$2 -DV_FNBF=$3 -c stacb_fast_variant.cu -o _obj/fast_$v.o > _obj/fast_$v.log 2>&1 &
  pids="$pids $!"
done
for w in 2 4 6 8; do
  $NVCC $FLAGS -Xptxas -v -DV_WW=$w -c stacb_wide_variant.cu -o _obj/wide_$w.o > _obj/wide_$w.log 2>&1 &
  pids="$pids $!"
done
$NVCC $FLAGS -c stacb_abi.cu -o _obj/abi.o > _obj/abi.log 2>&1 &
pids="$pids $!"
$NVCC $FLAGS -c stacb_post.cu -o _obj/post.o > _obj/post.log 2>&1 &
pids="$pids $!"
rc=0
for p in $pids; do wait $p || rc=1; done
cat _obj/*.log | grep -v -e '^ptxas info    : Function properties' -e 'bytes stack frame' -e '^ptxas info    : Compiling' -e '^ptxas info    : 0 bytes gmem' || true
[ $rc -eq 0 ] || { echo "stacb build failed"; exit 1; }
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libstacb.so _obj/*.o
echo "built $(cd .. && pwd)/libstacb.so"
# optional: XLA FFI handlers for a JAX host, only where jaxlib's FFI headers exist (not in the authoring image)
XLA_INC=$(python -c "import jax.ffi; print(jax.ffi.include_dir())" 2>/dev/null || true)
if [ -n "$XLA_INC" ] && [ -d "$XLA_INC" ]; then
  g++ -O2 -std=c++17 -fPIC -shared -I "$XLA_INC" -I ../../include -I /usr/local/cuda/include stacb_xla_ffi.cc \
      -L.. -l:libstacb.so -Wl,-rpath,'$ORIGIN' -L/usr/local/cuda/lib64 -lcudart -o ../libstacb_xla_ffi.so \
    && echo "built $(cd .. && pwd)/libstacb_xla_ffi.so"
else
  echo "jaxlib FFI headers not found: skipping libstacb_xla_ffi.so (stacb_xla_ffi.cc is not compiled in this image)"
fi
