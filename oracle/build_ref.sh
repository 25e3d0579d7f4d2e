#!/usr/bin/env bash
# oracle/build_ref.sh -- compile the UNMODIFIED reference from the sources where they lie
# (/root/reference) into oracle/_ref/ (git-ignored; ships to the GPU box with the snapshot).
# No reference source is copied into this repository.
#
#   oracle/_ref/llama2_q4_ref   the reference program itself (stock main, stock code path)
#   oracle/_ref/libq4ref.so     ref_harness.cu: `#define main ref_main` + #include of the
#                               reference TU, exporting its host wrappers through a C ABI so
#                               tests can run reference kernels on the same device buffers
#   oracle/_ref/weight_packer   the reference offline packer (format cross-check)
#
# `-include float.h` is needed because gpu_kernels.h:522 uses FLT_MAX without including it
# (SURVEY.md section 8c).  The stock CMake flow cannot be used in this image (GPU autodetect).
set -euo pipefail
REF=${REFERENCE_DIR:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -f "$REF/llama2_q4.cu" ]; then
  echo "build_ref: $REF not present; keeping prebuilt files in $OUT (if any)"; exit 0
fi
mkdir -p "$OUT"
NVCC=${NVCC:-nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
up_to_date() { [ -f "$1" ] && [ "$1" -nt "$REF/llama2_q4.cu" ] && [ "$1" -nt "$REF/gpu_kernels.h" ] && [ "$1" -nt "$HERE/ref_harness.cu" ] && [ "$1" -nt "$0" ]; }
if ! up_to_date "$OUT/llama2_q4_ref"; then
  $NVCC -O3 $ARCH -include float.h -o "$OUT/llama2_q4_ref" "$REF/llama2_q4.cu"
fi
if ! up_to_date "$OUT/libq4ref.so"; then
  $NVCC -O3 $ARCH -include float.h -I"$REF" -shared -Xcompiler -fPIC -o "$OUT/libq4ref.so" "$HERE/ref_harness.cu"
fi
if ! up_to_date "$OUT/weight_packer"; then
  g++ -O2 -o "$OUT/weight_packer" "$REF/weight_packer.cpp"
fi
echo "build_ref: ok -> $OUT"
