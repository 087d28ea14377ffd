#!/bin/bash
# A/B of alternative builds of the library (FORDE_LIB): isolated layer benchmark per build, then the GPU parity
# suite against each non-default build.  usage: [SIGN_BITS=1] tools/ab_libs.sh out.log lib1.so [lib2.so ...]
# SIGN_BITS=1 runs everything with the opt-in sign-bit backward (layer_bench --sign-bits 1, FORDE_SIGN_BITS=1).
# Example (the round-2 candidate of DESIGN.md section 7, built in-tree so it travels to the GPU box):
#   mkdir -p build && (cd forde-vision-language_b200 && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a \
#     -lineinfo -Xcompiler -fPIC -shared --threads 4 -DFORDE_UP_PACKED=1 -DFORDE_SIGN_WORD_MAJOR=1 -DFORDE_UP_DBL_H=1 \
#     -o ../build/libforde_r2.so csrc/forde_api.cu csrc/forde_aux.cu csrc/forde_norm.cu csrc/forde_moe.cu -lcuda)
#   gpurun -- 'SIGN_BITS=1 tools/ab_libs.sh gpurun_out/ab.log build/libforde_r2.so'
out=$1; shift
sb=${SIGN_BITS:-0}
export FORDE_SIGN_BITS=$sb
: > "$out"
for a in zeros blocks mixed; do
  echo "== shipped ($a, sign_bits=$sb)" >> "$out"
  env -u FORDE_LIB timeout 120 python tools/layer_bench.py --assign $a --sign-bits $sb >> "$out" 2>&1
  for lib in "$@"; do
    echo "== $lib ($a, sign_bits=$sb)" >> "$out"
    FORDE_LIB=$PWD/$lib timeout 120 python tools/layer_bench.py --assign $a --sign-bits $sb >> "$out" 2>&1
  done
done
for lib in "$@"; do
  echo "== parity suite with $lib (FORDE_SIGN_BITS=$sb)" >> "$out"
  FORDE_LIB=$PWD/$lib timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> "$out"
done
