#!/bin/bash
# A/B of alternative builds of the library (FORDE_LIB): isolated layer benchmark per build, then the GPU parity
# suite against each non-default build.  usage: tools/ab_libs.sh out.log lib1.so [lib2.so ...]
out=$1; shift
: > "$out"
for a in zeros mixed; do
  echo "== shipped ($a)" >> "$out"
  env -u FORDE_LIB timeout 120 python tools/layer_bench.py --assign $a >> "$out" 2>&1
  for lib in "$@"; do
    echo "== $lib ($a)" >> "$out"
    FORDE_LIB=$PWD/$lib timeout 120 python tools/layer_bench.py --assign $a >> "$out" 2>&1
  done
done
for lib in "$@"; do
  echo "== parity suite with $lib" >> "$out"
  FORDE_LIB=$PWD/$lib timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 >> "$out"
done
