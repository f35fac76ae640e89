#!/bin/sh
# Build the CPU oracle (float32 and float64 twins). Test infrastructure only.
set -e
cd "$(dirname "$0")"
mkdir -p _build
FLAGS="-O2 -fPIC -shared -ffp-contract=off -mfma -fopenmp -fno-fast-math -Wall -Wno-unused-function"
gcc $FLAGS -DREAL=float  stac_oracle.c -o _build/liboracle_f32.so -lm
gcc $FLAGS -DREAL=double stac_oracle.c -o _build/liboracle_f64.so -lm
echo "oracle built: $(ls _build)"
