#!/bin/bash
# Development helper: build a variant of the library with extra nvcc defines for ge2e_tc.cu only.
#   scripts/build_variant.sh <output.so> [-DNAME=VALUE ...]
# The other objects are reused from the last regular build (python -m speaker_embedding_ge2e_loss_b200.build).
set -e
cd "$(dirname "$0")/../speaker_embedding_ge2e_loss_b200/csrc"
out="$1"; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I ../../include -I . "$@" \
     -c ge2e_tc.cu -o /tmp/ge2e_tc_variant.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$out" ge2e_api.o ge2e_simt.o /tmp/ge2e_tc_variant.o ge2e_tail.o
echo "built $out"
