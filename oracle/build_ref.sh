#!/usr/bin/env bash
# Builds oracle/_ref/ from the parts of the real reference that compile stand-alone (only where
# /root/reference exists; the outputs travel to the GPU box, the sources never leave that tree).
#   libref_init.so : ransac_lib::LocallyOptimizedMSAC (lib/RansacLib, header-only) driving the
#                    product's init estimators — see oracle/ref/ref_init.cc
set -e
here="$(cd "$(dirname "$0")" && pwd)"
ref="${PPSFM_REFERENCE:-/root/reference}"
[ -d "$ref/lib/RansacLib/RansacLib" ] || { echo "reference tree not found: $ref"; exit 0; }
mkdir -p "$here/_ref"
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -shared \
    -I"$ref/lib/RansacLib" "$here/ref/ref_init.cc" -o "$here/_ref/libref_init.so"
echo "built $here/_ref/libref_init.so"
