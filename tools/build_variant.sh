#!/bin/bash
# Build an experimental variant of libgsb_b200.so with extra -D flags, without touching the shipped library:
#   tools/build_variant.sh NAME "-DGSB_FAST_UNROLL=8 ..."   ->   build/variants/libgsb_NAME.so
# Run it with GSB_LIB_PATH=build/variants/libgsb_NAME.so (read by intro_to_gaussian_splatting_b200/_lib.py).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
TMP=$(mktemp -d /tmp/gsb_variant_XXXX)
mkdir -p "$TMP/pkg/csrc" "$TMP/include" "$ROOT/build/variants"
cp "$ROOT"/intro_to_gaussian_splatting_b200/csrc/*.cu "$ROOT"/intro_to_gaussian_splatting_b200/csrc/*.cuh \
   "$ROOT"/intro_to_gaussian_splatting_b200/csrc/Makefile "$TMP/pkg/csrc/"
cp "$ROOT"/include/gsb.h "$TMP/include/"
make -C "$TMP/pkg/csrc" -j8 EXTRA="$*" > "$TMP/build.log" 2>&1 || { tail -30 "$TMP/build.log"; exit 1; }
cp "$TMP/pkg/libgsb_b200.so" "$ROOT/build/variants/libgsb_$NAME.so"
grep -h "composite_fast\|Used" "$TMP/pkg/csrc/composite.ptxas.log" | grep -A1 "composite_fast" | grep Used | head -2
rm -rf "$TMP"
echo "built build/variants/libgsb_$NAME.so"
