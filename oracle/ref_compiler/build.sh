#!/bin/bash
# Builds the reference's compiler half (IR + CKKS passes + reference executor)
# from the sources where they lie under /root/reference, against three stub
# headers, into oracle/_ref/eva_ref_compile.
# Fixture generator only: SEAL / protobuf parts of the reference cannot be built here.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_ref"
REF=/root/reference
[ -d "$REF" ] || { echo "no $REF: using prebuilt oracle/_ref"; exit 0; }
mkdir -p "$OUT"
if [ "$OUT/eva_ref_compile" -nt "$HERE/driver.cpp" ]; then exit 0; fi
g++ -std=c++17 -O1 -w -include map -include cassert -include cmath -include cstdint -include stdexcept \
    -include algorithm -include unordered_set -I"$HERE/stubs" -I"$REF" "$HERE/driver.cpp" \
    "$REF"/eva/ir/term.cpp "$REF"/eva/ir/program.cpp "$REF"/eva/ir/attribute_list.cpp "$REF"/eva/ir/attributes.cpp \
    "$REF"/eva/ckks/ckks_config.cpp "$REF"/eva/util/logging.cpp "$REF"/eva/common/reference_executor.cpp \
    -o "$OUT/eva_ref_compile"
echo "built $OUT/eva_ref_compile"
