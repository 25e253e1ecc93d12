#!/bin/sh
# Stage the reference's own tests and example for a GPU run: the GPU box has no /root/reference, and the
# reference's sources never enter this repository's history (_reftests/ is git-ignored; it travels with the
# gpurun snapshot only).  tests/test_reference_suite.py picks the files up from there.  Remove with --clean.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${EVA_REFERENCE_DIR:-/root/reference}"
rm -rf "$ROOT/_reftests"
[ "$1" = "--clean" ] && exit 0
mkdir -p "$ROOT/_reftests/tests" "$ROOT/_reftests/examples"
cp "$REF"/tests/*.py "$ROOT/_reftests/tests/"
cp "$REF"/examples/image_processing.py "$REF"/examples/serialization.py "$REF"/examples/baboon.png "$ROOT/_reftests/examples/"
echo "staged $(ls "$ROOT/_reftests/tests" | wc -l) test files and the examples under _reftests/"
