#!/usr/bin/env bash
# Stages the UNMODIFIED reference sources of the hot path under oracle/_ref/ so that bench.py's CPU arm
# (`--impl reference`, `cpu_baseline`, `reference_cuda`) can run the reference's own caldera() on the GPU box,
# where /root/reference does not exist.  oracle/_ref/ is git-ignored (no reference source enters this
# repository's history) but travels to the GPU box with the rest of the tree.  Run in the build container:
#
#     bash oracle/make_ref.sh            (also done by __graft_entry__.build() when /root/reference is present)
#
# Files staged (RCR = /root/reference/rank-constrained-regression-main/src):
#   RCR/__init__.py  RCR/caldera/decomposition/alg.py  RCR/caldera/utils/dataclasses.py  RCR/caldera/utils/quantization.py
# The reference is pure Python: nothing is compiled.  bench.py falls back to the numpy port in oracle/ when the
# directory is absent.
set -euo pipefail
REF="${1:-/root/reference/rank-constrained-regression-main}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$REF/src/caldera" ]; then
  echo "make_ref.sh: $REF not found (nothing staged)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/src/caldera/decomposition" "$DST/src/caldera/utils"
cp "$REF/src/__init__.py" "$DST/src/__init__.py"
cp "$REF/src/caldera/decomposition/alg.py" "$DST/src/caldera/decomposition/alg.py"
cp "$REF/src/caldera/utils/dataclasses.py" "$DST/src/caldera/utils/dataclasses.py"
cp "$REF/src/caldera/utils/quantization.py" "$DST/src/caldera/utils/quantization.py"
( cd "$REF" && sha256sum src/__init__.py src/caldera/decomposition/alg.py src/caldera/utils/dataclasses.py src/caldera/utils/quantization.py ) > "$DST/SHA256SUMS"
echo "staged $(wc -l < "$DST/SHA256SUMS") reference files under $DST"
