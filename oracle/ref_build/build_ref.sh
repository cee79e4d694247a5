#!/usr/bin/env bash
# Builds oracle/_ref/ from the reference sources WHERE THEY LIE (read-only /root/reference).
# Outputs only into oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
# Test infrastructure only.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${XREF_ROOT:-/root/reference}"
OUT="$HERE/../_ref"
mkdir -p "$OUT"
if [ ! -d "$REF/src/x/ekf" ]; then
  echo "build_ref: $REF not present; keeping prebuilt oracle/_ref" >&2
  exit 0
fi
SRC="$REF/src/x/ekf/propagator.cpp"
# Locate Propagator::discreteProcessNoiseCov (signature line .. closing brace at column 0).
START=$(grep -n '^CoreCovMatrix Propagator::discreteProcessNoiseCov' "$SRC" | cut -d: -f1)
END=$(awk -v s="$START" 'NR>s && /^}/ {print NR; exit}' "$SRC")
sed -n "${START},${END}p" "$SRC" > "$OUT/qd_body.inc"
g++ -O2 -fPIC -shared -I"$OUT" -o "$OUT/libxref_qd.so" "$HERE/qd_shim.cpp"
echo "build_ref: built $OUT/libxref_qd.so from $SRC:$START-$END"
