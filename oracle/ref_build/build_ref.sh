#!/usr/bin/env bash
# Builds oracle/_ref/ from the reference sources WHERE THEY LIE (read-only /root/reference).
# Outputs only into oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
# Test infrastructure only.
#
#  libxref_qd.so        Propagator::discreteProcessNoiseCov alone (propagator.cpp:207-840), no stand-in headers at all
#  libxref.so           the reference's filter back end (src/x/ekf, src/x/vio, src/x/vision/{feature,track,
#                       tiled_image,triangulation}.cpp, unmodified) against the stand-in headers of shim/ (Eigen,
#                       OpenCV, Boost, NLopt are not installed here) + xref_harness.cpp; strict IEEE arithmetic (-O2)
#  libxref_multi.so     the same with -DMULTI_UAV (+ ci.cpp, simple_state.cpp, multi_slam_update.cpp)
#  libxref_tm.so        the reference's own TrackManager (src/x/vio/track_manager.cpp + src/x/vision/{camera,feature,track,
#                       tiled_image}.cpp, unmodified; the front-end stubs of shim_stubs/ are NOT on its include path) +
#                       xref_tm_harness.cpp: pins oracle/track_manager.py and the product's xb_tm_* 
#  libxref_release.so   single-agent flavour with the reference's Release flags (CMakeLists.txt:185,194): the binary
#                       bench.py times as the CPU reference
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${XREF_ROOT:-/root/reference}"
OUT="$HERE/../_ref"
mkdir -p "$OUT"
# std::sort helper of oracle/track_manager.py (no reference source involved)
g++ -O2 -fPIC -shared -o "$OUT/libxsort.so" "$HERE/xsort.cpp"
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

TUS="ekf/state ekf/propagator ekf/state_buffer ekf/updater ekf/ekf vision/feature vision/track vision/tiled_image
     vision/triangulation vio/state_manager vio/msckf_update vio/slam_update vio/msckf_slam_update vio/vio_updater
     vio/range_update vio/solar_update"
TUS_MULTI="ekf/ci ekf/simple_state vio/multi_slam_update"
RELEASE="-O3 -funsafe-loop-optimizations -fsee -funroll-loops -fno-math-errno -funsafe-math-optimizations -ffinite-math-only -fno-signed-zeros -DNDEBUG"
COMMON="-std=c++17 -fPIC -w -DMULTI_THREAD -DEIGEN_MATRIXBASE_PLUGIN=<x/common/eigen_matrix_base_plugin.h> -I$HERE/shim_stubs -I$HERE/shim -I$REF/include"

build_flavour() {  # name, extra flags, extra TUs
  local name="$1" flags="$2" extra="$3" dir="$OUT/obj_$1"
  mkdir -p "$dir"
  local objs="" pids=""
  for tu in $TUS $extra; do
    local o="$dir/$(echo "$tu" | tr / _).o"
    objs="$objs $o"
    if [ ! -f "$o" ] || [ "$REF/src/x/$tu.cpp" -nt "$o" ] || [ "$HERE/shim/Eigen/Core" -nt "$o" ] \
       || [ "$HERE/shim/opencv2/xref_cv.hpp" -nt "$o" ] || [ "$HERE/build_ref.sh" -nt "$o" ] \
       || [ "$HERE/shim_stubs/x/vio/track_manager.h" -nt "$o" ] || [ "$HERE/shim_stubs/x/vision/tracker.h" -nt "$o" ]; then
      g++ $COMMON $flags -c "$REF/src/x/$tu.cpp" -o "$o" &
      pids="$pids $!"
    fi
  done
  local h="$dir/xref_harness.o"
  g++ $COMMON $flags -c "$HERE/xref_harness.cpp" -o "$h" &
  pids="$pids $!"
  for p in $pids; do wait "$p"; done
  g++ -shared -Wl,-Bsymbolic -o "$OUT/$name.so" $objs "$h" -ldl -lpthread
  echo "build_ref: built $OUT/$name.so from $REF/src/x/{ekf,vio,vision} (flags: $flags)"
}
build_flavour libxref "-O2" ""
build_flavour libxref_multi "-O2 -DMULTI_UAV" "$TUS_MULTI"
build_flavour libxref_release "$RELEASE" ""

# ---- the reference's TrackManager (real header, no front-end stubs)
TMDIR="$OUT/obj_tm"
mkdir -p "$TMDIR"
TMOBJS=""
for tu in vio/track_manager vision/camera vision/feature vision/track vision/tiled_image; do
  o="$TMDIR/$(echo "$tu" | tr / _).o"
  TMOBJS="$TMOBJS $o"
  g++ -std=c++17 -fPIC -w -O2 "-DEIGEN_MATRIXBASE_PLUGIN=<x/common/eigen_matrix_base_plugin.h>" -I"$HERE/shim" -I"$REF/include" -c "$REF/src/x/$tu.cpp" -o "$o" &
done
g++ -std=c++17 -fPIC -w -O2 "-DEIGEN_MATRIXBASE_PLUGIN=<x/common/eigen_matrix_base_plugin.h>" -I"$HERE/shim" -I"$REF/include" -c "$HERE/xref_tm_harness.cpp" -o "$TMDIR/harness.o" &
wait
g++ -shared -Wl,-Bsymbolic -o "$OUT/libxref_tm.so" $TMOBJS "$TMDIR/harness.o"
echo "build_ref: built $OUT/libxref_tm.so from $REF/src/x/vio/track_manager.cpp"
