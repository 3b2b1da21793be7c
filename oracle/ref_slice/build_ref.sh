#!/bin/sh
# Builds oracle/_ref/libref_slice.so: the reference's OWN ellipsoid / edge code (src/core/Ellipsoid.cpp, BasicEllipsoidEdges.cpp,
# Plane.cpp, src/Polygon/Polygon.cpp, src/utils/matrix_utils.cpp) + the vendored g2o slice they need, compiled from the sources where
# they lie under $REF (default /root/reference) with plain g++ -- not the reference's CMake build.  The only external requirement is
# the Eigen3 headers (EIGEN3_INCLUDE_DIR, /usr/include/eigen3, /usr/local/include/eigen3); OpenCV is replaced by a 30-line Point stub.
# In the round-2 image Eigen is absent: the script then reports "unbuildable" and exits 0, tests/test_ref_slice.py skips.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${REF:-/root/reference}
OUT="$HERE/../_ref"
EIG=${EIGEN3_INCLUDE_DIR:-}
for d in /usr/include/eigen3 /usr/local/include/eigen3 /opt/eigen3/include/eigen3; do
  [ -z "$EIG" ] && [ -f "$d/Eigen/Core" ] && EIG=$d
done
if [ ! -d "$REF" ]; then echo "ref_slice: $REF not present (GPU box): nothing to build"; exit 0; fi
if [ -z "$EIG" ]; then echo "ref_slice: unbuildable here -- Eigen3 headers not found (set EIGEN3_INCLUDE_DIR)"; exit 0; fi
mkdir -p "$OUT"
G2O="$REF/Thirdparty/g2o/g2o"
SRCS="$REF/src/core/Ellipsoid.cpp $REF/src/core/BasicEllipsoidEdges.cpp $REF/src/core/Plane.cpp $REF/src/Polygon/Polygon.cpp $REF/src/utils/matrix_utils.cpp"
SRCS="$SRCS $(ls $G2O/core/*.cpp $G2O/stuff/*.cpp $G2O/types/*.cpp)"
g++ -O2 -std=c++14 -fPIC -shared -w -I"$HERE/stub" -I"$REF" -I"$REF/Thirdparty/g2o" -I"$EIG" $SRCS "$HERE/ref_shim.cpp" -o "$OUT/libref_slice.so"
echo "ref_slice: built $OUT/libref_slice.so"
