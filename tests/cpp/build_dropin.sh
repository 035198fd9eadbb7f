#!/usr/bin/env bash
# Builds tests/cpp/dropin_pose_test and tests/cpp/dropin_init_test (a caller of the reference's
# init::initialize_reconstruction linked with cpp/dropin/init_initializer.cc, host build) and
# compile-checks cpp/dropin/estimators_triangulation.cc against the reference's declaration.
# dropin_pose_test: a caller written against the reference's own headers, the
# drop-in definitions (privacy_preserving_sfm_b200/cpp/dropin/estimators_pose_lines.cc) and the
# reference's Camera class, linked with libppsfm_b200.so.  Needs /root/reference (headers and
# base/camera.cc, base/camera_models.cc compiled from where they lie; Eigen / Ceres / glog / Boost
# are the stand-in headers of oracle/ref/shim).  Unused reference functions whose dependencies
# are not built (util/misc.cc, util/string.cc) are dropped by --gc-sections.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
ref="${PPSFM_REFERENCE:-/root/reference}"
[ -d "$ref/src/estimators" ] || { echo "reference tree not found: $ref"; exit 0; }
lib="$root/privacy_preserving_sfm_b200"
g++ -O1 -std=c++17 -w -ffunction-sections -fdata-sections \
    -I"$root/oracle/ref/shim" -I"$ref/src" -I"$root/include" -I"$lib/cpp" \
    "$here/dropin_pose_test.cc" "$lib/cpp/dropin/estimators_pose_lines.cc" \
    "$ref/src/base/camera.cc" "$ref/src/base/camera_models.cc" \
    -Wl,--gc-sections -L"$lib" -lppsfm_b200 -Wl,-rpath,"$lib" -o "$here/dropin_pose_test"
echo "built $here/dropin_pose_test"

g++ -O2 -std=c++17 -w -ffp-contract=off -I"$root/oracle/ref/shim" -I"$ref/src" -I"$lib/cpp" \
    "$here/dropin_init_test.cc" "$lib/cpp/dropin/init_initializer.cc" -o "$here/dropin_init_test"
echo "built $here/dropin_init_test"
g++ -O1 -std=c++17 -w -c -I"$root/oracle/ref/shim" -I"$ref/src" -I"$root/include" -I"$lib/cpp" \
    "$lib/cpp/dropin/estimators_triangulation.cc" -o "$here/dropin_triangulation.o"
g++ -O1 -std=c++17 -w -c -DPPSFM_INIT_ON_GPU -I"$root/oracle/ref/shim" -I"$ref/src" -I"$root/include" \
    -I"$lib/cpp" "$lib/cpp/dropin/init_initializer.cc" -o "$here/dropin_init_gpu.o"
echo "compiled $here/dropin_triangulation.o $here/dropin_init_gpu.o"
