#!/usr/bin/env bash
# Builds oracle/_ref/ from the parts of the real reference that compile stand-alone (only where
# /root/reference exists; the outputs travel to the GPU box, the sources never leave that tree).
#   libref_init.so : ransac_lib::LocallyOptimizedMSAC (lib/RansacLib, header-only) driving the
#                    product's init estimators — see oracle/ref/ref_init.cc
#   libref_p6l.so  : the reference's OWN P6L / re3q3 / line residual / RANSAC loop / sampler /
#                    support sources and the EstimateAbsolutePoseFromLines wrapper
#                    (src/estimators/{absolute_pose,pose}.cc, lib/re3q3/re3q3/re3q3.h,
#                    src/estimators/utils.cc, src/optim/{ransac.h,random_sampler.cc,
#                    support_measurement.cc}, src/util/random.cc) behind C entry points
#                    (oracle/ref/ref_p6l.cc), compiled against the Eigen / glog stand-ins of
#                    oracle/ref/shim/ (both libraries are absent in this image)
#   libref_tri.so  : the reference's OWN robust line triangulation (src/estimators/
#                    triangulation.cc, src/base/{triangulation,projection,camera}.cc,
#                    src/optim/{loransac.h,combination_sampler.cc}, src/util/math.cc) behind a C
#                    entry point (oracle/ref/ref_triangulation.cc); only the n x 4 JacobiSVD is
#                    the stand-in's (eigen_restated::NullVectorNx4)
#   libref_ba_setup.so : the reference's OWN bundle-adjustment assembly (src/optim/
#                    bundle_adjustment.cc + the classes it reads) against a RECORDING ceres::Problem,
#                    and the product's adaptor on the same colmap::Reconstruction
#                    (oracle/ref/ref_ba_setup.cc); links libppsfm_b200.so for the adaptor's option
#                    defaults (host-only entry points)
#   libref_filter.so : the reference's OWN Reconstruction::FilterPoints3D /
#                    FilterObservationsWithNegativeDepth (src/base/reconstruction.cc and the classes
#                    it uses) on a reconstruction built through its own Add* members, its
#                    ReadText / WriteText (text model format) (oracle/ref/ref_filter.cc), and
#                    the reference's CorrespondenceGraph (src/base/correspondence_graph.cc,
#                    oracle/ref/ref_corr_graph.cc)
#   libref_cost.so : the reference's OWN line cost functors (src/base/cost_functions.h) and
#                    camera models (src/base/camera_models.{h,cc}) behind C entry points
#                    (oracle/ref/ref_cost.cc), against the Ceres / Eigen / glog / Boost stand-ins
set -e
here="$(cd "$(dirname "$0")" && pwd)"
ref="${PPSFM_REFERENCE:-/root/reference}"
[ -d "$ref/lib/RansacLib/RansacLib" ] || { echo "reference tree not found: $ref"; exit 0; }
mkdir -p "$here/_ref"
pids=""   # the libraries are independent: build them side by side
(
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -shared \
    -I"$ref/lib/RansacLib" "$here/ref/ref_init.cc" -o "$here/_ref/libref_init.so"
echo "built $here/_ref/libref_init.so"
) &
pids="$pids $!"

(
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -shared -w \
    -fvisibility=hidden -ffunction-sections -fdata-sections -Wl,--gc-sections \
    -I"$here/ref/shim" -I"$ref/src" -I"$ref/lib" -I"$ref/lib/re3q3" \
    "$here/ref/ref_p6l.cc" "$ref/src/estimators/absolute_pose.cc" "$ref/src/estimators/utils.cc" \
    "$ref/src/estimators/pose.cc" "$ref/src/optim/random_sampler.cc" \
    "$ref/src/optim/support_measurement.cc" "$ref/src/util/random.cc" \
    "$ref/src/base/camera.cc" "$ref/src/base/camera_models.cc" "$ref/src/base/pose.cc" \
    "$ref/src/optim/bundle_adjustment.cc" \
    "$ref/src/util/misc.cc" "$ref/src/util/string.cc" "$ref/src/util/threading.cc" \
    "$ref/src/util/timer.cc" "$ref/src/util/logging.cc" \
    -o "$here/_ref/libref_p6l.so"
echo "built $here/_ref/libref_p6l.so"
) &
pids="$pids $!"

(
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -shared -w \
    -I"$here/ref/shim" -I"$ref/src" \
    "$here/ref/ref_cost.cc" "$ref/src/base/camera_models.cc" -o "$here/_ref/libref_cost.so"
echo "built $here/_ref/libref_cost.so"
) &
pids="$pids $!"

(
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -shared -w \
    -fvisibility=hidden -ffunction-sections -fdata-sections -Wl,--gc-sections \
    -I"$here/ref/shim" -I"$ref/src" -I"$ref/lib" \
    "$here/ref/ref_triangulation.cc" "$ref/src/estimators/triangulation.cc" \
    "$ref/src/base/triangulation.cc" "$ref/src/base/projection.cc" "$ref/src/base/camera.cc" \
    "$ref/src/base/camera_models.cc" "$ref/src/optim/combination_sampler.cc" \
    "$ref/src/optim/support_measurement.cc" "$ref/src/util/math.cc" \
    -o "$here/_ref/libref_tri.so"
echo "built $here/_ref/libref_tri.so"
) &
pids="$pids $!"

lib="$here/../privacy_preserving_sfm_b200"
if [ -f "$lib/libppsfm_b200.so" ]; then
(
g++ -O1 -std=c++17 -fPIC -w -shared -fvisibility=hidden -ffunction-sections -fdata-sections \
    -I"$here/ref/shim" -I"$ref/src" -I"$ref/lib" -I"$here/../include" -I"$lib/cpp" \
    "$here/ref/ref_ba_setup.cc" "$ref/src/optim/bundle_adjustment.cc" \
    "$ref/src/base/reconstruction.cc" "$ref/src/base/image.cc" \
    "$ref/src/base/point3d.cc" "$ref/src/base/track.cc" "$ref/src/base/camera.cc" \
    "$ref/src/base/camera_models.cc" "$ref/src/base/pose.cc" "$ref/src/base/projection.cc" \
    "$ref/src/base/triangulation.cc" "$ref/src/util/math.cc" \
    "$ref/src/util/string.cc" "$ref/src/util/misc.cc" \
    "$ref/src/util/threading.cc" "$ref/src/util/timer.cc" "$ref/src/util/logging.cc" \
    -Wl,--gc-sections -L"$lib" -lppsfm_b200 -Wl,-rpath,"$lib" -o "$here/_ref/libref_ba_setup.so"
echo "built $here/_ref/libref_ba_setup.so"
) &
pids="$pids $!"
fi

(
g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -shared -w \
    -fvisibility=hidden -ffunction-sections -fdata-sections -Wl,--gc-sections \
    -I"$here/ref/shim" -I"$ref/src" -I"$ref/lib" \
    "$here/ref/ref_filter.cc" "$here/ref/ref_corr_graph.cc" \
    "$ref/src/base/reconstruction.cc" "$ref/src/base/correspondence_graph.cc" "$ref/src/base/image.cc" \
    "$ref/src/base/point3d.cc" "$ref/src/base/track.cc" "$ref/src/base/camera.cc" \
    "$ref/src/base/camera_models.cc" "$ref/src/base/pose.cc" "$ref/src/base/projection.cc" \
    "$ref/src/base/triangulation.cc" "$ref/src/util/math.cc" "$ref/src/util/misc.cc" \
    "$ref/src/util/string.cc" "$ref/src/util/logging.cc" \
    -o "$here/_ref/libref_filter.so"
echo "built $here/_ref/libref_filter.so"
) &
pids="$pids $!"
rc=0
for p in $pids; do wait "$p" || rc=1; done
exit $rc
