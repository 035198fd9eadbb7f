// ref_corr_graph.cc — TEST INFRASTRUCTURE.  The REFERENCE's own CorrespondenceGraph (SURVEY §8 row
// f3: "minimal in-memory Reconstruction + CorrespondenceGraph"), compiled from where it lies under
// /root/reference:
//   src/base/correspondence_graph.{h,cc}   AddImage / AddCorrespondences (duplicate and
//                                          out-of-range matches dropped, counts corrected) /
//                                          Finalize / FindCorrespondences /
//                                          FindTransitiveCorrespondences /
//                                          FindCorrespondencesBetweenImages / IsTwoViewObservation
// behind C entry points, for privacy_preserving_sfm_b200/correspondence_graph.py
// (tests/test_ref_correspondence_graph.py).  Linked into oracle/_ref/libref_filter.so
// (oracle/build_ref.sh), which defines Database::kMaxNumImages (ref_filter.cc).
#include <algorithm>
#include <cstdint>
#include <vector>

#include "base/correspondence_graph.h"

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void* ref_cg_new() { return new colmap::CorrespondenceGraph; }
REF_API void ref_cg_free(void* h) { delete static_cast<colmap::CorrespondenceGraph*>(h); }

REF_API void ref_cg_add_image(void* h, uint32_t image_id, uint64_t num_lines) {
  static_cast<colmap::CorrespondenceGraph*>(h)->AddImage(image_id, num_lines);
}

REF_API void ref_cg_add_correspondences(void* h, uint32_t image_id1, uint32_t image_id2,
                                        const uint32_t* matches, uint64_t num_matches) {
  colmap::FeatureMatches m;
  m.reserve(num_matches);
  for (uint64_t k = 0; k < num_matches; ++k) m.emplace_back(matches[2 * k], matches[2 * k + 1]);
  static_cast<colmap::CorrespondenceGraph*>(h)->AddCorrespondences(image_id1, image_id2, m);
}

REF_API void ref_cg_finalize(void* h) { static_cast<colmap::CorrespondenceGraph*>(h)->Finalize(); }

// counts: NumImages, NumImagePairs
REF_API void ref_cg_sizes(void* h, uint64_t* out) {
  const auto* g = static_cast<colmap::CorrespondenceGraph*>(h);
  out[0] = g->NumImages();
  out[1] = g->NumImagePairs();
}

// per image: ExistsImage, NumObservationsForImage, NumCorrespondencesForImage (0, 0 if absent)
REF_API void ref_cg_image(void* h, uint32_t image_id, uint64_t* out) {
  const auto* g = static_cast<colmap::CorrespondenceGraph*>(h);
  out[0] = g->ExistsImage(image_id) ? 1 : 0;
  out[1] = out[0] ? g->NumObservationsForImage(image_id) : 0;
  out[2] = out[0] ? g->NumCorrespondencesForImage(image_id) : 0;
}

REF_API uint64_t ref_cg_num_between(void* h, uint32_t image_id1, uint32_t image_id2) {
  return static_cast<colmap::CorrespondenceGraph*>(h)->NumCorrespondencesBetweenImages(image_id1,
                                                                                        image_id2);
}

// all pairs, sorted by pair id; returns the number of pairs (writes at most cap)
REF_API uint64_t ref_cg_pairs(void* h, uint64_t* pair_id, uint32_t* num, uint64_t cap) {
  const auto m = static_cast<colmap::CorrespondenceGraph*>(h)->NumCorrespondencesBetweenImages();
  std::vector<std::pair<uint64_t, uint32_t>> v(m.begin(), m.end());
  std::sort(v.begin(), v.end());
  for (uint64_t k = 0; k < v.size() && k < cap; ++k) {
    pair_id[k] = v[k].first;
    num[k] = v[k].second;
  }
  return v.size();
}

// FindTransitiveCorrespondences (transitivity 1 = FindCorrespondences), in the reference's order
REF_API uint64_t ref_cg_find(void* h, uint32_t image_id, uint32_t line_idx, uint64_t transitivity,
                             uint32_t* out, uint64_t cap) {
  const auto v = static_cast<colmap::CorrespondenceGraph*>(h)->FindTransitiveCorrespondences(
      image_id, line_idx, transitivity);
  for (uint64_t k = 0; k < v.size() && k < cap; ++k) {
    out[2 * k] = v[k].image_id;
    out[2 * k + 1] = v[k].line_idx;
  }
  return v.size();
}

REF_API uint64_t ref_cg_between(void* h, uint32_t image_id1, uint32_t image_id2, uint32_t* out,
                                uint64_t cap) {
  const auto v = static_cast<colmap::CorrespondenceGraph*>(h)->FindCorrespondencesBetweenImages(
      image_id1, image_id2);
  for (uint64_t k = 0; k < v.size() && k < cap; ++k) {
    out[2 * k] = v[k].line_idx1;
    out[2 * k + 1] = v[k].line_idx2;
  }
  return v.size();
}

// HasCorrespondences, IsTwoViewObservation
REF_API void ref_cg_line(void* h, uint32_t image_id, uint32_t line_idx, uint64_t* out) {
  const auto* g = static_cast<colmap::CorrespondenceGraph*>(h);
  out[0] = g->HasCorrespondences(image_id, line_idx) ? 1 : 0;
  out[1] = g->IsTwoViewObservation(image_id, line_idx) ? 1 : 0;
}
