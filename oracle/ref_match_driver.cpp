// ref_match_driver.cpp — thin C entry points over the REFERENCE's own matcher code.
//
// TEST INFRASTRUCTURE ONLY.  This translation unit contains no restated algorithm: it includes
// openMVG's headers where they lie under /root/reference/src and drives the reference's public
// entry points on flat buffers:
//   * openMVG::matching_image_collection::Matcher_Regions::Match
//       (matching_image_collection/Matcher_Regions.cpp:32-107)  — the image-collection loop
//   * openMVG::matching::DistanceRatioMatch(…, BRUTE_FORCE_L2, …)
//       (matching/regions_matcher.cpp:37-52)                    — one image pair
//   * openMVG::matching::L2<uint8_t> / L2_AVX2 (matching/metric.hpp:55-93, metric_simd.hpp:34-67)
// Built by oracle/Makefile into oracle/_ref/libref_match.so (git-ignored; shipped by gpurun).
#include "openMVG/features/regions_factory.hpp"
#include "openMVG/matching/indMatch.hpp"
#include "openMVG/matching/metric.hpp"
#include "openMVG/matching/regions_matcher.hpp"
#include "openMVG/matching_image_collection/Matcher_Regions.hpp"
#include "openMVG/sfm/pipelines/sfm_regions_provider.hpp"

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

using namespace openMVG;

namespace {

// Same shape as the reference's own test double (sfm/pipelines/pipelines_test.hpp:25-69):
// a provider whose cache is filled in memory.
struct InMemory_Regions_Provider : public sfm::Regions_Provider
{
  void set(IndexT id, std::shared_ptr<features::Regions> r) { cache_[id] = std::move(r); }
  void set_type(features::Regions * r) { region_type_.reset(r); }
};

std::shared_ptr<features::SIFT_Regions> make_regions(const uint8_t * desc, uint32_t n)
{
  auto r = std::make_shared<features::SIFT_Regions>();
  r->Features().resize(n);
  r->Descriptors().resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    r->Features()[i] = features::SIOPointFeature(float(i), float(i), 1.f, 0.f);
    std::memcpy(r->Descriptors()[i].data(), desc + size_t(i) * 128, 128);
  }
  return r;
}

}  // namespace

extern "C" {

int ref_l2_u8(const uint8_t * a, const uint8_t * b, int n)
{
  return matching::L2<uint8_t>()(a, b, size_t(n));
}

// One pair through DistanceRatioMatch.  out_ij capacity: 2*n_j.  Returns #matches.
int64_t ref_match_pair(const uint8_t * desc_i, uint32_t n_i, const uint8_t * desc_j, uint32_t n_j,
                       float dist_ratio, uint32_t * out_ij)
{
  auto ri = make_regions(desc_i, n_i);
  auto rj = make_regions(desc_j, n_j);
  matching::IndMatches m;
  matching::DistanceRatioMatch(dist_ratio, matching::BRUTE_FORCE_L2, *ri, *rj, m);
  for (size_t k = 0; k < m.size(); ++k) { out_ij[2 * k] = m[k].i_; out_ij[2 * k + 1] = m[k].j_; }
  return int64_t(m.size());
}

// A collection through Matcher_Regions::Match.  desc = all images' rows back to back;
// row_start[k] = first row of image k; counts[k] = rows of image k.
// offsets[n_pairs+1] / ij[2*cap]: CSR over the pairs in the order given (pairs absent from the
// reference's output map get an empty row).  Returns total #matches, or -(needed) if cap too small.
int64_t ref_match_collection(const uint8_t * desc, const uint64_t * row_start, const uint32_t * counts,
                             uint32_t n_images, const uint32_t * pair_i, const uint32_t * pair_j,
                             uint64_t n_pairs, float dist_ratio, uint64_t * offsets, uint32_t * ij,
                             uint64_t cap)
{
  auto provider = std::make_shared<InMemory_Regions_Provider>();
  provider->set_type(new features::SIFT_Regions);
  for (uint32_t k = 0; k < n_images; ++k)
    provider->set(k, make_regions(desc + row_start[k] * 128, counts[k]));
  Pair_Set pairs;
  for (uint64_t p = 0; p < n_pairs; ++p) pairs.insert({pair_i[p], pair_j[p]});

  matching::PairWiseMatches out;
  matching_image_collection::Matcher_Regions matcher(dist_ratio, matching::BRUTE_FORCE_L2);
  matcher.Match(provider, pairs, out, nullptr);

  uint64_t total = 0;
  for (uint64_t p = 0; p < n_pairs; ++p) {
    auto it = out.find({pair_i[p], pair_j[p]});
    total += (it == out.end()) ? 0 : it->second.size();
  }
  if (total > cap) return -int64_t(total);
  uint64_t w = 0;
  for (uint64_t p = 0; p < n_pairs; ++p) {
    offsets[p] = w;
    auto it = out.find({pair_i[p], pair_j[p]});
    if (it == out.end()) continue;
    for (const auto & m : it->second) { ij[2 * w] = m.i_; ij[2 * w + 1] = m.j_; ++w; }
  }
  offsets[n_pairs] = w;
  return int64_t(total);
}

}  // extern "C"
