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
//   * openMVG::matching_image_collection::Cascade_Hashing_Matcher_Regions::Match
//       (matching_image_collection/Cascade_Hashing_Matcher_Regions.cpp:38-226) and
//     openMVG::matching::CascadeHasher (matching/cascade_hasher.hpp)  — the cascade-hashing variant
// Built by oracle/Makefile into oracle/_ref/libref_match.so (git-ignored; shipped by gpurun).
#include "openMVG/features/regions_factory.hpp"
#include "openMVG/matching/indMatch.hpp"
#include "openMVG/matching/indMatch_utils.hpp"
#include "openMVG/features/descriptor.hpp"
#include "openMVG/matching/metric.hpp"
#include "openMVG/matching/regions_matcher.hpp"
#include "openMVG/matching_image_collection/Matcher_Regions.hpp"
#include "openMVG/matching_image_collection/Cascade_Hashing_Matcher_Regions.hpp"
#include "openMVG/matching/cascade_hasher.hpp"
#include "openMVG/sfm/pipelines/sfm_regions_provider.hpp"

#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
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

// ---- cascade hashing
// The projections CascadeHasher::Init draws (cascade_hasher.hpp:142-162): the members are private, so the draw
// is repeated here with the same generator, distribution, seed and order — no openMVG code, only <random>.
void ref_cascade_projections(float * primary /*[128][128]*/, float * secondary /*[6][10][128]*/)
{
  std::mt19937 gen(std::mt19937::default_seed);
  std::normal_distribution<> d(0, 1);
  for (int i = 0; i < 128; ++i) for (int j = 0; j < 128; ++j) primary[i * 128 + j] = float(d(gen));
  for (int g = 0; g < 6; ++g) for (int j = 0; j < 10; ++j) for (int k = 0; k < 128; ++k) secondary[(g * 10 + j) * 128 + k] = float(d(gen));
}

// CascadeHasher::CreateHashedDescriptions on one image with a given zero-mean vector:
// codes[n][4] (bit j -> word j>>5, bit j&31), bids[n][6]
void ref_cascade_hash(const uint8_t * desc, uint32_t n, const float * zero_mean, uint32_t * codes, uint16_t * bids)
{
  using BaseMat = Eigen::Matrix<uint8_t, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>;
  matching::CascadeHasher hasher; hasher.Init(128);
  Eigen::Map<const BaseMat> mat(desc, n, 128);
  const Eigen::VectorXf zm = Eigen::Map<const Eigen::VectorXf>(zero_mean, 128);
  const matching::HashedDescriptions h = hasher.CreateHashedDescriptions(mat, zm);
  for (uint32_t r = 0; r < n; ++r) {
    for (int w = 0; w < 4; ++w) codes[4 * size_t(r) + w] = 0;
    for (int j = 0; j < 128; ++j) if (h.hashed_desc[r].hash_code[j]) codes[4 * size_t(r) + (j >> 5)] |= 1u << (j & 31);
    for (int g = 0; g < 6; ++g) bids[6 * size_t(r) + g] = h.hashed_desc[r].bucket_ids[g];
  }
}

// The collection's zero-mean descriptor exactly as Cascade_Hashing_Matcher_Regions.cpp:78-105 builds it
// (used images = the ones the pair list names).
void ref_cascade_zero_mean(const uint8_t * desc, const uint64_t * row_start, const uint32_t * counts, const uint32_t * used, uint32_t n_used, float * zm)
{
  using BaseMat = Eigen::Matrix<uint8_t, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>;
  Eigen::MatrixXf m(n_used, 128); m.fill(0.0f);
  for (uint32_t i = 0; i < n_used; ++i) {
    const uint32_t k = used[i];
    if (counts[k] > 0) { Eigen::Map<const BaseMat> mat(desc + row_start[k] * 128, counts[k], 128); m.row(i) = matching::CascadeHasher::GetZeroMeanDescriptor(mat); }
  }
  const Eigen::VectorXf z = matching::CascadeHasher::GetZeroMeanDescriptor(m);
  for (int k = 0; k < 128; ++k) zm[k] = z(k);
}

// A collection through Cascade_Hashing_Matcher_Regions::Match; same CSR convention as ref_match_collection.
int64_t ref_cascade_collection(const uint8_t * desc, const uint64_t * row_start, const uint32_t * counts,
                               uint32_t n_images, const uint32_t * pair_i, const uint32_t * pair_j,
                               uint64_t n_pairs, float dist_ratio, uint64_t * offsets, uint32_t * ij, uint64_t cap)
{
  auto provider = std::make_shared<InMemory_Regions_Provider>();
  provider->set_type(new features::SIFT_Regions);
  for (uint32_t k = 0; k < n_images; ++k)
    provider->set(k, make_regions(desc + row_start[k] * 128, counts[k]));
  Pair_Set pairs;
  for (uint64_t p = 0; p < n_pairs; ++p) pairs.insert({pair_i[p], pair_j[p]});
  matching::PairWiseMatches out;
  matching_image_collection::Cascade_Hashing_Matcher_Regions matcher(dist_ratio);
  matcher.Match(provider, pairs, out, nullptr);
  uint64_t total = 0;
  for (uint64_t p = 0; p < n_pairs; ++p) { auto it = out.find({pair_i[p], pair_j[p]}); total += (it == out.end()) ? 0 : it->second.size(); }
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

// ---- file formats: the reference's own writers / readers
int ref_save_descs(const char * path, const uint8_t * desc, uint32_t n)
{
  std::vector<features::Descriptor<unsigned char, 128>, Eigen::aligned_allocator<features::Descriptor<unsigned char, 128>>> v(n);
  for (uint32_t i = 0; i < n; ++i) std::memcpy(v[i].data(), desc + size_t(i) * 128, 128);
  return features::saveDescsToBinFile(path, v) ? 0 : 1;                // features/descriptor.hpp:206-228
}

int ref_save_matches(const char * path, uint64_t n_pairs, const uint32_t * pI, const uint32_t * pJ, const uint64_t * offsets, const uint32_t * ij)
{
  matching::PairWiseMatches m;
  for (uint64_t p = 0; p < n_pairs; ++p) {
    if (offsets[p + 1] == offsets[p]) continue;
    matching::IndMatches v;
    for (uint64_t k = offsets[p]; k < offsets[p + 1]; ++k) v.emplace_back(ij[2 * k], ij[2 * k + 1]);
    m[{pI[p], pJ[p]}] = std::move(v);
  }
  return matching::Save(m, path) ? 0 : 1;                              // matching/indMatch_utils.cpp:80-131
}

// matching::Load (indMatch_utils.cpp:31-78); returns the number of pairs or -1.
int64_t ref_load_matches(const char * path, uint64_t cap_pairs, uint32_t * pI, uint32_t * pJ, uint64_t * offsets, uint64_t cap_m, uint32_t * ij)
{
  matching::PairWiseMatches m;
  if (!matching::Load(m, path)) return -1;
  uint64_t p = 0, w = 0;
  for (const auto & kv : m) {
    if (p >= cap_pairs || w + kv.second.size() > cap_m) return -2;
    pI[p] = kv.first.first; pJ[p] = kv.first.second; offsets[p] = w;
    for (const auto & x : kv.second) { ij[2 * w] = x.i_; ij[2 * w + 1] = x.j_; ++w; }
    ++p;
  }
  offsets[p] = w;
  return int64_t(p);
}

}  // extern "C"
