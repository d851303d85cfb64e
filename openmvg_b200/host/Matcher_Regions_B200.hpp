// Matcher_Regions_B200 — drop-in replacement for
//   openMVG::matching_image_collection::Matcher_Regions(distRatio, BRUTE_FORCE_L2)
// (reference: src/openMVG/matching_image_collection/Matcher_Regions.{hpp,cpp}) implementing the
// abstract openMVG::matching_image_collection::Matcher (Matcher.hpp:34-48) on top of the C ABI of
// libomvg_b200.so.  Same constructor argument, same Match() signature, same observable behaviour:
//   * pairs whose regions are empty, or whose Type_id() differ, are skipped   (Matcher_Regions.cpp:65-69,85-90)
//   * database = regions of I, queries = regions of J                         (:73,93)
//   * only non-empty results are inserted under key {I,J}                      (:99-102)
//   * ++progress once per pair, including skipped ones; hasBeenCanceled() honoured (:59,67,88,104)
// Only 128-byte unsigned-char scalar regions (SIFT) go to the GPU; anything else returns without
// matches after logging — callers that need other region types keep using Matcher_Regions.
//
// Header-only; compile inside an openMVG build (needs openMVG headers) and link libomvg_b200.so.
#ifndef OPENMVG_B200_MATCHER_REGIONS_B200_HPP
#define OPENMVG_B200_MATCHER_REGIONS_B200_HPP

#include "openMVG/features/regions.hpp"
#include "openMVG/matching/indMatch.hpp"
#include "openMVG/matching_image_collection/Matcher.hpp"
#include "openMVG/sfm/pipelines/sfm_regions_provider.hpp"
#include "openMVG/system/logger.hpp"
#include "openMVG/system/progressinterface.hpp"

#include "omvg_b200.h"

#include <map>
#include <memory>
#include <set>
#include <vector>

namespace openMVG {
namespace matching_image_collection {

class Matcher_Regions_B200 : public Matcher
{
  public:
  explicit Matcher_Regions_B200(float dist_ratio, int device = 0)
    : Matcher(), f_dist_ratio_(dist_ratio), device_(device) {}

  void Match(
    const std::shared_ptr<sfm::Regions_Provider> & regions_provider,
    const Pair_Set & pairs,
    matching::PairWiseMatchesContainer & map_PutativeMatches,
    system::ProgressInterface * my_progress_bar = nullptr) const override
  {
    if (!my_progress_bar)
      my_progress_bar = &system::ProgressInterface::dummy();
    my_progress_bar->Restart(pairs.size(), "- Matching (B200) -");

    // Images referenced by the pair list, in ascending id order -> dense arena slots.
    std::set<IndexT> ids;
    for (const auto & p : pairs) { ids.insert(p.first); ids.insert(p.second); }
    std::map<IndexT, uint32_t> slot;
    std::vector<std::shared_ptr<features::Regions>> held;      // keeps DescriptorRawData() alive
    std::vector<uint32_t> counts;
    for (const IndexT id : ids)
    {
      std::shared_ptr<features::Regions> r = regions_provider->get(id);
      const bool usable = r && r->IsScalar() && r->Type_id() == typeid(unsigned char).name()
                          && r->DescriptorLength() == OMVG_DESC_LEN;
      if (r && r->RegionCount() != 0 && !usable)
      {
        OPENMVG_LOG_ERROR << "Matcher_Regions_B200 handles 128-D unsigned char descriptors only.";
        (*my_progress_bar) += pairs.size();
        return;
      }
      slot[id] = static_cast<uint32_t>(held.size());
      counts.push_back(r ? static_cast<uint32_t>(r->RegionCount()) : 0u);
      held.push_back(r);
    }

    omvg_match_ctx * ctx = nullptr;
    if (omvg_match_create(&ctx, device_) != OMVG_OK ||
        omvg_match_set_images(ctx, static_cast<uint32_t>(counts.size()), counts.data()) != OMVG_OK)
    {
      OPENMVG_LOG_ERROR << "omvg_b200: " << omvg_last_error();
      omvg_match_destroy(ctx);
      (*my_progress_bar) += pairs.size();
      return;
    }
    int rc = OMVG_OK;
    for (size_t k = 0; k < held.size() && rc == OMVG_OK; ++k)
      if (counts[k])
        rc = omvg_match_upload_host(ctx, static_cast<uint32_t>(k),
                                    static_cast<const uint8_t *>(held[k]->DescriptorRawData()));
    if (rc == OMVG_OK) rc = omvg_match_prepare(ctx);
    if (rc != OMVG_OK)
    {
      OPENMVG_LOG_ERROR << "omvg_b200: " << omvg_last_error();
      omvg_match_destroy(ctx);
      (*my_progress_bar) += pairs.size();
      return;
    }

    // Pair_Set is ordered: the CSR rows come back in the same lexicographic order the reference
    // iterates (map_Pairs by I, then J).
    std::vector<uint32_t> pi, pj;
    std::vector<Pair> order;
    for (const auto & p : pairs)
    {
      if (my_progress_bar->hasBeenCanceled()) break;
      pi.push_back(slot[p.first]); pj.push_back(slot[p.second]); order.push_back(p);
    }
    const uint64_t * offsets = nullptr; const uint32_t * ij = nullptr; uint64_t n_matches = 0;
    if (omvg_match_run(ctx, pi.data(), pj.data(), pi.size(), f_dist_ratio_) != OMVG_OK ||
        omvg_match_fetch(ctx, &offsets, &ij, &n_matches) != OMVG_OK)
    {
      OPENMVG_LOG_ERROR << "omvg_b200: " << omvg_last_error();
      omvg_match_destroy(ctx);
      (*my_progress_bar) += pairs.size();
      return;
    }
    for (size_t p = 0; p < order.size(); ++p)
    {
      const uint64_t b = offsets[p], e = offsets[p + 1];
      if (e > b)                                             // Matcher_Regions.cpp:99-102
      {
        matching::IndMatches m;
        m.reserve(e - b);
        for (uint64_t k = b; k < e; ++k) m.emplace_back(ij[2 * k], ij[2 * k + 1]);
        map_PutativeMatches.insert({order[p], std::move(m)});
      }
      ++(*my_progress_bar);
    }
    omvg_match_destroy(ctx);
  }

  private:
  float f_dist_ratio_;
  int device_;
};

}  // namespace matching_image_collection
}  // namespace openMVG

#endif  // OPENMVG_B200_MATCHER_REGIONS_B200_HPP
