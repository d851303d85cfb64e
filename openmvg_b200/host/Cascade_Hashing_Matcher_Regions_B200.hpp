// Cascade_Hashing_Matcher_Regions_B200 — drop-in replacement for
//   openMVG::matching_image_collection::Cascade_Hashing_Matcher_Regions(distRatio)
// (reference: src/openMVG/matching_image_collection/Cascade_Hashing_Matcher_Regions.{hpp,cpp}, the matcher
// openMVG_main_ComputeMatches picks by default for scalar descriptors: "FASTCASCADEHASHINGL2") on top of the
// C ABI of libomvg_b200.so.  Same constructor argument, same Match() signature, same observable behaviour:
//   * one CascadeHasher for the collection, projections drawn from std::mt19937(default_seed) +
//     std::normal_distribution<>(0,1) in the order of CascadeHasher::Init   (cascade_hasher.hpp:142-162)
//   * one zero-mean descriptor over the images the pair list names          (.cpp:78-105)
//   * database = regions of I, queries = regions of J, ratio test, output (i, j) (.cpp:150-189)
//   * IndMatch::getDeduplicated (sort by (i, j), unique) and the equal-coordinates filter of
//     IndMatchDecorator — host-side clean-ups done here with openMVG's own functions (.cpp:191-198)
//   * only non-empty results are inserted; pairs with an empty I are skipped, ++progress once per pair
// Only 128-byte unsigned-char scalar regions go to the GPU; anything else returns without matches after logging.
//
// Header-only; compile inside an openMVG build and link libomvg_b200.so.
#ifndef OPENMVG_B200_CASCADE_HASHING_MATCHER_REGIONS_B200_HPP
#define OPENMVG_B200_CASCADE_HASHING_MATCHER_REGIONS_B200_HPP

#include "openMVG/features/regions.hpp"
#include "openMVG/matching/indMatch.hpp"
#include "openMVG/matching/indMatchDecoratorXY.hpp"
#include "openMVG/matching_image_collection/Matcher.hpp"
#include "openMVG/sfm/pipelines/sfm_regions_provider.hpp"
#include "openMVG/system/logger.hpp"
#include "openMVG/system/progressinterface.hpp"

#include "omvg_b200.h"

#include <map>
#include <memory>
#include <random>
#include <set>
#include <vector>

namespace openMVG {
namespace matching_image_collection {

class Cascade_Hashing_Matcher_Regions_B200 : public Matcher
{
  public:
  explicit Cascade_Hashing_Matcher_Regions_B200(float dist_ratio, int device = 0)
    : Matcher(), f_dist_ratio_(dist_ratio), device_(device) {}

  void Match(
    const std::shared_ptr<sfm::Regions_Provider> & regions_provider,
    const Pair_Set & pairs,
    matching::PairWiseMatchesContainer & map_PutativeMatches,
    system::ProgressInterface * my_progress_bar = nullptr) const override
  {
    if (!my_progress_bar)
      my_progress_bar = &system::ProgressInterface::dummy();
    my_progress_bar->Restart(pairs.size(), "- Matching (B200, cascade hashing) -");

    std::set<IndexT> ids;
    for (const auto & p : pairs) { ids.insert(p.first); ids.insert(p.second); }
    std::map<IndexT, uint32_t> slot;
    std::vector<std::shared_ptr<features::Regions>> held;
    std::vector<uint32_t> counts;
    for (const IndexT id : ids)
    {
      std::shared_ptr<features::Regions> r = regions_provider->get(id);
      const bool usable = r && r->IsScalar() && r->Type_id() == typeid(unsigned char).name()
                          && r->DescriptorLength() == OMVG_DESC_LEN;
      if (r && r->RegionCount() != 0 && !usable)
      {
        OPENMVG_LOG_ERROR << "Cascade_Hashing_Matcher_Regions_B200 handles 128-D unsigned char descriptors only.";
        (*my_progress_bar) += pairs.size();
        return;
      }
      slot[id] = static_cast<uint32_t>(held.size());
      counts.push_back(r ? static_cast<uint32_t>(r->RegionCount()) : 0u);
      held.push_back(r);
    }

    // The projections of CascadeHasher::Init(128): same generator, distribution, seed and draw order.
    std::vector<float> primary(OMVG_DESC_LEN * OMVG_DESC_LEN), secondary(6 * 10 * OMVG_DESC_LEN);
    {
      std::mt19937 gen(std::mt19937::default_seed);
      std::normal_distribution<> d(0, 1);
      for (float & v : primary) v = static_cast<float>(d(gen));
      for (float & v : secondary) v = static_cast<float>(d(gen));
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

    std::vector<uint32_t> pi, pj;
    std::vector<Pair> order;
    for (const auto & p : pairs)
    {
      if (my_progress_bar->hasBeenCanceled()) break;
      pi.push_back(slot[p.first]); pj.push_back(slot[p.second]); order.push_back(p);
    }
    const uint64_t * offsets = nullptr; const uint32_t * ij = nullptr; uint64_t n_matches = 0;
    // every image of `ids` is named by a pair, so all of them enter the zero-mean descriptor (used = NULL)
    if (omvg_match_cascade_prepare(ctx, primary.data(), secondary.data(), nullptr) != OMVG_OK ||
        omvg_match_cascade_run(ctx, pi.data(), pj.data(), pi.size(), f_dist_ratio_) != OMVG_OK ||
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
      if (e > b)
      {
        const std::shared_ptr<features::Regions> & ri = held[pi[p]], & rj = held[pj[p]];
        if (ri->Type_id() == rj->Type_id())                    // .cpp:143-147
        {
          matching::IndMatches m;
          m.reserve(e - b);
          for (uint64_t k = b; k < e; ++k) m.emplace_back(ij[2 * k], ij[2 * k + 1]);
          matching::IndMatch::getDeduplicated(m);               // .cpp:192
          const std::vector<features::PointFeature> xi = ri->GetRegionsPositions(), xj = rj->GetRegionsPositions();
          matching::IndMatchDecorator<float> dedup(m, xi, xj);  // .cpp:195-198
          dedup.getDeduplicated(m);
          if (!m.empty())
            map_PutativeMatches.insert({order[p], std::move(m)});
        }
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

#endif  // OPENMVG_B200_CASCADE_HASHING_MATCHER_REGIONS_B200_HPP
