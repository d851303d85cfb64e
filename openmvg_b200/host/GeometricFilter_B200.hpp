// ImageCollectionGeometricFilter_B200 — the fundamental-matrix geometric filter of openMVG_main_GeometricFilter on the
// B200: drop-in for
//     ImageCollectionGeometricFilter(sfm_data, regions_provider)
//       .Robust_model_estimation(GeometricFilter_FMatrix_AC(dPrecision, iteration), putative_matches)
// (reference: matching_image_collection/GeometricFilter.hpp:44-133, F_ACRobust.hpp:45-106, called at
// software/SfM/main_GeometricFilter.cpp:300-308).  The per-pair feature positions come from openMVG's own
// MatchesPairToMat (un-distorted when the view has an intrinsic), all pairs go to the device in ONE call of
// omvg_geom_fundamental_acransac (one CTA per pair; same MT19937 sample sequence and NFA decisions as ACRANSAC), and a
// pair is kept iff it has more than 2.5 * 7 inliers (F_ACRobust.hpp:85).  Guided matching is not part of this path.
//
// Robust_model_estimation_H is the same for GeometricFilter_HMatrix_AC (H_ACRobust.hpp: 4-point homography).
// Header-only; compile inside an openMVG build and link libomvg_b200.so.
#ifndef OPENMVG_B200_GEOMETRIC_FILTER_B200_HPP
#define OPENMVG_B200_GEOMETRIC_FILTER_B200_HPP

#include "openMVG/matching/indMatch.hpp"
#include "openMVG/matching_image_collection/Geometric_Filter_utils.hpp"
#include "openMVG/sfm/pipelines/sfm_regions_provider.hpp"
#include "openMVG/sfm/sfm_data.hpp"
#include "openMVG/system/logger.hpp"
#include "openMVG/system/progressinterface.hpp"

#include "omvg_b200.h"

#include <memory>
#include <vector>

namespace openMVG {
namespace matching_image_collection {

struct ImageCollectionGeometricFilter_B200
{
  ImageCollectionGeometricFilter_B200(const sfm::SfM_Data * sfm_data, const std::shared_ptr<sfm::Regions_Provider> & regions_provider, int device = 0)
    : sfm_data_(sfm_data), regions_provider_(regions_provider), device_(device) {}

  // F model, a-contrario (GeometricFilter_FMatrix_AC(dPrecision, iteration)); returns false on a device error
  bool Robust_model_estimation_F(const matching::PairWiseMatches & putative_matches, double dPrecision = 4.0, uint32_t iteration = 2048,
                                 system::ProgressInterface * my_progress_bar = nullptr)
  { return Robust_model_estimation(OMVG_GEOM_FUNDAMENTAL, putative_matches, dPrecision, iteration, my_progress_bar); }
  // H model (GeometricFilter_HMatrix_AC(dPrecision, iteration), H_ACRobust.hpp:46-112): kept iff more than 2.5 * 4 inliers
  bool Robust_model_estimation_H(const matching::PairWiseMatches & putative_matches, double dPrecision = 4.0, uint32_t iteration = 2048,
                                 system::ProgressInterface * my_progress_bar = nullptr)
  { return Robust_model_estimation(OMVG_GEOM_HOMOGRAPHY, putative_matches, dPrecision, iteration, my_progress_bar); }

  bool Robust_model_estimation(int32_t model, const matching::PairWiseMatches & putative_matches, double dPrecision, uint32_t iteration,
                               system::ProgressInterface * my_progress_bar = nullptr)
  {
    if (!my_progress_bar) my_progress_bar = &system::ProgressInterface::dummy();
    my_progress_bar->Restart(putative_matches.size(), "- Geometric filtering (B200) -");
    std::vector<uint64_t> offsets(1, 0);
    std::vector<double> xI, xJ;
    std::vector<int32_t> size;
    std::vector<const std::pair<const Pair, matching::IndMatches> *> order;
    for (const auto & it : putative_matches)
    {
      Mat2X a, b;
      MatchesPairToMat(it.first, it.second, sfm_data_, regions_provider_, a, b);
      for (Mat2X::Index k = 0; k < a.cols(); ++k) { xI.push_back(a(0, k)); xI.push_back(a(1, k)); xJ.push_back(b(0, k)); xJ.push_back(b(1, k)); }
      offsets.push_back(offsets.back() + static_cast<uint64_t>(a.cols()));
      const sfm::View * vI = sfm_data_->GetViews().at(it.first.first).get(), * vJ = sfm_data_->GetViews().at(it.first.second).get();
      size.insert(size.end(), {static_cast<int32_t>(vI->ui_width), static_cast<int32_t>(vI->ui_height), static_cast<int32_t>(vJ->ui_width), static_cast<int32_t>(vJ->ui_height)});
      order.push_back(&it);
    }
    const size_t n_pairs = order.size();
    std::vector<uint32_t> inliers(xI.size() / 2 + 1), n_inliers(n_pairs + 1);
    F_.assign(9 * n_pairs, 0.0); stats_.assign(2 * n_pairs, 0.0);
    _map_GeometricMatches.clear();
    const double min_samples = model == OMVG_GEOM_FUNDAMENTAL ? 7.0 : 4.0;
    if (omvg_geom_acransac(device_, model, n_pairs, offsets.data(), xI.data(), xJ.data(), size.data(), dPrecision, iteration,
                                       inliers.data(), n_inliers.data(), F_.data(), stats_.data()) != OMVG_OK)
    {
      OPENMVG_LOG_ERROR << "omvg_b200: " << omvg_last_error();
      return false;
    }
    for (size_t p = 0; p < n_pairs; ++p)
    {
      if (n_inliers[p] > min_samples * 2.5)                      // F_ACRobust.hpp:85, H_ACRobust.hpp:97
      {
        matching::IndMatches geometric_inliers;
        geometric_inliers.reserve(n_inliers[p]);
        for (uint32_t k = 0; k < n_inliers[p]; ++k) geometric_inliers.push_back(order[p]->second[inliers[offsets[p] + k]]);
        _map_GeometricMatches.insert({order[p]->first, std::move(geometric_inliers)});
      }
      ++(*my_progress_bar);
    }
    return true;
  }

  const matching::PairWiseMatches & Get_geometric_matches() const { return _map_GeometricMatches; }

  const sfm::SfM_Data * sfm_data_;
  const std::shared_ptr<sfm::Regions_Provider> regions_provider_;   // (a copy: the reference keeps a reference, which dangles when bound to a converted temporary)
  int device_;
  matching::PairWiseMatches _map_GeometricMatches;
  std::vector<double> F_, stats_;                               // per pair (map order): un-normalised F, {errorMax, minNFA}
};

}  // namespace matching_image_collection
}  // namespace openMVG

#endif  // OPENMVG_B200_GEOMETRIC_FILTER_B200_HPP
