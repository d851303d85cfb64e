// Bundle_Adjustment_B200 — drop-in replacement for openMVG::sfm::Bundle_Adjustment_Ceres
// (reference: src/openMVG/sfm/sfm_data_BA_ceres.{hpp,cpp}) implementing the abstract
// openMVG::sfm::Bundle_Adjustment (sfm_data_BA.hpp:91-105) on top of the C ABI of libomvg_b200.so.
//
// Adjust() does what Bundle_Adjustment_Ceres::Adjust does around ceres::Solve:
//   * packs poses as angle-axis + t = -R*C            (sfm_data_BA_ceres.cpp:260-271)
//   * packs intrinsics with getParams()               (:310-317)
//   * one residual per (landmark, view) observation   (:354-396), points optimised in place
//   * Huber(4^2) unless bUse_loss_function_ is off    (:242-253)
//   * writes poses / intrinsics back by the same rules (:528-568) — done inside omvg_ba_solve
//   * returns false, leaving the scene untouched, when the solution is not usable (:503-507)
//   * ground control points (:398-451): fixed landmarks, residuals x weight, no loss
//   * motion priors (:181-240, 454-473, 570-573): the scene is registered to the pose-centre priors
//     with openMVG's own LeastMedianOfSquares / ApplySimilarity (host-side geometry, a few hundred
//     points), the prior residuals run on the GPU under HuberLoss(Square(median fitting error))
// NOT on the GPU path (Adjust logs an error and returns false — nothing is silently routed to a CPU
// solver): CAMERA_SPHERICAL, more than 32 intrinsic groups, more than 32768 poses.
//
// Header-only; compile inside an openMVG build (needs openMVG + ceres/rotation.h) and link
// libomvg_b200.so.
#ifndef OPENMVG_B200_BUNDLE_ADJUSTMENT_B200_HPP
#define OPENMVG_B200_BUNDLE_ADJUSTMENT_B200_HPP

#include "openMVG/cameras/Camera_Common.hpp"
#include "openMVG/cameras/Camera_Intrinsics.hpp"
#include "openMVG/sfm/sfm_data.hpp"
#include "openMVG/sfm/sfm_data_BA.hpp"
#include "openMVG/sfm/sfm_data_transform.hpp"
#include "openMVG/sfm/sfm_view_priors.hpp"
#include "openMVG/geometry/Similarity3.hpp"
#include "openMVG/geometry/Similarity3_Kernel.hpp"
#include "openMVG/robust_estimation/robust_estimator_LMeds.hpp"
#include "openMVG/system/logger.hpp"
#include "openMVG/types.hpp"

#include <ceres/rotation.h>

#include "omvg_b200.h"

#include <algorithm>
#include <limits>
#include <map>
#include <vector>

namespace openMVG {
namespace sfm {

class Bundle_Adjustment_B200 : public Bundle_Adjustment
{
  public:
  struct BA_B200_options            // the knobs of BA_Ceres_options that still mean something here
  {
    BA_B200_options()
      : bVerbose_(false), bUse_loss_function_(true), max_num_iterations_(50),
        parameter_tolerance_(1e-8), gradient_tolerance_(1e-10), device_(0) {}
    bool bVerbose_;
    bool bUse_loss_function_;
    unsigned int max_num_iterations_;
    double parameter_tolerance_;
    double gradient_tolerance_;
    int device_;
  };

  Bundle_Adjustment_B200() {}
  explicit Bundle_Adjustment_B200(const BA_B200_options & options) : options_(options) {}

  BA_B200_options & b200_options() { return options_; }
  const omvg_ba_summary & summary() const { return summary_; }

  bool Adjust(SfM_Data & sfm_data, const Optimize_Options & options) override
  {
    // ---- motion priors: register the scene to the prior frame first (sfm_data_BA_ceres.cpp:183-236)
    double prior_fit = 0.0;
    geometry::Similarity3 sim_to_center;
    const bool b_usable_prior = options.use_motion_priors_opt && sfm_data.GetViews().size() > 3
                                && RegisterToPriors(sfm_data, prior_fit, sim_to_center);
    // ---- dense indices
    std::map<IndexT, int32_t> pose_idx, intr_idx, view_idx;
    std::vector<double> poses, intrinsics, points, obs_xy;
    std::vector<int32_t> intr_model, view_pose, view_intr, obs_view, obs_point;
    for (const auto & pose_it : sfm_data.poses)
    {
      const geometry::Pose3 & pose = pose_it.second;
      const Mat3 R = pose.rotation();
      const Vec3 t = pose.translation();
      double aa[3];
      ceres::RotationMatrixToAngleAxis((const double *)R.data(), aa);    // as sfm_data_BA_ceres.cpp:268-269
      pose_idx[pose_it.first] = static_cast<int32_t>(poses.size() / 6);
      poses.insert(poses.end(), {aa[0], aa[1], aa[2], t(0), t(1), t(2)});
    }
    for (const auto & intr_it : sfm_data.intrinsics)
    {
      const cameras::EINTRINSIC type = intr_it.second->getType();
      if (!cameras::isPinhole(type))
      {
        OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: camera model " << int(type) << " is not on the GPU path.";
        return false;
      }
      const std::vector<double> p = intr_it.second->getParams();
      intr_idx[intr_it.first] = static_cast<int32_t>(intr_model.size());
      intr_model.push_back(static_cast<int32_t>(type));
      intrinsics.resize(intrinsics.size() + OMVG_BA_INTR_STRIDE, 0.0);
      for (size_t k = 0; k < p.size() && k < OMVG_BA_INTR_STRIDE; ++k)
        intrinsics[intrinsics.size() - OMVG_BA_INTR_STRIDE + k] = p[k];
    }
    std::vector<Landmark *> lm;
    for (auto & s : sfm_data.structure)
    {
      const int32_t j = static_cast<int32_t>(lm.size());
      lm.push_back(&s.second);
      points.insert(points.end(), {s.second.X(0), s.second.X(1), s.second.X(2)});
      for (const auto & obs_it : s.second.obs)
      {
        auto v = view_idx.find(obs_it.first);
        if (v == view_idx.end())
        {
          const View * view = sfm_data.views.at(obs_it.first).get();
          const auto p = pose_idx.find(view->id_pose);
          const auto q = intr_idx.find(view->id_intrinsic);
          if (p == pose_idx.end() || q == intr_idx.end())
          {
            OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: observation in a view without pose/intrinsic.";
            return false;
          }
          v = view_idx.emplace(obs_it.first, static_cast<int32_t>(view_pose.size())).first;
          view_pose.push_back(p->second);
          view_intr.push_back(q->second);
        }
        obs_view.push_back(v->second);
        obs_point.push_back(j);
        obs_xy.push_back(obs_it.second.x(0));
        obs_xy.push_back(obs_it.second.x(1));
      }
    }
    // ---- ground control points (:398-451): appended as fixed landmarks with weighted, loss-free residuals
    std::vector<double> obs_weight;
    std::vector<uint8_t> obs_no_loss, point_fixed;
    if (options.control_point_opt.bUse_control_points && !sfm_data.control_points.empty())
    {
      obs_weight.assign(obs_view.size(), 1.0);
      obs_no_loss.assign(obs_view.size(), 0);
      point_fixed.assign(lm.size(), 0);
      int32_t j = static_cast<int32_t>(lm.size());
      for (const auto & gcp : sfm_data.control_points)
      {
        if (gcp.second.obs.empty())
        {
          OPENMVG_LOG_ERROR << "Cannot use this GCP id: " << gcp.first << ". There is not linked image observation.";
          continue;
        }
        points.insert(points.end(), {gcp.second.X(0), gcp.second.X(1), gcp.second.X(2)});
        point_fixed.push_back(1);
        for (const auto & obs_it : gcp.second.obs)
        {
          auto v = view_idx.find(obs_it.first);
          if (v == view_idx.end())
          {
            const View * view = sfm_data.views.at(obs_it.first).get();
            const auto p = pose_idx.find(view->id_pose);
            const auto q = intr_idx.find(view->id_intrinsic);
            if (p == pose_idx.end() || q == intr_idx.end())
            {
              OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: GCP observation in a view without pose/intrinsic.";
              return false;
            }
            v = view_idx.emplace(obs_it.first, static_cast<int32_t>(view_pose.size())).first;
            view_pose.push_back(p->second);
            view_intr.push_back(q->second);
          }
          obs_view.push_back(v->second);
          obs_point.push_back(j);
          obs_xy.push_back(obs_it.second.x(0));
          obs_xy.push_back(obs_it.second.x(1));
          obs_weight.push_back(options.control_point_opt.weight);
          obs_no_loss.push_back(1);
        }
        ++j;
      }
    }
    // ---- pose-centre prior residuals (:455-472).  The reference keys the pose block by the prior's
    // id_view (map_poses.at(prior->id_view)); mirrored here, a prior whose id_view names no pose is an error.
    std::vector<int32_t> prior_pose;
    std::vector<double> prior_center, prior_weight;
    if (b_usable_prior)
    {
      for (const auto & view_it : sfm_data.GetViews())
      {
        const ViewPriors * prior = dynamic_cast<const ViewPriors *>(view_it.second.get());
        if (prior == nullptr || !prior->b_use_pose_center_ || !sfm_data.IsPoseAndIntrinsicDefined(prior)) continue;
        const auto p = pose_idx.find(prior->id_view);
        if (p == pose_idx.end())
        {
          OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: pose prior of view " << prior->id_view << " has no pose of that id.";
          return false;
        }
        prior_pose.push_back(p->second);
        prior_center.insert(prior_center.end(), {prior->pose_center_(0), prior->pose_center_(1), prior->pose_center_(2)});
        prior_weight.insert(prior_weight.end(), {prior->center_weight_(0), prior->center_weight_(1), prior->center_weight_(2)});
      }
    }
    omvg_ba_problem P = omvg_ba_problem();
    if (!obs_weight.empty()) { P.obs_weight = obs_weight.data(); P.obs_no_loss = obs_no_loss.data(); P.point_fixed = point_fixed.data(); }
    if (!prior_pose.empty())
    {
      P.n_priors = static_cast<int32_t>(prior_pose.size());
      P.prior_pose = prior_pose.data(); P.prior_center = prior_center.data(); P.prior_weight = prior_weight.data();
      P.prior_huber_a = prior_fit * prior_fit;            // HuberLoss(Square(pose_center_robust_fitting_error))
    }
    P.n_poses = static_cast<int32_t>(poses.size() / 6);
    P.n_intrinsics = static_cast<int32_t>(intr_model.size());
    P.n_points = static_cast<int32_t>(points.size() / 3);
    P.n_views = static_cast<int32_t>(view_pose.size());
    P.n_obs = static_cast<int64_t>(obs_view.size());
    P.poses = poses.data(); P.intrinsics = intrinsics.data(); P.intr_model = intr_model.data();
    P.points = points.data(); P.view_pose = view_pose.data(); P.view_intr = view_intr.data();
    P.obs_view = obs_view.data(); P.obs_point = obs_point.data(); P.obs_xy = obs_xy.data();

    omvg_ba_options O;
    omvg_ba_default_options(&O);
    O.intrinsics_opt = static_cast<int32_t>(options.intrinsics_opt);
    O.extrinsics_opt = static_cast<int32_t>(options.extrinsics_opt);
    O.structure_opt = options.structure_opt == Structure_Parameter_Type::ADJUST_ALL ? 1 : 0;
    O.use_loss = options_.bUse_loss_function_ ? 1 : 0;
    O.max_num_iterations = static_cast<int32_t>(options_.max_num_iterations_);
    O.parameter_tolerance = options_.parameter_tolerance_;
    O.gradient_tolerance = options_.gradient_tolerance_;
    O.verbose = options_.bVerbose_ ? 1 : 0;
    O.device = options_.device_;

    const int rc = omvg_ba_solve(&P, &O, &summary_);
    if (rc != OMVG_OK)
    {
      OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: " << omvg_last_error();
      return false;                                       // scene untouched (sfm_data_BA_ceres.cpp:503-507)
    }
    // ---- unpack (omvg_ba_solve already applied Adjust's write-back rules to the flat arrays)
    if (options.extrinsics_opt != Extrinsic_Parameter_Type::NONE)
    {
      for (auto & pose_it : sfm_data.poses)
      {
        const double * x = &poses[6 * pose_idx[pose_it.first]];
        Mat3 R;
        ceres::AngleAxisToRotationMatrix(x, R.data());
        const Vec3 t(x[3], x[4], x[5]);
        pose_it.second = geometry::Pose3(R, -R.transpose() * t);
      }
    }
    if (options.intrinsics_opt != cameras::Intrinsic_Parameter_Type::NONE)
    {
      for (auto & intr_it : sfm_data.intrinsics)
      {
        const double * x = &intrinsics[OMVG_BA_INTR_STRIDE * intr_idx[intr_it.first]];
        const size_t k = intr_it.second->getParams().size();
        intr_it.second->updateFromParams(std::vector<double>(x, x + k));
      }
    }
    if (options.structure_opt == Structure_Parameter_Type::ADJUST_ALL)
      for (size_t j = 0; j < lm.size(); ++j)
        lm[j]->X = Vec3(points[3 * j], points[3 * j + 1], points[3 * j + 2]);
    if (b_usable_prior)                                   // back to the original scene centroid (:570-573)
      ApplySimilarity(sim_to_center.inverse(), sfm_data, true);
    return true;
  }

  private:
  // The registration Adjust runs before it builds the problem when motion priors are on: robust similarity
  // SfM centres -> prior centres, median fitting error, scene moved into the prior frame and centred.
  // A few hundred 3-D points of host geometry; openMVG's own estimators are used so the frame (and the
  // LMedS sampling sequence) is the reference's.
  static bool RegisterToPriors(SfM_Data & sfm_data, double & fitting_error, geometry::Similarity3 & sim_to_center)
  {
    std::vector<Vec3> X_SfM, X_GPS;
    for (const auto & view_it : sfm_data.GetViews())
    {
      const ViewPriors * prior = dynamic_cast<const ViewPriors *>(view_it.second.get());
      if (prior == nullptr || !prior->b_use_pose_center_ || !sfm_data.IsPoseAndIntrinsicDefined(prior)) continue;
      X_SfM.push_back(sfm_data.GetPoses().at(prior->id_pose).center());
      X_GPS.push_back(prior->pose_center_);
    }
    if (!(X_GPS.size() > 3))
    {
      OPENMVG_LOG_WARNING << "Cannot used the motion prior, insufficient number of motion priors/poses";
      return false;
    }
    const Mat sfm_mat = Eigen::Map<Mat>(X_SfM[0].data(), 3, X_SfM.size());
    const Mat gps_mat = Eigen::Map<Mat>(X_GPS[0].data(), 3, X_GPS.size());
    geometry::kernel::Similarity3_Kernel kernel(sfm_mat, gps_mat);
    geometry::Similarity3 sim;
    if (robust::LeastMedianOfSquares(kernel, &sim) == std::numeric_limits<double>::max()) return false;
    std::vector<double> err(X_SfM.size());
    for (size_t i = 0; i < X_SfM.size(); ++i) err[i] = (sim(X_SfM[i]) - X_GPS[i]).norm();
    std::sort(err.begin(), err.end());
    fitting_error = err[err.size() / 2];
    ApplySimilarity(sim, sfm_data);
    Vec3 centroid = Vec3::Zero();
    for (const auto & pose_it : sfm_data.poses) centroid += (pose_it.second.center() / (double)sfm_data.poses.size());
    sim_to_center = geometry::Similarity3(geometry::Pose3(Mat3::Identity(), centroid), 1.0);
    ApplySimilarity(sim_to_center, sfm_data, true);
    return true;
  }

  private:
  BA_B200_options options_;
  omvg_ba_summary summary_{};
};

}  // namespace sfm
}  // namespace openMVG

#endif  // OPENMVG_B200_BUNDLE_ADJUSTMENT_B200_HPP
