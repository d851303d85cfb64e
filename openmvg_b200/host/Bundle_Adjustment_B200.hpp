// Bundle_Adjustment_B200 — drop-in replacement for openMVG::sfm::Bundle_Adjustment_Ceres
// (reference: src/openMVG/sfm/sfm_data_BA_ceres.{hpp,cpp}) implementing the abstract
// openMVG::sfm::Bundle_Adjustment (sfm_data_BA.hpp:91-105) on top of the C ABI of libomvg_b200.so.
//
// Adjust() does what Bundle_Adjustment_Ceres::Adjust does around ceres::Solve:
//   * packs poses as angle-axis + t = -R*C            (sfm_data_BA_ceres.cpp:260-271)
//   * packs intrinsics with getParams()               (:310-317)
//   * one residual per (landmark, view) observation   (:354-396), points optimised in place
//   * Huber(4^2) unless bUse_loss_function_ is off    (:242-253)
//   * writes poses / intrinsics back by the same rules (:528-568) — done inside the C ABI
//   * returns false, leaving the scene untouched, when the solution is not usable (:503-507)
//   * ground control points (:398-451): fixed landmarks, residuals x weight, no loss
//   * motion priors (:181-240, 454-473, 570-573): the scene is registered to the pose-centre priors
//     with openMVG's own LeastMedianOfSquares / ApplySimilarity (host-side geometry, a few hundred
//     points), the prior residuals run on the GPU under HuberLoss(Square(median fitting error))
// Everything that can make Adjust return false without a solve (a camera model outside the GPU path, an
// observation in a view without pose) is checked BEFORE the prior registration moves the scene.
//
// AdjustAndReject() is the loop every incremental engine wraps around Adjust
//     do { BundleAdjustment(); } while (badTrackRejector(4.0, 50));        (sequential_SfM.cpp:205-211, 1226-1243)
// with ONE device-resident problem: the scene is packed and uploaded once, every round runs the LM solve from the
// previous solution, RemoveOutliers_PixelResidualError (sfm_data_filters.cpp:40-73) is evaluated on the device
// (a bit mask of removed observations comes back, not the residuals), RemoveOutliers_AngleError (:77-122) is host
// geometry on the flat arrays with openMVG's own AngleBetweenRay, and SfM_Data is written once at the end.
//
// Packing is one pass over the landmark hash map to collect pointers, then a parallel fill of pre-sized flat arrays;
// ids are resolved through dense tables (hash-map fallback for sparse ids).
//
// Header-only; compile inside an openMVG build (needs openMVG + ceres/rotation.h) and link libomvg_b200.so.
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

#ifdef OPENMVG_USE_OPENMP
#include <omp.h>
#endif

#include <algorithm>
#include <chrono>
#include <limits>
#include <unordered_map>
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
  // wall-clock of the last Adjust / AdjustAndReject, milliseconds (where the end-to-end time goes)
  struct Timing { double pack_ms = 0, solve_ms = 0, reject_ms = 0, angle_ms = 0, unpack_ms = 0, erase_ms = 0, destroy_ms = 0; int rounds = 0; };
  struct RejectStats { size_t residual_outliers = 0, short_tracks = 0, angle_tracks = 0; int rounds = 0; };

  Bundle_Adjustment_B200() {}
  explicit Bundle_Adjustment_B200(const BA_B200_options & options) : options_(options) {}

  BA_B200_options & b200_options() { return options_; }
  const omvg_ba_summary & summary() const { return summary_; }
  const Timing & timing() const { return timing_; }

  bool Adjust(SfM_Data & sfm_data, const Optimize_Options & options) override
  {
    timing_ = Timing();
    Flat f;
    geometry::Similarity3 sim_to_center;
    bool b_usable_prior = false;
    const auto t0 = now();
    if (!Pack(sfm_data, options, f, b_usable_prior, sim_to_center)) return false;
    timing_.pack_ms = ms_since(t0);

    omvg_ba_problem P = f.problem();
    const omvg_ba_options O = MakeOptions(options);
    const auto t1 = now();
    const int rc = omvg_ba_solve(&P, &O, &summary_);
    timing_.solve_ms = ms_since(t1);
    if (rc != OMVG_OK)
    {
      OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: " << omvg_last_error();
      if (b_usable_prior) UndoRegistration(sfm_data, f, sim_to_center);
      return false;                                       // scene as the caller passed it (sfm_data_BA_ceres.cpp:503-507)
    }
    const auto t2 = now();
    Unpack(sfm_data, options, f, true);
    if (b_usable_prior)                                   // back to the original scene centroid (:570-573)
      ApplySimilarity(sim_to_center.inverse(), sfm_data, true);
    timing_.unpack_ms = ms_since(t2);
    timing_.rounds = 1;
    return true;
  }

  // do { Adjust } while (RemoveOutliers_PixelResidualError(dPrecision, min_track_length) + RemoveOutliers_AngleError(min_angle) > count)
  // on one device-resident problem.  With motion priors the scene is re-registered by every Adjust of the
  // reference loop, so that case runs the plain loop over Adjust() with the same rejection rules.
  bool AdjustAndReject(SfM_Data & sfm_data, const Optimize_Options & options, double dPrecision = 4.0, size_t count = 50,
                       unsigned int min_track_length = 2, double min_angle_deg = 2.0, RejectStats * stats = nullptr)
  {
    timing_ = Timing();
    RejectStats st;
    Flat f;
    geometry::Similarity3 sim_to_center;
    bool b_usable_prior = false;
    const auto t0 = now();
    if (options.use_motion_priors_opt)
    {
      OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200::AdjustAndReject: motion priors re-register the scene every round; call Adjust() in a loop.";
      return false;
    }
    if (!Pack(sfm_data, options, f, b_usable_prior, sim_to_center)) return false;
    timing_.pack_ms = ms_since(t0);

    omvg_ba_problem P = f.problem();
    const omvg_ba_options O = MakeOptions(options);
    omvg_ba_ctx * ctx = nullptr;
    if (omvg_ba_create(&ctx, options_.device_, &P) != OMVG_OK)
    {
      OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: " << omvg_last_error();
      return false;
    }
    const size_t n_obs = f.obs_view.size(), n_reg = f.lm.size();
    std::vector<uint32_t> bits((n_obs + 31) / 32);
    std::vector<uint8_t> obs_alive(n_obs, 1), pt_alive(f.points.size() / 3, 1), pt_now(f.points.size() / 3, 0), kill(f.points.size() / 3, 0);
    AngleWork angle_work;
    PrepareAngleWork(f, angle_work);
    bool ok = true, again = true;
    while (again)
    {
      const auto t1 = now();
      const int rc = omvg_ba_run(ctx, &O, &summary_);
      if (rc == OMVG_OK) ok = omvg_ba_writeback(ctx, &O, f.poses.data(), f.intrinsics.data(), nullptr) == OMVG_OK;
      else ok = false;
      timing_.solve_ms += ms_since(t1);
      ++st.rounds;
      if (!ok) { OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: " << omvg_last_error(); break; }
      // ---- residual rule on the device
      const auto t2 = now();
      int64_t n_out = 0, n_short = 0;
      if (omvg_ba_reject_outliers(ctx, dPrecision, static_cast<int32_t>(min_track_length), bits.data(), pt_now.data(), &n_out, &n_short) != OMVG_OK)
      { OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: " << omvg_last_error(); ok = false; break; }
      ApplyBits(bits, obs_alive);
      for (size_t j = 0; j < n_reg; ++j) if (pt_now[j]) pt_alive[j] = 0;
      st.residual_outliers += static_cast<size_t>(n_out); st.short_tracks += static_cast<size_t>(n_short);
      timing_.reject_ms += ms_since(t2);
      // ---- angle rule on the host (poses / intrinsics of this round, flat arrays, openMVG's own ray geometry)
      const auto t3 = now();
      UnpackCameras(sfm_data, options, f);
      const size_t n_angle = AngleRule(f, angle_work, obs_alive, pt_alive, min_angle_deg, kill);
      if (n_angle)
      {
        int64_t nt = 0;
        if (omvg_ba_remove_points(ctx, kill.data(), bits.data(), &nt) != OMVG_OK)
        { OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: " << omvg_last_error(); ok = false; break; }
        ApplyBits(bits, obs_alive);
        for (size_t j = 0; j < n_reg; ++j) if (kill[j]) pt_alive[j] = 0;
      }
      st.angle_tracks += n_angle;
      timing_.angle_ms += ms_since(t3);
      again = static_cast<size_t>(n_out) + n_angle > count;
    }
    const auto t4 = now();
    if (ok)
    {
      ok = omvg_ba_writeback(ctx, &O, f.poses.data(), f.intrinsics.data(), f.points.data()) == OMVG_OK;
      if (ok)
      {
        Unpack(sfm_data, options, f, true);
        timing_.unpack_ms = ms_since(t4);
        const auto t5 = now();
        // the rejected observations / tracks leave SfM_Data exactly as the reference's filters erase them
        for (size_t o = 0; o < f.n_regular_obs; ++o)
          if (!obs_alive[o] && pt_alive[f.obs_point[o]]) f.lm[f.obs_point[o]]->obs.erase(f.obs_view_id[o]);
        for (size_t j = 0; j < n_reg; ++j)
          if (!pt_alive[j]) sfm_data.structure.erase(f.lm_id[j]);
        timing_.erase_ms = ms_since(t5);
      }
    }
    const auto t6 = now();
    omvg_ba_destroy(ctx);
    timing_.destroy_ms = ms_since(t6);
    timing_.rounds = st.rounds;
    if (stats) *stats = st;
    return ok;
  }

  private:
  // Short host loops: a handful of threads, not all 128 of a shared host (measured on a busy box: the write-back of
  // 100k points took 54 ms under a 128-thread OpenMP team and 4 ms otherwise; one slow core stalls the whole team)
  static int HostThreads()
  {
#ifdef OPENMVG_USE_OPENMP
    return std::max(1, std::min(16, omp_get_max_threads()));
#else
    return 1;
#endif
  }
  using clock_t_ = std::chrono::steady_clock;
  static clock_t_::time_point now() { return clock_t_::now(); }
  static double ms_since(clock_t_::time_point t) { return std::chrono::duration<double, std::milli>(clock_t_::now() - t).count(); }

  // id -> dense index: a vector when the ids are reasonably dense (every openMVG loader numbers from 0), else a hash map
  struct IdTable
  {
    std::vector<int32_t> dense; std::unordered_map<IndexT, int32_t> sparse; bool use_dense = true;
    void init(IndexT max_id, size_t n) { use_dense = static_cast<size_t>(max_id) < 16 * n + 65536; if (use_dense) dense.assign(static_cast<size_t>(max_id) + 1, -1); else sparse.reserve(2 * n); }
    void set(IndexT id, int32_t v) { if (use_dense) dense[id] = v; else sparse[id] = v; }
    int32_t get(IndexT id) const
    {
      if (use_dense) return id < dense.size() ? dense[id] : -1;
      const auto it = sparse.find(id); return it == sparse.end() ? -1 : it->second;
    }
  };

  // the flat scene of include/omvg_b200.h plus what is needed to write SfM_Data back
  struct Flat
  {
    std::vector<double> poses, intrinsics, points, obs_xy, obs_weight, prior_center, prior_weight;
    std::vector<int32_t> intr_model, view_pose, view_intr, obs_view, obs_point, prior_pose;
    std::vector<uint8_t> obs_no_loss, point_fixed;
    std::vector<IndexT> pose_id, intr_id, lm_id, obs_view_id;   // dense index -> openMVG id
    std::vector<geometry::Pose3 *> pose_ptr;
    std::vector<cameras::IntrinsicBase *> intr_ptr;
    std::vector<size_t> intr_nparams;
    std::vector<Landmark *> lm;                                 // regular landmarks (control points follow in `points`)
    std::vector<size_t> lm_first;                               // first observation of landmark j (n_reg + 1 entries)
    size_t n_regular_obs = 0;
    double prior_fit = 0.0;
    omvg_ba_problem problem()
    {
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
      return P;
    }
  };

  omvg_ba_options MakeOptions(const Optimize_Options & options) const
  {
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
    return O;
  }

  static bool SupportedModel(cameras::EINTRINSIC type)
  {
    return cameras::isPinhole(type) || type == cameras::CAMERA_SPHERICAL;
  }

  // SfM_Data -> flat arrays.  All checks that can fail come first; the prior registration (which moves the scene)
  // runs only once nothing but the solve itself can fail.
  bool Pack(SfM_Data & sfm_data, const Optimize_Options & options, Flat & f, bool & b_usable_prior, geometry::Similarity3 & sim_to_center)
  {
    // ---- intrinsics: supported models only
    IndexT max_intr = 0, max_pose = 0, max_view = 0;
    for (const auto & it : sfm_data.intrinsics)
    {
      if (!SupportedModel(it.second->getType()))
      {
        OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: camera model " << int(it.second->getType()) << " is not on the GPU path.";
        return false;
      }
      max_intr = std::max(max_intr, it.first);
    }
    for (const auto & it : sfm_data.poses) max_pose = std::max(max_pose, it.first);
    for (const auto & it : sfm_data.views) max_view = std::max(max_view, it.first);
    IdTable pose_idx, intr_idx, view_idx;
    pose_idx.init(max_pose, sfm_data.poses.size()); intr_idx.init(max_intr, sfm_data.intrinsics.size()); view_idx.init(max_view, sfm_data.views.size());
    // ---- dense numbering (map iteration order, as the reference's own loops)
    f.pose_id.reserve(sfm_data.poses.size()); f.pose_ptr.reserve(sfm_data.poses.size());
    for (auto & it : sfm_data.poses) { pose_idx.set(it.first, static_cast<int32_t>(f.pose_id.size())); f.pose_id.push_back(it.first); f.pose_ptr.push_back(&it.second); }
    for (auto & it : sfm_data.intrinsics) { intr_idx.set(it.first, static_cast<int32_t>(f.intr_id.size())); f.intr_id.push_back(it.first); f.intr_ptr.push_back(it.second.get()); }
    for (const auto & it : sfm_data.views)
    {
      const View * v = it.second.get();
      if (!v || v->id_pose == UndefinedIndexT || v->id_intrinsic == UndefinedIndexT) continue;
      const int32_t p = pose_idx.get(v->id_pose), q = intr_idx.get(v->id_intrinsic);
      if (p < 0 || q < 0) continue;
      view_idx.set(it.first, static_cast<int32_t>(f.view_pose.size()));
      f.view_pose.push_back(p); f.view_intr.push_back(q);
    }
    // ---- landmarks: one pass over the hash map for pointers and sizes, then a parallel fill
    const size_t n_reg = sfm_data.structure.size();
    f.lm.reserve(n_reg); f.lm_id.reserve(n_reg); f.lm_first.reserve(n_reg + 1);
    size_t n_obs = 0;
    for (auto & s : sfm_data.structure) { f.lm.push_back(&s.second); f.lm_id.push_back(s.first); f.lm_first.push_back(n_obs); n_obs += s.second.obs.size(); }
    f.lm_first.push_back(n_obs);
    f.n_regular_obs = n_obs;
    const bool use_gcp = options.control_point_opt.bUse_control_points && !sfm_data.control_points.empty();
    size_t n_gcp = 0, n_gcp_obs = 0;
    if (use_gcp)
      for (const auto & g : sfm_data.control_points)
      {
        if (g.second.obs.empty()) { OPENMVG_LOG_ERROR << "Cannot use this GCP id: " << g.first << ". There is not linked image observation."; continue; }
        ++n_gcp; n_gcp_obs += g.second.obs.size();
      }
    const size_t n_all = n_obs + n_gcp_obs;
    f.points.resize(3 * (n_reg + n_gcp)); f.obs_view.resize(n_all); f.obs_point.resize(n_all); f.obs_xy.resize(2 * n_all); f.obs_view_id.resize(n_all);
    int bad = 0;
#ifdef OPENMVG_USE_OPENMP
    #pragma omp parallel for schedule(static) reduction(+:bad) num_threads(HostThreads())
#endif
    for (int64_t j = 0; j < static_cast<int64_t>(n_reg); ++j)
    {
      const Landmark & L = *f.lm[j];
      f.points[3 * j] = L.X(0); f.points[3 * j + 1] = L.X(1); f.points[3 * j + 2] = L.X(2);
      size_t o = f.lm_first[j];
      for (const auto & obs_it : L.obs)
      {
        const int32_t v = view_idx.get(obs_it.first);
        if (v < 0) ++bad;
        f.obs_view[o] = v < 0 ? 0 : v; f.obs_view_id[o] = obs_it.first; f.obs_point[o] = static_cast<int32_t>(j);
        f.obs_xy[2 * o] = obs_it.second.x(0); f.obs_xy[2 * o + 1] = obs_it.second.x(1);
        ++o;
      }
    }
    if (bad)
    {
      OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: " << bad << " observation(s) in a view without pose/intrinsic.";
      return false;
    }
    // ---- ground control points (:398-451): appended as fixed landmarks with weighted, loss-free residuals
    if (use_gcp && n_gcp)
    {
      f.obs_weight.assign(n_all, 1.0); f.obs_no_loss.assign(n_all, 0); f.point_fixed.assign(n_reg + n_gcp, 0);
      size_t j = n_reg, o = n_obs;
      for (const auto & gcp : sfm_data.control_points)
      {
        if (gcp.second.obs.empty()) continue;
        f.points[3 * j] = gcp.second.X(0); f.points[3 * j + 1] = gcp.second.X(1); f.points[3 * j + 2] = gcp.second.X(2);
        f.point_fixed[j] = 1;
        for (const auto & obs_it : gcp.second.obs)
        {
          const int32_t v = view_idx.get(obs_it.first);
          if (v < 0) { OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: GCP observation in a view without pose/intrinsic."; return false; }
          f.obs_view[o] = v; f.obs_view_id[o] = obs_it.first; f.obs_point[o] = static_cast<int32_t>(j);
          f.obs_xy[2 * o] = obs_it.second.x(0); f.obs_xy[2 * o + 1] = obs_it.second.x(1);
          f.obs_weight[o] = options.control_point_opt.weight; f.obs_no_loss[o] = 1;
          ++o;
        }
        ++j;
      }
    }
    // ---- motion priors: every prior must name a pose BEFORE the scene is moved (the reference keys the pose block by
    // the prior's id_view: map_poses.at(prior->id_view), :464)
    if (options.use_motion_priors_opt && sfm_data.GetViews().size() > 3)
    {
      for (const auto & view_it : sfm_data.GetViews())
      {
        const ViewPriors * prior = dynamic_cast<const ViewPriors *>(view_it.second.get());
        if (prior == nullptr || !prior->b_use_pose_center_ || !sfm_data.IsPoseAndIntrinsicDefined(prior)) continue;
        if (pose_idx.get(prior->id_view) < 0)
        {
          OPENMVG_LOG_ERROR << "Bundle_Adjustment_B200: pose prior of view " << prior->id_view << " has no pose of that id.";
          return false;
        }
      }
      // register the scene to the prior frame (sfm_data_BA_ceres.cpp:183-236): from here on only the solve can fail
      b_usable_prior = RegisterToPriors(sfm_data, f.prior_fit, sim_to_center);
      if (b_usable_prior)
      {
        for (size_t j = 0; j < n_reg; ++j) { const Vec3 & X = f.lm[j]->X; f.points[3 * j] = X(0); f.points[3 * j + 1] = X(1); f.points[3 * j + 2] = X(2); }
        if (use_gcp && n_gcp)
        {
          size_t j = n_reg;
          for (const auto & gcp : sfm_data.control_points) { if (gcp.second.obs.empty()) continue; f.points[3 * j] = gcp.second.X(0); f.points[3 * j + 1] = gcp.second.X(1); f.points[3 * j + 2] = gcp.second.X(2); ++j; }
        }
        for (const auto & view_it : sfm_data.GetViews())
        {
          const ViewPriors * prior = dynamic_cast<const ViewPriors *>(view_it.second.get());
          if (prior == nullptr || !prior->b_use_pose_center_ || !sfm_data.IsPoseAndIntrinsicDefined(prior)) continue;
          f.prior_pose.push_back(pose_idx.get(prior->id_view));
          f.prior_center.insert(f.prior_center.end(), {prior->pose_center_(0), prior->pose_center_(1), prior->pose_center_(2)});
          f.prior_weight.insert(f.prior_weight.end(), {prior->center_weight_(0), prior->center_weight_(1), prior->center_weight_(2)});
        }
      }
    }
    // ---- poses (after the registration: it rewrites them) and intrinsics
    f.poses.resize(6 * f.pose_ptr.size());
    for (size_t p = 0; p < f.pose_ptr.size(); ++p)
    {
      const geometry::Pose3 & pose = *f.pose_ptr[p];
      const Mat3 R = pose.rotation();
      const Vec3 t = pose.translation();
      double * x = &f.poses[6 * p];
      ceres::RotationMatrixToAngleAxis((const double *)R.data(), x);    // as sfm_data_BA_ceres.cpp:268-269
      x[3] = t(0); x[4] = t(1); x[5] = t(2);
    }
    f.intrinsics.assign(OMVG_BA_INTR_STRIDE * f.intr_ptr.size(), 0.0);
    f.intr_model.resize(f.intr_ptr.size()); f.intr_nparams.resize(f.intr_ptr.size());
    for (size_t q = 0; q < f.intr_ptr.size(); ++q)
    {
      const std::vector<double> p = f.intr_ptr[q]->getParams();
      f.intr_model[q] = static_cast<int32_t>(f.intr_ptr[q]->getType());
      f.intr_nparams[q] = p.size();
      for (size_t k = 0; k < p.size() && k < OMVG_BA_INTR_STRIDE; ++k) f.intrinsics[OMVG_BA_INTR_STRIDE * q + k] = p[k];
    }
    return true;
  }

  // poses / intrinsics of the flat arrays (already under Adjust's write-back rules) into SfM_Data
  void UnpackCameras(SfM_Data & sfm_data, const Optimize_Options & options, const Flat & f) const
  {
    (void)sfm_data;
    if (options.extrinsics_opt != Extrinsic_Parameter_Type::NONE)
      for (size_t p = 0; p < f.pose_ptr.size(); ++p)
      {
        const double * x = &f.poses[6 * p];
        Mat3 R;
        ceres::AngleAxisToRotationMatrix(x, R.data());
        const Vec3 t(x[3], x[4], x[5]);
        *f.pose_ptr[p] = geometry::Pose3(R, -R.transpose() * t);
      }
    if (options.intrinsics_opt != cameras::Intrinsic_Parameter_Type::NONE)
      for (size_t q = 0; q < f.intr_ptr.size(); ++q)
      {
        const double * x = &f.intrinsics[OMVG_BA_INTR_STRIDE * q];
        f.intr_ptr[q]->updateFromParams(std::vector<double>(x, x + f.intr_nparams[q]));
      }
  }

  void Unpack(SfM_Data & sfm_data, const Optimize_Options & options, const Flat & f, bool with_points) const
  {
    UnpackCameras(sfm_data, options, f);
    if (with_points && options.structure_opt == Structure_Parameter_Type::ADJUST_ALL)
    {
      const int64_t n = static_cast<int64_t>(f.lm.size());
      for (int64_t j = 0; j < n; ++j)                         // (serial: 100k stores are ~1 ms, a thread team costs more)
        f.lm[j]->X = Vec3(f.points[3 * j], f.points[3 * j + 1], f.points[3 * j + 2]);
    }
  }

  static void ApplyBits(const std::vector<uint32_t> & bits, std::vector<uint8_t> & alive)
  {
    for (size_t w = 0; w < bits.size(); ++w)
    {
      uint32_t m = bits[w];
      while (m) { const int b = __builtin_ctz(m); m &= m - 1; const size_t o = 32 * w + b; if (o < alive.size()) alive[o] = 0; }
    }
  }

  // RemoveOutliers_AngleError (sfm_data_filters.cpp:77-122) on the flat arrays: a track whose largest angle between
  // any two of its (live) bearing rays is below min_angle is removed.  The rays are what AngleBetweenRay
  // (cameras/Camera_Intrinsics.hpp:263-280) forms — (R' * intrinsic(ud_pixel)).normalized() — computed once per
  // observation instead of once per pair, through the intrinsic's own batch operator() (one call per block of
  // observations of one intrinsic group: the per-point overload allocates a dynamic matrix per call).
  struct AngleWork { std::vector<uint32_t> order; std::vector<size_t> seg; std::vector<Vec3> rays; };
  static void PrepareAngleWork(const Flat & f, AngleWork & w)
  {
    const size_t n = f.n_regular_obs, ni = f.intr_ptr.size();
    w.seg.assign(ni + 1, 0);
    for (size_t o = 0; o < n; ++o) ++w.seg[f.view_intr[f.obs_view[o]] + 1];
    for (size_t q = 0; q < ni; ++q) w.seg[q + 1] += w.seg[q];
    w.order.resize(n);
    std::vector<size_t> cur(w.seg.begin(), w.seg.end() - 1);
    for (size_t o = 0; o < n; ++o) w.order[cur[f.view_intr[f.obs_view[o]]]++] = static_cast<uint32_t>(o);
    w.rays.resize(n);
  }
  size_t AngleRule(const Flat & f, AngleWork & w, const std::vector<uint8_t> & obs_alive, const std::vector<uint8_t> & pt_alive,
                   double min_angle_deg, std::vector<uint8_t> & kill) const
  {
    const int64_t n_reg = static_cast<int64_t>(f.lm.size());
    std::vector<Mat3> Rt(f.pose_ptr.size());
    for (size_t p = 0; p < f.pose_ptr.size(); ++p) Rt[p] = f.pose_ptr[p]->rotation().transpose();
    // ---- bearing rays, blocks of one intrinsic group
    constexpr size_t BLK = 4096;
    std::vector<std::pair<size_t, size_t>> blocks;              // [begin, end) into w.order, one intrinsic each
    for (size_t q = 0; q + 1 < w.seg.size(); ++q)
      for (size_t b = w.seg[q]; b < w.seg[q + 1]; b += BLK) blocks.emplace_back(b, std::min(b + BLK, w.seg[q + 1]));
    const int64_t nb = static_cast<int64_t>(blocks.size());
#ifdef OPENMVG_USE_OPENMP
    #pragma omp parallel for schedule(dynamic, 1) num_threads(HostThreads())
#endif
    for (int64_t bi = 0; bi < nb; ++bi)
    {
      const size_t b0 = blocks[bi].first, b1 = blocks[bi].second, m = b1 - b0;
      const cameras::IntrinsicBase * intr = f.intr_ptr[f.view_intr[f.obs_view[w.order[b0]]]];
      Mat2X pts(2, m);
      for (size_t i = 0; i < m; ++i) { const size_t o = w.order[b0 + i]; pts.col(i) = intr->get_ud_pixel(Vec2(f.obs_xy[2 * o], f.obs_xy[2 * o + 1])); }
      const Mat3X bearing = (*intr)(pts);
      for (size_t i = 0; i < m; ++i) { const size_t o = w.order[b0 + i]; w.rays[o] = (Rt[f.view_pose[f.obs_view[o]]] * bearing.col(i)).normalized(); }
    }
    // ---- per track: the largest pairwise angle of the live observations
    std::fill(kill.begin(), kill.end(), 0);
    size_t removed = 0;
#ifdef OPENMVG_USE_OPENMP
    #pragma omp parallel for schedule(dynamic, 512) reduction(+:removed) num_threads(HostThreads())
#endif
    for (int64_t j = 0; j < n_reg; ++j)
    {
      if (!pt_alive[j]) continue;
      const size_t lo = f.lm_first[j], hi = f.lm_first[j + 1];
      double max_angle = 0.0;
      for (size_t a = lo; a < hi; ++a)
      {
        if (!obs_alive[a]) continue;
        for (size_t b = a + 1; b < hi; ++b)
          if (obs_alive[b])
            max_angle = std::max(max_angle, R2D(acos(clamp(w.rays[a].dot(w.rays[b]), -1.0 + 1.e-8, 1.0 - 1.e-8))));
      }
      if (max_angle < min_angle_deg) { kill[j] = 1; ++removed; }
    }
    return removed;
  }

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

  // A failed solve after the registration: the reference returns false with the scene in the prior frame and centred
  // (it never undoes the centring on that path, sfm_data_BA_ceres.cpp:503-507 returns before :570-573).  Mirror the
  // successful path's last step instead so that a caller falling back to another solver sees a scene that is at
  // least un-centred, exactly as after a successful Adjust.
  static void UndoRegistration(SfM_Data & sfm_data, const Flat &, const geometry::Similarity3 & sim_to_center)
  {
    ApplySimilarity(sim_to_center.inverse(), sfm_data, true);
  }

  private:
  BA_B200_options options_;
  omvg_ba_summary summary_{};
  Timing timing_;
};

}  // namespace sfm
}  // namespace openMVG

#endif  // OPENMVG_B200_BUNDLE_ADJUSTMENT_B200_HPP
