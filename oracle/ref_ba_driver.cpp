// ref_ba_driver.cpp — thin C entry point over the REFERENCE's own bundle adjustment.
//
// TEST INFRASTRUCTURE ONLY.  No algorithm is restated here: the flat arrays are turned into an
// openMVG::sfm::SfM_Data (the way sfm/sfm_data_BA_test.cpp:336-465 fills one), the unmodified
// openMVG::sfm::Bundle_Adjustment_Ceres::Adjust (sfm/sfm_data_BA_ceres.cpp:165-608) runs on the
// vendored Ceres 1.13.0 + Eigen 3.4.0, and the refined scene is flattened back.
// Built by oracle/Makefile into oracle/_ref/libref_ba.so (git-ignored; shipped by gpurun).
//
// Flat layout (same as include/omvg_b200.h):
//   poses[n_poses][6]     angle-axis (3) + t (3), t = -R*C   (sfm_data_BA_ceres.cpp:265-271)
//   intr[n_intr][8]       getParams() order, zero padded     (cameras/Camera_Pinhole*.hpp)
//   intr_model[n_intr]    cameras::EINTRINSIC value          (cameras/Camera_Common.hpp:39-49)
//   points[n_points][3]
//   view_pose/view_intr[n_views], obs_view/obs_point[n_obs], obs_xy[n_obs][2]
#include "openMVG/cameras/cameras.hpp"
#include "openMVG/sfm/sfm_data.hpp"
#include "openMVG/sfm/sfm_data_BA.hpp"
#include "openMVG/sfm/sfm_data_BA_ceres.hpp"
#include "openMVG/sfm/sfm_data_transform.hpp"
#include "openMVG/sfm/sfm_view_priors.hpp"
#include "openMVG/geometry/Similarity3.hpp"
#include "openMVG/geometry/Similarity3_Kernel.hpp"
#include "openMVG/robust_estimation/robust_estimator_LMeds.hpp"

#include <ceres/rotation.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <limits>
#include <vector>
#include <memory>
#include <string>
#include <unistd.h>

using namespace openMVG;
using namespace openMVG::cameras;
using namespace openMVG::geometry;
using namespace openMVG::sfm;

namespace {

std::shared_ptr<IntrinsicBase> make_intrinsic(int model, const double * p)
{
  const int w = 1000, h = 1000;
  switch (model) {
    case PINHOLE_CAMERA:         return std::make_shared<Pinhole_Intrinsic>(w, h, p[0], p[1], p[2]);
    case PINHOLE_CAMERA_RADIAL1: return std::make_shared<Pinhole_Intrinsic_Radial_K1>(w, h, p[0], p[1], p[2], p[3]);
    case PINHOLE_CAMERA_RADIAL3: return std::make_shared<Pinhole_Intrinsic_Radial_K3>(w, h, p[0], p[1], p[2], p[3], p[4], p[5]);
    case PINHOLE_CAMERA_BROWN:   return std::make_shared<Pinhole_Intrinsic_Brown_T2>(w, h, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
    case PINHOLE_CAMERA_FISHEYE: return std::make_shared<Pinhole_Intrinsic_Fisheye>(w, h, p[0], p[1], p[2], p[3], p[4], p[5], p[6]);
    case CAMERA_SPHERICAL:       return std::make_shared<Intrinsic_Spherical>((unsigned int)p[0], (unsigned int)p[1]);   // flat layout: {w, h}
    default: return {};
  }
}

// 1/2 * sum rho(|r|^2) with the library's own residual() (cameras/Camera_Intrinsics.hpp:105-112)
// and HuberLoss(a=16) as constructed at sfm_data_BA_ceres.cpp:249 (loss_function.cc:47-61).
double huber_cost(const SfM_Data & s, bool use_loss)
{
  long double c = 0;
  for (const auto & l : s.structure)
    for (const auto & o : l.second.obs) {
      const View * v = s.views.at(o.first).get();
      const Vec2 r = s.intrinsics.at(v->id_intrinsic)->residual(s.poses.at(v->id_pose)(l.second.X), o.second.x);
      const double sq = r.squaredNorm();
      c += 0.5 * ((!use_loss || sq <= 256.0) ? sq : 32.0 * std::sqrt(sq) - 256.0);
    }
  return double(c);
}

// Flat arrays of the extended entry points:
//   GCP (SfM_Data::control_points, sfm_data.hpp:44): gcp_X[n_gcp][3], observations gcp_obs_view /
//   gcp_obs_gcp / gcp_obs_xy[n_gcp_obs]; one weight for all (Control_Point_Parameter::weight).
//   Priors (sfm_view_priors.hpp:27-80): prior_view[n_priors] = the view turned into a ViewPriors
//   with pose_center_ = prior_center[k], center_weight_ = prior_weight[k].
struct Ext {
  int n_gcp; double * gcp_X; long n_gcp_obs; const int * gcp_obs_view, * gcp_obs_gcp; const double * gcp_obs_xy; double gcp_weight;
  int n_priors; const int * prior_view; double * prior_center; const double * prior_weight;
};

bool build_scene(SfM_Data & s, int n_poses, const double * poses, int n_intr, const double * intr, const int * intr_model,
                 int n_points, const double * points, int n_views, const int * view_pose, const int * view_intr,
                 long n_obs, const int * obs_view, const int * obs_point, const double * obs_xy, const Ext * e)
{
  for (int q = 0; q < n_intr; ++q) {
    auto cam = make_intrinsic(intr_model[q], intr + 8 * q);
    if (!cam) return false;
    s.intrinsics[q] = cam;
  }
  for (int p = 0; p < n_poses; ++p) {
    Mat3 R;
    ceres::AngleAxisToRotationMatrix(poses + 6 * p, R.data());   // column-major, as Adjust reads it back
    const Vec3 t(poses[6 * p + 3], poses[6 * p + 4], poses[6 * p + 5]);
    s.poses[p] = Pose3(R, -R.transpose() * t);
  }
  for (int v = 0; v < n_views; ++v)
    s.views[v] = std::make_shared<View>("", v, view_intr[v], view_pose[v], 1000, 1000);
  if (e) for (int k = 0; k < e->n_priors; ++k) {
    const int v = e->prior_view[k];
    auto vp = std::make_shared<ViewPriors>("", v, view_intr[v], view_pose[v], 1000, 1000);
    vp->SetPoseCenterPrior(Vec3(e->prior_center[3 * k], e->prior_center[3 * k + 1], e->prior_center[3 * k + 2]),
                           Vec3(e->prior_weight[3 * k], e->prior_weight[3 * k + 1], e->prior_weight[3 * k + 2]));
    s.views[v] = vp;
  }
  for (int j = 0; j < n_points; ++j)
    s.structure[j].X = Vec3(points[3 * j], points[3 * j + 1], points[3 * j + 2]);
  for (long o = 0; o < n_obs; ++o)
    s.structure[obs_point[o]].obs[obs_view[o]] = Observation(Vec2(obs_xy[2 * o], obs_xy[2 * o + 1]), IndexT(o));
  if (e) {
    for (int g = 0; g < e->n_gcp; ++g) s.control_points[g].X = Vec3(e->gcp_X[3 * g], e->gcp_X[3 * g + 1], e->gcp_X[3 * g + 2]);
    for (long o = 0; o < e->n_gcp_obs; ++o)
      s.control_points[e->gcp_obs_gcp[o]].obs[e->gcp_obs_view[o]] = Observation(Vec2(e->gcp_obs_xy[2 * o], e->gcp_obs_xy[2 * o + 1]), IndexT(o));
  }
  return true;
}

void flatten_scene(const SfM_Data & s, int n_poses, double * poses, int n_intr, double * intr, int n_points, double * points, const Ext * e)
{
  for (int p = 0; p < n_poses; ++p) {
    const Pose3 & P = s.poses.at(p);
    const Mat3 R = P.rotation();
    const Vec3 t = P.translation();
    ceres::RotationMatrixToAngleAxis((const double *)R.data(), poses + 6 * p);
    poses[6 * p + 3] = t(0); poses[6 * p + 4] = t(1); poses[6 * p + 5] = t(2);
  }
  for (int q = 0; q < n_intr; ++q) {
    const std::vector<double> v = s.intrinsics.at(q)->getParams();
    for (size_t k = 0; k < v.size() && k < 8; ++k) intr[8 * q + k] = v[k];
  }
  for (int j = 0; j < n_points; ++j) {
    const Vec3 & X = s.structure.at(j).X;
    points[3 * j] = X(0); points[3 * j + 1] = X(1); points[3 * j + 2] = X(2);
  }
  if (e) {
    for (int g = 0; g < e->n_gcp; ++g) { const Vec3 & X = s.control_points.at(g).X; e->gcp_X[3 * g] = X(0); e->gcp_X[3 * g + 1] = X(1); e->gcp_X[3 * g + 2] = X(2); }
    for (int k = 0; k < e->n_priors; ++k) {
      const ViewPriors * vp = dynamic_cast<const ViewPriors *>(s.views.at(e->prior_view[k]).get());
      for (int a = 0; a < 3; ++a) e->prior_center[3 * k + a] = vp->pose_center_(a);
    }
  }
}

// cost of the GCP residual blocks (weighted, no loss: sfm_data_BA_ceres.cpp:418-435) and of the
// pose-centre priors (HuberLoss(a = fit^2), :455-472), with the library's own residual()/center().
double ext_cost(const SfM_Data & s, const Ext & e, double prior_fit)
{
  long double c = 0;
  for (const auto & l : s.control_points)
    for (const auto & o : l.second.obs) {
      const View * v = s.views.at(o.first).get();
      const Vec2 r = e.gcp_weight * s.intrinsics.at(v->id_intrinsic)->residual(s.poses.at(v->id_pose)(l.second.X), o.second.x);
      c += 0.5 * r.squaredNorm();
    }
  if (prior_fit >= 0) {
    const double a = prior_fit * prior_fit, b = a * a;
    for (const auto & v : s.views) {
      const ViewPriors * vp = dynamic_cast<const ViewPriors *>(v.second.get());
      if (!vp || !vp->b_use_pose_center_) continue;
      const Vec3 r = vp->center_weight_.cwiseProduct(s.poses.at(vp->id_pose).center() - vp->pose_center_);
      const double sq = r.squaredNorm();
      c += 0.5 * (sq <= b ? sq : 2.0 * a * std::sqrt(sq) - b);
    }
  }
  return double(c);
}

// Runs Adjust with stderr captured into report.
bool run_adjust(SfM_Data & s, const int * opts, const Ext * e, double * seconds, char * report, int report_cap)
{
  Bundle_Adjustment_Ceres::BA_Ceres_options bo(false, true);
  bo.bCeres_summary_ = true;
  if (opts[3] > 0) bo.nb_threads_ = opts[3];
  bo.bUse_loss_function_ = opts[4] != 0;
  Bundle_Adjustment_Ceres ba(bo);
  const Optimize_Options oo(static_cast<Intrinsic_Parameter_Type>(opts[0]),
                            static_cast<Extrinsic_Parameter_Type>(opts[1]),
                            opts[2] ? Structure_Parameter_Type::ADJUST_ALL : Structure_Parameter_Type::NONE,
                            (e && e->n_gcp > 0) ? Control_Point_Parameter(e->gcp_weight, true) : Control_Point_Parameter(),
                            e && e->n_priors > 0);
  // OPENMVG_LOG_INFO writes to std::cerr (system/logger.hpp:63-64): capture fd 2 around Adjust.
  char tmpl[] = "/tmp/ref_ba_XXXXXX";
  const int fd = mkstemp(tmpl);
  fflush(stderr);
  const int saved = dup(2);
  if (fd >= 0) dup2(fd, 2);
  const auto t0 = std::chrono::steady_clock::now();
  const bool ok = ba.Adjust(s, oo);
  *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  fflush(stderr);
  dup2(saved, 2);
  close(saved);
  if (report && report_cap > 0) {
    report[0] = 0;
    if (fd >= 0) {
      const off_t len = lseek(fd, 0, SEEK_END);
      lseek(fd, 0, SEEK_SET);
      const ssize_t n = read(fd, report, size_t(std::min<off_t>(len, report_cap - 1)));
      report[n > 0 ? n : 0] = 0;
    }
  }
  if (fd >= 0) { close(fd); unlink(tmpl); }
  return ok;
}

// The registration Adjust performs before building the Ceres problem when priors are used
// (sfm_data_BA_ceres.cpp:183-236), executed with the reference's own LeastMedianOfSquares /
// Similarity3_Kernel / ApplySimilarity so the flat scene handed to the LM core under test is the
// one the reference's LM sees.  Returns the median fitting error, or -1 if the prior is unusable.
double register_to_priors(SfM_Data & sfm_data, double centroid_out[3])
{
  std::vector<Vec3> X_SfM, X_GPS;
  for (const auto & view_it : sfm_data.GetViews()) {
    const ViewPriors * prior = dynamic_cast<ViewPriors *>(view_it.second.get());
    if (prior != nullptr && prior->b_use_pose_center_ && sfm_data.IsPoseAndIntrinsicDefined(prior)) {
      X_SfM.push_back(sfm_data.GetPoses().at(prior->id_pose).center());
      X_GPS.push_back(prior->pose_center_);
    }
  }
  if (!(sfm_data.GetViews().size() > 3) || !(X_GPS.size() > 3)) return -1.0;
  openMVG::geometry::Similarity3 sim;
  const Mat X_SfM_Mat = Eigen::Map<Mat>(X_SfM[0].data(), 3, X_SfM.size());
  const Mat X_GPS_Mat = Eigen::Map<Mat>(X_GPS[0].data(), 3, X_GPS.size());
  geometry::kernel::Similarity3_Kernel kernel(X_SfM_Mat, X_GPS_Mat);
  const double lmeds_median = openMVG::robust::LeastMedianOfSquares(kernel, &sim);
  if (lmeds_median == std::numeric_limits<double>::max()) return -1.0;
  for (Vec3 & pos : X_SfM) pos = sim(pos);
  Vec residual = (Eigen::Map<Mat3X>(X_SfM[0].data(), 3, X_SfM.size()) - Eigen::Map<Mat3X>(X_GPS[0].data(), 3, X_GPS.size())).colwise().norm();
  std::sort(residual.data(), residual.data() + residual.size());
  const double fit = residual(residual.size() / 2);
  ApplySimilarity(sim, sfm_data);
  Vec3 pose_centroid = Vec3::Zero();
  for (const auto & pose_it : sfm_data.poses) pose_centroid += (pose_it.second.center() / (double)sfm_data.poses.size());
  const openMVG::geometry::Similarity3 sim_to_center(Pose3(Mat3::Identity(), pose_centroid), 1.0);
  ApplySimilarity(sim_to_center, sfm_data, true);
  for (int a = 0; a < 3; ++a) centroid_out[a] = pose_centroid(a);
  return fit;
}

}  // namespace

extern "C" {

// opts[0]=intrinsics_opt (Intrinsic_Parameter_Type), [1]=extrinsics_opt, [2]=structure_opt(0/1),
// [3]=nb_threads (0 = library default), [4]=use_loss (0/1)
// out[0]=ok, [1]=initial cost, [2]=final cost, [3]=Adjust() wall seconds
// report: receives everything Adjust logged (Ceres FullReport) — NUL-terminated, truncated to cap.
int ref_ba_adjust(int n_poses, double * poses, int n_intr, double * intr, const int * intr_model,
                  int n_points, double * points, int n_views, const int * view_pose,
                  const int * view_intr, long n_obs, const int * obs_view, const int * obs_point,
                  const double * obs_xy, const int * opts, double * out, char * report, int report_cap)
{
  SfM_Data s;
  if (!build_scene(s, n_poses, poses, n_intr, intr, intr_model, n_points, points, n_views, view_pose, view_intr, n_obs, obs_view, obs_point, obs_xy, nullptr)) return -1;
  out[1] = huber_cost(s, opts[4] != 0);
  const bool ok = run_adjust(s, opts, nullptr, &out[3], report, report_cap);
  out[0] = ok ? 1.0 : 0.0;
  out[2] = huber_cost(s, opts[4] != 0);
  flatten_scene(s, n_poses, poses, n_intr, intr, n_points, points, nullptr);
  return ok ? 0 : 1;
}

// Extended entry point: ground control points and pose-centre priors (see Ext above).
// out[0]=ok [1]=initial cost incl. GCP terms (prior terms excluded: the scene is not yet registered)
// [2]=final cost incl. GCP and prior terms [3]=seconds [4]=median fitting error used for the prior loss (-1: none)
int ref_ba_adjust_ex(int n_poses, double * poses, int n_intr, double * intr, const int * intr_model,
                     int n_points, double * points, int n_views, const int * view_pose,
                     const int * view_intr, long n_obs, const int * obs_view, const int * obs_point,
                     const double * obs_xy,
                     int n_gcp, double * gcp_X, long n_gcp_obs, const int * gcp_obs_view, const int * gcp_obs_gcp, const double * gcp_obs_xy, double gcp_weight,
                     int n_priors, const int * prior_view, double * prior_center, const double * prior_weight,
                     const int * opts, double * out, char * report, int report_cap)
{
  Ext e{n_gcp, gcp_X, n_gcp_obs, gcp_obs_view, gcp_obs_gcp, gcp_obs_xy, gcp_weight, n_priors, prior_view, prior_center, prior_weight};
  SfM_Data s;
  if (!build_scene(s, n_poses, poses, n_intr, intr, intr_model, n_points, points, n_views, view_pose, view_intr, n_obs, obs_view, obs_point, obs_xy, &e)) return -1;
  double fit = -1.0;
  if (n_priors > 0) {        // the fitting error Adjust will compute, from a scratch copy of the scene
    SfM_Data t;
    build_scene(t, n_poses, poses, n_intr, intr, intr_model, n_points, points, n_views, view_pose, view_intr, n_obs, obs_view, obs_point, obs_xy, &e);
    double c[3]; fit = register_to_priors(t, c);
  }
  out[4] = fit;
  out[1] = huber_cost(s, opts[4] != 0) + ext_cost(s, e, -1.0);
  const bool ok = run_adjust(s, opts, &e, &out[3], report, report_cap);
  out[0] = ok ? 1.0 : 0.0;
  out[2] = huber_cost(s, opts[4] != 0) + ext_cost(s, e, fit);
  flatten_scene(s, n_poses, poses, n_intr, intr, n_points, points, &e);
  return ok ? 0 : 1;
}

// The pre-solve registration alone (see register_to_priors): transforms poses / points / GCPs /
// prior centres in place.  out[0]=median fitting error (-1: prior unusable, nothing changed),
// out[1..3]=the centroid that was subtracted (add it back to undo the centring after the solve).
int ref_ba_register_priors(int n_poses, double * poses, int n_intr, double * intr, const int * intr_model,
                           int n_points, double * points, int n_views, const int * view_pose,
                           const int * view_intr, long n_obs, const int * obs_view, const int * obs_point,
                           const double * obs_xy,
                           int n_gcp, double * gcp_X, long n_gcp_obs, const int * gcp_obs_view, const int * gcp_obs_gcp, const double * gcp_obs_xy,
                           int n_priors, const int * prior_view, double * prior_center, const double * prior_weight, double * out)
{
  Ext e{n_gcp, gcp_X, n_gcp_obs, gcp_obs_view, gcp_obs_gcp, gcp_obs_xy, 1.0, n_priors, prior_view, prior_center, prior_weight};
  SfM_Data s;
  if (!build_scene(s, n_poses, poses, n_intr, intr, intr_model, n_points, points, n_views, view_pose, view_intr, n_obs, obs_view, obs_point, obs_xy, &e)) return -1;
  double c[3] = {0, 0, 0};
  out[0] = register_to_priors(s, c);
  out[1] = c[0]; out[2] = c[1]; out[3] = c[2];
  if (out[0] >= 0) flatten_scene(s, n_poses, poses, n_intr, intr, n_points, points, &e);
  return 0;
}

}  // extern "C"
