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

#include <ceres/rotation.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
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
  for (int q = 0; q < n_intr; ++q) {
    auto cam = make_intrinsic(intr_model[q], intr + 8 * q);
    if (!cam) return -1;
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
  for (int j = 0; j < n_points; ++j)
    s.structure[j].X = Vec3(points[3 * j], points[3 * j + 1], points[3 * j + 2]);
  for (long o = 0; o < n_obs; ++o)
    s.structure[obs_point[o]].obs[obs_view[o]] = Observation(Vec2(obs_xy[2 * o], obs_xy[2 * o + 1]), IndexT(o));

  Bundle_Adjustment_Ceres::BA_Ceres_options bo(false, true);
  bo.bCeres_summary_ = true;
  if (opts[3] > 0) bo.nb_threads_ = opts[3];
  bo.bUse_loss_function_ = opts[4] != 0;
  Bundle_Adjustment_Ceres ba(bo);
  const Optimize_Options oo(static_cast<Intrinsic_Parameter_Type>(opts[0]),
                            static_cast<Extrinsic_Parameter_Type>(opts[1]),
                            opts[2] ? Structure_Parameter_Type::ADJUST_ALL : Structure_Parameter_Type::NONE);

  out[1] = huber_cost(s, bo.bUse_loss_function_);

  // OPENMVG_LOG_INFO writes to std::cerr (system/logger.hpp:63-64): capture fd 2 around Adjust.
  char tmpl[] = "/tmp/ref_ba_XXXXXX";
  const int fd = mkstemp(tmpl);
  fflush(stderr);
  const int saved = dup(2);
  if (fd >= 0) dup2(fd, 2);
  const auto t0 = std::chrono::steady_clock::now();
  const bool ok = ba.Adjust(s, oo);
  out[3] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
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

  out[0] = ok ? 1.0 : 0.0;
  out[2] = huber_cost(s, bo.bUse_loss_function_);

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
  return ok ? 0 : 1;
}

}  // extern "C"
