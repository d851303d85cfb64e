// patched_ba_test.cpp — TEST INFRASTRUCTURE ONLY.  Links the PATCHED reference translation unit
// (integration/openmvg_b200.patch applied to sfm/sfm_data_BA_ceres.cpp, compiled with -DOPENMVG_USE_B200) and calls
// Bundle_Adjustment_Ceres::Adjust exactly as the engines do (by name, poking linear_solver_type_):
//   * default                      -> the solve runs on the B200 through Bundle_Adjustment_B200 (no caller edit)
//   * OPENMVG_B200_DISABLE=1       -> the same call runs Ceres (the fallback the patch keeps)
// Prints the Huber cost of the returned scene; tests/test_integration_gpu.py compares the two runs.
#include "openMVG/cameras/cameras.hpp"
#include "openMVG/numeric/numeric.h"
#include "openMVG/sfm/sfm_data.hpp"
#include "openMVG/sfm/sfm_data_BA_ceres.hpp"

#include <ceres/types.h>
#include <cmath>
#include <cstdio>
#include <random>

using namespace openMVG;
using namespace openMVG::cameras;
using namespace openMVG::geometry;
using namespace openMVG::sfm;

int main()
{
  std::mt19937 g(7);
  std::uniform_real_distribution<double> U(-0.6, 0.6);
  std::normal_distribution<double> N(0, 1);
  SfM_Data s;
  const int C = 40, P = 3000, K = 6;
  s.intrinsics[0] = std::make_shared<Pinhole_Intrinsic_Radial_K3>(1000, 1000, 1000, 500, 500, 0.02, -0.005, 0.001);
  std::vector<Pose3> gt(C);
  for (int i = 0; i < C; ++i) {
    const double th = i * 2 * M_PI / C;
    const Vec3 c(1.5 * std::sin(th), 0.2 * std::sin(3 * th), 1.5 * std::cos(th));
    gt[i] = Pose3(LookAt(Vec3(-c)), c);
    s.views[i] = std::make_shared<View>("", i, 0, i, 1000, 1000);
    s.poses[i] = Pose3(gt[i].rotation(), c + Vec3(N(g), N(g), N(g)) * 0.005);
  }
  std::uniform_int_distribution<int> start(0, C - 1);
  for (int j = 0; j < P; ++j) {
    const Vec3 X(U(g), U(g), U(g));
    Landmark L;
    const int s0 = start(g);
    for (int k = 0; k < K; ++k) {
      const int i = (s0 + k) % C;
      L.obs[i] = Observation(s.intrinsics.at(0)->project(gt[i](X)) + Vec2(0.5 * N(g), 0.5 * N(g)), j);
    }
    L.X = X + Vec3(N(g), N(g), N(g)) * 0.01;
    s.structure[j] = L;
  }
  Bundle_Adjustment_Ceres::BA_Ceres_options options(false, true);
  options.linear_solver_type_ = ceres::DENSE_SCHUR;                  // as sequential_SfM.cpp:594 does
  Bundle_Adjustment_Ceres ba(options);
  const bool ok = ba.Adjust(s, Optimize_Options(Intrinsic_Parameter_Type::ADJUST_ALL, Extrinsic_Parameter_Type::ADJUST_ALL, Structure_Parameter_Type::ADJUST_ALL));
  long double c = 0;
  for (const auto & l : s.structure)
    for (const auto & o : l.second.obs) {
      const View * v = s.views.at(o.first).get();
      const double sq = s.intrinsics.at(v->id_intrinsic)->residual(s.poses.at(v->id_pose)(l.second.X), o.second.x).squaredNorm();
      c += 0.5 * (sq <= 256.0 ? sq : 32.0 * std::sqrt(sq) - 256.0);
    }
  std::printf("PATCHED_BA ok %d cost %.12f\n", int(ok), double(c));
  return ok ? 0 : 1;
}
