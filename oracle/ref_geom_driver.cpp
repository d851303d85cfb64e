// ref_geom_driver.cpp — C entry points over the REFERENCE's own a-contrario RANSAC (test infrastructure only).
// Compiled by oracle/Makefile from /root/reference/src where it lies.  What GeometricFilter_FMatrix_AC::Robust_estimation
// (matching_image_collection/F_ACRobust.hpp:45-106) runs per image pair after MatchesPairToMat:
//   ACKernelAdaptor<SevenPointSolver, EpipolarDistanceError, UnnormalizerT, Mat3> kernel(xI, wI, hI, xJ, wJ, hJ, true);
//   ACRANSAC(kernel, inliers, iterations, &F, Square(precision));           (robust_estimator_ACRansac.hpp:303-490)
#include "openMVG/multiview/solver_fundamental_kernel.hpp"
#include "openMVG/multiview/solver_homography_kernel.hpp"
#include "openMVG/multiview/conditioning.hpp"
#include "openMVG/numeric/numeric.h"
#include "openMVG/robust_estimation/robust_estimator_ACRansac.hpp"
#include "openMVG/robust_estimation/robust_estimator_ACRansacKernelAdaptator.hpp"

#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

using namespace openMVG;
using namespace openMVG::robust;

extern "C" {

// xI, xJ: n points each, interleaved (x, y) doubles.  precision = upper bound in pixels (<= 0: infinity).
// Returns the number of inliers (> 2.5 * 7 for the pair to be kept by Robust_estimation), or 0.
// out: inliers[n] indices (ascending as ACRANSAC leaves them), F[9] row-major (un-normalised), stats = {errorMax (pixels), minNFA}
int ref_acransac_fundamental(const double * xI, const double * xJ, int n, int wI, int hI, int wJ, int hJ, double precision, unsigned int iterations,
                             uint32_t * inliers, double * F, double * stats)
{
  Mat2X a(2, n), b(2, n);
  for (int i = 0; i < n; ++i) { a(0, i) = xI[2 * i]; a(1, i) = xI[2 * i + 1]; b(0, i) = xJ[2 * i]; b(1, i) = xJ[2 * i + 1]; }
  using KernelType = ACKernelAdaptor<openMVG::fundamental::kernel::SevenPointSolver, openMVG::fundamental::kernel::EpipolarDistanceError, UnnormalizerT, Mat3>;
  const KernelType kernel(a, wI, hI, b, wJ, hJ, true);
  Mat3 Fm = Mat3::Identity();
  std::vector<uint32_t> vec_inliers;
  const double upper = precision > 0 ? Square(precision) : std::numeric_limits<double>::infinity();
  const std::pair<double, double> out = ACRANSAC(kernel, vec_inliers, iterations, &Fm, upper);
  for (size_t i = 0; i < vec_inliers.size(); ++i) inliers[i] = vec_inliers[i];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F[3 * r + c] = Fm(r, c);
  stats[0] = out.first; stats[1] = out.second;
  return (int)vec_inliers.size();
}

// GeometricFilter_HMatrix_AC::Robust_estimation (matching_image_collection/H_ACRobust.hpp:46-112):
//   ACKernelAdaptor<FourPointSolver, AsymmetricError, UnnormalizerI, Mat3> kernel(xI, wI, hI, xJ, wJ, hJ, false /* point to point */);
// Same contract; H[9] row-major un-normalised; the pair is kept iff more than 2.5 * 4 inliers.
int ref_acransac_homography(const double * xI, const double * xJ, int n, int wI, int hI, int wJ, int hJ, double precision, unsigned int iterations,
                            uint32_t * inliers, double * H, double * stats)
{
  Mat2X a(2, n), b(2, n);
  for (int i = 0; i < n; ++i) { a(0, i) = xI[2 * i]; a(1, i) = xI[2 * i + 1]; b(0, i) = xJ[2 * i]; b(1, i) = xJ[2 * i + 1]; }
  using KernelType = ACKernelAdaptor<openMVG::homography::kernel::FourPointSolver, openMVG::homography::kernel::AsymmetricError, UnnormalizerI, Mat3>;
  KernelType kernel(a, wI, hI, b, wJ, hJ, false);
  Mat3 Hm = Mat3::Identity();
  std::vector<uint32_t> vec_inliers;
  const double upper = precision > 0 ? Square(precision) : std::numeric_limits<double>::infinity();
  const std::pair<double, double> out = ACRANSAC(kernel, vec_inliers, iterations, &Hm, upper);
  for (size_t i = 0; i < vec_inliers.size(); ++i) inliers[i] = vec_inliers[i];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[3 * r + c] = Hm(r, c);
  stats[0] = out.first; stats[1] = out.second;
  return (int)vec_inliers.size();
}

}  // extern "C"
