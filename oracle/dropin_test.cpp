// dropin_test.cpp — the drop-in boundary exercised through openMVG's OWN types.
//
// TEST INFRASTRUCTURE ONLY.  Links the reference's compiled code (oracle/_ref objects) and
// libomvg_b200.so, then runs, on identical inputs,
//   reference  Matcher_Regions(0.8f, BRUTE_FORCE_L2)::Match   vs   Matcher_Regions_B200::Match
//   reference  Bundle_Adjustment_Ceres::Adjust                vs   Bundle_Adjustment_B200::Adjust
// through the abstract interfaces Matcher (matching_image_collection/Matcher.hpp:34-48) and
// Bundle_Adjustment (sfm/sfm_data_BA.hpp:91-105).  MATCH must be identical; BA final Huber cost
// must agree to 1e-6 relative.  Needs a B200; run by tests/test_dropin_gpu.py on the GPU box.
#include "openMVG/cameras/cameras.hpp"
#include "openMVG/features/regions_factory.hpp"
#include "openMVG/matching_image_collection/Matcher_Regions.hpp"
#include "openMVG/numeric/numeric.h"
#include "openMVG/sfm/sfm_data.hpp"
#include "openMVG/sfm/sfm_data_BA_ceres.hpp"
#include "openMVG/sfm/sfm_data_filters.hpp"
#include "openMVG/sfm/sfm_view_priors.hpp"

#include "../openmvg_b200/host/Bundle_Adjustment_B200.hpp"
#include "../openmvg_b200/host/Cascade_Hashing_Matcher_Regions_B200.hpp"
#include "openMVG/matching_image_collection/Cascade_Hashing_Matcher_Regions.hpp"
#include "../openmvg_b200/host/Matcher_Regions_B200.hpp"
#include "../openmvg_b200/host/GeometricFilter_B200.hpp"
#include "openMVG/matching_image_collection/GeometricFilter.hpp"
#include "openMVG/matching_image_collection/F_ACRobust.hpp"
#include "openMVG/matching_image_collection/H_ACRobust.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <set>

using namespace openMVG;
using namespace openMVG::cameras;
using namespace openMVG::geometry;
using namespace openMVG::sfm;

namespace {

struct InMemory_Regions_Provider : public sfm::Regions_Provider
{
  void set(IndexT id, std::shared_ptr<features::Regions> r) { cache_[id] = std::move(r); }
  void set_type(features::Regions * r) { region_type_.reset(r); }
};

std::shared_ptr<features::SIFT_Regions> random_regions(int n, unsigned seed, const features::SIFT_Regions * base)
{
  std::mt19937 g(seed);
  std::uniform_int_distribution<int> u(0, 255), nz(-8, 8);
  std::uniform_real_distribution<double> p(0, 1);
  auto r = std::make_shared<features::SIFT_Regions>();
  r->Features().resize(n); r->Descriptors().resize(n);
  for (int i = 0; i < n; ++i) {
    r->Features()[i] = features::SIOPointFeature(float(i), float(i), 1.f, 0.f);
    if (base && i < int(base->RegionCount()) && p(g) < 0.3)
      for (int k = 0; k < 128; ++k) { const int v = int(base->Descriptors()[i][k]) + nz(g); r->Descriptors()[i][k] = (unsigned char)std::min(255, std::max(0, v)); }
    else
      for (int k = 0; k < 128; ++k) r->Descriptors()[i][k] = (unsigned char)(u(g) * u(g) / 255);
  }
  return r;
}

double huber_cost(const SfM_Data & s)
{
  long double c = 0;
  for (const auto & l : s.structure)
    for (const auto & o : l.second.obs) {
      const View * v = s.views.at(o.first).get();
      const Vec2 r = s.intrinsics.at(v->id_intrinsic)->residual(s.poses.at(v->id_pose)(l.second.X), o.second.x);
      const double sq = r.squaredNorm();
      c += 0.5 * (sq <= 256.0 ? sq : 32.0 * std::sqrt(sq) - 256.0);
    }
  return double(c);
}

// priors_and_gcp: views become ViewPriors (GPS = true centre * 3 + offset + noise; id_pose == id_view as in
// every openMVG loader) and 6 ground control points, surveyed in the GPS frame, are observed by 4 views each.
SfM_Data make_scene(int C, int P, int K, bool priors_and_gcp = false, double outlier_frac = 0.0, int stride = 1, int short_every = 0)
{
  std::mt19937 g(42);
  std::uniform_real_distribution<double> U(-0.6, 0.6);
  std::normal_distribution<double> N(0, 1);
  SfM_Data s;
  const double f = 1000, cx = 500, cy = 500;
  s.intrinsics[7] = std::make_shared<Pinhole_Intrinsic>(1000, 1000, f, cx, cy);     // non-dense ids on purpose
  std::vector<Pose3> gt(C);
  for (int i = 0; i < C; ++i) {
    const double th = i * 2 * M_PI / C;
    const Vec3 c(1.5 * std::sin(th), 0.2 * std::sin(3 * th), 1.5 * std::cos(th));
    const Mat3 R = LookAt(Vec3(-c));
    gt[i] = Pose3(R, c);
    const IndexT pose_id = priors_and_gcp ? 10 + i : 100 + i;
    if (priors_and_gcp) {
      auto vp = std::make_shared<ViewPriors>("", 10 + i, 7, pose_id, 1000, 1000);
      vp->SetPoseCenterPrior(3.0 * c + Vec3(50, -20, 10) + Vec3(N(g), N(g), N(g)) * 0.03, Vec3(1, 1, 1));
      s.views[10 + i] = vp;
    } else
    s.views[10 + i] = std::make_shared<View>("", 10 + i, 7, 100 + i, 1000, 1000);
    const Vec3 aa(N(g) * 0.005, N(g) * 0.005, N(g) * 0.005);
    const Mat3 dR = Eigen::AngleAxisd(aa.norm(), aa.normalized()).toRotationMatrix();
    s.poses[pose_id] = Pose3(dR * R, c + Vec3(N(g), N(g), N(g)) * 0.005);
  }
  std::uniform_int_distribution<int> start(0, C - 1);
  for (int j = 0; j < P; ++j) {
    const Vec3 X(U(g), U(g), U(g));
    Landmark L;
    const int s0 = start(g);
    const int Kj = (short_every > 0 && j % short_every == 0) ? 2 : K;   // two neighbouring views: little parallax
    for (int k = 0; k < Kj; ++k) {
      const int i = (s0 + k * stride) % C;
      const Vec3 Xc = gt[i](X);
      Vec2 x(cx + f * Xc(0) / Xc(2) + 0.5 * N(g), cy + f * Xc(1) / Xc(2) + 0.5 * N(g));
      if (outlier_frac > 0 && U(g) + 0.6 < 1.2 * outlier_frac) x += Vec2(60.0 * N(g), 60.0 * N(g));     // gross outliers
      L.obs[10 + i] = Observation(x, j);
    }
    L.X = X + Vec3(N(g), N(g), N(g)) * 0.01;
    s.structure[1000 + j] = L;
  }
  if (priors_and_gcp)
    for (int q = 0; q < 6; ++q) {
      const Vec3 X(U(g), U(g), U(g));
      Landmark L;
      for (int k = 0; k < 4; ++k) {
        const int i = (5 * q + 6 * k) % C;
        const Vec3 Xc = gt[i](X);
        L.obs[10 + i] = Observation(Vec2(cx + f * Xc(0) / Xc(2) + 0.3 * N(g), cy + f * Xc(1) / Xc(2) + 0.3 * N(g)), q);
      }
      L.X = 3.0 * X + Vec3(50, -20, 10);        // surveyed position, GPS frame
      s.control_points[q] = L;
    }
  return s;
}

}  // namespace

int main(int argc, char ** argv)
{
  std::setvbuf(stdout, nullptr, _IOLBF, 0);          // line-buffered even into a pipe: a crash must not swallow earlier results
  int failures = 0;
  // ------------------------------------------------------------------ MATCH
  {
    auto provider = std::make_shared<InMemory_Regions_Provider>();
    provider->set_type(new features::SIFT_Regions);
    const int counts[5] = {700, 650, 0, 300, 1};
    std::shared_ptr<features::SIFT_Regions> prev;
    for (int k = 0; k < 5; ++k) { auto r = random_regions(counts[k], 100 + k, prev.get()); provider->set(3 * k + 1, r); if (counts[k] > 1) prev = r; }
    Pair_Set pairs;
    for (int a = 0; a < 5; ++a) for (int b = a + 1; b < 5; ++b) pairs.insert({3 * a + 1, 3 * b + 1});
    matching::PairWiseMatches ref, ours;
    std::unique_ptr<matching_image_collection::Matcher> m_ref(new matching_image_collection::Matcher_Regions(0.8f, matching::BRUTE_FORCE_L2));
    std::unique_ptr<matching_image_collection::Matcher> m_b200(new matching_image_collection::Matcher_Regions_B200(0.8f));
    m_ref->Match(provider, pairs, ref, nullptr);
    m_b200->Match(provider, pairs, ours, nullptr);
    bool same = ref.size() == ours.size();
    size_t total = 0;
    for (const auto & kv : ref) {
      const auto it = ours.find(kv.first);
      if (it == ours.end() || it->second.size() != kv.second.size()) { same = false; break; }
      for (size_t i = 0; i < kv.second.size(); ++i)
        if (kv.second[i].i_ != it->second[i].i_ || kv.second[i].j_ != it->second[i].j_) { same = false; break; }
      total += kv.second.size();
    }
    std::printf("MATCH drop-in: %zu pairs with matches, %zu matches, %s\n", ref.size(), total, same ? "IDENTICAL" : "DIFFERENT");
    if (!same || total == 0) ++failures;
  }
  // ------------------------------------------------------------------ MATCH, cascade hashing (the CLI default)
  {
    auto provider = std::make_shared<InMemory_Regions_Provider>();
    provider->set_type(new features::SIFT_Regions);
    const int counts[5] = {1400, 1250, 0, 900, 3};
    std::shared_ptr<features::SIFT_Regions> prev;
    for (int k = 0; k < 5; ++k) { auto r = random_regions(counts[k], 300 + k, prev.get()); provider->set(3 * k + 1, r); if (counts[k] > 3) prev = r; }
    Pair_Set pairs;
    for (int a = 0; a < 5; ++a) for (int b = a + 1; b < 5; ++b) pairs.insert({3 * a + 1, 3 * b + 1});
    matching::PairWiseMatches ref, ours;
    std::unique_ptr<matching_image_collection::Matcher> m_ref(new matching_image_collection::Cascade_Hashing_Matcher_Regions(0.8f));
    std::unique_ptr<matching_image_collection::Matcher> m_b200(new matching_image_collection::Cascade_Hashing_Matcher_Regions_B200(0.8f));
    m_ref->Match(provider, pairs, ref, nullptr);
    m_b200->Match(provider, pairs, ours, nullptr);
    size_t total = 0, same = 0, uni = 0;
    for (const auto & kv : ref) {
      total += kv.second.size();
      const auto it = ours.find(kv.first);
      std::set<std::pair<IndexT, IndexT>> a, b;
      for (const auto & m : kv.second) a.insert({m.i_, m.j_});
      if (it != ours.end()) for (const auto & m : it->second) b.insert({m.i_, m.j_});
      for (const auto & x : a) if (b.count(x)) ++same;
      std::set<std::pair<IndexT, IndexT>> u(a); u.insert(b.begin(), b.end()); uni += u.size();
    }
    for (const auto & kv : ours) if (!ref.count(kv.first)) uni += kv.second.size();
    std::printf("MATCH cascade drop-in: %zu pairs with matches, %zu reference matches, %zu identical of %zu in the union (%s)\n",
                ref.size(), total, same, uni, same == uni ? "IDENTICAL" : "statistical");
    // the hashing is a float mat-vec whose summation order differs from Eigen's: identical up to rare sign flips
    if (total == 0 || same < 0.999 * uni) ++failures;
  }
  // ------------------------------------------------------------------ BA
  {
    SfM_Data a = make_scene(24, 1200, 6), b = a;
    // deep-copy the shared intrinsic so the two runs do not alias
    b.intrinsics[7] = std::shared_ptr<IntrinsicBase>(a.intrinsics.at(7)->clone());
    const double c0 = huber_cost(a);
    std::unique_ptr<Bundle_Adjustment> ba_ref(new Bundle_Adjustment_Ceres(Bundle_Adjustment_Ceres::BA_Ceres_options(false, true)));
    std::unique_ptr<Bundle_Adjustment> ba_b200(new Bundle_Adjustment_B200());
    const Optimize_Options opt(Intrinsic_Parameter_Type::ADJUST_ALL, Extrinsic_Parameter_Type::ADJUST_ALL, Structure_Parameter_Type::ADJUST_ALL);
    const bool ok_ref = ba_ref->Adjust(a, opt);
    const bool ok_b200 = ba_b200->Adjust(b, opt);
    const double ca = huber_cost(a), cb = huber_cost(b);
    const double rel = std::fabs(ca - cb) / ca;
    std::printf("BA drop-in: initial %.6f  reference %.9f  B200 %.9f  rel %.3e  ok %d/%d\n", c0, ca, cb, rel, int(ok_ref), int(ok_b200));
    if (!ok_ref || !ok_b200 || !(rel <= 1e-6) || !(ca < c0)) ++failures;
  }
  // ------------------------------------------------------------------ BA with control points + motion priors
  {
    SfM_Data a = make_scene(24, 1200, 6, /*priors_and_gcp=*/true), b = a;
    b.intrinsics[7] = std::shared_ptr<IntrinsicBase>(a.intrinsics.at(7)->clone());
    for (auto & v : b.views)        // deep-copy the views too: the registration step rewrites the priors in place
      v.second = std::make_shared<ViewPriors>(*dynamic_cast<ViewPriors *>(a.views.at(v.first).get()));
    const double c0 = huber_cost(a);
    std::unique_ptr<Bundle_Adjustment> ba_ref(new Bundle_Adjustment_Ceres(Bundle_Adjustment_Ceres::BA_Ceres_options(false, true)));
    std::unique_ptr<Bundle_Adjustment> ba_b200(new Bundle_Adjustment_B200());
    const Optimize_Options opt(Intrinsic_Parameter_Type::ADJUST_ALL, Extrinsic_Parameter_Type::ADJUST_ALL, Structure_Parameter_Type::ADJUST_ALL,
                               Control_Point_Parameter(20.0, true), true);
    const bool ok_ref = ba_ref->Adjust(a, opt);
    const bool ok_b200 = ba_b200->Adjust(b, opt);
    const double ca = huber_cost(a), cb = huber_cost(b);
    const double rel = std::fabs(ca - cb) / ca;
    double dc = 0, dg = 0;
    for (const auto & p : a.poses) dc = std::max(dc, (p.second.center() - b.poses.at(p.first).center()).norm());
    for (const auto & g : a.control_points) dg = std::max(dg, (g.second.X - b.control_points.at(g.first).X).norm());
    std::printf("BA+GCP+priors drop-in: initial %.6f  reference %.9f  B200 %.9f  rel %.3e  max centre diff %.3e  gcp diff %.3e  ok %d/%d\n",
                c0, ca, cb, rel, dc, dg, int(ok_ref), int(ok_b200));
    if (!ok_ref || !ok_b200 || !(rel <= 1e-6) || !(dc <= 1e-6) || !(dg <= 1e-12)) ++failures;
  }
  // ------------------------------------------------------------------ geometric filtering, F model (N4)
  // reference: ImageCollectionGeometricFilter::Robust_model_estimation(GeometricFilter_FMatrix_AC(4.0, 2048), putative)
  {
    std::mt19937 g(11);
    std::uniform_real_distribution<double> U(-1.0, 1.0), Upx(0.0, 1000.0), U01(0.0, 1.0);
    std::normal_distribution<double> N(0, 1);
    SfM_Data s;
    s.intrinsics[0] = std::make_shared<Pinhole_Intrinsic_Radial_K1>(1000, 1000, 1100, 500, 500, 0.03);      // positions get un-distorted
    const int V = 4, P = 900;
    std::vector<Pose3> poses;
    for (int v = 0; v < V; ++v) {
      const Vec3 c(0.6 * v, 0.05 * v, 0.1 * v);
      const Vec3 aa(0.01 * v, -0.08 * v, 0.005 * v);
      const Mat3 R = aa.norm() > 0 ? Eigen::AngleAxisd(aa.norm(), aa.normalized()).toRotationMatrix() : Mat3(Mat3::Identity());
      poses.emplace_back(R, c);
      s.views[v] = std::make_shared<View>("", v, v == 3 ? UndefinedIndexT : 0, v, 1000, 1000);   // view 3 has no intrinsic: raw positions
    }
    auto provider = std::make_shared<InMemory_Regions_Provider>();
    provider->set_type(new features::SIFT_Regions);
    std::vector<Vec3> X(P);
    for (auto & x : X) x = Vec3(1.5 * U(g), 1.2 * U(g), 6.5 + 2.5 * U(g));
    for (int v = 0; v < V; ++v) {
      auto r = std::make_shared<features::SIFT_Regions>();
      r->Features().resize(P); r->Descriptors().resize(P);
      for (int j = 0; j < P; ++j) {
        const Vec2 x = s.intrinsics.at(0)->project(poses[v](X[j])) + Vec2(0.4 * N(g), 0.4 * N(g));
        r->Features()[j] = features::SIOPointFeature(float(x(0)), float(x(1)), 1.f, 0.f);
      }
      provider->set(v, r);
    }
    matching::PairWiseMatches putative;
    for (int a = 0; a < V; ++a) for (int b = a + 1; b < V; ++b) {
      matching::IndMatches m;
      const int n = (a == 0 && b == 1) ? P : (a == 1 && b == 2) ? 12 : 300 + 100 * a;      // one pair with too few matches
      for (int j = 0; j < n; ++j) m.emplace_back(j, U01(g) < 0.35 ? int(U01(g) * (P - 1)) : j);   // 35 % wrong correspondences
      putative.insert({{a, b}, m});
    }
    // (both filters keep a REFERENCE to the shared_ptr<Regions_Provider>: hand them an lvalue of exactly that type)
    const std::shared_ptr<sfm::Regions_Provider> base_provider = provider;
    matching_image_collection::ImageCollectionGeometricFilter ref_filter(&s, base_provider);
    ref_filter.Robust_model_estimation(matching_image_collection::GeometricFilter_FMatrix_AC(4.0, 2048), putative);
    matching_image_collection::ImageCollectionGeometricFilter_B200 gpu_filter(&s, base_provider);
    const bool ok = gpu_filter.Robust_model_estimation_F(putative, 4.0, 2048);
    const auto & A = ref_filter.Get_geometric_matches(); const auto & B = gpu_filter.Get_geometric_matches();
    bool same = ok && A.size() == B.size();
    size_t total = 0;
    for (const auto & kv : A) {
      const auto it = B.find(kv.first);
      if (it == B.end() || it->second.size() != kv.second.size()) { same = false; break; }
      for (size_t i = 0; i < kv.second.size(); ++i) if (kv.second[i].i_ != it->second[i].i_ || kv.second[i].j_ != it->second[i].j_) { same = false; break; }
      total += kv.second.size();
    }
    std::printf("GEOMETRIC FILTER (F, AC-RANSAC) drop-in: %zu of %zu pairs kept, %zu geometric matches, %s\n", A.size(), putative.size(), total, same ? "IDENTICAL" : "DIFFERENT");
    if (!same || total == 0 || A.size() == putative.size()) ++failures;
    // the homography model on the same putative matches (a general scene: few pairs survive, which is the point of the test)
    matching_image_collection::ImageCollectionGeometricFilter ref_h(&s, base_provider);
    ref_h.Robust_model_estimation(matching_image_collection::GeometricFilter_HMatrix_AC(4.0, 2048), putative);
    matching_image_collection::ImageCollectionGeometricFilter_B200 gpu_h(&s, base_provider);
    const bool okh = gpu_h.Robust_model_estimation_H(putative, 4.0, 2048);
    const auto & AH = ref_h.Get_geometric_matches(); const auto & BH = gpu_h.Get_geometric_matches();
    bool sameh = okh && AH.size() == BH.size(); size_t totalh = 0;
    for (const auto & kv : AH) {
      const auto it = BH.find(kv.first);
      if (it == BH.end() || it->second.size() != kv.second.size()) { sameh = false; break; }
      for (size_t i = 0; i < kv.second.size(); ++i) if (kv.second[i].i_ != it->second[i].i_ || kv.second[i].j_ != it->second[i].j_) { sameh = false; break; }
      totalh += kv.second.size();
    }
    std::printf("GEOMETRIC FILTER (H, AC-RANSAC) drop-in: %zu of %zu pairs kept, %zu geometric matches, %s\n", AH.size(), putative.size(), totalh, sameh ? "IDENTICAL" : "DIFFERENT");
    if (!sameh) ++failures;
  }
  // ------------------------------------------------------------------ the BA / outlier-rejection loop (N1)
  // reference: do { Bundle_Adjustment_Ceres::Adjust } while (RemoveOutliers_PixelResidualError(4.0, 2) +
  // RemoveOutliers_AngleError(2.0) > 50)  (sequential_SfM.cpp:205-211, 1226-1243) vs AdjustAndReject on ONE resident problem
  {
    SfM_Data a = make_scene(200, 6000, 6, false, 0.04, 1, 10);      // every 10th track: two neighbouring views, ~1.8 degrees of parallax
    SfM_Data b = a;
    b.intrinsics[7] = std::shared_ptr<IntrinsicBase>(a.intrinsics.at(7)->clone());
    const Optimize_Options opt(Intrinsic_Parameter_Type::ADJUST_ALL, Extrinsic_Parameter_Type::ADJUST_ALL, Structure_Parameter_Type::ADJUST_ALL);
    size_t ref_res = 0, ref_ang = 0; int ref_rounds = 0; bool again = true;
    while (again) {
      Bundle_Adjustment_Ceres ba_ref(Bundle_Adjustment_Ceres::BA_Ceres_options(false, true));
      ba_ref.Adjust(a, opt); ++ref_rounds;
      const size_t r1 = RemoveOutliers_PixelResidualError(a, 4.0, 2), r2 = RemoveOutliers_AngleError(a, 2.0);
      ref_res += r1; ref_ang += r2; again = r1 + r2 > 50;
    }
    Bundle_Adjustment_B200 ba_b200; Bundle_Adjustment_B200::RejectStats st;
    const bool ok = ba_b200.AdjustAndReject(b, opt, 4.0, 50, 2, 2.0, &st);
    bool same = ok && a.structure.size() == b.structure.size();
    size_t n_obs_a = 0, n_obs_b = 0;
    for (const auto & l : a.structure) n_obs_a += l.second.obs.size();
    for (const auto & l : b.structure) n_obs_b += l.second.obs.size();
    if (same) for (const auto & l : a.structure) {
      const auto it = b.structure.find(l.first);
      if (it == b.structure.end() || it->second.obs.size() != l.second.obs.size()) { same = false; break; }
      for (const auto & o : l.second.obs) if (!it->second.obs.count(o.first)) { same = false; break; }
      if (!same) break;
    }
    const double ca = huber_cost(a), cb = huber_cost(b), rel = std::fabs(ca - cb) / ca;
    std::printf("BA reject loop: reference %d rounds, %zu residual outliers, %zu angle tracks -> %zu tracks / %zu obs, cost %.9f | B200 %d rounds, %zu + %zu short + %zu angle -> %zu tracks / %zu obs, cost %.9f  rel %.3e  %s\n",
                ref_rounds, ref_res, ref_ang, a.structure.size(), n_obs_a, ca, st.rounds, st.residual_outliers, st.short_tracks, st.angle_tracks, b.structure.size(), n_obs_b, cb, rel,
                same ? "SAME STRUCTURE" : "DIFFERENT STRUCTURE");
    if (!same || !(rel <= 1e-6) || ref_res != st.residual_outliers || ref_ang != st.angle_tracks || ref_res == 0 || ref_ang == 0) ++failures;
  }
  // ------------------------------------------------------------------ Adjust() wall at config 2 through SfM_Data (--config2 [ref])
  if (argc > 1 && std::strcmp(argv[1], "--config2") == 0)
  {
    const bool with_ref = argc > 2 && std::strcmp(argv[2], "ref") == 0;
    SfM_Data a = make_scene(1000, 100000, 10, false, 0.0, 1);
    const Optimize_Options opt(Intrinsic_Parameter_Type::ADJUST_ALL, Extrinsic_Parameter_Type::ADJUST_ALL, Structure_Parameter_Type::ADJUST_ALL);
    double best = 1e30; Bundle_Adjustment_B200::Timing bt; double cost_b = 0;
    for (int rep = 0; rep < 4; ++rep) {
      SfM_Data b = a; b.intrinsics[7] = std::shared_ptr<IntrinsicBase>(a.intrinsics.at(7)->clone());
      Bundle_Adjustment_B200 ba;
      const auto t0 = std::chrono::steady_clock::now();
      const bool ok = ba.Adjust(b, opt);
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (!ok) { std::printf("config2 Adjust (B200) failed after pack %.2f ms\n", ba.timing().pack_ms); ++failures; break; }
      if (rep > 0 && ms < best) { best = ms; bt = ba.timing(); }
      cost_b = huber_cost(b);
      std::printf("config2 Adjust (B200) rep %d: %.2f ms  [pack %.2f, omvg_ba_solve %.2f, unpack %.2f]  device %.2f ms, %d iterations\n", rep, ms, ba.timing().pack_ms, ba.timing().solve_ms,
                  ba.timing().unpack_ms, ba.summary().device_ms, ba.summary().iterations);
    }
    std::printf("CONFIG2 B200 Adjust wall best %.2f ms (pack %.2f + solve %.2f + unpack %.2f), final cost %.9f\n", best, bt.pack_ms, bt.solve_ms, bt.unpack_ms, cost_b);
    {  // the resident loop on the same scene with 2 %% gross outliers
      SfM_Data b = make_scene(1000, 100000, 10, false, 0.02, 1);
      Bundle_Adjustment_B200 ba; Bundle_Adjustment_B200::RejectStats st;
      const auto t0 = std::chrono::steady_clock::now();
      const bool ok = ba.AdjustAndReject(b, opt, 4.0, 50, 2, 2.0, &st);
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      const auto & t = ba.timing();
      std::printf("CONFIG2 B200 AdjustAndReject wall %.2f ms, %d rounds (pack %.2f, solves %.2f, device rejection %.2f, host angle rule %.2f, unpack %.2f, erase %.2f, destroy %.2f): %zu residual outliers, %zu short, %zu angle tracks, ok %d\n",
                  ms, st.rounds, t.pack_ms, t.solve_ms, t.reject_ms, t.angle_ms, t.unpack_ms, t.erase_ms, t.destroy_ms, st.residual_outliers, st.short_tracks, st.angle_tracks, int(ok));
      if (!ok) ++failures;
    }
    if (with_ref) {
      SfM_Data r = a; r.intrinsics[7] = std::shared_ptr<IntrinsicBase>(a.intrinsics.at(7)->clone());
      Bundle_Adjustment_Ceres::BA_Ceres_options ro(false, true);
      ro.nb_threads_ = 8;                                        // the reference's best thread count on these hosts (bench.py sweeps it)
      Bundle_Adjustment_Ceres ba_ref(ro);
      const auto t0 = std::chrono::steady_clock::now();
      const bool ok = ba_ref.Adjust(r, opt);
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      const double cr = huber_cost(r);
      std::printf("CONFIG2 reference Adjust (8 threads) wall %.1f ms, final cost %.9f, rel diff %.3e, speed-up %.1fx, ok %d\n", ms, cr, std::fabs(cr - cost_b) / cr, ms / best, int(ok));
      if (!ok || !(std::fabs(cr - cost_b) <= 1e-6 * cr)) ++failures;
    }
  }
  std::printf(failures ? "DROPIN FAILED\n" : "DROPIN OK\n");
  return failures;
}
