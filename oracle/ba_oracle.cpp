// ba_oracle.cpp — CPU restatement of the bundle-adjustment hot path that
// openMVG::sfm::Bundle_Adjustment_Ceres::Adjust delegates to Ceres 1.13.0.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference leg may load this.  The product path (openmvg_b200/csrc) never calls it.
//
// Parity status: PINNED.  tests/test_oracle_ba.py checks this file against
//   (a) oracle/_ref/libref_ba.so — the unmodified reference Adjust + vendored Ceres/Eigen compiled
//       where they lie — on seeded scenes (final cost, iteration count, accepted/rejected steps);
//   (b) the committed golden vectors tests/golden/ba_ref_*.json generated from (a) by
//       tests/golden/make_golden.py.
// The reference's own BA tests hold no golden final cost (sfm/sfm_data_BA_test.cpp asserts only
// that the RMSE decreases), so (a)/(b) are the pin.
//
// Reference lines restated (ceres = /root/reference/src/third_party/ceres-solver):
//   openMVG/sfm/sfm_data_BA_ceres_camera_functor.hpp:103-194,207-300,313-412,425-533,548-660
//                                             residual functors (pinhole, K1, K3, Brown T2, fisheye), :662-760 spherical
//   ceres/include/ceres/rotation.h:563-622    AngleAxisRotatePoint incl. the theta^2<=eps branch
//   ceres/include/ceres/jet.h, internal/autodiff.h:207-319   forward-mode AD (Jet below)
//   ceres/internal/ceres/loss_function.cc:47-61   HuberLoss;  corrector.cc:41-155  Corrector
//   ceres/internal/ceres/residual_block.cc:68-196 cost = rho/2, correction only when J wanted
//   ceres/internal/ceres/local_parameterization.cc:91-154 SubsetParameterization (column drop)
//   ceres/internal/ceres/trust_region_minimizer.cc:66-119,226-279,355-424,667-786 LM outer loop
//   ceres/internal/ceres/levenberg_marquardt_strategy.cc:65-160   damping policy
//   ceres/internal/ceres/trust_region_step_evaluator.cc:51-59     step quality
//   ceres/internal/ceres/schur_eliminator_impl.h:176-366  Eliminate / BackSubstitute (the algebra;
//       any exact solve of (J'J + D'D) y = J'r yields the same step, so the elimination order
//       chosen by reorder_program.cc is irrelevant to the result)
//   ceres/internal/ceres/schur_complement_solver.cc:181-224,310-347  exact solve of the reduced system
//   openMVG/sfm/sfm_data_BA_ceres.cpp:242-253,275-305,321-344,394-395,477-493  options -> blocks
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>

namespace {

constexpr int KI = 8;            // intrinsic stride
constexpr int NJ = KI + 6 + 3;   // derivative lanes: [intr(8) | pose(6) | point(3)]

// ---- forward-mode dual number (ceres/jet.h) -------------------------------------------------
struct Jet {
  double a;
  double v[NJ];
  Jet() : a(0) { std::memset(v, 0, sizeof v); }
  Jet(double s) : a(s) { std::memset(v, 0, sizeof v); }
  Jet(double s, int k) : a(s) { std::memset(v, 0, sizeof v); v[k] = 1.0; }
};
inline Jet operator+(const Jet &f, const Jet &g) { Jet h; h.a = f.a + g.a; for (int i = 0; i < NJ; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
inline Jet operator-(const Jet &f, const Jet &g) { Jet h; h.a = f.a - g.a; for (int i = 0; i < NJ; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
inline Jet operator-(const Jet &f) { Jet h; h.a = -f.a; for (int i = 0; i < NJ; ++i) h.v[i] = -f.v[i]; return h; }
inline Jet operator*(const Jet &f, const Jet &g) { Jet h; h.a = f.a * g.a; for (int i = 0; i < NJ; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
inline Jet operator/(const Jet &f, const Jet &g) {   // jet.h: g_inv = 1/g.a; f/g = (f.a*g_inv, (f.v - f.a*g_inv*g.v)*g_inv)
  Jet h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q;
  for (int i = 0; i < NJ; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
inline Jet jsqrt(const Jet &f) { Jet h; h.a = std::sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < NJ; ++i) h.v[i] = t * f.v[i]; return h; }
inline Jet jcos(const Jet &f) { Jet h; h.a = std::cos(f.a); const double s = -std::sin(f.a); for (int i = 0; i < NJ; ++i) h.v[i] = s * f.v[i]; return h; }
inline Jet jsin(const Jet &f) { Jet h; h.a = std::sin(f.a); const double c = std::cos(f.a); for (int i = 0; i < NJ; ++i) h.v[i] = c * f.v[i]; return h; }
inline Jet jatan(const Jet &f) { Jet h; h.a = std::atan(f.a); const double t = 1.0 / (1.0 + f.a * f.a); for (int i = 0; i < NJ; ++i) h.v[i] = t * f.v[i]; return h; }
// jet.h:682-695  atan2(g, f) ~= atan2(b, a) + (-b da + a db) / (a^2 + b^2)
inline Jet jatan2(const Jet &g, const Jet &f) { Jet h; h.a = std::atan2(g.a, f.a); const double t = 1.0 / (f.a * f.a + g.a * g.a); for (int i = 0; i < NJ; ++i) h.v[i] = t * (-g.a * f.v[i] + f.a * g.v[i]); return h; }
inline double jatan2(double y, double x) { return std::atan2(y, x); }
inline double jsqrt(double x) { return std::sqrt(x); }
inline double jcos(double x) { return std::cos(x); }
inline double jsin(double x) { return std::sin(x); }
inline double jatan(double x) { return std::atan(x); }
inline double val(double x) { return x; }
inline double val(const Jet &x) { return x.a; }

// rotation.h:563-622
template <typename T>
void angle_axis_rotate_point(const T aa[3], const T pt[3], T out[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (val(theta2) > std::numeric_limits<double>::epsilon()) {
    const T theta = jsqrt(theta2);
    const T costheta = jcos(theta);
    const T sintheta = jsin(theta);
    const T theta_inverse = T(1.0) / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - costheta);
    out[0] = pt[0] * costheta + wxp[0] * sintheta + w[0] * tmp;
    out[1] = pt[1] * costheta + wxp[1] * sintheta + w[1] * tmp;
    out[2] = pt[2] * costheta + wxp[2] * sintheta + w[2] * tmp;
  } else {
    const T wxp[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2], aa[0] * pt[1] - aa[1] * pt[0]};
    out[0] = pt[0] + wxp[0]; out[1] = pt[1] + wxp[1]; out[2] = pt[2] + wxp[2];
  }
}

// sfm_data_BA_ceres_camera_functor.hpp — one template per EINTRINSIC model.
template <typename T>
bool residual_functor(int model, const T *K, const T *ext, const T *X, const double *xy, T *res) {
  T p[3];
  angle_axis_rotate_point(ext, X, p);
  p[0] = p[0] + ext[3]; p[1] = p[1] + ext[4]; p[2] = p[2] + ext[5];
  if (model == 7) {  // CAMERA_SPHERICAL :662-717 — no intrinsic block; K[0], K[1] carry the image size (constants)
    const T lon = jatan2(p[0], p[2]);
    const T lat = jatan2(-p[1], jsqrt(p[0] * p[0] + p[2] * p[2]));
    const T c0 = lon / T(2 * M_PI), c1 = -lat / T(2 * M_PI);
    const double w = val(K[0]), h = val(K[1]);
    const T size(std::max(w, h));
    res[0] = c0 * size + T(w / 2.0) - T(xy[0]);
    res[1] = c1 * size + T(h / 2.0) - T(xy[1]);
    return true;
  }
  const T x = p[0] / p[2], y = p[1] / p[2];           // hnormalized()
  const T &focal = K[0]; const T &ppx = K[1]; const T &ppy = K[2];
  switch (model) {
    case 1: {  // PINHOLE_CAMERA :135-163
      res[0] = ppx + x * focal - T(xy[0]);
      res[1] = ppy + y * focal - T(xy[1]);
      return true; }
    case 2: {  // PINHOLE_CAMERA_RADIAL1 :236-268
      const T r2 = x * x + y * y;
      const T rc = T(1.0) + K[3] * r2;
      res[0] = ppx + (x * rc) * focal - T(xy[0]);
      res[1] = ppy + (y * rc) * focal - T(xy[1]);
      return true; }
    case 3: {  // PINHOLE_CAMERA_RADIAL3 :345-380
      const T r2 = x * x + y * y; const T r4 = r2 * r2; const T r6 = r4 * r2;
      const T rc = T(1.0) + K[3] * r2 + K[4] * r4 + K[5] * r6;
      res[0] = ppx + (x * rc) * focal - T(xy[0]);
      res[1] = ppy + (y * rc) * focal - T(xy[1]);
      return true; }
    case 4: {  // PINHOLE_CAMERA_BROWN :458-501
      const T r2 = x * x + y * y; const T r4 = r2 * r2; const T r6 = r4 * r2;
      const T rc = T(1.0) + K[3] * r2 + K[4] * r4 + K[5] * r6;
      const T &t1 = K[6]; const T &t2 = K[7];
      const T tx = t2 * (r2 + T(2.0) * x * x) + T(2.0) * t1 * x * y;
      const T ty = t1 * (r2 + T(2.0) * y * y) + T(2.0) * t2 * x * y;
      res[0] = ppx + (x * rc + tx) * focal - T(xy[0]);
      res[1] = ppy + (y * rc + ty) * focal - T(xy[1]);
      return true; }
    case 5: {  // PINHOLE_CAMERA_FISHEYE :580-627
      const T r2 = x * x + y * y;
      const T r = jsqrt(r2);
      const T theta = jatan(r), theta2 = theta * theta, theta3 = theta2 * theta, theta4 = theta2 * theta2,
              theta5 = theta4 * theta, theta7 = theta3 * theta3 * theta, theta8 = theta4 * theta4,
              theta9 = theta8 * theta;
      const T theta_dist = theta + K[3] * theta3 + K[4] * theta5 + K[5] * theta7 + K[6] * theta9;
      const T inv_r = val(r) > 1e-8 ? T(1.0) / r : T(1.0);
      const T cdist = val(r) > 1e-8 ? theta_dist * inv_r : T(1.0);
      res[0] = ppx + (x * cdist) * focal - T(xy[0]);
      res[1] = ppy + (y * cdist) * focal - T(xy[1]);
      return true; }
    default: return false;
  }
}

int model_nparams(int model) {
  switch (model) { case 1: return 3; case 2: return 4; case 3: return 6; case 4: return 8; case 5: return 7; case 7: return 0; default: return -1; }
}

// loss_function.cc:47-61 with a = huber_a (b = a^2); use_loss == 0 -> TrivialLoss
inline void loss_eval(bool use_loss, double a, double s, double rho[3]) {
  const double b = a * a;
  if (use_loss && s > b) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

struct Problem {
  int n_poses, n_intr, n_points, n_views; long n_obs;
  const int *intr_model, *view_pose, *view_intr, *obs_view, *obs_point;
  const double *obs_xy;
  // optional (all may be null / 0): ground control points and pose-centre priors
  const double *obs_weight;          // WeightedCostFunction weight per observation (functor.hpp:35-90)
  const unsigned char *obs_no_loss;  // residual block added with loss == nullptr (sfm_data_BA_ceres.cpp:418-435)
  const unsigned char *point_fixed;  // SetParameterBlockConstant on the landmark (:447)
  int n_priors; const int *prior_pose; const double *prior_center, *prior_weight; double prior_huber_a;  // :455-472
};

struct Options {   // mirrors include/omvg_b200.h omvg_ba_options
  int intrinsics_opt, extrinsics_opt, structure_opt, use_loss;
  double huber_a;
  int max_num_iterations;
  double function_tolerance, gradient_tolerance, parameter_tolerance;
  double initial_radius, max_radius, min_radius, min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
  int max_consecutive_invalid_steps;
};

// residual + (optionally) corrected Jacobian of one observation.  J layout: [2][NJ].
// Returns 1/2 rho(s).   residual_block.cc:68-196
double eval_obs(const Problem &P, const Options &O, const double *poses, const double *intr,
                const double *pts, long o, double r[2], double (*J)[NJ]) {
  const int v = P.obs_view[o], ip = P.view_pose[v], iq = P.view_intr[v], j = P.obs_point[o];
  const int model = P.intr_model[iq];
  double rho[3];
  const double wgt = P.obs_weight ? P.obs_weight[o] : 1.0;
  const bool use_loss = O.use_loss && !(P.obs_no_loss && P.obs_no_loss[o]);
  if (!J) {
    residual_functor<double>(model, intr + KI * iq, poses + 6 * ip, pts + 3 * j, P.obs_xy + 2 * o, r);
    if (P.obs_weight) { r[0] *= wgt; r[1] *= wgt; }
    const double s = r[0] * r[0] + r[1] * r[1];
    loss_eval(use_loss, O.huber_a, s, rho);
    return 0.5 * rho[0];                       // cost-only: no correction (residual_block.cc:170-172)
  }
  Jet K[KI], E[6], X[3], R[2];
  for (int k = 0; k < KI; ++k) K[k] = Jet(intr[KI * iq + k], k);
  for (int k = 0; k < 6; ++k) E[k] = Jet(poses[6 * ip + k], KI + k);
  for (int k = 0; k < 3; ++k) X[k] = Jet(pts[3 * j + k], KI + 6 + k);
  residual_functor<Jet>(model, K, E, X, P.obs_xy + 2 * o, R);
  if (P.obs_weight) { R[0] = R[0] * Jet(wgt); R[1] = R[1] * Jet(wgt); }
  r[0] = R[0].a; r[1] = R[1].a;
  const double s = r[0] * r[0] + r[1] * r[1];
  loss_eval(use_loss, O.huber_a, s, rho);
  // corrector.cc:41-110 — Huber has rho'' <= 0 everywhere, so the common case applies; the general
  // branch is kept for completeness.
  const double sqrt_rho1 = std::sqrt(rho[1]);
  double residual_scaling, alpha_sq_norm;
  if (s == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
  else {
    const double D = 1.0 + 2.0 * s * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha); alpha_sq_norm = alpha / s;
  }
  for (int c = 0; c < NJ; ++c) {
    if (alpha_sq_norm == 0.0) { J[0][c] = sqrt_rho1 * R[0].v[c]; J[1][c] = sqrt_rho1 * R[1].v[c]; }
    else {
      const double rtj = R[0].v[c] * r[0] + R[1].v[c] * r[1];
      J[0][c] = sqrt_rho1 * (R[0].v[c] - alpha_sq_norm * r[0] * rtj);
      J[1][c] = sqrt_rho1 * (R[1].v[c] - alpha_sq_norm * r[1] * rtj);
    }
  }
  r[0] *= residual_scaling; r[1] *= residual_scaling;
  return 0.5 * rho[0];
}

// Pose-centre prior (sfm_data_BA_ceres.cpp:44-80 PoseCenterConstraintCostFunction, added at :455-472
// with its own HuberLoss(Square(fitting_error))):  r = w .* (-R(-aa) t - c0).  J: [3][6] (pose lanes).
double eval_prior(const Problem &P, const double *poses, int k, double r[3], double (*J)[6]) {
  const double *pose = poses + 6 * P.prior_pose[k];
  const double *c0 = P.prior_center + 3 * k, *w = P.prior_weight + 3 * k;
  double rho[3];
  if (!J) {
    const double naa[3] = {-pose[0], -pose[1], -pose[2]};
    double c[3]; angle_axis_rotate_point<double>(naa, pose + 3, c);
    for (int i = 0; i < 3; ++i) r[i] = w[i] * (c[i] * -1.0 - c0[i]);
    loss_eval(true, P.prior_huber_a, r[0] * r[0] + r[1] * r[1] + r[2] * r[2], rho);
    return 0.5 * rho[0];
  }
  Jet E[6], naa[3], c[3], R[3];
  for (int i = 0; i < 6; ++i) E[i] = Jet(pose[i], KI + i);
  for (int i = 0; i < 3; ++i) naa[i] = -E[i];
  angle_axis_rotate_point<Jet>(naa, E + 3, c);
  for (int i = 0; i < 3; ++i) { R[i] = Jet(w[i]) * (c[i] * Jet(-1.0) - Jet(c0[i])); r[i] = R[i].a; }
  const double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  loss_eval(true, P.prior_huber_a, s, rho);
  const double sqrt_rho1 = std::sqrt(rho[1]);
  double residual_scaling, alpha_sq_norm;
  if (s == 0.0 || rho[2] <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
  else { const double D = 1.0 + 2.0 * s * rho[2] / rho[1]; const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha); alpha_sq_norm = alpha / s; }
  for (int c6 = 0; c6 < 6; ++c6) {
    const double rtj = R[0].v[KI + c6] * r[0] + R[1].v[KI + c6] * r[1] + R[2].v[KI + c6] * r[2];
    for (int i = 0; i < 3; ++i) J[i][c6] = sqrt_rho1 * (R[i].v[KI + c6] - alpha_sq_norm * r[i] * rtj);
  }
  for (int i = 0; i < 3; ++i) r[i] *= residual_scaling;
  return 0.5 * rho[0];
}

// in-place dense Cholesky A = L L^T (lower), returns false if not PD.
bool cholesky(std::vector<double> &A, int n) {
  const int B = 64;
  for (int k0 = 0; k0 < n; k0 += B) {
    const int kb = std::min(B, n - k0);
    for (int k = k0; k < k0 + kb; ++k) {          // factor diagonal block + panel column by column
      double d = A[(size_t)k * n + k];
      for (int p = k0; p < k; ++p) d -= A[(size_t)k * n + p] * A[(size_t)k * n + p];
      if (!(d > 0.0) || !std::isfinite(d)) return false;
      d = std::sqrt(d); A[(size_t)k * n + k] = d;
      #pragma omp parallel for schedule(static) if (n - k > 512)
      for (int i = k + 1; i < n; ++i) {
        double s = A[(size_t)i * n + k];
        for (int p = k0; p < k; ++p) s -= A[(size_t)i * n + p] * A[(size_t)k * n + p];
        A[(size_t)i * n + k] = s / d;
      }
    }
    const int r0 = k0 + kb;                       // trailing update A22 -= L21 L21^T (lower part)
    #pragma omp parallel for schedule(dynamic, 8)
    for (int i = r0; i < n; ++i) {
      const double *li = &A[(size_t)i * n + k0];
      for (int j = r0; j <= i; ++j) {
        const double *lj = &A[(size_t)j * n + k0];
        double s = 0; for (int p = 0; p < kb; ++p) s += li[p] * lj[p];
        A[(size_t)i * n + j] -= s;
      }
    }
  }
  return true;
}
void chol_solve(const std::vector<double> &L, int n, double *b) {
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int p = 0; p < i; ++p) s -= L[(size_t)i * n + p] * b[p]; b[i] = s / L[(size_t)i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int p = i + 1; p < n; ++p) s -= L[(size_t)p * n + i] * b[p]; b[i] = s / L[(size_t)i * n + i]; }
}

struct Layout {              // effective (tangent) columns: [points | poses | intrinsics]
  std::vector<int> pose_free;             // free coordinates of a pose block (same for all poses)
  std::vector<std::vector<int>> intr_free;
  std::vector<int> pose_col, intr_col;    // first reduced column of each block, -1 if constant
  int pt_cols, n_red, n_eff;
  bool pts_var;
};

// sfm_data_BA_ceres.cpp:275-305 (poses), 321-344 + Camera_Pinhole*.hpp subsetParameterization
// (intrinsics), 394-395 (structure)
Layout make_layout(const Problem &P, const Options &O) {
  Layout L;
  if (O.extrinsics_opt != 1) {
    const bool rot = (O.extrinsics_opt & 2) != 0, tr = (O.extrinsics_opt & 4) != 0;
    if (O.extrinsics_opt == 4) { L.pose_free = {3, 4, 5}; }
    else if (O.extrinsics_opt == 2) { L.pose_free = {0, 1, 2}; }
    else { (void)rot; (void)tr; L.pose_free = {0, 1, 2, 3, 4, 5}; }
  }
  L.intr_free.resize(P.n_intr);
  for (int q = 0; q < P.n_intr; ++q) {
    const int k = model_nparams(P.intr_model[q]);
    if (O.intrinsics_opt == 1 || (O.intrinsics_opt & 1)) continue;   // NONE: whole block constant
    for (int i = 0; i < k; ++i) {
      bool constant;
      if (i == 0) constant = !(O.intrinsics_opt & 2);
      else if (i <= 2) constant = !(O.intrinsics_opt & 4);
      else constant = !(O.intrinsics_opt & 8);
      if (!constant) L.intr_free[q].push_back(i);
    }
  }
  L.pts_var = O.structure_opt != 0;
  L.pt_cols = L.pts_var ? 3 * P.n_points : 0;
  int c = 0;
  L.pose_col.assign(P.n_poses, -1);
  for (int p = 0; p < P.n_poses; ++p) if (!L.pose_free.empty()) { L.pose_col[p] = c; c += (int)L.pose_free.size(); }
  L.intr_col.assign(P.n_intr, -1);
  for (int q = 0; q < P.n_intr; ++q) if (!L.intr_free[q].empty()) { L.intr_col[q] = c; c += (int)L.intr_free[q].size(); }
  L.n_red = c; L.n_eff = L.pt_cols + c;
  return L;
}

// ceres/include/ceres/rotation.h:377-420 AngleAxisToRotationMatrix (row-major R here)
void aa_to_R(const double *aa, double R[9]) {
  const double theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (theta2 > std::numeric_limits<double>::epsilon()) {
    const double theta = std::sqrt(theta2), wx = aa[0] / theta, wy = aa[1] / theta, wz = aa[2] / theta;
    const double c = std::cos(theta), s = std::sin(theta);
    R[0] = c + wx * wx * (1.0 - c); R[3] = wz * s + wx * wy * (1.0 - c); R[6] = -wy * s + wx * wz * (1.0 - c);
    R[1] = wx * wy * (1.0 - c) - wz * s; R[4] = c + wy * wy * (1.0 - c); R[7] = wx * s + wy * wz * (1.0 - c);
    R[2] = wy * s + wx * wz * (1.0 - c); R[5] = -wx * s + wy * wz * (1.0 - c); R[8] = c + wz * wz * (1.0 - c);
  } else {
    R[0] = 1; R[3] = aa[2]; R[6] = -aa[1]; R[1] = -aa[2]; R[4] = 1; R[7] = aa[0]; R[2] = aa[1]; R[5] = -aa[0]; R[8] = 1;
  }
}

// Write-back rule of Adjust (sfm_data_BA_ceres.cpp:528-555): with ADJUST_ROTATION only the rotation
// of the Pose3 is replaced, its CENTRE is kept, so the returned t is -R_new * C_old (not the t that
// Ceres held constant).  ADJUST_TRANSLATION / ADJUST_ALL return (R_refined, t_refined) unchanged.
void write_back_poses(int extrinsics_opt, int n_poses, const double *old_poses, const double *x_pose, double *out) {
  for (int p = 0; p < n_poses; ++p) {
    for (int k = 0; k < 6; ++k) out[6 * p + k] = x_pose[6 * p + k];
    if (extrinsics_opt != 2) continue;
    double Ro[9], Rn[9], C[3];
    aa_to_R(old_poses + 6 * p, Ro); aa_to_R(x_pose + 6 * p, Rn);
    const double *t = old_poses + 6 * p + 3;
    for (int i = 0; i < 3; ++i) C[i] = -(Ro[0 * 3 + i] * t[0] + Ro[1 * 3 + i] * t[1] + Ro[2 * 3 + i] * t[2]);
    for (int i = 0; i < 3; ++i) out[6 * p + 3 + i] = -(Rn[i * 3 + 0] * C[0] + Rn[i * 3 + 1] * C[1] + Rn[i * 3 + 2] * C[2]);
  }
}

}  // namespace

extern "C" {

// 1/2 sum rho(|r|^2)
double oracle_ba_cost(int n_poses, const double *poses, int n_intr, const double *intr, const int *intr_model,
                      int n_points, const double *points, int n_views, const int *view_pose, const int *view_intr,
                      long n_obs, const int *obs_view, const int *obs_point, const double *obs_xy,
                      int use_loss, double huber_a) {
  Problem P{n_poses, n_intr, n_points, n_views, n_obs, intr_model, view_pose, view_intr, obs_view, obs_point, obs_xy};
  Options O{}; O.use_loss = use_loss; O.huber_a = huber_a;
  long double c = 0;
  for (long o = 0; o < n_obs; ++o) { double r[2]; c += eval_obs(P, O, poses, intr, points, o, r, nullptr); }
  return (double)c;
}

// corrected residuals r[n_obs][2] and Jacobian blocks (row-major 2 x k): J_intr[n_obs][2][8],
// J_pose[n_obs][2][6], J_point[n_obs][2][3]; returns the cost.  (unscaled: no Jacobi scaling)
double oracle_ba_eval(int n_poses, const double *poses, int n_intr, const double *intr, const int *intr_model,
                      int n_points, const double *points, int n_views, const int *view_pose, const int *view_intr,
                      long n_obs, const int *obs_view, const int *obs_point, const double *obs_xy,
                      int use_loss, double huber_a, double *r, double *J_intr, double *J_pose, double *J_point) {
  Problem P{n_poses, n_intr, n_points, n_views, n_obs, intr_model, view_pose, view_intr, obs_view, obs_point, obs_xy};
  Options O{}; O.use_loss = use_loss; O.huber_a = huber_a;
  long double c = 0;
  for (long o = 0; o < n_obs; ++o) {
    double J[2][NJ];
    c += eval_obs(P, O, poses, intr, points, o, r + 2 * o, J);
    for (int row = 0; row < 2; ++row) {
      for (int k = 0; k < KI; ++k) J_intr[(o * 2 + row) * KI + k] = J[row][k];
      for (int k = 0; k < 6; ++k) J_pose[(o * 2 + row) * 6 + k] = J[row][KI + k];
      for (int k = 0; k < 3; ++k) J_point[(o * 2 + row) * 3 + k] = J[row][KI + 6 + k];
    }
  }
  return (double)c;
}

// opts (doubles, same order as omvg_ba_options fields): see Options above.
// summary[0]=initial_cost [1]=final_cost [2]=iterations (Ceres' "Minimizer iterations" = recorded
// iterations incl. iteration 0, excl. the one that returns on a tolerance test; [8]=LM steps attempted) [3]=successful steps (incl. iteration 0) [4]=unsuccessful steps
// [5]=termination (0 CONVERGENCE fn-tol, 1 param-tol, 2 gradient-tol, 3 NO_CONVERGENCE max iters,
// 4 min radius, -1 FAILURE) [6]=usable (1/0)
// trace (optional, may be NULL): per LM iteration 4 doubles {cost, cost_change, radius, rho}, cap trace_cap rows.
static int solve_impl(const Problem &P, double *poses, double *intr, double *points,
                      const double *opts, double *summary, double *trace, int trace_cap) {
  const int n_poses = P.n_poses, n_intr = P.n_intr, n_points = P.n_points; const long n_obs = P.n_obs;
  const int *intr_model = P.intr_model, *view_pose = P.view_pose, *view_intr = P.view_intr, *obs_view = P.obs_view, *obs_point = P.obs_point;
  const int n_pri = P.n_priors;
  Options O;
  O.intrinsics_opt = (int)opts[0]; O.extrinsics_opt = (int)opts[1]; O.structure_opt = (int)opts[2];
  O.use_loss = (int)opts[3]; O.huber_a = opts[4]; O.max_num_iterations = (int)opts[5];
  O.function_tolerance = opts[6]; O.gradient_tolerance = opts[7]; O.parameter_tolerance = opts[8];
  O.initial_radius = opts[9]; O.max_radius = opts[10]; O.min_radius = opts[11];
  O.min_relative_decrease = opts[12]; O.min_lm_diagonal = opts[13]; O.max_lm_diagonal = opts[14];
  O.max_consecutive_invalid_steps = (int)opts[15];
  const Layout L = make_layout(P, O);
  const int npf = (int)L.pose_free.size();
  const int n_eff = L.n_eff, n_red = L.n_red;

  // observations grouped by point (for the Schur elimination)
  std::vector<long> pt_start(n_points + 1, 0);
  for (long o = 0; o < n_obs; ++o) pt_start[obs_point[o] + 1]++;
  for (int j = 0; j < n_points; ++j) pt_start[j + 1] += pt_start[j];
  std::vector<long> by_pt(n_obs);
  { std::vector<long> cur(pt_start.begin(), pt_start.end() - 1);
    for (long o = 0; o < n_obs; ++o) by_pt[cur[obs_point[o]]++] = o; }

  std::vector<double> x_pose(poses, poses + 6 * n_poses), x_intr(intr, intr + KI * n_intr), x_pt(points, points + 3 * n_points);
  std::vector<double> c_pose, c_intr, c_pt;                         // candidates
  std::vector<double> res(2 * n_obs), Jall((size_t)n_obs * 2 * NJ);  // corrected r and J (scaled after eval)
  std::vector<double> scale(n_eff, 1.0), grad(n_eff), diag(n_eff), lmD(n_eff), step(n_eff), delta(n_eff);

  auto col_of = [&](long o, int lane) -> int {   // derivative lane -> effective column (or -1)
    const int v = obs_view[o];
    if (lane < KI) { const int q = view_intr[v]; if (L.intr_col[q] < 0) return -1;
      const auto &f = L.intr_free[q]; for (size_t t = 0; t < f.size(); ++t) if (f[t] == lane) return L.pt_cols + L.intr_col[q] + (int)t; return -1; }
    if (lane < KI + 6) { const int p = view_pose[v]; if (L.pose_col[p] < 0) return -1;
      for (int t = 0; t < npf; ++t) if (L.pose_free[t] == lane - KI) return L.pt_cols + L.pose_col[p] + t; return -1; }
    return (L.pts_var && !(P.point_fixed && P.point_fixed[obs_point[o]])) ? 3 * obs_point[o] + (lane - KI - 6) : -1;
  };
  // pose-centre priors: residual rows appended after the observations; they touch pose columns only
  std::vector<double> resP(3 * (size_t)n_pri), JP((size_t)n_pri * 18);
  std::vector<int> colP((size_t)n_pri * 6, -1);
  for (int k = 0; k < n_pri; ++k) { const int p = P.prior_pose[k]; if (L.pose_col[p] < 0) continue;
    for (int t = 0; t < npf; ++t) colP[6 * k + L.pose_free[t]] = L.pt_cols + L.pose_col[p] + t; }
  std::vector<int> colmap((size_t)n_obs * NJ);
  for (long o = 0; o < n_obs; ++o) for (int l = 0; l < NJ; ++l) colmap[o * NJ + l] = col_of(o, l);

  double x_cost = 0;
  bool have_scale = false;
  auto evaluate_jac = [&](const double *xp, const double *xi, const double *xq) {
    long double c = 0;
    std::fill(grad.begin(), grad.end(), 0.0);
    for (long o = 0; o < n_obs; ++o) {
      double J[2][NJ];
      c += eval_obs(P, O, xp, xi, xq, o, &res[2 * o], J);
      for (int l = 0; l < NJ; ++l) {
        const int col = colmap[o * NJ + l];
        Jall[(o * 2 + 0) * NJ + l] = col >= 0 ? J[0][l] : 0.0; Jall[(o * 2 + 1) * NJ + l] = col >= 0 ? J[1][l] : 0.0;   // constant lanes have no column
        if (col >= 0) grad[col] += J[0][l] * res[2 * o] + J[1][l] * res[2 * o + 1];   // program_evaluator.h:239-256 (unscaled J)
      }
    }
    for (int k = 0; k < n_pri; ++k) {
      double J[3][6];
      c += eval_prior(P, xp, k, &resP[3 * k], J);
      for (int l = 0; l < 6; ++l) { const int col = colP[6 * k + l];
        for (int i = 0; i < 3; ++i) JP[(k * 3 + i) * 6 + l] = J[i][l];
        if (col >= 0) grad[col] += J[0][l] * resP[3 * k] + J[1][l] * resP[3 * k + 1] + J[2][l] * resP[3 * k + 2]; }
    }
    x_cost = (double)c;
    if (!have_scale) {            // trust_region_minimizer.cc:239-250, iteration 0 only
      std::vector<double> n2(n_eff, 0.0);
      for (long o = 0; o < n_obs; ++o) for (int l = 0; l < NJ; ++l) { const int col = colmap[o * NJ + l];
        if (col >= 0) n2[col] += Jall[(o * 2) * NJ + l] * Jall[(o * 2) * NJ + l] + Jall[(o * 2 + 1) * NJ + l] * Jall[(o * 2 + 1) * NJ + l]; }
      for (int k = 0; k < n_pri; ++k) for (int l = 0; l < 6; ++l) { const int col = colP[6 * k + l];
        if (col >= 0) for (int i = 0; i < 3; ++i) n2[col] += JP[(k * 3 + i) * 6 + l] * JP[(k * 3 + i) * 6 + l]; }
      for (int i = 0; i < n_eff; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(n2[i]));
      have_scale = true;
    }
    for (long o = 0; o < n_obs; ++o) for (int l = 0; l < NJ; ++l) { const int col = colmap[o * NJ + l];   // :253 ScaleColumns
      if (col >= 0) { Jall[(o * 2) * NJ + l] *= scale[col]; Jall[(o * 2 + 1) * NJ + l] *= scale[col]; } }
    for (int k = 0; k < n_pri; ++k) for (int l = 0; l < 6; ++l) { const int col = colP[6 * k + l];
      if (col >= 0) for (int i = 0; i < 3; ++i) JP[(k * 3 + i) * 6 + l] *= scale[col]; }
  };
  auto cost_only = [&](const double *xp, const double *xi, const double *xq) {
    long double c = 0; for (long o = 0; o < n_obs; ++o) { double r[2]; c += eval_obs(P, O, xp, xi, xq, o, r, nullptr); }
    for (int k = 0; k < n_pri; ++k) { double r[3]; c += eval_prior(P, xp, k, r, nullptr); }
    return (double)c; };
  auto x_norm_of = [&]() {    // norm over the reduced program's parameter vector (constant blocks are removed)
    long double s = 0;
    if (L.pts_var) for (int j = 0; j < n_points; ++j) if (!(P.point_fixed && P.point_fixed[j])) for (int a = 0; a < 3; ++a) s += x_pt[3 * j + a] * x_pt[3 * j + a];
    if (npf) for (double v : x_pose) s += v * v;
    for (int q = 0; q < n_intr; ++q) if (L.intr_col[q] >= 0) for (int k = 0; k < model_nparams(intr_model[q]); ++k) s += x_intr[KI * q + k] * x_intr[KI * q + k];
    return std::sqrt((double)s); };

  // ---- linear solve:  min |J y - r|^2 + |D y|^2  (levenberg_marquardt_strategy.cc:65-145) ----
  std::vector<double> S, rhs(n_red), y(n_eff);
  auto solve_step = [&]() -> bool {
    S.assign((size_t)n_red * n_red, 0.0);
    std::fill(rhs.begin(), rhs.end(), 0.0);
    std::vector<double> Einv(L.pts_var ? 9 * (size_t)n_points : 0), Etb(L.pts_var ? 3 * (size_t)n_points : 0);
    const int base = L.pt_cols;
    bool ok = true;
    if (L.pts_var) {
      #pragma omp parallel for schedule(dynamic, 64)
      for (int j = 0; j < n_points; ++j) {
        // EtE + D^2, Etb    (schur_eliminator_impl.h:434-490)
        double ete[9] = {0}, etb[3] = {0};
        const long a0 = pt_start[j], a1 = pt_start[j + 1];
        for (long t = a0; t < a1; ++t) { const long o = by_pt[t];
          for (int row = 0; row < 2; ++row) { const double *Jr = &Jall[(o * 2 + row) * NJ + KI + 6];
            for (int a = 0; a < 3; ++a) { etb[a] += Jr[a] * res[2 * o + row]; for (int b = 0; b < 3; ++b) ete[a * 3 + b] += Jr[a] * Jr[b]; } } }
        for (int a = 0; a < 3; ++a) ete[a * 3 + a] += lmD[3 * j + a] * lmD[3 * j + a];
        // 3x3 inverse via LLT (invert_psd_matrix.h:56-59)
        double l00 = std::sqrt(ete[0]), l10 = ete[3] / l00, l20 = ete[6] / l00;
        double l11 = std::sqrt(ete[4] - l10 * l10), l21 = (ete[7] - l20 * l10) / l11;
        double l22 = std::sqrt(ete[8] - l20 * l20 - l21 * l21);
        if (!(std::isfinite(l00) && std::isfinite(l11) && std::isfinite(l22) && l00 > 0 && l11 > 0 && l22 > 0)) { ok = false; continue; }
        double inv[9];
        for (int c = 0; c < 3; ++c) { double b[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
          b[0] /= l00; b[1] = (b[1] - l10 * b[0]) / l11; b[2] = (b[2] - l20 * b[0] - l21 * b[1]) / l22;
          b[2] /= l22; b[1] = (b[1] - l21 * b[2]) / l11; b[0] = (b[0] - l10 * b[1] - l20 * b[2]) / l00;
          inv[0 * 3 + c] = b[0]; inv[1 * 3 + c] = b[1]; inv[2 * 3 + c] = b[2]; }
        std::memcpy(&Einv[9 * (size_t)j], inv, sizeof inv); std::memcpy(&Etb[3 * (size_t)j], etb, sizeof etb);
        // per observation: F row (free cam + intr columns), EtF
        const int nobs_j = (int)(a1 - a0);
        std::vector<int> idx; std::vector<double> F, EtF;   // F: [obs][2][w], EtF: [obs][3][w], idx: [obs][w]
        const int W = 6 + KI;
        idx.assign((size_t)nobs_j * W, -1); F.assign((size_t)nobs_j * 2 * W, 0.0); EtF.assign((size_t)nobs_j * 3 * W, 0.0);
        for (int t = 0; t < nobs_j; ++t) { const long o = by_pt[a0 + t];
          for (int l = 0; l < KI + 6; ++l) { const int col = colmap[o * NJ + l]; if (col < 0) continue;
            idx[t * W + l] = col - base;
            const double j0 = Jall[(o * 2) * NJ + l], j1 = Jall[(o * 2 + 1) * NJ + l];
            F[(t * 2) * W + l] = j0; F[(t * 2 + 1) * W + l] = j1;
            for (int a = 0; a < 3; ++a) EtF[(t * 3 + a) * W + l] = Jall[(o * 2) * NJ + KI + 6 + a] * j0 + Jall[(o * 2 + 1) * NJ + KI + 6 + a] * j1; } }
        double ie[3]; for (int a = 0; a < 3; ++a) ie[a] = inv[a * 3] * etb[0] + inv[a * 3 + 1] * etb[1] + inv[a * 3 + 2] * etb[2];
        for (int t = 0; t < nobs_j; ++t) { const long o = by_pt[a0 + t];
          for (int l = 0; l < W; ++l) { const int ci = idx[t * W + l]; if (ci < 0) continue;
            // rhs += F'b - (EtF)' Einv Etb     (:374-410)
            double v = F[(t * 2) * W + l] * res[2 * o] + F[(t * 2 + 1) * W + l] * res[2 * o + 1];
            for (int a = 0; a < 3; ++a) v -= EtF[(t * 3 + a) * W + l] * ie[a];
            #pragma omp atomic
            rhs[ci] += v;
            // S += F'F  (same observation)
            for (int m = 0; m < W; ++m) { const int cj = idx[t * W + m]; if (cj < 0) continue;
              const double f = F[(t * 2) * W + l] * F[(t * 2) * W + m] + F[(t * 2 + 1) * W + l] * F[(t * 2 + 1) * W + m];
              #pragma omp atomic
              S[(size_t)ci * n_red + cj] += f; }
            // S -= (EtF_t)' Einv (EtF_u) for every observation u of this point   (:499-548)
            double g[3]; for (int a = 0; a < 3; ++a) g[a] = EtF[(t * 3) * W + l] * inv[a] + EtF[(t * 3 + 1) * W + l] * inv[3 + a] + EtF[(t * 3 + 2) * W + l] * inv[6 + a];
            for (int u = 0; u < nobs_j; ++u) for (int m = 0; m < W; ++m) { const int cj = idx[u * W + m]; if (cj < 0) continue;
              const double f = g[0] * EtF[(u * 3) * W + m] + g[1] * EtF[(u * 3 + 1) * W + m] + g[2] * EtF[(u * 3 + 2) * W + m];
              #pragma omp atomic
              S[(size_t)ci * n_red + cj] -= f; } } }
      }
    } else {
      for (long o = 0; o < n_obs; ++o) for (int l = 0; l < KI + 6; ++l) { const int ci = colmap[o * NJ + l]; if (ci < 0) continue;
        rhs[ci] += Jall[(o * 2) * NJ + l] * res[2 * o] + Jall[(o * 2 + 1) * NJ + l] * res[2 * o + 1];
        for (int m = 0; m < KI + 6; ++m) { const int cj = colmap[o * NJ + m]; if (cj < 0) continue;
          S[(size_t)ci * n_red + cj] += Jall[(o * 2) * NJ + l] * Jall[(o * 2) * NJ + m] + Jall[(o * 2 + 1) * NJ + l] * Jall[(o * 2 + 1) * NJ + m]; } }
    }
    if (!ok) return false;
    for (int k = 0; k < n_pri; ++k) for (int l = 0; l < 6; ++l) { const int ci = colP[6 * k + l]; if (ci < 0) continue;   // rows without an e-block
      for (int i = 0; i < 3; ++i) rhs[ci - base] += JP[(k * 3 + i) * 6 + l] * resP[3 * k + i];
      for (int m = 0; m < 6; ++m) { const int cj = colP[6 * k + m]; if (cj < 0) continue;
        for (int i = 0; i < 3; ++i) S[(size_t)(ci - base) * n_red + (cj - base)] += JP[(k * 3 + i) * 6 + l] * JP[(k * 3 + i) * 6 + m]; } }
    for (int i = 0; i < n_red; ++i) S[(size_t)i * n_red + i] += lmD[base + i] * lmD[base + i];
    std::vector<double> z(rhs);
    if (n_red > 0) { if (!cholesky(S, n_red)) return false; chol_solve(S, n_red, z.data()); }
    for (int i = 0; i < n_red; ++i) y[base + i] = z[i];
    if (L.pts_var) {        // back substitution (:303-366)
      for (int j = 0; j < n_points; ++j) {
        double b[3] = {Etb[3 * (size_t)j], Etb[3 * (size_t)j + 1], Etb[3 * (size_t)j + 2]};
        for (long t = pt_start[j]; t < pt_start[j + 1]; ++t) { const long o = by_pt[t];
          for (int l = 0; l < KI + 6; ++l) { const int col = colmap[o * NJ + l]; if (col < 0) continue;
            const double zz = z[col - base];
            for (int a = 0; a < 3; ++a) b[a] -= (Jall[(o * 2) * NJ + KI + 6 + a] * Jall[(o * 2) * NJ + l] + Jall[(o * 2 + 1) * NJ + KI + 6 + a] * Jall[(o * 2 + 1) * NJ + l]) * zz; } }
        const double *inv = &Einv[9 * (size_t)j];
        for (int a = 0; a < 3; ++a) y[3 * j + a] = inv[a * 3] * b[0] + inv[a * 3 + 1] * b[1] + inv[a * 3 + 2] * b[2];
      }
    }
    for (int i = 0; i < n_eff; ++i) if (!std::isfinite(y[i])) return false;
    for (int i = 0; i < n_eff; ++i) step[i] = -y[i];
    return true;
  };

  auto plus = [&](const std::vector<double> &d) {   // program.cc:115-127 + SubsetParameterization::Plus
    c_pose = x_pose; c_intr = x_intr; c_pt = x_pt;
    if (L.pts_var) for (int i = 0; i < 3 * n_points; ++i) c_pt[i] += d[i];
    for (int p = 0; p < n_poses; ++p) if (L.pose_col[p] >= 0) for (int t = 0; t < npf; ++t) c_pose[6 * p + L.pose_free[t]] += d[L.pt_cols + L.pose_col[p] + t];
    for (int q = 0; q < n_intr; ++q) if (L.intr_col[q] >= 0) for (size_t t = 0; t < L.intr_free[q].size(); ++t) c_intr[KI * q + L.intr_free[q][t]] += d[L.pt_cols + L.intr_col[q] + (int)t];
  };

  // ---------------- trust_region_minimizer.cc:66-119 ----------------
  double radius = O.initial_radius, decrease_factor = 2.0; bool reuse_diagonal = false;
  int iteration = 0, n_success = 0, n_fail = 0, n_invalid_consec = 0, termination = 3;
  double x_norm = -1.0;                 // Init(): "x_norm_ = -1;  // Invalid value"
  evaluate_jac(x_pose.data(), x_intr.data(), x_pt.data());
  const double initial_cost = x_cost;
  bool step_is_successful = true;
  double gradient_max_norm = 0; for (double g : grad) gradient_max_norm = std::max(gradient_max_norm, std::fabs(g));
  double reference_cost = x_cost, accumulated_reference = 0.0, current_cost = x_cost;   // step evaluator
  int n_trace = 0;
  if (trace && n_trace < trace_cap) { trace[0] = x_cost; trace[1] = 0; trace[2] = radius; trace[3] = 0; n_trace = 1; }
  bool failure = false;
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue (:291-335)
    if (step_is_successful) ++n_success; else ++n_fail;
    if (iteration >= O.max_num_iterations) { termination = 3; break; }
    if (step_is_successful && gradient_max_norm <= O.gradient_tolerance) { termination = 2; break; }
    if (radius <= O.min_radius) { termination = 4; break; }
    ++iteration;
    // ComputeTrustRegionStep (:355-424)
    if (!reuse_diagonal) {
      std::fill(diag.begin(), diag.end(), 0.0);
      for (long o = 0; o < n_obs; ++o) for (int l = 0; l < NJ; ++l) { const int col = colmap[o * NJ + l];
        if (col >= 0) diag[col] += Jall[(o * 2) * NJ + l] * Jall[(o * 2) * NJ + l] + Jall[(o * 2 + 1) * NJ + l] * Jall[(o * 2 + 1) * NJ + l]; }
      for (int k = 0; k < n_pri; ++k) for (int l = 0; l < 6; ++l) { const int col = colP[6 * k + l];
        if (col >= 0) for (int i = 0; i < 3; ++i) diag[col] += JP[(k * 3 + i) * 6 + l] * JP[(k * 3 + i) * 6 + l]; }
      for (int i = 0; i < n_eff; ++i) diag[i] = std::min(std::max(diag[i], O.min_lm_diagonal), O.max_lm_diagonal);
    }
    for (int i = 0; i < n_eff; ++i) lmD[i] = std::sqrt(diag[i] / radius);
    const bool solved = solve_step();
    reuse_diagonal = true;
    bool step_is_valid = false; double model_cost_change = 0;
    if (solved) {
      long double m = 0;
      for (long o = 0; o < n_obs; ++o) for (int row = 0; row < 2; ++row) {
        double mr = 0; for (int l = 0; l < NJ; ++l) { const int col = colmap[o * NJ + l]; if (col >= 0) mr += Jall[(o * 2 + row) * NJ + l] * step[col]; }
        m += -mr * (res[2 * o + row] + mr / 2.0); }
      for (int k = 0; k < n_pri; ++k) for (int i = 0; i < 3; ++i) {
        double mr = 0; for (int l = 0; l < 6; ++l) { const int col = colP[6 * k + l]; if (col >= 0) mr += JP[(k * 3 + i) * 6 + l] * step[col]; }
        m += -mr * (resP[3 * k + i] + mr / 2.0); }
      model_cost_change = (double)m;
      step_is_valid = model_cost_change > 0.0;
    }
    if (!step_is_valid) {        // HandleInvalidStep (:429-462)
      if (++n_invalid_consec >= O.max_consecutive_invalid_steps) { failure = true; termination = -1; break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;   // StepIsInvalid == StepRejected
      step_is_successful = false;
      if (trace && n_trace < trace_cap) { trace[4 * n_trace] = x_cost; trace[4 * n_trace + 1] = 0; trace[4 * n_trace + 2] = radius; trace[4 * n_trace + 3] = 0; ++n_trace; }
      continue;
    }
    n_invalid_consec = 0;
    for (int i = 0; i < n_eff; ++i) delta[i] = step[i] * scale[i];
    plus(delta);
    const double candidate_cost = cost_only(c_pose.data(), c_intr.data(), c_pt.data());
    // ParameterToleranceReached (:667-685)
    long double sn = 0;
    if (L.pts_var) for (int i = 0; i < 3 * n_points; ++i) sn += (x_pt[i] - c_pt[i]) * (x_pt[i] - c_pt[i]);
    if (npf) for (int i = 0; i < 6 * n_poses; ++i) sn += (x_pose[i] - c_pose[i]) * (x_pose[i] - c_pose[i]);
    for (int i = 0; i < KI * n_intr; ++i) sn += (x_intr[i] - c_intr[i]) * (x_intr[i] - c_intr[i]);
    const double step_norm = std::sqrt((double)sn);
    if (step_norm <= O.parameter_tolerance * (x_norm + O.parameter_tolerance)) { termination = 1; break; }
    // FunctionToleranceReached (:688-705)
    const double cost_change = x_cost - candidate_cost;
    if (std::fabs(cost_change) <= O.function_tolerance * x_cost) {
      if (trace && n_trace < trace_cap) { trace[4 * n_trace] = candidate_cost; trace[4 * n_trace + 1] = cost_change; trace[4 * n_trace + 2] = radius; trace[4 * n_trace + 3] = 0; ++n_trace; }
      termination = 0; break; }
    // IsStepSuccessful (:736-765), trust_region_step_evaluator.cc:51-59
    const double rel = (current_cost - candidate_cost) / model_cost_change;
    const double hist = (reference_cost - candidate_cost) / (accumulated_reference + model_cost_change);
    const double rho = std::max(rel, hist);
    if (rho > O.min_relative_decrease) {     // HandleSuccessfulStep (:767-780)
      x_pose = c_pose; x_intr = c_intr; x_pt = c_pt; x_norm = x_norm_of();
      evaluate_jac(x_pose.data(), x_intr.data(), x_pt.data());
      gradient_max_norm = 0; for (double g : grad) gradient_max_norm = std::max(gradient_max_norm, std::fabs(g));
      step_is_successful = true;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)); radius = std::min(O.max_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      current_cost = candidate_cost; reference_cost = candidate_cost; accumulated_reference = 0.0;   // monotonic mode
    } else {                                  // HandleUnsuccessfulStep (:782-786)
      step_is_successful = false;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
    if (trace && n_trace < trace_cap) { trace[4 * n_trace] = step_is_successful ? x_cost : candidate_cost; trace[4 * n_trace + 1] = cost_change; trace[4 * n_trace + 2] = radius; trace[4 * n_trace + 3] = rho; ++n_trace; }
  }
  const bool usable = !failure;
  if (usable) {      // solver.cc: user state is updated only when the solution is usable
    if (O.extrinsics_opt != 1) {                 // sfm_data_BA_ceres.cpp:529: poses untouched when extrinsics are NONE
      std::vector<double> old(poses, poses + 6 * n_poses);
      write_back_poses(O.extrinsics_opt, n_poses, old.data(), x_pose.data(), poses);
    }
    std::memcpy(intr, x_intr.data(), sizeof(double) * KI * n_intr);
    std::memcpy(points, x_pt.data(), sizeof(double) * 3 * n_points);
  }
  summary[0] = initial_cost; summary[1] = x_cost; summary[2] = n_success + n_fail; summary[3] = n_success;
  summary[4] = n_fail; summary[5] = termination; summary[6] = usable ? 1 : 0; summary[7] = n_trace;
  summary[8] = iteration;   // linear solves attempted (includes the terminating iteration)
  return usable ? 0 : 1;
}

extern "C" {
int oracle_ba_solve(int n_poses, double *poses, int n_intr, double *intr, const int *intr_model,
                    int n_points, double *points, int n_views, const int *view_pose, const int *view_intr,
                    long n_obs, const int *obs_view, const int *obs_point, const double *obs_xy,
                    const double *opts, double *summary, double *trace, int trace_cap) {
  Problem P{n_poses, n_intr, n_points, n_views, n_obs, intr_model, view_pose, view_intr, obs_view, obs_point, obs_xy};
  return solve_impl(P, poses, intr, points, opts, summary, trace, trace_cap);
}

// Same with ground control points (per-observation weight / no-loss flag, fixed landmarks) and
// pose-centre priors.  Any of the extension pointers may be NULL.
int oracle_ba_solve_ex(int n_poses, double *poses, int n_intr, double *intr, const int *intr_model,
                       int n_points, double *points, int n_views, const int *view_pose, const int *view_intr,
                       long n_obs, const int *obs_view, const int *obs_point, const double *obs_xy,
                       const double *obs_weight, const unsigned char *obs_no_loss, const unsigned char *point_fixed,
                       int n_priors, const int *prior_pose, const double *prior_center, const double *prior_weight, double prior_huber_a,
                       const double *opts, double *summary, double *trace, int trace_cap) {
  Problem P{n_poses, n_intr, n_points, n_views, n_obs, intr_model, view_pose, view_intr, obs_view, obs_point, obs_xy,
            obs_weight, obs_no_loss, point_fixed, n_priors, prior_pose, prior_center, prior_weight, prior_huber_a};
  return solve_impl(P, poses, intr, points, opts, summary, trace, trace_cap);
}

double oracle_ba_cost_ex(int n_poses, const double *poses, int n_intr, const double *intr, const int *intr_model,
                         int n_points, const double *points, int n_views, const int *view_pose, const int *view_intr,
                         long n_obs, const int *obs_view, const int *obs_point, const double *obs_xy,
                         const double *obs_weight, const unsigned char *obs_no_loss,
                         int n_priors, const int *prior_pose, const double *prior_center, const double *prior_weight, double prior_huber_a,
                         int use_loss, double huber_a) {
  Problem P{n_poses, n_intr, n_points, n_views, n_obs, intr_model, view_pose, view_intr, obs_view, obs_point, obs_xy,
            obs_weight, obs_no_loss, nullptr, n_priors, prior_pose, prior_center, prior_weight, prior_huber_a};
  Options O{}; O.use_loss = use_loss; O.huber_a = huber_a;
  long double c = 0;
  for (long o = 0; o < n_obs; ++o) { double r[2]; c += eval_obs(P, O, poses, intr, points, o, r, nullptr); }
  for (int k = 0; k < n_priors; ++k) { double r[3]; c += eval_prior(P, poses, k, r, nullptr); }
  return (double)c;
}
}  // extern "C"

}  // extern "C"
