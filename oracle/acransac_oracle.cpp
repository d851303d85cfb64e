// acransac_oracle.cpp — CPU restatement of the per-pair a-contrario RANSAC openMVG runs after putative matching
// (geometric filtering, fundamental-matrix model).
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
//
// Parity status: PINNED by tests/test_oracle_geom.py against oracle/_ref/libref_geom.so — the reference's own
// ACRANSAC + ACKernelAdaptor<SevenPointSolver, EpipolarDistanceError, UnnormalizerT> compiled where it lies — on seeded
// pairs: identical inlier lists, identical (errorMax, minNFA) to 1e-9, F equal up to scale to 1e-7.
//
// Reference lines restated (/root/reference/src/openMVG unless noted):
//   robust_estimation/robust_estimator_ACRansac.hpp:57-119   logcombi tables (float arithmetic)
//   robust_estimation/robust_estimator_ACRansac.hpp:190-300  ComputeNFA_and_inliers (quantified: 20-bin histogram; exhaustive: sorted)
//   robust_estimation/robust_estimator_ACRansac.hpp:303-490  ACRANSAC main loop (early exit, focused sampling, reserve iterations)
//   robust_estimation/rand_sampling.hpp:35-95                UniformSample (rejection) / UniformSample (Fisher-Yates prefix)
//   robust_estimation/robust_estimator_ACRansacKernelAdaptator.hpp:43-63,120-200   logalpha0, multError, normalisation, unnormalise
//   multiview/conditioning.cpp:54-77                         PreconditionerFromPoints(w, h), NormalizePoints
//   multiview/solver_fundamental_kernel.cpp:38-95            SevenPointSolver::Solve (null space of A'A, cubic in alpha)
//   multiview/solver_fundamental_kernel.hpp:83-93            EncodeEpipolarEquation
//   multiview/solver_fundamental_kernel.cpp:157-166          EpipolarDistanceError
//   numeric/poly.h:32-96                                     SolveCubicPolynomial
//   third_party/histogram/histogram.hpp:44-113               Histogram::Add / GetXbinsValue (bin edges /nBins, bin values /(nBins-1))
// Third-party pieces the reference gets from its toolchain, restated from their published definitions:
//   std::mt19937 (seed 5489, MT19937 of Matsumoto & Nishimura) and libstdc++ 13's
//   std::uniform_int_distribution<uint32_t> on a 32-bit generator = Lemire's nearly divisionless method
//   (/usr/include/c++/13/bits/uniform_int_dist.h:251-283, 313-321);
//   Eigen 3.4 SelfAdjointEigenSolver<9x9> is replaced by a cyclic Jacobi eigen-solver: any orthonormal basis of the
//   two-dimensional null space gives the same pencil F1 + a F2, hence the same models up to scale.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

namespace {

struct MT19937 {
  uint32_t mt[624]; int idx;
  explicit MT19937(uint32_t seed = 5489u) { mt[0] = seed; for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i; idx = 624; }
  uint32_t next() {
    if (idx >= 624) {
      for (int i = 0; i < 624; ++i) {
        const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
  }
};

// uniform_int_distribution<uint32_t>(a, a + range - 1) - a on a 32-bit generator: _S_nd<uint64_t>
uint32_t lemire(MT19937 &g, uint32_t range) {
  uint64_t product = (uint64_t)g.next() * (uint64_t)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    const uint32_t threshold = (0u - range) % range;
    while (low < threshold) { product = (uint64_t)g.next() * (uint64_t)range; low = (uint32_t)product; }
  }
  return (uint32_t)(product >> 32);
}

// rand_sampling.hpp:35-58
void uniform_sample_reject(uint32_t num, uint32_t total, MT19937 &g, std::vector<uint32_t> &samples) {
  samples.resize(0);
  while (samples.size() < num) {
    const uint32_t s = lemire(g, total);                      // distribution(0, total - 1)
    bool found = false;
    for (size_t j = 0; j < samples.size() && !found; ++j) found = samples[j] == s;
    if (!found) samples.push_back(s);
  }
}
// rand_sampling.hpp:71-95
bool uniform_sample_shuffle(size_t num, MT19937 &g, std::vector<uint32_t> &vec_index, std::vector<uint32_t> &samples) {
  if (num > vec_index.size()) return false;
  const uint32_t last = (uint32_t)(vec_index.size() - 1);
  for (uint32_t i = 0; i < num; ++i) {
    const uint32_t s = i + lemire(g, last - i + 1);           // distribution(i, last)
    std::swap(vec_index[i], vec_index[s]);
  }
  samples.resize(num);
  for (size_t i = 0; i < num; ++i) samples[i] = vec_index[i];
  return true;
}

// numeric/poly.h:32-96
int solve_cubic(double a, double b, double c, double x[3]) {
  const double eps = std::numeric_limits<double>::epsilon();
  a /= 3;
  double p = (b - 3 * a * a) / 3;
  double q = (2 * a * a * a - a * b + c) / 2;
  double d = q * q + p * p * p;
  const double tolq = std::max(std::abs(2 * a * a * a), std::max(std::abs(a * b), std::abs(c)));
  const double tolp = std::max(std::abs(b), std::abs(3 * a * a));
  int n = (d > eps * std::max(p * p * tolp, std::abs(q) * tolq) ? 1 : 3);
  if (n == 1) {
    d = std::pow(std::abs(q) + std::sqrt(d), 1 / (double)3);
    x[0] = d - p / d;
    if (q > 0) x[0] = -x[0];
  } else {
    if (3 * p >= -eps * tolp) { n = 1; x[0] = 0; }
    else {
      p = std::sqrt(-p);
      q /= p * p * p;
      d = (q <= -1) ? M_PI : (q >= 1) ? 0 : std::acos(q);
      for (int i = 0; i < 3; ++i) x[i] = -2 * p * std::cos((d + 2 * M_PI * i) / 3);
    }
  }
  for (int i = 0; i < n; ++i) x[i] -= a;
  return n;
}
int solve_cubic_coeffs(const double P[4], double roots[3]) {
  if (P[0] == 0.0) return 0;
  return solve_cubic(P[2] / P[3], P[1] / P[3], P[0] / P[3], roots);
}

// eigenvectors of the two smallest eigenvalues of a symmetric 9x9 (cyclic Jacobi, threshold sweeps)
void smallest_two_eigvecs(double A[9][9], double f1[9], double f2[9]) {
  double V[9][9];
  for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < 9; ++i) { diag += A[i][i] * A[i][i]; for (int j = i + 1; j < 9; ++j) off += A[i][j] * A[i][j]; }
    if (off <= 1e-60 * diag || off == 0.0) break;
    for (int p = 0; p < 8; ++p)
      for (int q = p + 1; q < 9; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 9; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 9; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 9; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  int i1 = 0; for (int i = 1; i < 9; ++i) if (A[i][i] < A[i1][i1]) i1 = i;
  int i2 = i1 == 0 ? 1 : 0; for (int i = 0; i < 9; ++i) if (i != i1 && A[i][i] < A[i2][i2]) i2 = i;
  for (int k = 0; k < 9; ++k) { f1[k] = V[k][i1]; f2[k] = V[k][i2]; }
}

// model 0: fundamental matrix (7-point solver, point-to-line error, up to 3 models per sample; F_ACRobust.hpp:66-83)
// model 1: homography (4-point DLT, asymmetric transfer error, point-to-point; H_ACRobust.hpp:78-89)
struct Kernel {
  int n; std::vector<double> x1, x2;           // normalised points, interleaved
  double N1[9], N2[9], logalpha0, mult_error;
  int model, min_samples, max_models;
};

// conditioning.cpp:54-77 ; ACKernelAdaptator.hpp:120-140
void make_kernel(const double *xI, const double *xJ, int n, int wI, int hI, int wJ, int hJ, int model, Kernel &K) {
  K.n = n; K.x1.resize(2 * n); K.x2.resize(2 * n);
  K.model = model; K.min_samples = model == 0 ? 7 : 4; K.max_models = model == 0 ? 3 : 1; K.mult_error = model == 0 ? 0.5 : 1.0;
  auto norm = [](int w, int h, double T[9]) {
    const double d = 1.0 / std::sqrt((double)(w * h));
    for (int i = 0; i < 9; ++i) T[i] = 0; T[0] = T[4] = d; T[8] = 1.0;
    T[2] = (double)(-.5f * w) * d; T[5] = -.5 * h * d;
  };
  norm(wI, hI, K.N1); norm(wJ, hJ, K.N2);
  for (int i = 0; i < n; ++i) {
    K.x1[2 * i] = (K.N1[0] * xI[2 * i] + K.N1[2]) / 1.0; K.x1[2 * i + 1] = (K.N1[4] * xI[2 * i + 1] + K.N1[5]) / 1.0;
    K.x2[2 * i] = (K.N2[0] * xJ[2 * i] + K.N2[2]) / 1.0; K.x2[2 * i + 1] = (K.N2[4] * xJ[2 * i + 1] + K.N2[5]) / 1.0;
  }
  const double D = std::hypot((double)wJ, (double)hJ), A = wJ * (double)hJ;
  K.logalpha0 = model == 0 ? std::log10(2. * D / A / K.N2[0])                      // point to line (ACKernelAdaptator.hpp:47-56)
                           : std::log10(M_PI / (wJ * (double)hJ) / (K.N2[0] * K.N2[0]));   // point to point (:64-74)
}

// solver_homography_kernel.cpp:36-84: the right singular vector of the 8 x 9 DLT matrix for the smallest singular
// value = the eigenvector of L'L for the smallest eigenvalue (sign / scale are immaterial to the transfer error)
int four_point(const Kernel &K, const std::vector<uint32_t> &s, double H[3][9]) {
  double LtL[9][9];
  for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) LtL[i][j] = 0;
  for (int t = 0; t < 4; ++t) {
    const double x = K.x1[2 * s[t]], y = K.x1[2 * s[t] + 1], u = K.x2[2 * s[t]], v = K.x2[2 * s[t] + 1];
    const double r0[9] = {x, y, 1.0, 0, 0, 0, -u * x, -u * y, -u * 1.0};
    const double r1[9] = {0, 0, 0, x, y, 1.0, -v * x, -v * y, -v * 1.0};
    for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) LtL[i][j] += r0[i] * r0[j] + r1[i] * r1[j];
  }
  double f1[9], f2[9];
  smallest_two_eigvecs(LtL, f1, f2);
  for (int t = 0; t < 9; ++t) H[0][t] = f1[t];
  return 1;
}
// solver_homography_kernel.hpp:59-63
inline double transfer_error(const double H[9], double x, double y, double u, double v) {
  const double p0 = H[0] * x + H[1] * y + H[2], p1 = H[3] * x + H[4] * y + H[5], p2 = H[6] * x + H[7] * y + H[8];
  const double dx = u - p0 / p2, dy = v - p1 / p2;
  return dx * dx + dy * dy;
}

// solver_fundamental_kernel.cpp:38-95: up to 3 models (row-major 3x3)
int seven_point(const Kernel &K, const std::vector<uint32_t> &s, double F[3][9]) {
  double AtA[9][9];
  for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) AtA[i][j] = 0;
  for (int t = 0; t < 7; ++t) {
    const double x = K.x1[2 * s[t]], y = K.x1[2 * s[t] + 1], u = K.x2[2 * s[t]], v = K.x2[2 * s[t] + 1];
    const double r[9] = {u * x, u * y, u, v * x, v * y, v, x, y, 1.0};
    for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) AtA[i][j] += r[i] * r[j];
  }
  double f1[9], f2[9];
  smallest_two_eigvecs(AtA, f1, f2);
  const double a = f1[0], j = f2[0], b = f1[1], k = f2[1], c = f1[2], l = f2[2], d = f1[3], m = f2[3], e = f1[4], n = f2[4],
               f = f1[5], o = f2[5], g = f1[6], p = f2[6], h = f1[7], q = f2[7], i = f1[8], r = f2[8];
  const double P[4] = {
    a*e*i + b*f*g + c*d*h - a*f*h - b*d*i - c*e*g,
    a*e*r + a*i*n + b*f*p + b*g*o + c*d*q + c*h*m + d*h*l + e*i*j + f*g*k -
    a*f*q - a*h*o - b*d*r - b*i*m - c*e*p - c*g*n - d*i*k - e*g*l - f*h*j,
    a*n*r + b*o*p + c*m*q + d*l*q + e*j*r + f*k*p + g*k*o + h*l*m + i*j*n -
    a*o*q - b*m*r - c*n*p - d*k*r - e*l*p - f*j*q - g*l*n - h*j*o - i*k*m,
    j*n*r + k*o*p + l*m*q - j*o*q - k*m*r - l*n*p};
  double roots[3];
  const int nr = solve_cubic_coeffs(P, roots);
  for (int kk = 0; kk < nr; ++kk) for (int t = 0; t < 9; ++t) F[kk][t] = f1[t] + roots[kk] * f2[t];
  return nr;
}

// solver_fundamental_kernel.cpp:157-166
inline double epipolar_error(const double F[9], double x, double y, double u, double v) {
  const double fx0 = F[0] * x + F[1] * y + F[2], fx1 = F[3] * x + F[4] * y + F[5], fx2 = F[6] * x + F[7] * y + F[8];
  const double dt = fx0 * u + fx1 * v + fx2;
  return dt * dt / (fx0 * fx0 + fx1 * fx1);
}

struct NFA {
  const Kernel &K; std::vector<double> residuals; std::vector<float> logc_n, logc_k; double loge0, max_threshold; bool quantified;
  std::vector<std::pair<double, uint32_t>> sorted;
  static float logcombi(uint32_t k, uint32_t n, const std::vector<float> &l10) {
    if (k >= n) return 0.f;
    if (n - k < k) k = n - k;
    float r = 0.f;
    for (uint32_t i = 1; i <= k; ++i) r += l10[n - i + 1] - l10[i];
    return r;
  }
  NFA(const Kernel &k_, double maxthr, bool q) : K(k_), residuals(k_.n), max_threshold(maxthr), quantified(q) {
    const uint32_t n = (uint32_t)K.n;
    loge0 = std::log10((double)K.max_models * (double)(K.n - K.min_samples));   // MAX_MODELS * (NumSamples - MINIMUM_SAMPLES)
    std::vector<float> l10(n + 1);
    for (uint32_t i = 0; i <= n; ++i) l10[i] = (float)std::log10((double)(float)i);
    logc_n.resize(n + 1); logc_k.resize(n + 1);
    for (uint32_t k = 0; k <= n; ++k) logc_n[k] = logcombi(k, n, l10);
    for (uint32_t m = 0; m <= n; ++m) logc_k[m] = logcombi((uint32_t)K.min_samples, m, l10);
  }
  bool compute(std::vector<uint32_t> &inliers, std::pair<double, double> &nfa_threshold) {
    const double feps = std::numeric_limits<float>::epsilon();
    if (quantified) {
      const int nBins = 20;
      const double by_interval = nBins / (max_threshold - 0.0);
      size_t freq[20] = {0};
      for (int i = 0; i < K.n; ++i) {
        const double x = residuals[i];
        if (x < 0.0) continue;
        const size_t b = (size_t)((x - 0.0) * by_interval);
        if (b < (size_t)nBins) ++freq[b];
      }
      const double val = (max_threshold - 0.0) / (double)(nBins - 1);
      std::pair<double, double> best(std::numeric_limits<double>::infinity(), 0.0);
      unsigned int cum = 0;
      for (int bin = 0; bin < nBins; ++bin) {
        cum += (unsigned int)freq[bin];
        const double rv = val * (double)bin + 0.0;
        if (cum > (unsigned int)K.min_samples && rv > feps) {
          const double logalpha = K.logalpha0 + K.mult_error * std::log10(rv + feps);
          const std::pair<double, double> cur(loge0 + logalpha * (double)(cum - K.min_samples) + logc_n[cum] + logc_k[cum], rv);
          if (cur.first < best.first && cur.first < 0) best = cur;
        }
      }
      if (best.first < nfa_threshold.first) {
        nfa_threshold = best;
        inliers.clear();
        for (uint32_t i = 0; i < (uint32_t)K.n; ++i) if (residuals[i] <= nfa_threshold.second) inliers.push_back(i);
        return inliers.size() > (size_t)K.min_samples;
      }
    } else {
      sorted.clear();
      for (uint32_t i = 0; i < (uint32_t)K.n; ++i) sorted.emplace_back(residuals[i], i);
      std::sort(sorted.begin(), sorted.end());
      std::pair<double, uint32_t> best(std::numeric_limits<double>::infinity(), (uint32_t)K.min_samples);
      const size_t n = K.n;
      for (size_t k = K.min_samples + 1; k <= n && sorted[k - 1].first <= max_threshold; ++k) {
        const double logalpha = K.logalpha0 + K.mult_error * std::log10(sorted[k - 1].first + feps);
        const std::pair<double, uint32_t> cur(loge0 + logalpha * (double)(k - K.min_samples) + logc_n[k] + logc_k[k], (uint32_t)k);
        if (cur.first < best.first) best = cur;
      }
      if (best.first < nfa_threshold.first) {
        nfa_threshold.first = best.first; nfa_threshold.second = sorted[best.second - 1].first;
        inliers.resize(best.second);
        for (size_t i = 0; i < best.second; ++i) inliers[i] = sorted[i].second;
        return true;
      }
    }
    return false;
  }
};

}  // namespace

extern "C" {

// Same contract as ref_acransac_fundamental / ref_acransac_homography (oracle/ref_geom_driver.cpp).  trace (optional):
// {iterations run, models evaluated, iteration at which AC-RANSAC mode was entered or -1}.
static int acransac(int model, const double *xI, const double *xJ, int n, int wI, int hI, int wJ, int hJ, double precision, unsigned int iterations,
                    uint32_t *inliers_out, double *F_out, double *stats, int *trace) {
  std::vector<uint32_t> vec_inliers;
  double best_model[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  stats[0] = 0.0; stats[1] = 0.0;
  for (int i = 0; i < 9; ++i) F_out[i] = best_model[i];
  if (trace) { trace[0] = trace[1] = 0; trace[2] = -1; }
  Kernel K; make_kernel(xI, xJ, n, wI, hI, wJ, hJ, model, K);
  const unsigned int sizeSample = (unsigned int)K.min_samples, nData = (unsigned int)n;
  if (nData <= sizeSample) return 0;
  const double inf = std::numeric_limits<double>::infinity();
  const double prec2 = precision > 0 ? precision * precision : inf;
  std::vector<uint32_t> vec_index(nData); for (unsigned int i = 0; i < nData; ++i) vec_index[i] = i;
  std::vector<uint32_t> vec_sample(sizeSample);
  const double maxThreshold = prec2 == inf ? inf : prec2 * K.N2[0] * K.N2[0];
  NFA nfa(K, maxThreshold, prec2 != inf);
  double minNFA = inf, errorMax = inf;
  int nIterReserve = (int)(iterations / 10);
  unsigned int nIter = iterations - nIterReserve;
  bool bACRansacMode = prec2 == inf;
  MT19937 rng(5489u);
  unsigned int iter = 0; int n_models = 0;
  for (iter = 0; iter < nIter && iter < iterations; ++iter) {
    if (bACRansacMode) uniform_sample_shuffle(sizeSample, rng, vec_index, vec_sample);
    else uniform_sample_reject(sizeSample, nData, rng, vec_sample);
    double Fs[3][9];
    const int nm = model == 0 ? seven_point(K, vec_sample, Fs) : four_point(K, vec_sample, Fs);
    bool better = false;
    for (int mi = 0; mi < nm; ++mi) {
      ++n_models;
      for (unsigned int i = 0; i < nData; ++i)
        nfa.residuals[i] = model == 0 ? epipolar_error(Fs[mi], K.x1[2 * i], K.x1[2 * i + 1], K.x2[2 * i], K.x2[2 * i + 1])
                                      : transfer_error(Fs[mi], K.x1[2 * i], K.x1[2 * i + 1], K.x2[2 * i], K.x2[2 * i + 1]);
      if (!bACRansacMode) {
        unsigned int nInlier = 0;
        for (unsigned int i = 0; i < nData; ++i) if (nfa.residuals[i] <= maxThreshold) ++nInlier;
        if (nInlier > 2.5 * sizeSample) { bACRansacMode = true; if (trace) trace[2] = (int)iter; }
      }
      if (bACRansacMode) {
        std::pair<double, double> nfa_threshold(minNFA, 0.0);
        if (nfa.compute(vec_inliers, nfa_threshold)) {
          better = true; minNFA = nfa_threshold.first; errorMax = nfa_threshold.second;
          std::memcpy(best_model, Fs[mi], sizeof best_model);
        }
      }
    }
    if (!bACRansacMode && (int)iter > nIterReserve * 2) { nIter = 0; continue; }
    if (bACRansacMode && ((better && minNFA < 0) || ((iter + 1) == nIter && nIterReserve > 0))) {
      if (vec_inliers.empty()) { ++nIter; --nIterReserve; }
      else {
        vec_index = vec_inliers;
        if (nIterReserve) { nIter = iter + 1 + nIterReserve; nIterReserve = 0; }
      }
    }
  }
  if (trace) { trace[0] = (int)iter; trace[1] = n_models; }
  if (minNFA >= 0) vec_inliers.clear();
  if (!vec_inliers.empty()) {
    double T[9], U[9];
    if (model == 0) {      // UnnormalizerT: F = N2' * F * N1
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double v = 0; for (int k = 0; k < 3; ++k) v += K.N2[3 * k + r] * best_model[3 * k + c]; T[3 * r + c] = v; }
    } else {               // UnnormalizerI: H = N2^-1 * H * N1 ; N2 = [[d,0,a],[0,d,b],[0,0,1]] -> N2^-1 = [[1/d,0,-a/d],[0,1/d,-b/d],[0,0,1]]
      const double d = K.N2[0], a2 = K.N2[2], b2 = K.N2[5];
      const double I2[9] = {1.0 / d, 0, -a2 / d, 0, 1.0 / d, -b2 / d, 0, 0, 1.0};
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double v = 0; for (int k = 0; k < 3; ++k) v += I2[3 * r + k] * best_model[3 * k + c]; T[3 * r + c] = v; }
    }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double v = 0; for (int k = 0; k < 3; ++k) v += T[3 * r + k] * K.N1[3 * k + c]; U[3 * r + c] = v; }
    std::memcpy(best_model, U, sizeof best_model);
    errorMax = std::sqrt(errorMax) / K.N2[0];
  }
  for (size_t i = 0; i < vec_inliers.size(); ++i) inliers_out[i] = vec_inliers[i];
  for (int i = 0; i < 9; ++i) F_out[i] = best_model[i];
  stats[0] = errorMax; stats[1] = minNFA;
  return (int)vec_inliers.size();
}

int oracle_acransac_fundamental(const double *xI, const double *xJ, int n, int wI, int hI, int wJ, int hJ, double precision, unsigned int iterations,
                                uint32_t *inliers_out, double *F_out, double *stats, int *trace) {
  return acransac(0, xI, xJ, n, wI, hI, wJ, hJ, precision, iterations, inliers_out, F_out, stats, trace);
}
int oracle_acransac_homography(const double *xI, const double *xJ, int n, int wI, int hI, int wJ, int hJ, double precision, unsigned int iterations,
                               uint32_t *inliers_out, double *H_out, double *stats, int *trace) {
  return acransac(1, xI, xJ, n, wI, hI, wJ, hJ, precision, iterations, inliers_out, H_out, stats, trace);
}

}  // extern "C"
