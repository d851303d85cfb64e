// geom.cu — per-pair a-contrario RANSAC (fundamental-matrix model) on sm_100a: the geometric filter openMVG runs
// right after putative matching.
//
// Replaces (reference, /root/reference/src/openMVG):
//   matching_image_collection/GeometricFilter.hpp:66-128          the pair loop (OpenMP over pairs)
//   matching_image_collection/F_ACRobust.hpp:45-106               GeometricFilter_FMatrix_AC::Robust_estimation
//   robust_estimation/robust_estimator_ACRansac.hpp:57-119,190-300,303-490   logcombi, NFA, ACRANSAC
//   robust_estimation/rand_sampling.hpp:35-95                     the two UniformSample variants
//   robust_estimation/robust_estimator_ACRansacKernelAdaptator.hpp:43-63,120-200
//   multiview/conditioning.cpp:54-77, solver_fundamental_kernel.cpp:38-95,157-166, numeric/poly.h:32-96
//
// AC-RANSAC is sequential per pair (the sampling set of iteration k is the inlier set of the best model so far, the
// std::mt19937 stream is consumed in order), so the parallelism is ACROSS pairs — one CTA per pair, thousands of
// pairs per launch — and, inside a pair, across the matches (residuals, the 20-bin NFA histogram, the ordered
// inlier compaction).  The control flow, the generator (MT19937 + libstdc++'s Lemire down-scaling) and every
// floating-point expression that feeds a comparison follow the CPU restatement used by the tests (acransac_oracle.cpp) line by line; this file is
// compiled with -fmad=false so those expressions round exactly as the CPU evaluates them.
#include "common.cuh"

#include <cfloat>
#include <cmath>
#include <vector>

namespace omvg { namespace geom {

constexpr int THREADS = 128;
constexpr int NBINS = 20;

struct Rng { uint32_t *mt; int idx; };        // state in shared memory, driven by thread 0
__device__ void rng_seed(Rng &g, uint32_t seed) { g.mt[0] = seed; for (int i = 1; i < 624; ++i) g.mt[i] = 1812433253u * (g.mt[i - 1] ^ (g.mt[i - 1] >> 30)) + (uint32_t)i; g.idx = 624; }
__device__ uint32_t rng_next(Rng &g) {
  if (g.idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (g.mt[i] & 0x80000000u) | (g.mt[(i + 1) % 624] & 0x7fffffffu);
      g.mt[i] = g.mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g.idx = 0;
  }
  uint32_t y = g.mt[g.idx++];
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}
// std::uniform_int_distribution<uint32_t> of libstdc++ 13 on a 32-bit generator (Lemire), value - a
__device__ uint32_t lemire(Rng &g, uint32_t range) {
  unsigned long long product = (unsigned long long)rng_next(g) * (unsigned long long)range;
  uint32_t low = (uint32_t)product;
  if (low < range) {
    const uint32_t threshold = (0u - range) % range;
    while (low < threshold) { product = (unsigned long long)rng_next(g) * (unsigned long long)range; low = (uint32_t)product; }
  }
  return (uint32_t)(product >> 32);
}

// numeric/poly.h:32-96
__device__ int solve_cubic(double a, double b, double c, double x[3]) {
  const double eps = DBL_EPSILON;
  a /= 3;
  double p = (b - 3 * a * a) / 3;
  double q = (2 * a * a * a - a * b + c) / 2;
  double d = q * q + p * p * p;
  const double tolq = fmax(fabs(2 * a * a * a), fmax(fabs(a * b), fabs(c)));
  const double tolp = fmax(fabs(b), fabs(3 * a * a));
  int n = (d > eps * fmax(p * p * tolp, fabs(q) * tolq) ? 1 : 3);
  if (n == 1) {
    d = pow(fabs(q) + sqrt(d), 1 / (double)3);
    x[0] = d - p / d;
    if (q > 0) x[0] = -x[0];
  } else {
    if (3 * p >= -eps * tolp) { n = 1; x[0] = 0; }
    else {
      p = sqrt(-p);
      q /= p * p * p;
      d = (q <= -1) ? 3.14159265358979323846 : (q >= 1) ? 0 : acos(q);
      for (int i = 0; i < 3; ++i) x[i] = -2 * p * cos((d + 2 * 3.14159265358979323846 * i) / 3);
    }
  }
  for (int i = 0; i < n; ++i) x[i] -= a;
  return n;
}

// cyclic Jacobi on a symmetric 9x9 in shared memory (thread 0), eigenvectors of the two smallest eigenvalues
__device__ void smallest_two_eigvecs(double (*A)[9], double (*V)[9], double f1[9], double f2[9]) {
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
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 9; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
        for (int k = 0; k < 9; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
        for (int k = 0; k < 9; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
      }
  }
  int i1 = 0; for (int i = 1; i < 9; ++i) if (A[i][i] < A[i1][i1]) i1 = i;
  int i2 = i1 == 0 ? 1 : 0; for (int i = 0; i < 9; ++i) if (i != i1 && A[i][i] < A[i2][i2]) i2 = i;
  for (int k = 0; k < 9; ++k) { f1[k] = V[k][i1]; f2[k] = V[k][i2]; }
}

__device__ __forceinline__ double epipolar_error(const double *F, double x, double y, double u, double v) {
  const double fx0 = F[0] * x + F[1] * y + F[2], fx1 = F[3] * x + F[4] * y + F[5], fx2 = F[6] * x + F[7] * y + F[8];
  const double dt = fx0 * u + fx1 * v + fx2;
  return dt * dt / (fx0 * fx0 + fx1 * fx1);
}

// solver_homography_kernel.hpp:59-63
__device__ __forceinline__ double transfer_error(const double *H, double x, double y, double u, double v) {
  const double p0 = H[0] * x + H[1] * y + H[2], p1 = H[3] * x + H[4] * y + H[5], p2 = H[6] * x + H[7] * y + H[8];
  const double dx = u - p0 / p2, dy = v - p1 / p2;
  return dx * dx + dy * dy;
}
__device__ __forceinline__ double model_error(int model, const double *M, double x, double y, double u, double v) {
  return model == 0 ? epipolar_error(M, x, y, u, v) : transfer_error(M, x, y, u, v);
}

struct Args {
  const unsigned long long *offsets;       // [n_pairs + 1] into the match arrays
  const double *xI, *xJ;                   // [n_matches][2] pixels
  const int *image_size;                   // [n_pairs][4] wI hI wJ hJ
  double precision; unsigned int iterations; int model;   // model 0: fundamental (7 points, <= 3 models), 1: homography (4 points, 1 model)
  double *x1, *x2;                         // scratch [n_matches][2]: normalised points
  float *logc_n, *logc_k, *l10;            // scratch [n_matches + n_pairs] each (n + 1 entries per pair)
  uint32_t *vec_index;                     // scratch [n_matches]
  uint32_t *inliers;                       // out [n_matches]: first n_inliers[p] entries of pair p's segment
  uint32_t *n_inliers; double *F; double *stats;
  unsigned int n_pairs;
};

struct Smem {
  uint32_t mt[624];
  double AtA[9][9], V[9][9], f1[9], f2[9], models[3][9];
  double N1[9], N2[9];
  int hist[NBINS]; int n_le; int warp_cnt[THREADS / 32];
  uint32_t sample[7];
  int n_models, better_now;
  double nfa_thr, nfa_val;                 // threshold / NFA of the model being accepted
};

// rand_sampling.hpp:35-58 / 71-95 (thread 0)
__device__ void sample_reject(Rng &g, int ms, uint32_t total, uint32_t *s) {
  int cnt = 0;
  while (cnt < ms) {
    const uint32_t v = lemire(g, total);
    bool found = false;
    for (int j = 0; j < cnt && !found; ++j) found = s[j] == v;
    if (!found) s[cnt++] = v;
  }
}
__device__ void sample_shuffle(Rng &g, int ms, uint32_t *vec_index, uint32_t size, uint32_t *s) {
  if ((uint32_t)ms > size) return;                             // UniformSample returns false and leaves the sample as it was
  const uint32_t last = size - 1;
  for (uint32_t i = 0; i < (uint32_t)ms; ++i) {
    const uint32_t k = i + lemire(g, last - i + 1);
    const uint32_t a = vec_index[i], b = vec_index[k]; vec_index[i] = b; vec_index[k] = a;
  }
  for (int i = 0; i < ms; ++i) s[i] = vec_index[i];
}

__global__ void __launch_bounds__(THREADS) acransac_f_kernel(Args A) {
  __shared__ Smem S;
  const double INF = __longlong_as_double(0x7ff0000000000000ll);
  for (unsigned int pair = blockIdx.x; pair < A.n_pairs; pair += gridDim.x) {
    __syncthreads();
    const unsigned long long off = A.offsets[pair];
    const unsigned int nData = (unsigned int)(A.offsets[pair + 1] - off);
    const double *xI = A.xI + 2 * off, *xJ = A.xJ + 2 * off;
    double *x1 = A.x1 + 2 * off, *x2 = A.x2 + 2 * off;
    uint32_t *vec_index = A.vec_index + off, *inl = A.inliers + off;
    float *logc_n = A.logc_n + off + pair, *logc_k = A.logc_k + off + pair, *l10 = A.l10 + off + pair;
    if (threadIdx.x == 0) {
      A.n_inliers[pair] = 0; A.stats[2 * pair] = 0.0; A.stats[2 * pair + 1] = 0.0;
      for (int i = 0; i < 9; ++i) A.F[9 * (size_t)pair + i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
    const int ms = A.model == 0 ? 7 : 4;                          // MINIMUM_SAMPLES; MAX_MODELS = 3 / 1
    if (nData <= (unsigned int)ms) continue;                      // ACRANSAC: nData <= sizeSample -> {0, 0}
    // ---- kernel adaptor: normalisation (conditioning.cpp:54-77), logalpha0
    const int wI = A.image_size[4 * pair], hI = A.image_size[4 * pair + 1], wJ = A.image_size[4 * pair + 2], hJ = A.image_size[4 * pair + 3];
    if (threadIdx.x == 0) {
      const double d1 = 1.0 / sqrt((double)(wI * hI)), d2 = 1.0 / sqrt((double)(wJ * hJ));
      for (int i = 0; i < 9; ++i) { S.N1[i] = 0; S.N2[i] = 0; }
      S.N1[0] = S.N1[4] = d1; S.N1[8] = 1.0; S.N1[2] = (double)(-.5f * wI) * d1; S.N1[5] = -.5 * hI * d1;
      S.N2[0] = S.N2[4] = d2; S.N2[8] = 1.0; S.N2[2] = (double)(-.5f * wJ) * d2; S.N2[5] = -.5 * hJ * d2;
    }
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < nData; i += THREADS) {
      x1[2 * i] = (S.N1[0] * xI[2 * i] + S.N1[2]) / 1.0; x1[2 * i + 1] = (S.N1[4] * xI[2 * i + 1] + S.N1[5]) / 1.0;
      x2[2 * i] = (S.N2[0] * xJ[2 * i] + S.N2[2]) / 1.0; x2[2 * i + 1] = (S.N2[4] * xJ[2 * i + 1] + S.N2[5]) / 1.0;
      vec_index[i] = i;
    }
    const double logalpha0 = A.model == 0 ? log10(2. * hypot((double)wJ, (double)hJ) / (wJ * (double)hJ) / S.N2[0])       // point to line
                                          : log10(3.14159265358979323846 / (wJ * (double)hJ) / (S.N2[0] * S.N2[0]));   // point to point
    const double mult_error = A.model == 0 ? 0.5 : 1.0;
    const double maxThreshold = A.precision * A.precision * S.N2[0] * S.N2[0];
    // ---- NFA tables (robust_estimator_ACRansac.hpp:57-119), float arithmetic in the reference's order
    for (unsigned int i = threadIdx.x; i <= nData; i += THREADS) l10[i] = (float)log10((double)(float)i);
    __syncthreads();
    for (unsigned int k = threadIdx.x; k <= nData; k += THREADS) {
      { unsigned int kk = k; float r = 0.f;                         // logcombi(k, n)
        if (kk < nData) { if (nData - kk < kk) kk = nData - kk; for (unsigned int i = 1; i <= kk; ++i) r += l10[nData - i + 1] - l10[i]; }
        logc_n[k] = r; }
      { unsigned int kk = (unsigned int)ms; const unsigned int m = k; float r = 0.f;   // logcombi(MINIMUM_SAMPLES, m)
        if (kk < m) { if (m - kk < kk) kk = m - kk; for (unsigned int i = 1; i <= kk; ++i) r += l10[m - i + 1] - l10[i]; }
        logc_k[k] = r; }
    }
    const double loge0 = log10((double)(A.model == 0 ? 3 : 1) * (double)(nData - ms));
    __syncthreads();
    // ---- ACRANSAC main loop (robust_estimator_ACRansac.hpp:330-480); the scalars live in every thread (uniform)
    Rng g{S.mt, 624};
    if (threadIdx.x == 0) rng_seed(g, 5489u);
    double minNFA = INF, errorMax = INF;
    int nIterReserve = (int)(A.iterations / 10);
    unsigned int nIter = A.iterations - nIterReserve;
    bool bACRansacMode = false;                                    // a finite precision is required (see the C entry point)
    unsigned int n_inl = 0, index_size = nData;
    for (unsigned int iter = 0; iter < nIter && iter < A.iterations; ++iter) {
      __syncthreads();
      if (threadIdx.x == 0) {
        if (bACRansacMode) sample_shuffle(g, ms, vec_index, index_size, S.sample); else sample_reject(g, ms, nData, S.sample);
      }
      __syncthreads();
      // A'A of the 7 epipolar constraints, summed in sample order
      if (threadIdx.x < 81) {
        const int i = threadIdx.x / 9, j = threadIdx.x % 9;
        double acc = 0;
        for (int t = 0; t < ms; ++t) {
          const unsigned int s = S.sample[t];
          const double x = x1[2 * s], y = x1[2 * s + 1], u = x2[2 * s], v = x2[2 * s + 1];
          if (A.model == 0) {                                     // EncodeEpipolarEquation (solver_fundamental_kernel.hpp:83-93)
            const double r[9] = {u * x, u * y, u, v * x, v * y, v, x, y, 1.0};
            acc += r[i] * r[j];
          } else {                                                // BuildActionMatrix (solver_homography_kernel.cpp:38-58): two rows per point
            const double r0[9] = {x, y, 1.0, 0, 0, 0, -u * x, -u * y, -u * 1.0};
            const double r1[9] = {0, 0, 0, x, y, 1.0, -v * x, -v * y, -v * 1.0};
            acc += r0[i] * r0[j] + r1[i] * r1[j];
          }
        }
        S.AtA[i][j] = acc;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        smallest_two_eigvecs(S.AtA, S.V, S.f1, S.f2);
        if (A.model != 0) { for (int t = 0; t < 9; ++t) S.models[0][t] = S.f1[t]; S.n_models = 1; }
        else {
        const double a = S.f1[0], j = S.f2[0], b = S.f1[1], k = S.f2[1], c = S.f1[2], l = S.f2[2], d = S.f1[3], m = S.f2[3], e = S.f1[4], n = S.f2[4],
                     f = S.f1[5], o = S.f2[5], gg = S.f1[6], p = S.f2[6], h = S.f1[7], q = S.f2[7], i = S.f1[8], r = S.f2[8];
        const double P[4] = {
          a*e*i + b*f*gg + c*d*h - a*f*h - b*d*i - c*e*gg,
          a*e*r + a*i*n + b*f*p + b*gg*o + c*d*q + c*h*m + d*h*l + e*i*j + f*gg*k -
          a*f*q - a*h*o - b*d*r - b*i*m - c*e*p - c*gg*n - d*i*k - e*gg*l - f*h*j,
          a*n*r + b*o*p + c*m*q + d*l*q + e*j*r + f*k*p + gg*k*o + h*l*m + i*j*n -
          a*o*q - b*m*r - c*n*p - d*k*r - e*l*p - f*j*q - gg*l*n - h*j*o - i*k*m,
          j*n*r + k*o*p + l*m*q - j*o*q - k*m*r - l*n*p};
        double roots[3]; int nr = 0;
        if (P[0] != 0.0) nr = solve_cubic(P[2] / P[3], P[1] / P[3], P[0] / P[3], roots);
        for (int kk = 0; kk < nr; ++kk) for (int t = 0; t < 9; ++t) S.models[kk][t] = S.f1[t] + roots[kk] * S.f2[t];
        S.n_models = nr;
        }
      }
      __syncthreads();
      const int n_models = S.n_models;
      bool better = false;
      for (int mi = 0; mi < n_models; ++mi) {
        // residuals -> 20-bin histogram (histogram.hpp:65-77) and the max-consensus count
        if (threadIdx.x < NBINS) S.hist[threadIdx.x] = 0;
        if (threadIdx.x == 0) { S.n_le = 0; S.better_now = 0; }
        __syncthreads();
        const double by_interval = NBINS / (maxThreshold - 0.0);
        int le = 0;
        for (unsigned int i = threadIdx.x; i < nData; i += THREADS) {
          const double e = model_error(A.model, S.models[mi], x1[2 * i], x1[2 * i + 1], x2[2 * i], x2[2 * i + 1]);
          if (e <= maxThreshold) ++le;
          if (e >= 0.0) { const unsigned long long b = (unsigned long long)((e - 0.0) * by_interval); if (b < (unsigned long long)NBINS) atomicAdd(&S.hist[(int)b], 1); }
        }
        if (!bACRansacMode) { for (int o = 16; o > 0; o >>= 1) le += __shfl_xor_sync(0xffffffffu, le, o); if ((threadIdx.x & 31) == 0) atomicAdd(&S.n_le, le); }
        __syncthreads();
        if (!bACRansacMode && (double)S.n_le > 2.5 * ms) bACRansacMode = true;       // (uniform: every thread reads the same count)
        if (bACRansacMode) {
          if (threadIdx.x == 0) {
            const double feps = (double)FLT_EPSILON;
            const double val = (maxThreshold - 0.0) / (double)(NBINS - 1);
            double best_nfa = INF, best_thr = 0.0;
            unsigned int cum = 0;
            for (int bin = 0; bin < NBINS; ++bin) {
              cum += (unsigned int)S.hist[bin];
              const double rv = val * (double)bin + 0.0;
              if (cum > (unsigned int)ms && rv > feps) {
                const double logalpha = logalpha0 + mult_error * log10(rv + feps);
                const double cur = loge0 + logalpha * (double)(cum - ms) + logc_n[cum] + logc_k[cum];
                if (cur < best_nfa && cur < 0) { best_nfa = cur; best_thr = rv; }
              }
            }
            if (best_nfa < minNFA) { S.better_now = 1; S.nfa_thr = best_thr; S.nfa_val = best_nfa; }
          }
          __syncthreads();
          if (S.better_now) {
            // inliers = { i : residual <= threshold } in ascending order (ordered block compaction)
            const double thr = S.nfa_thr;
            unsigned int base = 0;
            for (unsigned int i0 = 0; i0 < nData; i0 += THREADS) {
              const unsigned int i = i0 + threadIdx.x;
              bool in = false;
              if (i < nData) in = model_error(A.model, S.models[mi], x1[2 * i], x1[2 * i + 1], x2[2 * i], x2[2 * i + 1]) <= thr;
              const unsigned int bal = __ballot_sync(0xffffffffu, in);
              if ((threadIdx.x & 31) == 0) S.warp_cnt[threadIdx.x >> 5] = __popc(bal);
              __syncthreads();
              unsigned int pre = base;
              for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) pre += S.warp_cnt[w];
              if (in) inl[pre + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u))] = i;
              unsigned int tot = 0; for (int w = 0; w < THREADS / 32; ++w) tot += S.warp_cnt[w];
              base += tot;
              __syncthreads();
            }
            n_inl = base;
            // (the inlier list is overwritten even when the support is too small: ComputeNFA_and_inliers returns false
            //  then and the score is NOT updated, robust_estimator_ACRansac.hpp:243-256)
            if (n_inl > (unsigned int)ms) {
              better = true; minNFA = S.nfa_val; errorMax = thr;
              if (threadIdx.x < 9) A.F[9 * (size_t)pair + threadIdx.x] = S.models[mi][threadIdx.x];
            }
          }
        }
        __syncthreads();
      }
      if (!bACRansacMode && (int)iter > nIterReserve * 2) { nIter = 0; continue; }
      if (bACRansacMode && ((better && minNFA < 0) || ((iter + 1) == nIter && nIterReserve > 0))) {
        if (n_inl == 0) { ++nIter; --nIterReserve; }
        else {
          for (unsigned int i = threadIdx.x; i < n_inl; i += THREADS) vec_index[i] = inl[i];
          index_size = n_inl;
          if (nIterReserve) { nIter = iter + 1 + nIterReserve; nIterReserve = 0; }
        }
      }
    }
    __syncthreads();
    if (minNFA >= 0) n_inl = 0;
    if (threadIdx.x == 0) {
      A.n_inliers[pair] = n_inl;
      if (n_inl > 0) {
        double M[9], T[9], U[9];
        for (int i = 0; i < 9; ++i) M[i] = A.F[9 * (size_t)pair + i];
        if (A.model == 0) {      // UnnormalizerT: N2' * F * N1
          for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double v = 0; for (int k = 0; k < 3; ++k) v += S.N2[3 * k + r] * M[3 * k + c]; T[3 * r + c] = v; }
        } else {                 // UnnormalizerI: N2^-1 * H * N1
          const double d = S.N2[0], a2 = S.N2[2], b2 = S.N2[5];
          const double I2[9] = {1.0 / d, 0, -a2 / d, 0, 1.0 / d, -b2 / d, 0, 0, 1.0};
          for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double v = 0; for (int k = 0; k < 3; ++k) v += I2[3 * r + k] * M[3 * k + c]; T[3 * r + c] = v; }
        }
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { double v = 0; for (int k = 0; k < 3; ++k) v += T[3 * r + k] * S.N1[3 * k + c]; U[3 * r + c] = v; }
        for (int i = 0; i < 9; ++i) A.F[9 * (size_t)pair + i] = U[i];
        errorMax = sqrt(errorMax) / S.N2[0];
      }
      A.stats[2 * pair] = errorMax; A.stats[2 * pair + 1] = minNFA;
    }
  }
}

}}  // namespace omvg::geom

extern "C" int omvg_geom_acransac(int device, int32_t model, uint64_t n_pairs, const uint64_t *offsets, const double *xI, const double *xJ,
                                  const int32_t *image_size, double precision, uint32_t max_iterations,
                                  uint32_t *inliers, uint32_t *n_inliers, double *F, double *stats) {
  using namespace omvg;
  if (model != OMVG_GEOM_FUNDAMENTAL && model != OMVG_GEOM_HOMOGRAPHY) return fail(OMVG_E_UNSUPPORTED, "geometric model %d is not implemented on the GPU path (0 = fundamental, 1 = homography)", model);
  if (n_pairs && (!offsets || !image_size || !n_inliers || !F || !stats)) return fail(OMVG_E_ARG, "null argument");
  if (!(precision > 0.0) || !std::isfinite(precision)) return fail(OMVG_E_UNSUPPORTED, "the GPU filter needs a finite upper bound of the precision (main_GeometricFilter uses 4.0)");
  if (max_iterations < 1) return fail(OMVG_E_ARG, "max_iterations must be >= 1");
  int n = 0; OMVG_CUDA(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return fail(OMVG_E_CUDA, "no CUDA device %d (found %d)", device, n);
  int cc_major = 0, n_sms = 0;
  OMVG_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
  OMVG_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
  if (cc_major != 10) return fail(OMVG_E_CUDA, "device %d is not sm_100; this library is sm_100a only", device);
  if (!n_pairs) return OMVG_OK;
  for (uint64_t p = 0; p < n_pairs; ++p) {
    if (offsets[p + 1] < offsets[p]) return fail(OMVG_E_ARG, "offsets must be non-decreasing");
    for (int k = 0; k < 4; ++k) if (image_size[4 * p + k] < 1 || image_size[4 * p + k] > 46340) return fail(OMVG_E_ARG, "pair %llu: bad image size", (unsigned long long)p);
  }
  const uint64_t nm = offsets[n_pairs];
  if (nm && (!xI || !xJ || !inliers)) return fail(OMVG_E_ARG, "null match arrays");
  if (nm >= (1ull << 31)) return fail(OMVG_E_UNSUPPORTED, "more than 2^31 putative matches in one call");
  OMVG_CUDA(cudaSetDevice(device));
  cudaStream_t st; OMVG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct Bufs { std::vector<void *> p; cudaStream_t s; ~Bufs() { for (void *q : p) pool_free(q); cudaStreamDestroy(s); } } B{{}, st};
  auto dalloc = [&](size_t bytes) -> void * { void *q = nullptr; if (pool_malloc(&q, std::max<size_t>(bytes, 8)) != cudaSuccess) return nullptr; B.p.push_back(q); return q; };
  geom::Args A{};
  unsigned long long *d_off = (unsigned long long *)dalloc((n_pairs + 1) * 8);
  double *d_xI = (double *)dalloc(nm * 16), *d_xJ = (double *)dalloc(nm * 16), *d_x1 = (double *)dalloc(nm * 16), *d_x2 = (double *)dalloc(nm * 16);
  int *d_sz = (int *)dalloc(n_pairs * 16);
  float *d_ln = (float *)dalloc((nm + n_pairs) * 4), *d_lk = (float *)dalloc((nm + n_pairs) * 4), *d_l10 = (float *)dalloc((nm + n_pairs) * 4);
  uint32_t *d_idx = (uint32_t *)dalloc(nm * 4), *d_inl = (uint32_t *)dalloc(nm * 4), *d_ninl = (uint32_t *)dalloc(n_pairs * 4);
  double *d_F = (double *)dalloc(n_pairs * 72), *d_stats = (double *)dalloc(n_pairs * 16);
  if (!d_off || !d_xI || !d_xJ || !d_x1 || !d_x2 || !d_sz || !d_ln || !d_lk || !d_l10 || !d_idx || !d_inl || !d_ninl || !d_F || !d_stats) return fail(OMVG_E_CUDA, "out of device memory");
  OMVG_CUDA(cudaMemcpyAsync(d_off, offsets, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
  OMVG_CUDA(cudaMemcpyAsync(d_sz, image_size, n_pairs * 16, cudaMemcpyHostToDevice, st));
  if (nm) { OMVG_CUDA(cudaMemcpyAsync(d_xI, xI, nm * 16, cudaMemcpyHostToDevice, st)); OMVG_CUDA(cudaMemcpyAsync(d_xJ, xJ, nm * 16, cudaMemcpyHostToDevice, st)); }
  A.model = model;
  A.offsets = d_off; A.xI = d_xI; A.xJ = d_xJ; A.image_size = d_sz; A.precision = precision; A.iterations = max_iterations;
  A.x1 = d_x1; A.x2 = d_x2; A.logc_n = d_ln; A.logc_k = d_lk; A.l10 = d_l10; A.vec_index = d_idx; A.inliers = d_inl; A.n_inliers = d_ninl; A.F = d_F; A.stats = d_stats;
  A.n_pairs = (unsigned int)n_pairs;
  int per_sm = 1; OMVG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, geom::acransac_f_kernel, geom::THREADS, 0));
  const unsigned int grid = (unsigned int)std::min<uint64_t>(n_pairs, (uint64_t)n_sms * std::max(1, per_sm));
  geom::acransac_f_kernel<<<grid, geom::THREADS, 0, st>>>(A);
  OMVG_CUDA(cudaGetLastError());
  if (nm) OMVG_CUDA(cudaMemcpyAsync(inliers, d_inl, nm * 4, cudaMemcpyDeviceToHost, st));
  OMVG_CUDA(cudaMemcpyAsync(n_inliers, d_ninl, n_pairs * 4, cudaMemcpyDeviceToHost, st));
  OMVG_CUDA(cudaMemcpyAsync(F, d_F, n_pairs * 72, cudaMemcpyDeviceToHost, st));
  OMVG_CUDA(cudaMemcpyAsync(stats, d_stats, n_pairs * 16, cudaMemcpyDeviceToHost, st));
  OMVG_CUDA(cudaStreamSynchronize(st));
  return OMVG_OK;
}

extern "C" int omvg_geom_fundamental_acransac(int device, uint64_t n_pairs, const uint64_t *offsets, const double *xI, const double *xJ,
                                              const int32_t *image_size, double precision, uint32_t max_iterations,
                                              uint32_t *inliers, uint32_t *n_inliers, double *F, double *stats) {
  return omvg_geom_acransac(device, OMVG_GEOM_FUNDAMENTAL, n_pairs, offsets, xI, xJ, image_size, precision, max_iterations, inliers, n_inliers, F, stats);
}
