// Build configuration for compiling the reference's vendored Ceres 1.13.0 with
// oracle/Makefile instead of the reference's CMake (which would generate this file).
// Mirrors openMVG's default internal-Ceres configuration
// (reference: src/third_party/ceres-solver/CMakeLists.txt:67-82 — SUITESPARSE/CXSPARSE/LAPACK
// OFF, EIGENSPARSE ON, OPENMP ON, SCHUR_SPECIALIZATIONS ON, MINIGLOG ON).
// TEST INFRASTRUCTURE ONLY (oracle/_ref); never part of the product path.
#ifndef CERES_PUBLIC_INTERNAL_CONFIG_H_
#define CERES_PUBLIC_INTERNAL_CONFIG_H_
#define CERES_USE_EIGEN_SPARSE
#define CERES_NO_LAPACK
#define CERES_NO_SUITESPARSE
#define CERES_NO_CXSPARSE
#define CERES_USE_OPENMP
#define CERES_HAVE_PTHREAD
#define CERES_HAVE_RWLOCK
#define CERES_STD_UNORDERED_MAP
#endif
