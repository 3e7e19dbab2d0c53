// lio_mapping_b200 — C-ABI seams of the host shell's dense fp64 kernels (hostmath.cc), so that the CPU test-suite can pin
// them against LAPACK without a device: the blocked Cholesky used by the dogleg step (Ceres DENSE_SCHUR equivalent,
// Estimator.cc:1909-1921) and the symmetric eigen-solver used by the marginalisation (Eigen::SelfAdjointEigenSolver call
// sites MarginalizationFactor.cc:276, :293).
#include <cstring>
#include "../../include/lio_b200.h"
#include "hostmath.h"

void lio_set_last_error(const char *file, int line, const char *msg);  // capi_common.cu

using namespace lio::hm;

extern "C" int lio_host_cholesky_solve(int n, const double *A, const double *b, double *L_out, double *x) {
  if (n <= 0 || !A || !b || !x) return LIO_ERR_INVALID;
  Mat M(n, n);
  std::memcpy(M.d.data(), A, sizeof(double) * n * n);
  if (!cholesky(M)) {
    lio_set_last_error(__FILE__, __LINE__, "matrix is not positive definite");
    return LIO_ERR_NUMERIC;
  }
  Vec v(b, b + n);
  cholesky_solve(M, v);
  std::memcpy(x, v.data(), sizeof(double) * n);
  if (L_out)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) L_out[(size_t)i * n + j] = j <= i ? M(i, j) : 0.0;  // only the lower triangle is defined
  return LIO_OK;
}

extern "C" int lio_host_sym_eigen(int n, const double *A, double *evals, double *evecs, int threads) {
  if (n <= 0 || !A || !evals || !evecs) return LIO_ERR_INVALID;
  Mat M(n, n), Z;
  std::memcpy(M.d.data(), A, sizeof(double) * n * n);
  Vec d;
  sym_eigen(M, d, Z, threads < 1 ? 1 : threads);
  std::memcpy(evals, d.data(), sizeof(double) * n);
  std::memcpy(evecs, Z.d.data(), sizeof(double) * n * n);
  return LIO_OK;
}
