// lio_mapping_b200 — C-ABI seams of the host shell's dense fp64 kernels (hostmath.cc), so that the CPU test-suite can pin
// them against LAPACK without a device: the blocked Cholesky used by the dogleg step (Ceres DENSE_SCHUR equivalent,
// Estimator.cc:1909-1921) and the symmetric eigen-solver used by the marginalisation (Eigen::SelfAdjointEigenSolver call
// sites MarginalizationFactor.cc:276, :293).
#include <cstring>
#include "../../include/lio_b200.h"
#include "hostmath.h"
#include "solver_host.h"
#include <cmath>

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

// The dogleg controller (solver_host.cc, stand-in for ceres::Solve as configured at Estimator.cc:1909-1921) on a toy
// nonlinear least-squares problem assembled on the host:  r_k = a_k . x + amp sin(b_k . x) - y_k, optional CauchyLoss(1.0)
// with the Ceres corrector (rho'' <= 0: residual and Jacobian scaled by sqrt(rho')).  The normal equations are built
// exactly the way the window solver consumes them (H = J^T J, g = J^T r, cost = 1/2 sum rho).  summary = {iterations,
// successful steps, termination, initial cost, final cost, evaluations}.
extern "C" int lio_host_dogleg_toy(int n, int m, const double *A, const double *B, const double *y, double amp, int use_cauchy,
                                   double *x, int max_iter, double *summary) {
  if (n <= 0 || m <= 0 || !A || !B || !y || !x || !summary || max_iter < 0) return LIO_ERR_INVALID;
  using namespace lio;
  Vec state(x, x + n);
  DoglegProblem P;
  P.n = n;
  P.get_state = [&](Vec &o) { o = state; };
  P.set_state = [&](const Vec &v) { state = v; };
  P.plus = [&](const Vec &a, const Vec &d, Vec &o) { o = a; for (int i = 0; i < n; ++i) o[i] += d[i]; };
  P.linearize = [&](Mat &H, Vec &g, double &cost) {
    if (H.r != n) H = Mat(n, n); else H.zero();
    g.assign(n, 0.0);
    cost = 0.0;
    Vec J(n);
    for (int k = 0; k < m; ++k) {
      const double *a = A + (size_t)k * n, *b = B + (size_t)k * n;
      double s = 0, t = 0;
      for (int j = 0; j < n; ++j) { s += a[j] * state[j]; t += b[j] * state[j]; }
      double r = s + amp * std::sin(t) - y[k];
      const double cb = amp * std::cos(t);
      for (int j = 0; j < n; ++j) J[j] = a[j] + cb * b[j];
      if (use_cauchy) {
        const double sq = r * r, rho1 = 1.0 / (1.0 + sq);
        cost += 0.5 * std::log(1.0 + sq);
        const double w = std::sqrt(rho1);
        r *= w;
        for (int j = 0; j < n; ++j) J[j] *= w;
      } else {
        cost += 0.5 * r * r;
      }
      for (int i = 0; i < n; ++i) {
        const double ji = J[i];
        g[i] += ji * r;
        double *row = &H.d[(size_t)i * n];
        for (int j = 0; j < n; ++j) row[j] += ji * J[j];
      }
    }
    return std::isfinite(cost);
  };
  DoglegOptions opt;
  opt.max_num_iterations = max_iter;
  DoglegSummary sum;
  dogleg_solve(opt, P, &sum);
  for (int i = 0; i < n; ++i) x[i] = state[i];
  summary[0] = sum.iterations; summary[1] = sum.successful_steps; summary[2] = sum.termination;
  summary[3] = sum.initial_cost; summary[4] = sum.final_cost; summary[5] = sum.evaluations;
  return LIO_OK;
}
