// lio_mapping_b200 — trust-region / traditional-dogleg controller (see solver_host.h).
// Algorithm: Ceres-Solver 1.14.0 TrustRegionMinimizer + DoglegStrategy(TRADITIONAL) with Jacobi
// scaling fixed at iteration zero, expressed on the normal equations H = J^T J, g = J^T r (the
// DENSE_SCHUR solve is exact, so a dense Cholesky of H + mu D^2 gives the same Gauss-Newton step).
// One device evaluation per iteration: the candidate evaluation already carries H and g, which are
// reused when the step is accepted.
#include "solver_host.h"
#include <cstring>

namespace lio {
using namespace hm;

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("arch=x86-64-v3", "default")))
#endif
void dogleg_solve(const DoglegOptions &opt, DoglegProblem &P, DoglegSummary *sum) {
  const int n = P.n;
  *sum = DoglegSummary();
  if (n == 0) { sum->termination = 1; return; }
  Vec x, cand;
  P.get_state(x);
  Mat H(n, n), Hc(n, n), A(n, n);
  Vec g(n), gc(n), rhs(n);
  double x_cost = 0;
  if (!P.linearize(H, g, x_cost)) { sum->termination = 2; return; }
  sum->evaluations = 1;
  sum->initial_cost = sum->final_cost = x_cost;
  Vec scale(n);
  for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(H(i, i)));
  auto grad_max = [](const Vec &v) { double m = 0; for (double a : v) m = std::max(m, std::fabs(a)); return m; };
  auto apply_scale = [&](Mat &A, Vec &b) {
    for (int i = 0; i < n; ++i) {
      b[i] *= scale[i];
      double *row = &A.d[(size_t)i * n];
      const double si = scale[i];
      for (int j = 0; j < n; ++j) row[j] *= si * scale[j];
    }
  };
  double gmax = grad_max(g);
  apply_scale(H, g);
  double x_norm = vnorm(x);
  double radius = opt.initial_radius;
  const double min_diagonal = 1e-6, max_diagonal = 1e32, min_mu = 1e-8, max_mu = 1.0, mu_factor = 10.0;
  double mu = min_mu, dogleg_step_norm = 0.0, alpha = 0.0;
  bool reuse = false;
  Vec diagonal(n), gradient(n), gn(n), step(n), delta(n);
  int invalid = 0, iteration = 0;
  if (gmax <= opt.gradient_tolerance) { sum->termination = 1; return; }
  while (true) {
    if (iteration >= opt.max_num_iterations) { sum->termination = 0; break; }
    if (radius < opt.min_radius) { sum->termination = 1; break; }
    ++iteration;
    sum->iterations = iteration;
    bool linear_ok = true;
    if (!reuse) {
      reuse = true;
      for (int i = 0; i < n; ++i) diagonal[i] = std::sqrt(std::min(std::max(H(i, i), min_diagonal), max_diagonal));
      for (int i = 0; i < n; ++i) gradient[i] = g[i] / diagonal[i];
      {
        Vec sg(n);
        for (int i = 0; i < n; ++i) sg[i] = gradient[i] / diagonal[i];
        Vec Hsg;
        matvec(H, sg, Hsg);
        alpha = vdot(gradient, gradient) / vdot(sg, Hsg);
      }
      linear_ok = false;
      while (mu < max_mu) {
        std::memcpy(A.d.data(), H.d.data(), sizeof(double) * (size_t)n * n);   // workspace reused across iterations
        for (int i = 0; i < n; ++i) A(i, i) += mu * diagonal[i] * diagonal[i];
        rhs = g;
        bool ok = cholesky(A);
        if (ok) {
          cholesky_solve(A, rhs);
          for (double v : rhs) if (!std::isfinite(v)) ok = false;
        }
        if (!ok) { mu *= mu_factor; continue; }
        for (int i = 0; i < n; ++i) gn[i] = -diagonal[i] * rhs[i];
        linear_ok = true;
        break;
      }
    }
    bool step_valid = linear_ok;
    double model_cost_change = 0;
    if (linear_ok) {
      const double gradient_norm = vnorm(gradient), gn_norm = vnorm(gn);
      if (gn_norm <= radius) {
        step = gn; dogleg_step_norm = gn_norm;
      } else if (gradient_norm * alpha >= radius) {
        for (int i = 0; i < n; ++i) step[i] = -(radius / gradient_norm) * gradient[i];
        dogleg_step_norm = radius;
      } else {
        const double b_dot_a = -alpha * vdot(gradient, gn);
        const double a2 = std::pow(alpha * gradient_norm, 2.0);
        const double bma2 = a2 - 2 * b_dot_a + std::pow(gn_norm, 2);
        const double c = b_dot_a - a2;
        const double d = std::sqrt(c * c + bma2 * (std::pow(radius, 2.0) - a2));
        const double beta = (c <= 0) ? (d - c) / bma2 : (radius * radius - a2) / (d + c);
        for (int i = 0; i < n; ++i) step[i] = (-alpha * (1.0 - beta)) * gradient[i] + beta * gn[i];
        dogleg_step_norm = vnorm(step);
      }
      for (int i = 0; i < n; ++i) step[i] /= diagonal[i];
      Vec Hs;
      matvec(H, step, Hs);
      model_cost_change = -vdot(step, g) - 0.5 * vdot(step, Hs);
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {
      if (++invalid >= opt.max_consecutive_invalid_steps) { sum->termination = 2; break; }
      mu *= mu_factor;
      reuse = false;
      continue;
    }
    invalid = 0;
    for (int i = 0; i < n; ++i) delta[i] = step[i] * scale[i];
    P.plus(x, delta, cand);
    P.set_state(cand);
    double cand_cost = 0;
    bool ok = P.linearize(Hc, gc, cand_cost);
    ++sum->evaluations;
    if (!ok || !std::isfinite(cand_cost)) cand_cost = 1e300;
    double step_norm = 0;
    for (size_t i = 0; i < x.size(); ++i) step_norm += (x[i] - cand[i]) * (x[i] - cand[i]);
    step_norm = std::sqrt(step_norm);
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { sum->termination = 1; break; }
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= opt.function_tolerance * x_cost) { sum->termination = 1; break; }
    const double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > opt.min_relative_decrease) {
      x = cand;
      x_norm = vnorm(x);
      x_cost = cand_cost;
      H.d.swap(Hc.d);
      g.swap(gc);
      ++sum->successful_steps;
      gmax = grad_max(g);
      apply_scale(H, g);
      if (relative_decrease < 0.25) radius *= 0.5;
      if (relative_decrease > 0.75) radius = std::min(opt.max_radius, std::max(radius, 3.0 * dogleg_step_norm));
      mu = std::max(min_mu, 2.0 * mu / mu_factor);
      reuse = false;
      if (gmax <= opt.gradient_tolerance) { sum->termination = 1; break; }
    } else {
      radius *= 0.5;
      reuse = true;
    }
  }
  P.set_state(x);
  sum->final_cost = x_cost;
}

}  // namespace lio
