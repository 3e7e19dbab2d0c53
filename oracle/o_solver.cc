// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
// Ceres-1.14-equivalent dense trust-region (TRADITIONAL_DOGLEG) solver, see o_solver.h.
// Follows the published algorithm of ceres-solver 1.14.0 internal/ceres/trust_region_minimizer.cc,
// dogleg_strategy.cc, corrector.cc, loss_function.cc (library defaults unless set at
// Estimator.cc:1909-1921).  DENSE_SCHUR is an exact solve of the regularised normal equations, so
// one dense Cholesky of the tangent-space system reproduces its step up to round-off.
#include "o_solver.h"
#include <cmath>

namespace orc {

void Problem::AddParameterBlock(double *p, int size, bool pose) {
  if (index.count(p)) return;
  index[p] = (int)params.size();
  params.push_back({p, size, pose, false, -1, -1});
}
void Problem::SetParameterBlockConstant(double *p) { params[index.at(p)].constant = true; }
int Problem::AddResidualBlock(std::shared_ptr<CostFunction> cost, const CauchyLoss *loss, const std::vector<double *> &ps) {
  res.push_back({cost, loss, ps, false});
  return (int)res.size() - 1;
}

static double eval_block_cost(const Problem::ResBlock &rb) {
  int nr = rb.cost->num_residuals;
  VecX r(nr);
  rb.cost->Evaluate(rb.params.data(), r.data(), nullptr);
  double sq = 0;
  for (double v : r) sq += v * v;
  if (!rb.loss) return 0.5 * sq;
  double rho[3];
  rb.loss->Evaluate(sq, rho);
  return 0.5 * rho[0];
}

double Problem::EvaluateCost(const std::vector<int> *ids) const {
  double c = 0;
  if (ids) {
    for (int id : *ids) if (!res[id].removed) c += eval_block_cost(res[id]);
  } else {
    for (const ResBlock &rb : res) if (!rb.removed) c += eval_block_cost(rb);
  }
  return c;
}

int Problem::Prepare() {
  int t = 0, a = 0;
  for (ParamBlock &pb : params) {
    if (pb.constant) { pb.toff = pb.aoff = -1; continue; }
    pb.toff = t; pb.aoff = a;
    t += pb.pose ? 6 : pb.size;
    a += pb.size;
  }
  tangent_dim = t; ambient_dim = a;
  return t;
}

void Problem::Linearize(MatX *H, VecX *g, double *cost) const {
  if (H) { *H = MatX(tangent_dim, tangent_dim); }
  if (g) g->assign(tangent_dim, 0.0);
  double c = 0;
  std::vector<MatX> J;
  std::vector<double *> raw;
  for (const ResBlock &rb : res) {
    if (rb.removed) continue;
    bool any_free = false;
    for (double *p : rb.params) if (!params[index.at(p)].constant) any_free = true;
    if (!any_free) continue;  // goes to Ceres' fixed_cost
    const int nr = rb.cost->num_residuals;
    const std::vector<int> &bs = rb.cost->block_sizes;
    VecX r(nr);
    if (!H && !g) {
      c += eval_block_cost(rb);
      continue;
    }
    J.clear(); raw.resize(bs.size());
    for (size_t i = 0; i < bs.size(); ++i) J.push_back(MatX(nr, bs[i]));
    for (size_t i = 0; i < bs.size(); ++i) raw[i] = params[index.at(rb.params[i])].constant ? nullptr : J[i].d.data();
    rb.cost->Evaluate(rb.params.data(), r.data(), raw.data());
    double sq = 0;
    for (double v : r) sq += v * v;
    if (rb.loss) {  // ceres::internal::Corrector
      double rho[3];
      rb.loss->Evaluate(sq, rho);
      c += 0.5 * rho[0];
      const double sqrt_rho1 = std::sqrt(rho[1]);
      double residual_scaling, alpha_sq_norm;
      if ((sq == 0.0) || (rho[2] <= 0.0)) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
      else {
        const double D = 1.0 + 2.0 * sq * rho[2] / rho[1];
        const double alpha = 1.0 - std::sqrt(D);
        residual_scaling = sqrt_rho1 / (1 - alpha);
        alpha_sq_norm = alpha / sq;
      }
      for (size_t i = 0; i < bs.size(); ++i) {
        if (!raw[i]) continue;
        MatX &Ji = J[i];
        if (alpha_sq_norm == 0.0) { for (double &v : Ji.d) v *= sqrt_rho1; }
        else {
          for (int cc = 0; cc < Ji.c; ++cc) {
            double rtj = 0;
            for (int rr = 0; rr < nr; ++rr) rtj += r[rr] * Ji(rr, cc);
            for (int rr = 0; rr < nr; ++rr) Ji(rr, cc) = sqrt_rho1 * (Ji(rr, cc) - alpha_sq_norm * r[rr] * rtj);
          }
        }
      }
      for (double &v : r) v *= residual_scaling;
    } else {
      c += 0.5 * sq;
    }
    for (size_t i = 0; i < bs.size(); ++i) {
      if (!raw[i]) continue;
      const ParamBlock &pi = params[index.at(rb.params[i])];
      const int li = pi.pose ? 6 : pi.size;
      for (int a = 0; a < li; ++a) {
        double s = 0;
        for (int rr = 0; rr < nr; ++rr) s += J[i](rr, a) * r[rr];
        (*g)[pi.toff + a] += s;
      }
      if (!H) continue;
      for (size_t j = 0; j < bs.size(); ++j) {
        if (!raw[j]) continue;
        const ParamBlock &pj = params[index.at(rb.params[j])];
        const int lj = pj.pose ? 6 : pj.size;
        for (int a = 0; a < li; ++a)
          for (int b = 0; b < lj; ++b) {
            double s = 0;
            for (int rr = 0; rr < nr; ++rr) s += J[i](rr, a) * J[j](rr, b);
            (*H)(pi.toff + a, pj.toff + b) += s;
          }
      }
    }
  }
  if (cost) *cost = c;
}

void Problem::GetState(VecX &x) const {
  x.assign(ambient_dim, 0.0);
  for (const ParamBlock &pb : params) if (!pb.constant) for (int k = 0; k < pb.size; ++k) x[pb.aoff + k] = pb.ptr[k];
}
void Problem::SetState(const VecX &x) {
  for (ParamBlock &pb : params) if (!pb.constant) for (int k = 0; k < pb.size; ++k) pb.ptr[k] = x[pb.aoff + k];
}
void Problem::Plus(const VecX &x, const VecX &delta, VecX &out) const {
  out = x;
  for (const ParamBlock &pb : params) {
    if (pb.constant) continue;
    if (pb.pose) PosePlus(&x[pb.aoff], &delta[pb.toff], &out[pb.aoff]);
    else for (int k = 0; k < pb.size; ++k) out[pb.aoff + k] = x[pb.aoff + k] + delta[pb.toff + k];
  }
}

static double vnorm(const VecX &v) { double s = 0; for (double a : v) s += a * a; return std::sqrt(s); }
static double vdot(const VecX &a, const VecX &b) { double s = 0; for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i]; return s; }

void Solve(const SolverOptions &opt, Problem *problem, SolverSummary *sum) {
  Problem &P = *problem;
  const int n = P.Prepare();
  *sum = SolverSummary();
  if (n == 0) { sum->termination = 1; return; }
  VecX x, cand;
  P.GetState(x);
  MatX H;
  VecX g;
  double x_cost;
  P.Linearize(&H, &g, &x_cost);  // iteration zero
  sum->num_linearizations = 1;
  sum->initial_cost = x_cost;
  sum->cost_trace.push_back(x_cost);
  sum->H_initial = H;
  sum->g_initial = g;
  VecX scale(n, 1.0);
  if (opt.jacobi_scaling) for (int i = 0; i < n; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(H(i, i)));
  auto apply_scale = [&]() {
    for (int i = 0; i < n; ++i) { g[i] *= scale[i]; for (int j = 0; j < n; ++j) H(i, j) *= scale[i] * scale[j]; }
  };
  auto grad_max = [&]() { double m = 0; for (double v : g) m = std::max(m, std::fabs(v)); return m; };
  double gmax = grad_max();  // unscaled tangent-space gradient, checked before scaling
  apply_scale();
  double x_norm = vnorm(x);
  // DoglegStrategy state
  double radius = opt.initial_trust_region_radius;
  const double min_diagonal = 1e-6, max_diagonal = 1e32, min_mu = 1e-8, max_mu = 1.0, mu_increase_factor = 10.0;
  double mu = min_mu, dogleg_step_norm = 0.0, alpha = 0.0;
  bool reuse = false;
  VecX diagonal(n), gradient(n), gn(n), step(n);
  int num_consecutive_invalid_steps = 0;
  int iteration = 0;
  sum->termination = 0;
  if (gmax <= opt.gradient_tolerance) { sum->termination = 1; sum->final_cost = x_cost; return; }
  while (true) {
    if (iteration >= opt.max_num_iterations) { sum->termination = 0; break; }
    if (radius < opt.min_trust_region_radius) { sum->termination = 1; break; }
    ++iteration;
    sum->num_iterations = iteration;
    // ---- DoglegStrategy::ComputeStep
    bool linear_ok = true;
    if (!reuse) {
      reuse = true;
      for (int i = 0; i < n; ++i) diagonal[i] = std::sqrt(std::min(std::max(H(i, i), min_diagonal), max_diagonal));
      for (int i = 0; i < n; ++i) gradient[i] = g[i] / diagonal[i];
      {  // Cauchy point
        VecX sg(n);
        for (int i = 0; i < n; ++i) sg[i] = gradient[i] / diagonal[i];
        VecX Hsg = matvec(H, sg);
        double jg2 = vdot(sg, Hsg);
        alpha = vdot(gradient, gradient) / jg2;
      }
      // Gauss-Newton step with the mu * D^2 regularisation, mu raised on Cholesky failure
      linear_ok = false;
      while (mu < max_mu) {
        MatX A = H;
        for (int i = 0; i < n; ++i) A(i, i) += mu * diagonal[i] * diagonal[i];
        VecX rhs = g;
        bool ok = cholesky_lower(A);
        if (ok) {
          chol_solve(A, rhs);
          for (double v : rhs) if (!std::isfinite(v)) ok = false;
        }
        if (!ok) { mu *= mu_increase_factor; continue; }
        for (int i = 0; i < n; ++i) gn[i] = -diagonal[i] * rhs[i];
        linear_ok = true;
        break;
      }
    }
    bool step_is_valid = linear_ok;
    double model_cost_change = 0;
    if (linear_ok) {
      // ComputeTraditionalDoglegStep
      const double gradient_norm = vnorm(gradient), gauss_newton_norm = vnorm(gn);
      if (gauss_newton_norm <= radius) {
        step = gn; dogleg_step_norm = gauss_newton_norm;
      } else if (gradient_norm * alpha >= radius) {
        for (int i = 0; i < n; ++i) step[i] = -(radius / gradient_norm) * gradient[i];
        dogleg_step_norm = radius;
      } else {
        const double b_dot_a = -alpha * vdot(gradient, gn);
        const double a_squared_norm = std::pow(alpha * gradient_norm, 2.0);
        const double b_minus_a_squared_norm = a_squared_norm - 2 * b_dot_a + std::pow(gauss_newton_norm, 2);
        const double c = b_dot_a - a_squared_norm;
        const double d = std::sqrt(c * c + b_minus_a_squared_norm * (std::pow(radius, 2.0) - a_squared_norm));
        double beta = (c <= 0) ? (d - c) / b_minus_a_squared_norm : (radius * radius - a_squared_norm) / (d + c);
        for (int i = 0; i < n; ++i) step[i] = (-alpha * (1.0 - beta)) * gradient[i] + beta * gn[i];
        dogleg_step_norm = vnorm(step);
      }
      for (int i = 0; i < n; ++i) step[i] /= diagonal[i];
      // model_cost_change = -(J s)'(r + J s / 2) = -s'g - s'Hs/2  (scaled J)
      VecX Hs = matvec(H, step);
      model_cost_change = -vdot(step, g) - 0.5 * vdot(step, Hs);
      step_is_valid = model_cost_change > 0.0;
    }
    if (!step_is_valid) {
      if (++num_consecutive_invalid_steps >= opt.max_num_consecutive_invalid_steps) { sum->termination = 2; break; }
      mu *= mu_increase_factor;  // StepIsInvalid
      reuse = false;
      continue;
    }
    num_consecutive_invalid_steps = 0;
    VecX delta(n);
    for (int i = 0; i < n; ++i) delta[i] = step[i] * scale[i];
    P.Plus(x, delta, cand);
    P.SetState(cand);
    double cand_cost;
    P.Linearize(nullptr, nullptr, &cand_cost);
    ++sum->num_cost_evaluations;
    if (!std::isfinite(cand_cost)) cand_cost = std::numeric_limits<double>::max();
    // ParameterToleranceReached
    double step_norm = 0;
    for (int i = 0; i < (int)x.size(); ++i) step_norm += (x[i] - cand[i]) * (x[i] - cand[i]);
    step_norm = std::sqrt(step_norm);
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { P.SetState(x); sum->termination = 1; break; }
    // FunctionToleranceReached
    const double cost_change = x_cost - cand_cost;
    if (std::fabs(cost_change) <= opt.function_tolerance * x_cost) { P.SetState(x); sum->termination = 1; break; }
    const double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > opt.min_relative_decrease) {  // HandleSuccessfulStep
      x = cand;
      x_norm = vnorm(x);
      P.Linearize(&H, &g, &x_cost);
      ++sum->num_linearizations;
      ++sum->num_successful_steps;
      sum->cost_trace.push_back(x_cost);
      gmax = grad_max();
      apply_scale();
      if (relative_decrease < 0.25) radius *= 0.5;
      if (relative_decrease > 0.75) radius = std::min(opt.max_trust_region_radius, std::max(radius, 3.0 * dogleg_step_norm));
      mu = std::max(min_mu, 2.0 * mu / mu_increase_factor);
      reuse = false;
      if (gmax <= opt.gradient_tolerance) { sum->termination = 1; break; }
    } else {  // HandleUnsuccessfulStep
      P.SetState(x);
      radius *= 0.5;
      reuse = true;
    }
  }
  P.SetState(x);
  sum->final_cost = x_cost;
}

}  // namespace orc
