// Host-only micro-benchmark of the dogleg controller's per-iteration algebra at the window solver's size (n = 171):
// a nonlinear least-squares toy (residuals r_k = a_k . x + 0.05 sin(b_k . x) - y_k) whose linearisation cost is negligible,
// so the time per iteration is the controller's dense algebra (Jacobi scaling, Cholesky, dogleg step, products).
// build: g++ -O3 -std=c++17 -fopenmp-simd -pthread -I lio_mapping_b200/csrc scripts/dogleg_bench.cc \
//        lio_mapping_b200/csrc/hostmath.cc lio_mapping_b200/csrc/solver_host.cc -o /tmp/dogleg_bench
#include <chrono>
#include <cstdio>
#include <random>
#include "solver_host.h"
using namespace lio;
using namespace lio::hm;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 171, m = 2 * n;
  std::mt19937_64 rng(3);
  std::normal_distribution<double> nd;
  Mat Aa(m, n), Bb(m, n);
  for (auto &v : Aa.d) v = nd(rng);
  for (auto &v : Bb.d) v = 0.3 * nd(rng);
  Vec xt(n), y(m);
  for (auto &v : xt) v = nd(rng);
  for (int k = 0; k < m; ++k) { double a = 0, b = 0; for (int j = 0; j < n; ++j) { a += Aa(k, j) * xt[j]; b += Bb(k, j) * xt[j]; } y[k] = a + 0.05 * std::sin(b); }
  // precomputed Gauss-Newton pieces of the linear part: the toy linearise only adds a cheap correction
  Mat AtA(n, n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < m; ++k) s += Aa(k, i) * Aa(k, j); AtA(i, j) = s; }
  double t_lin = 0;
  int evals = 0;
  Vec x(n, 0.0);
  DoglegProblem P;
  P.n = n;
  P.get_state = [&](Vec &o) { o = x; };
  P.set_state = [&](const Vec &v) { x = v; };
  P.plus = [&](const Vec &a, const Vec &d, Vec &o) { o = a; for (int i = 0; i < n; ++i) o[i] += d[i]; };
  P.linearize = [&](Mat &H, Vec &g, double &cost) {
    const double t0 = now();
    if (H.r != n) H = Mat(n, n);
    H.d = AtA.d;
    g.assign(n, 0.0);
    cost = 0;
    for (int k = 0; k < m; ++k) {
      double a = 0, b = 0;
      for (int j = 0; j < n; ++j) { a += Aa(k, j) * x[j]; b += Bb(k, j) * x[j]; }
      const double r = a + 0.05 * std::sin(b) - y[k], cb = 0.05 * std::cos(b);
      cost += 0.5 * r * r;
      for (int j = 0; j < n; ++j) g[j] += (Aa(k, j) + cb * Bb(k, j)) * r;
    }
    ++evals;
    t_lin += now() - t0;
    return true;
  };
  DoglegOptions opt;
  DoglegSummary sum;
  double best = 1e9;
  int its = 0;
  for (int rep = 0; rep < 30; ++rep) {
    std::fill(x.begin(), x.end(), 0.0);
    t_lin = 0;
    const double t0 = now();
    dogleg_solve(opt, P, &sum);
    const double t = now() - t0 - t_lin;
    if (t / sum.iterations < best) { best = t / sum.iterations; its = sum.iterations; }
  }
  printf("n=%d  iterations %d  controller algebra %.1f us per iteration (min over 30 solves)  final cost %.3g\n", n, its, best * 1e6, sum.final_cost);
  return 0;
}
