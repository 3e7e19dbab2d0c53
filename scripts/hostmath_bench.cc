// Micro-benchmark of the host shell's dense kernels at the solver's sizes (n = 15 (O+1) [+6]).
// build: g++ -O3 -std=c++17 -fopenmp-simd -pthread -I lio_mapping_b200/csrc scripts/hostmath_bench.cc lio_mapping_b200/csrc/hostmath.cc -o /tmp/hm_bench
#include <chrono>
#include <cstdio>
#include <random>
#include "hostmath.h"
using namespace lio::hm;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 171, reps = 200;
  std::mt19937_64 rng(1);
  std::normal_distribution<double> nd;
  Mat B(n, n), H(n, n);
  for (auto &v : B.d) v = nd(rng);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { double s = 0; for (int k = 0; k < n; ++k) s += B(i, k) * B(j, k); H(i, j) = s + (i == j ? n : 0); }
  Vec g(n, 1.0);
  double t0 = now();
  for (int r = 0; r < reps; ++r) { Mat A = H; volatile double x = A.d[5]; (void)x; }
  double t_copy = (now() - t0) / reps;
  t0 = now();
  double acc = 0;
  for (int r = 0; r < reps; ++r) { Mat A = H; cholesky(A); acc += A.d[7]; }
  double t_chol = (now() - t0) / reps - t_copy;
  Mat L = H; cholesky(L);
  t0 = now();
  for (int r = 0; r < reps; ++r) { Vec b = g; cholesky_solve(L, b); acc += b[3]; }
  double t_solve = (now() - t0) / reps;
  t0 = now();
  for (int r = 0; r < reps; ++r) { Vec y; matvec(H, g, y); acc += y[3]; }
  double t_mv = (now() - t0) / reps;
  t0 = now();
  for (int r = 0; r < reps; ++r) { Vec y = mul(H, g); acc += y[3]; }
  double t_mul = (now() - t0) / reps;
  Vec d; Mat Z;
  double t_eig = 0;
  for (int th : {4, 2, 1}) {
    t0 = now();
    for (int r = 0; r < 20; ++r) { sym_eigen(H, d, Z, th); acc += d[0]; }
    t_eig = (now() - t0) / 20;
    printf("sym_eigen n=%d threads=%d: %.1f us\n", n, th, t_eig * 1e6);
  }
  printf("n=%d copy %.1f us  chol %.1f us  trisolve %.1f us  matvec %.1f us  mul %.1f us  sym_eigen %.1f us  (%g)\n", n, t_copy * 1e6, t_chol * 1e6,
         t_solve * 1e6, t_mv * 1e6, t_mul * 1e6, t_eig * 1e6, acc);
  return 0;
}
