// lio_mapping_b200 — dense fp64 kernels of the host shell (n <= 15*(O+1)+6 <= 256).
#include "hostmath.h"

namespace lio {
namespace hm {

bool cholesky(Mat &a) {
  const int n = a.r;
  for (int j = 0; j < n; ++j) {
    double *rj = &a.d[(size_t)j * n];
    double s = rj[j];
    for (int k = 0; k < j; ++k) s -= rj[k] * rj[k];
    if (!(s > 0.0) || !std::isfinite(s)) return false;
    const double l = std::sqrt(s), il = 1.0 / l;
    rj[j] = l;
    for (int i = j + 1; i < n; ++i) {
      double *ri = &a.d[(size_t)i * n];
      double t = ri[j];
      for (int k = 0; k < j; ++k) t -= ri[k] * rj[k];
      ri[j] = t * il;
    }
  }
  return true;
}

void cholesky_solve(const Mat &L, Vec &b) {
  const int n = L.r;
  for (int i = 0; i < n; ++i) {
    const double *ri = &L.d[(size_t)i * n];
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= ri[k] * b[k];
    b[i] = s / ri[i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L.d[(size_t)k * n + i] * b[k];
    b[i] = s / L.d[(size_t)i * n + i];
  }
}

// Householder reduction to tridiagonal form followed by implicit-shift QL iterations with
// accumulated transformations (the classical EISPACK tred2/tql2 pair).  Z holds A on entry.
void sym_eigen(const Mat &A, Vec &d, Mat &Z) {
  const int n = A.r;
  Z = A;
  d.assign(n, 0.0);
  if (n == 0) return;
  Vec e(n, 0.0);
  auto z = [&](int i, int j) -> double & { return Z.d[(size_t)i * n + j]; };
  for (int j = 0; j < n; ++j) d[j] = z(n - 1, j);
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) { d[j] = z(i - 1, j); z(i, j) = 0.0; z(j, i) = 0.0; }
    } else {
      const double inv = 1.0 / scale;
      for (int k = 0; k < i; ++k) { d[k] *= inv; h += d[k] * d[k]; }
      double f = d[i - 1];
      double g = f > 0 ? -std::sqrt(h) : std::sqrt(h);
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j];
        z(j, i) = f;
        g = e[j] + z(j, j) * f;
        for (int k = j + 1; k < i; ++k) { g += z(k, j) * d[k]; e[k] += z(k, j) * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j]; g = e[j];
        for (int k = j; k < i; ++k) z(k, j) -= (f * e[k] + g * d[k]);
        d[j] = z(i - 1, j);
        z(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; ++i) {
    z(n - 1, i) = z(i, i);
    z(i, i) = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; ++k) d[k] = z(k, i + 1) / h;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        for (int k = 0; k <= i; ++k) g += z(k, i + 1) * z(k, j);
        for (int k = 0; k <= i; ++k) z(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) z(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) { d[j] = z(n - 1, j); z(n - 1, j) = 0.0; }
  z(n - 1, n - 1) = 1.0;
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n - 1 && std::fabs(e[m]) > eps * tst1) ++m;
    if (m > l) {
      for (int iter = 0; iter < 300; ++iter) {
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
        const double el1 = e[l + 1];
        for (int i = m - 1; i >= l; --i) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; ++k) {
            double *row = &Z.d[(size_t)k * n];
            h = row[i + 1];
            row[i + 1] = s * row[i] + c * h;
            row[i] = c * row[i] - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
        if (std::fabs(e[l]) <= eps * tst1) break;
      }
    }
    d[l] += f;
    e[l] = 0.0;
  }
  // ascending order
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return d[a] < d[b]; });
  Vec ds(n);
  Mat Zs(n, n);
  for (int j = 0; j < n; ++j) {
    ds[j] = d[idx[j]];
    for (int k = 0; k < n; ++k) Zs.d[(size_t)k * n + j] = Z.d[(size_t)k * n + idx[j]];
  }
  d.swap(ds);
  Z = Zs;
}

}  // namespace hm
}  // namespace lio
