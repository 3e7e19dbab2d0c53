// lio_mapping_b200 — dense fp64 kernels of the host shell (n <= 15*(O+1)+6 <= 256).
// All inner loops are stride-1 (right-looking Cholesky on rows, column-major eigen-solver) and the hot
// functions are multiversioned (x86-64-v3 = AVX2+FMA when the CPU has it, baseline otherwise).
#include "hostmath.h"

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define LIO_MV __attribute__((target_clones("arch=x86-64-v3", "default")))
#else
#define LIO_MV
#endif

namespace lio {
namespace hm {

// Right-looking (outer-product) lower Cholesky: after step k the trailing lower triangle has the rank-1
// update  a[i][j] -= l_ik * l_jk  applied row by row (contiguous axpy, no reduction to reassociate).
LIO_MV bool cholesky(Mat &a) {
  const int n = a.r;
  std::vector<double> col(n);
  double *A = a.d.data();
  for (int k = 0; k < n; ++k) {
    const double s = A[(size_t)k * n + k];
    if (!(s > 0.0) || !std::isfinite(s)) return false;
    const double l = std::sqrt(s), il = 1.0 / l;
    A[(size_t)k * n + k] = l;
    for (int i = k + 1; i < n; ++i) {
      const double v = A[(size_t)i * n + k] * il;
      A[(size_t)i * n + k] = v;
      col[i] = v;
    }
    for (int i = k + 1; i < n; ++i) {
      double *ri = A + (size_t)i * n;
      const double lik = col[i];
      const double *c = col.data();
      for (int j = k + 1; j <= i; ++j) ri[j] -= lik * c[j];
    }
  }
  return true;
}

LIO_MV void cholesky_solve(const Mat &L, Vec &b) {
  const int n = L.r;
  const double *A = L.d.data();
  for (int i = 0; i < n; ++i) {
    const double *ri = A + (size_t)i * n;
    double s = 0;
#pragma omp simd reduction(+ : s)
    for (int k = 0; k < i; ++k) s += ri[k] * b[k];
    b[i] = (b[i] - s) / ri[i];
  }
  // back substitution in axpy form: x_i known -> subtract its column (= row i of L) from the rest
  for (int i = n - 1; i >= 0; --i) {
    const double *ri = A + (size_t)i * n;
    const double xi = b[i] / ri[i];
    b[i] = xi;
    for (int k = 0; k < i; ++k) b[k] -= ri[k] * xi;
  }
}

// Householder tridiagonalisation + implicit-shift QL with accumulated transformations (EISPACK
// tred2/tql2), on column-major storage: z(i,j) = Zt[j*n + i], so every inner loop walks a column.
LIO_MV void sym_eigen(const Mat &A, Vec &d, Mat &Zout) {
  const int n = A.r;
  d.assign(n, 0.0);
  Zout = Mat(n, n);
  if (n == 0) return;
  std::vector<double> Zt((size_t)n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Zt[(size_t)j * n + i] = A.d[(size_t)i * n + j];
  Vec e(n, 0.0);
  auto col = [&](int j) { return Zt.data() + (size_t)j * n; };
  for (int j = 0; j < n; ++j) d[j] = col(j)[n - 1];
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) { d[j] = col(j)[i - 1]; col(j)[i] = 0.0; col(i)[j] = 0.0; }
    } else {
      const double inv = 1.0 / scale;
      for (int k = 0; k < i; ++k) { d[k] *= inv; h += d[k] * d[k]; }
      double f = d[i - 1];
      double g = f > 0 ? -std::sqrt(h) : std::sqrt(h);
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      double *ci = col(i);
      for (int j = 0; j < i; ++j) {
        f = d[j];
        ci[j] = f;
        double *cj = col(j);
        double gg = 0.0;
        const double *dp = d.data();
        double *ep = e.data();
#pragma omp simd reduction(+ : gg)
        for (int k = j + 1; k < i; ++k) { gg += cj[k] * dp[k]; ep[k] += cj[k] * f; }
        e[j] = e[j] + cj[j] * f + gg;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j]; g = e[j];
        double *cj = col(j);
        const double *dp = d.data(), *ep = e.data();
        for (int k = j; k < i; ++k) cj[k] -= (f * ep[k] + g * dp[k]);
        d[j] = cj[i - 1];
        cj[i] = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; ++i) {
    col(i)[n - 1] = col(i)[i];
    col(i)[i] = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      const double *c1 = col(i + 1);
      for (int k = 0; k <= i; ++k) d[k] = c1[k] / h;
      for (int j = 0; j <= i; ++j) {
        double *cj = col(j);
        double g = 0.0;
#pragma omp simd reduction(+ : g)
        for (int k = 0; k <= i; ++k) g += c1[k] * cj[k];
        const double *dp = d.data();
        for (int k = 0; k <= i; ++k) cj[k] -= g * dp[k];
      }
    }
    double *c1 = col(i + 1);
    for (int k = 0; k <= i; ++k) c1[k] = 0.0;
  }
  for (int j = 0; j < n; ++j) { d[j] = col(j)[n - 1]; col(j)[n - 1] = 0.0; }
  col(n - 1)[n - 1] = 1.0;
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
          double *ca = col(i), *cb = col(i + 1);
          for (int k = 0; k < n; ++k) {
            const double hb = cb[k], ha = ca[k];
            cb[k] = s * ha + c * hb;
            ca[k] = c * ha - s * hb;
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
  std::vector<int> idx(n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return d[a] < d[b]; });
  Vec ds(n);
  for (int j = 0; j < n; ++j) {
    ds[j] = d[idx[j]];
    const double *c = col(idx[j]);
    for (int k = 0; k < n; ++k) Zout.d[(size_t)k * n + j] = c[k];
  }
  d.swap(ds);
}

LIO_MV void add_JtJ_mapped(const double *J, const double *r, int rows, int cols, const int *cm, Mat &H, Vec &g) {
  // contiguous fast path when the mapped columns are consecutive
  bool contiguous = true;
  for (int a = 1; a < cols; ++a) if (cm[a] != cm[0] + a) contiguous = false;
  const int n = H.c;
  double tmp[64];
  for (int a = 0; a < cols; ++a) {
    double gs = 0;
    for (int b = 0; b < cols; ++b) tmp[b] = 0.0;
    for (int k = 0; k < rows; ++k) {
      const double ja = J[k * cols + a];
      const double *jr = J + k * cols;
      for (int b = 0; b < cols; ++b) tmp[b] += ja * jr[b];
      gs += ja * r[k];
    }
    g[cm[a]] += gs;
    double *hrow = &H.d[(size_t)cm[a] * n];
    if (contiguous) { double *dst = hrow + cm[0]; for (int b = 0; b < cols; ++b) dst[b] += tmp[b]; }
    else for (int b = 0; b < cols; ++b) hrow[cm[b]] += tmp[b];
  }
}

LIO_MV void matvec(const Mat &A, const Vec &x, Vec &y) {
  y.assign(A.r, 0.0);
  for (int i = 0; i < A.r; ++i) {
    const double *row = &A.d[(size_t)i * A.c];
    double s = 0;
#pragma omp simd reduction(+ : s)
    for (int j = 0; j < A.c; ++j) s += row[j] * x[j];
    y[i] = s;
  }
}

// C = A * A^T restricted to the first `kc` columns of A scaled by w: C(r,c) = sum_k A(r,k) w[k] A(c,k)
LIO_MV void weighted_gram(const Mat &A, const Vec &w, const std::vector<int> &cols, Mat &C) {
  const int n = A.r, kc = (int)cols.size();
  Mat B(n, kc), Bw(n, kc);  // gathered columns, contiguous per row
  for (int r = 0; r < n; ++r)
    for (int k = 0; k < kc; ++k) { double v = A.d[(size_t)r * A.c + cols[k]]; B.d[(size_t)r * kc + k] = v; Bw.d[(size_t)r * kc + k] = v * w[cols[k]]; }
  C = Mat(n, n);
  for (int r = 0; r < n; ++r) {
    const double *br = &Bw.d[(size_t)r * kc];
    for (int c = r; c < n; ++c) {
      const double *bc = &B.d[(size_t)c * kc];
      double s = 0;
#pragma omp simd reduction(+ : s)
      for (int k = 0; k < kc; ++k) s += br[k] * bc[k];
      C.d[(size_t)r * n + c] = s;
      C.d[(size_t)c * n + r] = s;
    }
  }
}

}  // namespace hm
}  // namespace lio
