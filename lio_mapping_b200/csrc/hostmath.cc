// lio_mapping_b200 — dense fp64 kernels of the host shell (n <= 15*(O+1)+6 <= 256).
// All inner loops are stride-1 (right-looking Cholesky on rows, column-major eigen-solver) and the hot
// functions are multiversioned (x86-64-v3 = AVX2+FMA when the CPU has it, baseline otherwise).
#include "hostmath.h"
#include <thread>

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define LIO_MV __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
#else
#define LIO_MV
#endif

namespace lio {
namespace hm {

// Unblocked right-looking (outer-product) lower Cholesky on the square [k0, k1) diagonal block: the rank-1 update
// a[i][j] -= l_ik * l_jk is applied row by row (contiguous axpy, no reduction to reassociate).
static inline bool chol_diag(double *A, int n, int k0, int k1) {
  double col[64];
  for (int k = k0; k < k1; ++k) {
    const double s = A[(size_t)k * n + k];
    if (!(s > 0.0) || !std::isfinite(s)) return false;
    const double l = std::sqrt(s), il = 1.0 / l;
    A[(size_t)k * n + k] = l;
    for (int i = k + 1; i < k1; ++i) {
      const double v = A[(size_t)i * n + k] * il;
      A[(size_t)i * n + k] = v;
      col[i - k0] = v;
    }
    for (int i = k + 1; i < k1; ++i) {
      double *ri = A + (size_t)i * n;
      const double lik = col[i - k0];
      for (int j = k + 1; j <= i; ++j) ri[j] -= lik * col[j - k0];
    }
  }
  return true;
}

// Trailing update of the blocked Cholesky: C[i][j] -= sum_k P[i][k] P[j][k] for i0 <= j <= i < n, with the panel P (rows
// i0.., columns kb..kb+nb of A) read row-wise for the left factor and through its transposed copy PT (nb x m) for the
// right one.  Register tile: 4 rows x (2 vectors) columns; tiles that straddle the diagonal also write above it.
#define LIO_TRAIL_BODY(VT, VTU, VW)                                                                                     \
  for (int i = i0; i < n; i += 4) {                                                                                    \
    const int rows = std::min(4, n - i), jmax = i + rows - 1;                                                          \
    /* a short last row group re-reads its last valid row: the duplicates are computed and dropped */                  \
    const double *p0 = A + (size_t)i * n + kb, *p1 = A + (size_t)std::min(i + 1, n - 1) * n + kb,                        \
                 *p2 = A + (size_t)std::min(i + 2, n - 1) * n + kb, *p3 = A + (size_t)std::min(i + 3, n - 1) * n + kb;   \
    for (int jb = i0; jb <= jmax; jb += 2 * VW) {                                                                      \
      const int w = std::min(2 * VW, n - jb);                                                                          \
      VT c00 = {}, c01 = {}, c10 = {}, c11 = {}, c20 = {}, c21 = {}, c30 = {}, c31 = {};                                 \
      const double *pt = PT + (jb - i0);   /* rows of PT are padded to ldpt: lanes past m are finite and dropped */     \
      for (int k = 0; k < nb; ++k, pt += ldpt) {                                                                       \
        const VT b0 = *(const VTU *)pt, b1 = *(const VTU *)(pt + VW);                                                  \
        const double a0 = p0[k], a1 = p1[k], a2 = p2[k], a3 = p3[k];                                                   \
        c00 += a0 * b0; c01 += a0 * b1;                                                                                \
        c10 += a1 * b0; c11 += a1 * b1;                                                                                \
        c20 += a2 * b0; c21 += a2 * b1;                                                                                \
        c30 += a3 * b0; c31 += a3 * b1;                                                                                \
      }                                                                                                                \
      if (rows == 4 && w == 2 * VW) {                                                                                  \
        double *r0 = A + (size_t)i * n + jb, *r1 = r0 + n, *r2 = r1 + n, *r3 = r2 + n;                                  \
        *(VTU *)r0 -= c00; *(VTU *)(r0 + VW) -= c01;                                                                   \
        *(VTU *)r1 -= c10; *(VTU *)(r1 + VW) -= c11;                                                                   \
        *(VTU *)r2 -= c20; *(VTU *)(r2 + VW) -= c21;                                                                   \
        *(VTU *)r3 -= c30; *(VTU *)(r3 + VW) -= c31;                                                                   \
      } else {                                                                                                         \
        double tmp[4][2 * VW];                                                                                         \
        *(VTU *)&tmp[0][0] = c00; *(VTU *)&tmp[0][VW] = c01; *(VTU *)&tmp[1][0] = c10; *(VTU *)&tmp[1][VW] = c11;       \
        *(VTU *)&tmp[2][0] = c20; *(VTU *)&tmp[2][VW] = c21; *(VTU *)&tmp[3][0] = c30; *(VTU *)&tmp[3][VW] = c31;       \
        for (int r = 0; r < rows; ++r) {                                                                               \
          double *cr = A + (size_t)(i + r) * n + jb;                                                                   \
          for (int cidx = 0; cidx < w; ++cidx) cr[cidx] -= tmp[r][cidx];                                               \
        }                                                                                                              \
      }                                                                                                                \
    }                                                                                                                  \
  }

typedef double v4d __attribute__((vector_size(32)));
typedef double v4du __attribute__((vector_size(32), aligned(8)));
typedef double v8d __attribute__((vector_size(64)));
typedef double v8du __attribute__((vector_size(64), aligned(8)));

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target("arch=x86-64-v4"))) static void trailing_v4(double *A, int n, int kb, int nb, int i0, const double *PT, int ldpt) {
  LIO_TRAIL_BODY(v8d, v8du, 8)
}
__attribute__((target("arch=x86-64-v3"))) static void trailing_v3(double *A, int n, int kb, int nb, int i0, const double *PT, int ldpt) {
  LIO_TRAIL_BODY(v4d, v4du, 4)
}
#endif
static void trailing_base(double *A, int n, int kb, int nb, int i0, const double *PT, int ldpt) {
  LIO_TRAIL_BODY(v4d, v4du, 4)
}
typedef void (*trailing_fn)(double *, int, int, int, int, const double *, int);
static trailing_fn pick_trailing() {
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq") &&
      __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512cd"))
    return trailing_v4;
  if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return trailing_v3;
#endif
  return trailing_base;
}
static const trailing_fn trailing = pick_trailing();

// Blocked lower Cholesky (block width kCholNb): diagonal block, panel triangular solve, then the trailing update
// C -= P P^T through a 4 x 8 register tile over the transposed panel, so the trailing matrix is read and written once
// per block instead of once per column.  Only the lower triangle of the result is meaningful (tiles that straddle the
// diagonal also write above it).
constexpr int kCholNb = 24;

LIO_MV bool cholesky(Mat &a) {
  const int n = a.r;
  double *A = a.d.data();
  if (n <= 2 * kCholNb) return chol_diag(A, n, 0, n);
  const int ldpt = ((n + 31) / 32) * 32 + 32;   // padded row length of the transposed panel (vector tiles over-read / over-write the pad)
  std::vector<double> PTbuf((size_t)kCholNb * ldpt + 16, 0.0);
  double *PT = PTbuf.data();
  for (int kb = 0; kb < n; kb += kCholNb) {
    const int nb = std::min(kCholNb, n - kb), i0 = kb + nb, m = n - i0;
    if (!chol_diag(A, n, kb, i0)) return false;
    if (m <= 0) break;
    // panel X L_kk^T = A[i0:, kb:i0], solved on the transposed copy: L_kk PT = B^T is a forward substitution whose
    // updates are axpys over contiguous rows of PT (no short reductions).  Column chunks of kPc keep the nb x kPc working
    // set in L1 and the accumulators of row j in registers across its j updates.
    for (int i = i0; i < n; ++i) {
      const double *x = A + (size_t)i * n + kb;
      for (int k = 0; k < nb; ++k) PT[(size_t)k * ldpt + (i - i0)] = x[k];
    }
    {
      constexpr int kPc = 32;
      double inv[kCholNb];
      for (int j = 0; j < nb; ++j) inv[j] = 1.0 / A[(size_t)(kb + j) * n + kb + j];
      for (int c0 = 0; c0 < m; c0 += kPc) {   // the padded tail of PT (finite values) is processed along and ignored
        for (int j = 0; j < nb; ++j) {
          const double *lj = A + (size_t)(kb + j) * n + kb;
          double *pj = PT + (size_t)j * ldpt + c0;
          double acc[kPc];
#pragma omp simd
          for (int c2 = 0; c2 < kPc; ++c2) acc[c2] = pj[c2];
          for (int t = 0; t < j; ++t) {
            const double l0 = lj[t];
            const double *q0 = PT + (size_t)t * ldpt + c0;
#pragma omp simd
            for (int c2 = 0; c2 < kPc; ++c2) acc[c2] -= l0 * q0[c2];
          }
          const double iv = inv[j];
#pragma omp simd
          for (int c2 = 0; c2 < kPc; ++c2) pj[c2] = acc[c2] * iv;
        }
      }
    }
    for (int i = i0; i < n; ++i) {
      double *x = A + (size_t)i * n + kb;
      for (int k = 0; k < nb; ++k) x[k] = PT[(size_t)k * ldpt + (i - i0)];
    }
    trailing(A, n, kb, nb, i0, PT, ldpt);
  }
  return true;
}

LIO_MV void cholesky_solve(const Mat &L, Vec &b) {
  const int n = L.r;
  const double *A = L.d.data();
  for (int i = 0; i < n; ++i) {
    const double *ri = A + (size_t)i * n;
    const double *bp = b.data();
    // four interleaved partial sums: the reduction is not one latency chain
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    const int q = i / 4;
#pragma omp simd reduction(+ : s0, s1, s2, s3)
    for (int k = 0; k < q; ++k) {
      s0 += ri[k] * bp[k]; s1 += ri[q + k] * bp[q + k]; s2 += ri[2 * q + k] * bp[2 * q + k]; s3 += ri[3 * q + k] * bp[3 * q + k];
    }
    double s = (s0 + s1) + (s2 + s3);
    for (int k = 4 * q; k < i; ++k) s += ri[k] * bp[k];
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
// Applies the recorded QL plane rotations to the rows [k0, k1) of Z (column-major: Zt[j * n + k]).  Rows are independent,
// so the range is cut into chunks of KC rows that stay in L1 while ALL sweeps run over them; within a sweep the column
// shared by two successive rotations is carried in registers (one load + one store per element and rotation).
LIO_MV static void ql_apply_rows(double *Zt, int n, const double *rc, const double *rs, const std::vector<int> &sw_l,
                                 const std::vector<int> &sw_m, const std::vector<long long> &sw_off, int k0, int k1) {
  constexpr int KC = 32;
  const size_t ns = sw_l.size();
  for (int kb = k0; kb < k1; kb += KC) {
    const int kc = std::min(KC, k1 - kb);
    for (size_t sidx = 0; sidx < ns; ++sidx) {
      const int l = sw_l[sidx], m = sw_m[sidx];
      const double *c = rc + sw_off[sidx], *s = rs + sw_off[sidx];
      double carry[KC];
      const double *cm = Zt + (size_t)m * n + kb;
      if (kc == KC) {
        for (int k = 0; k < KC; ++k) carry[k] = cm[k];
        for (int i = m - 1; i >= l; --i) {
          const double ci = c[i], si = s[i];
          const double *ca = Zt + (size_t)i * n + kb;
          double *cb = Zt + (size_t)(i + 1) * n + kb;
#pragma omp simd
          for (int k = 0; k < KC; ++k) {
            const double ha = ca[k], hb = carry[k];
            cb[k] = si * ha + ci * hb;
            carry[k] = ci * ha - si * hb;
          }
        }
        double *cl = Zt + (size_t)l * n + kb;
        for (int k = 0; k < KC; ++k) cl[k] = carry[k];
      } else {
        for (int k = 0; k < kc; ++k) carry[k] = cm[k];
        for (int i = m - 1; i >= l; --i) {
          const double ci = c[i], si = s[i];
          const double *ca = Zt + (size_t)i * n + kb;
          double *cb = Zt + (size_t)(i + 1) * n + kb;
          for (int k = 0; k < kc; ++k) {
            const double ha = ca[k], hb = carry[k];
            cb[k] = si * ha + ci * hb;
            carry[k] = ci * ha - si * hb;
          }
        }
        double *cl = Zt + (size_t)l * n + kb;
        for (int k = 0; k < kc; ++k) cl[k] = carry[k];
      }
    }
  }
}

static void ql_apply(double *Zt, int n, const double *rc, const double *rs, const std::vector<int> &sw_l, const std::vector<int> &sw_m,
                     const std::vector<long long> &sw_off, int threads) {
  const int chunks = (n + 31) / 32;
  threads = std::max(1, std::min(threads, chunks));
  if (threads == 1) { ql_apply_rows(Zt, n, rc, rs, sw_l, sw_m, sw_off, 0, n); return; }
  std::vector<std::thread> pool;
  int k0 = 0;
  for (int t = 0; t < threads; ++t) {
    const int nch = chunks / threads + (t < chunks % threads ? 1 : 0);
    const int k1 = std::min(n, k0 + 32 * nch);
    if (t + 1 < threads) pool.emplace_back([=, &sw_l, &sw_m, &sw_off]() { ql_apply_rows(Zt, n, rc, rs, sw_l, sw_m, sw_off, k0, k1); });
    else ql_apply_rows(Zt, n, rc, rs, sw_l, sw_m, sw_off, k0, k1);
    k0 = k1;
  }
  for (auto &th : pool) th.join();
}

LIO_MV void sym_eigen(const Mat &A, Vec &d, Mat &Zout, int threads) {
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
  Vec rc, rs;                       // recorded rotations of every QL sweep
  std::vector<int> sw_l, sw_m;
  std::vector<long long> sw_off;
  rc.reserve((size_t)n * n); rs.reserve((size_t)n * n);
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n - 1 && std::fabs(e[m]) > eps * tst1) ++m;
    if (m > l) {
      for (int iter = 0; iter < 300; ++iter) {
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::sqrt(p * p + 1.0);
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
        // scalar QL recurrence only (it never reads Z): the plane rotations (i, i+1), i = m-1 .. l, are recorded and
        // applied to Z after the whole spectrum is known (ql_apply below)
        sw_l.push_back(l); sw_m.push_back(m); sw_off.push_back((long long)rc.size() - l);  // rc[off + i] is rotation i
        rc.resize(rc.size() + (size_t)(m - l)); rs.resize(rc.size());
        double *rcp = rc.data() + sw_off.back(), *rsp = rs.data() + sw_off.back();
        for (int i = m - 1; i >= l; --i) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::sqrt(p * p + e[i] * e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          rcp[i] = c; rsp[i] = s;
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
  ql_apply(Zt.data(), n, rc.data(), rs.data(), sw_l, sw_m, sw_off, threads);
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

LIO_MV void JtJ_dense(const double *J, const double *r, int rows, int cols, double *JtJ, double *Jtr) {
  for (int a = 0; a < cols; ++a) {
    double *out = JtJ + (size_t)a * cols;
    double gs = 0;
    for (int b = 0; b < cols; ++b) out[b] = 0.0;
    for (int k = 0; k < rows; ++k) {
      const double ja = J[k * cols + a];
      const double *jr = J + k * cols;
      for (int b = 0; b < cols; ++b) out[b] += ja * jr[b];
      gs += ja * r[k];
    }
    Jtr[a] = gs;
  }
}

// Four rows at a time: x is loaded once per four dot products and the four reductions are independent chains.
LIO_MV void matvec(const Mat &A, const Vec &x, Vec &y) {
  const int nr = A.r, nc = A.c;
  y.assign(nr, 0.0);
  const double *xp = x.data();
  int i = 0;
  for (; i + 4 <= nr; i += 4) {
    const double *r0 = &A.d[(size_t)i * nc], *r1 = r0 + nc, *r2 = r1 + nc, *r3 = r2 + nc;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma omp simd reduction(+ : s0, s1, s2, s3)
    for (int j = 0; j < nc; ++j) {
      const double xj = xp[j];
      s0 += r0[j] * xj; s1 += r1[j] * xj; s2 += r2[j] * xj; s3 += r3[j] * xj;
    }
    y[i] = s0; y[i + 1] = s1; y[i + 2] = s2; y[i + 3] = s3;
  }
  for (; i < nr; ++i) {
    const double *row = &A.d[(size_t)i * nc];
    double s = 0;
#pragma omp simd reduction(+ : s)
    for (int j = 0; j < nc; ++j) s += row[j] * xp[j];
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
    int cc = r;
    for (; cc + 4 <= n; cc += 4) {  // four columns per pass: br is loaded once, four independent reductions
      const double *b0 = &B.d[(size_t)cc * kc], *b1 = b0 + kc, *b2 = b1 + kc, *b3 = b2 + kc;
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma omp simd reduction(+ : s0, s1, s2, s3)
      for (int k = 0; k < kc; ++k) {
        const double v = br[k];
        s0 += v * b0[k]; s1 += v * b1[k]; s2 += v * b2[k]; s3 += v * b3[k];
      }
      const double sv[4] = {s0, s1, s2, s3};
      for (int q = 0; q < 4; ++q) { C.d[(size_t)r * n + cc + q] = sv[q]; C.d[(size_t)(cc + q) * n + r] = sv[q]; }
    }
    for (; cc < n; ++cc) {
      const double *bc = &B.d[(size_t)cc * kc];
      double s = 0;
#pragma omp simd reduction(+ : s)
      for (int k = 0; k < kc; ++k) s += br[k] * bc[k];
      C.d[(size_t)r * n + cc] = s;
      C.d[(size_t)cc * n + r] = s;
    }
  }
}

}  // namespace hm
}  // namespace lio
