// lio_mapping_b200 — small host-side math for the estimator shell around the CUDA kernels
// (window state, IMU / prior factors, dense trust-region step).  fp64 unless noted.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#ifdef __CUDACC__
#define LIO_HD __host__ __device__
#else
#define LIO_HD
#endif

namespace lio {
namespace hm {

struct V3 {
  double x = 0, y = 0, z = 0;
  LIO_HD V3() {}
  LIO_HD V3(double a, double b, double c) : x(a), y(b), z(c) {}
  LIO_HD explicit V3(const double *p) : x(p[0]), y(p[1]), z(p[2]) {}
  LIO_HD double &operator[](int i) { return (&x)[i]; }
  LIO_HD double operator[](int i) const { return (&x)[i]; }
};
LIO_HD inline V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
LIO_HD inline V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
LIO_HD inline V3 operator-(const V3 &a) { return V3(-a.x, -a.y, -a.z); }
LIO_HD inline V3 operator*(const V3 &a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
LIO_HD inline V3 operator*(double s, const V3 &a) { return a * s; }
LIO_HD inline double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
LIO_HD inline V3 cross(const V3 &a, const V3 &b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
LIO_HD inline double norm(const V3 &a) { return std::sqrt(dot(a, a)); }

struct M3 {
  double m[3][3];
  LIO_HD M3() { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = 0.0; }
  LIO_HD static M3 I() { M3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1; return r; }
  LIO_HD double &operator()(int i, int j) { return m[i][j]; }
  LIO_HD double operator()(int i, int j) const { return m[i][j]; }
};
LIO_HD inline M3 operator*(const M3 &a, const M3 &b) {
  M3 r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += a.m[i][k] * b.m[k][j]; r.m[i][j] = s; }
  return r;
}
LIO_HD inline V3 operator*(const M3 &a, const V3 &v) {
  return V3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
LIO_HD inline M3 operator*(const M3 &a, double s) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] * s; return r; }
LIO_HD inline M3 operator+(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r; }
LIO_HD inline M3 operator-(const M3 &a, const M3 &b) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j]; return r; }
LIO_HD inline M3 operator-(const M3 &a) { return a * -1.0; }
LIO_HD inline M3 T(const M3 &a) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i]; return r; }
LIO_HD inline M3 skew(const V3 &v) {
  M3 r;
  r(0, 1) = -v.z; r(0, 2) = v.y; r(1, 0) = v.z; r(1, 2) = -v.x; r(2, 0) = -v.y; r(2, 1) = v.x;
  return r;
}

struct Q {  // Hamilton quaternion, (x,y,z,w) storage
  double x = 0, y = 0, z = 0, w = 1;
  LIO_HD Q() {}
  LIO_HD Q(double w_, double x_, double y_, double z_) : x(x_), y(y_), z(z_), w(w_) {}
  LIO_HD V3 vec() const { return V3(x, y, z); }
};
LIO_HD inline Q conj(const Q &q) { return Q(q.w, -q.x, -q.y, -q.z); }
LIO_HD inline double norm(const Q &q) { return std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w); }
LIO_HD inline Q normalized(const Q &q) { double n = norm(q); return Q(q.w / n, q.x / n, q.y / n, q.z / n); }
LIO_HD inline Q inverse(const Q &q) { double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w; return Q(q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2); }
LIO_HD inline Q operator*(const Q &a, const Q &b) {
  return Q(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
           a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x);
}
LIO_HD inline V3 rotate(const Q &q, const V3 &v) {
  V3 u = cross(q.vec(), v);
  u = u + u;
  return v + u * q.w + cross(q.vec(), u);
}
LIO_HD inline M3 toR(const Q &q) {
  M3 r;
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
  r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
  return r;
}
LIO_HD inline Q fromR(const M3 &m) {
  Q q;
  double t = m(0, 0) + m(1, 1) + m(2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m(2, 1) - m(1, 2)) * t; q.y = (m(0, 2) - m(2, 0)) * t; q.z = (m(1, 0) - m(0, 1)) * t;
  } else {
    int i = 0;
    if (m(1, 1) > m(0, 0)) i = 1;
    if (m(2, 2) > m(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
    double c[3];
    c[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m(k, j) - m(j, k)) * t;
    c[j] = (m(j, i) + m(i, j)) * t;
    c[k] = (m(k, i) + m(i, k)) * t;
    q.x = c[0]; q.y = c[1]; q.z = c[2];
  }
  return q;
}
LIO_HD inline Q deltaQ(const V3 &th) { return Q(1.0, th.x / 2, th.y / 2, th.z / 2); }  // mathutils::DeltaQ (not normalised)

// rigid transform with the reference's Twist<T> composition rules (include/utils/Twist.h:40-97)
struct Tw {
  Q rot;
  V3 pos;
  LIO_HD Tw() {}
  LIO_HD Tw(const Q &r, const V3 &p) : rot(r), pos(p) {}
};
LIO_HD inline M3 linear(const Tw &t) { return toR(normalized(t.rot)); }
LIO_HD inline Tw tw_from_affine(const M3 &R, const V3 &t) { return Tw(normalized(fromR(R)), t); }
LIO_HD inline Tw tw_inverse(const Tw &t) {
  M3 Rt = T(linear(t));
  return Tw(fromR(Rt), -(Rt * t.pos));
}
LIO_HD inline Tw tw_mul(const Tw &a, const Tw &b) { return tw_from_affine(linear(a) * linear(b), linear(a) * b.pos + a.pos); }

// ---- dense row-major matrix ------------------------------------------------------------------
struct Mat {
  int r = 0, c = 0;
  std::vector<double> d;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, 0.0) {}
  double &operator()(int i, int j) { return d[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return d[(size_t)i * c + j]; }
  void zero() { std::fill(d.begin(), d.end(), 0.0); }
};
typedef std::vector<double> Vec;

inline Vec mul(const Mat &a, const Vec &x) {
  Vec o(a.r, 0.0);
  for (int i = 0; i < a.r; ++i) { const double *row = &a.d[(size_t)i * a.c]; double s = 0; for (int j = 0; j < a.c; ++j) s += row[j] * x[j]; o[i] = s; }
  return o;
}
inline double vdot(const Vec &a, const Vec &b) { double s = 0; for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i]; return s; }
inline double vnorm(const Vec &a) { return std::sqrt(vdot(a, a)); }

// lower Cholesky in place; false when not positive definite
bool cholesky(Mat &a);
void cholesky_solve(const Mat &L, Vec &b);
// symmetric eigen-decomposition (ascending eigenvalues, eigenvectors in columns)
// threads > 1: the QL rotations are applied to disjoint row ranges of the eigenvector matrix by that many threads
// (same arithmetic per element, bit-identical to threads = 1).
void sym_eigen(const Mat &A, Vec &evals, Mat &evecs, int threads = 1);
// H(cm[a], cm[b]) += sum_k J[k][a] J[k][b];  g[cm[a]] += sum_k J[k][a] r[k]   (J: rows x cols row-major, cm: column map)
void add_JtJ_mapped(const double *J, const double *r, int rows, int cols, const int *cm, Mat &H, Vec &g);
// JtJ (cols x cols, row-major) = J^T J,  Jtr (cols) = J^T r   (J: rows x cols row-major)
void JtJ_dense(const double *J, const double *r, int rows, int cols, double *JtJ, double *Jtr);
// y = A x (multiversioned)
void matvec(const Mat &A, const Vec &x, Vec &y);
// C(r,c) = sum_{k in cols} A(r,k) w[k] A(c,k)
void weighted_gram(const Mat &A, const Vec &w, const std::vector<int> &cols, Mat &C);

}  // namespace hm
}  // namespace lio
