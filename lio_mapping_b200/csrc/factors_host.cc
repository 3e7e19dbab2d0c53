// lio_mapping_b200 — host fp64 factors (see factors_host.h).
#include "factors_host.h"

namespace lio {
using namespace hm;

enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };

static inline void put33(double *dst, int ld, int r0, int c0, const M3 &m) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) dst[(r0 + i) * ld + c0 + j] = m(i, j);
}
static inline M3 get33(const double *src, int ld, int r0, int c0) {
  M3 m;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(i, j) = src[(r0 + i) * ld + c0 + j];
  return m;
}
static inline Q pose_q(const double *p) { return Q(p[6], p[3], p[4], p[5]); }

// ---- pre-integration (IntegrationBase.h:77-101 ctor, :127-208 MidPointIntegration, :278-307 Propagate)
Preintegration::Preintegration(const V3 &a0, const V3 &g0, const V3 &ba, const V3 &bg, const ImuNoise &n)
    : acc0(a0), gyr0(g0), lin_acc(a0), lin_gyr(g0), lin_ba(ba), lin_bg(bg), g_norm(n.g_norm) {
  std::memset(jac, 0, sizeof(jac));
  std::memset(cov, 0, sizeof(cov));
  for (int i = 0; i < 15; ++i) jac[i][i] = 1.0;
  for (int k = 0; k < 3; ++k) {
    noise_diag[0 + k] = n.acc_n * n.acc_n; noise_diag[3 + k] = n.gyr_n * n.gyr_n;
    noise_diag[6 + k] = n.acc_n * n.acc_n; noise_diag[9 + k] = n.gyr_n * n.gyr_n;
    noise_diag[12 + k] = n.acc_w * n.acc_w; noise_diag[15 + k] = n.gyr_w * n.gyr_w;
  }
}

void Preintegration::push_back(double dt, const V3 &acc1, const V3 &gyr1) {
  sqrt_info_valid = false;
  const V3 un_acc_0 = rotate(delta_q, acc0 - lin_ba);
  const V3 un_gyr = 0.5 * (gyr0 + gyr1) - lin_bg;
  const Q rq = delta_q * Q(1, un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2);
  const V3 un_acc_1 = rotate(rq, acc1 - lin_ba);
  const V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
  const V3 rp = delta_p + delta_v * dt + 0.5 * un_acc * dt * dt;
  const V3 rv = delta_v + un_acc * dt;
  {
    const V3 w_x = un_gyr, a_0_x = acc0 - lin_ba, a_1_x = acc1 - lin_ba;
    const M3 R_w_x = skew(w_x), R_a_0_x = skew(a_0_x), R_a_1_x = skew(a_1_x);
    const M3 I3 = M3::I();
    const M3 dR = toR(delta_q), rR = toR(rq);
    double F[15][15], V[15][18];
    std::memset(F, 0, sizeof(F));
    std::memset(V, 0, sizeof(V));
    const M3 IwR = I3 - R_w_x * dt;
    put33(&F[0][0], 15, 0, 0, I3);
    put33(&F[0][0], 15, 0, 3, dR * R_a_0_x * (-0.25 * dt * dt) + rR * R_a_1_x * IwR * (-0.25 * dt * dt));
    put33(&F[0][0], 15, 0, 6, I3 * dt);
    put33(&F[0][0], 15, 0, 9, (dR + rR) * (-0.25 * dt * dt));
    put33(&F[0][0], 15, 0, 12, rR * R_a_1_x * (-0.1667 * dt * dt * -dt));
    put33(&F[0][0], 15, 3, 3, IwR);
    put33(&F[0][0], 15, 3, 12, I3 * (-1.0 * dt));
    put33(&F[0][0], 15, 6, 3, dR * R_a_0_x * (-0.5 * dt) + rR * R_a_1_x * IwR * (-0.5 * dt));
    put33(&F[0][0], 15, 6, 6, I3);
    put33(&F[0][0], 15, 6, 9, (dR + rR) * (-0.5 * dt));
    put33(&F[0][0], 15, 6, 12, rR * R_a_1_x * (-0.5 * dt * -dt));
    put33(&F[0][0], 15, 9, 9, I3);
    put33(&F[0][0], 15, 12, 12, I3);
    const M3 v03 = (-rR) * R_a_1_x * (0.25 * dt * dt * 0.5 * dt);
    const M3 v63 = (-rR) * R_a_1_x * (0.5 * dt * 0.5 * dt);
    put33(&V[0][0], 18, 0, 0, dR * (0.5 * dt * dt));
    put33(&V[0][0], 18, 0, 3, v03);
    put33(&V[0][0], 18, 0, 6, rR * (0.5 * dt * dt));
    put33(&V[0][0], 18, 0, 9, v03);
    put33(&V[0][0], 18, 3, 3, I3 * (0.5 * dt));
    put33(&V[0][0], 18, 3, 9, I3 * (0.5 * dt));
    put33(&V[0][0], 18, 6, 0, dR * (0.5 * dt));
    put33(&V[0][0], 18, 6, 3, v63);
    put33(&V[0][0], 18, 6, 6, rR * (0.5 * dt));
    put33(&V[0][0], 18, 6, 9, v63);
    put33(&V[0][0], 18, 9, 12, I3 * dt);
    put33(&V[0][0], 18, 12, 15, I3 * dt);
    double nj[15][15], fc[15][15], nc[15][15];
    for (int i = 0; i < 15; ++i)
      for (int j = 0; j < 15; ++j) {
        double s = 0, s2 = 0;
        for (int k = 0; k < 15; ++k) { s += F[i][k] * jac[k][j]; s2 += F[i][k] * cov[k][j]; }
        nj[i][j] = s; fc[i][j] = s2;
      }
    for (int i = 0; i < 15; ++i)
      for (int j = 0; j < 15; ++j) {
        double s = 0;
        for (int k = 0; k < 15; ++k) s += fc[i][k] * F[j][k];
        double s2 = 0;
        for (int k = 0; k < 18; ++k) s2 += V[i][k] * noise_diag[k] * V[j][k];
        nc[i][j] = s + s2;
      }
    std::memcpy(jac, nj, sizeof(jac));
    std::memcpy(cov, nc, sizeof(cov));
  }
  delta_p = rp; delta_v = rv;
  delta_q = normalized(rq);
  sum_dt += dt;
  acc0 = acc1; gyr0 = gyr1;
}

// ImuFactor.h:74-75: sqrt_info = LLT(covariance^-1).matrixL().transpose()
void Preintegration::ensure_sqrt_info() {
  if (sqrt_info_valid) return;
  Mat L(15, 15);
  for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) L(i, j) = cov[i][j];
  Mat inv(15, 15);
  if (cholesky(L)) {
    for (int c = 0; c < 15; ++c) {
      Vec e(15, 0.0);
      e[c] = 1.0;
      cholesky_solve(L, e);
      for (int r = 0; r < 15; ++r) inv(r, c) = e[r];
    }
    for (int i = 0; i < 15; ++i) for (int j = i + 1; j < 15; ++j) { double s = 0.5 * (inv(i, j) + inv(j, i)); inv(i, j) = inv(j, i) = s; }
  }
  Mat Li = inv;
  bool ok = cholesky(Li);
  for (int i = 0; i < 15; ++i) for (int j = 0; j < 15; ++j) sqrt_info[i][j] = (ok && j >= i) ? Li(j, i) : 0.0;
  sqrt_info_valid = true;
}

static inline M3 left_tl(const Q &q) { return M3::I() * q.w + skew(q.vec()); }    // LeftQuatMatrix top-left 3x3
static inline M3 right_tl(const Q &p) { return M3::I() * p.w - skew(p.vec()); }   // RightQuatMatrix top-left 3x3

void imu_factor_evaluate(Preintegration &pim, const double *pose_i, const double *sb_i, const double *pose_j, const double *sb_j,
                         double r[15], double (*Ji)[6], double (*Jsi)[9], double (*Jj)[6], double (*Jsj)[9]) {
  pim.ensure_sqrt_info();
  const V3 Pi(pose_i), Pj(pose_j), Vi(sb_i), Bai(sb_i + 3), Bgi(sb_i + 6), Vj(sb_j), Baj(sb_j + 3), Bgj(sb_j + 6);
  const Q Qi = pose_q(pose_i), Qj = pose_q(pose_j);
  const V3 g_vec(0, 0, -pim.g_norm);
  const double sum_dt = pim.sum_dt;
  const M3 dp_dba = get33(&pim.jac[0][0], 15, O_P, O_BA), dp_dbg = get33(&pim.jac[0][0], 15, O_P, O_BG);
  const M3 dq_dbg = get33(&pim.jac[0][0], 15, O_R, O_BG);
  const M3 dv_dba = get33(&pim.jac[0][0], 15, O_V, O_BA), dv_dbg = get33(&pim.jac[0][0], 15, O_V, O_BG);
  const V3 dba = Bai - pim.lin_ba, dbg = Bgi - pim.lin_bg;
  const Q corrected_delta_q = pim.delta_q * deltaQ(dq_dbg * dbg);
  const V3 corrected_delta_v = pim.delta_v + dv_dba * dba + dv_dbg * dbg;
  const V3 corrected_delta_p = pim.delta_p + dp_dba * dba + dp_dbg * dbg;
  const Q Qi_inv = inverse(Qi);
  const V3 rP = rotate(Qi_inv, -0.5 * g_vec * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt) - corrected_delta_p;
  const V3 rR = 2.0 * (inverse(corrected_delta_q) * (Qi_inv * Qj)).vec();
  const V3 rV = rotate(Qi_inv, -1.0 * g_vec * sum_dt + Vj - Vi) - corrected_delta_v;
  double raw[15];
  for (int k = 0; k < 3; ++k) { raw[O_P + k] = rP[k]; raw[O_R + k] = rR[k]; raw[O_V + k] = rV[k]; raw[O_BA + k] = Baj[k] - Bai[k]; raw[O_BG + k] = Bgj[k] - Bgi[k]; }
  for (int i = 0; i < 15; ++i) { double s = 0; for (int j = i; j < 15; ++j) s += pim.sqrt_info[i][j] * raw[j]; r[i] = s; }
  if (!Ji) return;
  const M3 RiT = toR(Qi_inv);
  double A0[15][6], A1[15][9], A2[15][6], A3[15][9];
  std::memset(A0, 0, sizeof(A0)); std::memset(A1, 0, sizeof(A1)); std::memset(A2, 0, sizeof(A2)); std::memset(A3, 0, sizeof(A3));
  put33(&A0[0][0], 6, O_P, 0, -RiT);
  put33(&A0[0][0], 6, O_P, 3, skew(rotate(Qi_inv, -0.5 * g_vec * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt)));
  {  // -(L(Qj^-1 Qi) R(corrected_delta_q)) top-left 3x3
    const Q ql = inverse(Qj) * Qi;
    M3 m = left_tl(ql) * right_tl(corrected_delta_q);
    const V3 qv = ql.vec(), pv = corrected_delta_q.vec();
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) m(a, b) -= qv[a] * pv[b];
    put33(&A0[0][0], 6, O_R, 3, -m);
  }
  put33(&A0[0][0], 6, O_V, 3, skew(rotate(Qi_inv, -1.0 * g_vec * sum_dt + Vj - Vi)));
  put33(&A1[0][0], 9, O_P, 0, -RiT * sum_dt);
  put33(&A1[0][0], 9, O_P, 3, -dp_dba);
  put33(&A1[0][0], 9, O_P, 6, -dp_dbg);
  put33(&A1[0][0], 9, O_R, 6, -(left_tl(inverse(Qj) * Qi * corrected_delta_q) * dq_dbg));
  put33(&A1[0][0], 9, O_V, 0, -RiT);
  put33(&A1[0][0], 9, O_V, 3, -dv_dba);
  put33(&A1[0][0], 9, O_V, 6, -dv_dbg);
  put33(&A1[0][0], 9, O_BA, 3, -M3::I());
  put33(&A1[0][0], 9, O_BG, 6, -M3::I());
  put33(&A2[0][0], 6, O_P, 0, RiT);
  put33(&A2[0][0], 6, O_R, 3, left_tl(inverse(corrected_delta_q) * Qi_inv * Qj));
  put33(&A3[0][0], 9, O_V, 0, RiT);
  put33(&A3[0][0], 9, O_BA, 3, M3::I());
  put33(&A3[0][0], 9, O_BG, 6, M3::I());
  for (int i = 0; i < 15; ++i) {
    for (int c = 0; c < 6; ++c) {
      double s0 = 0, s2 = 0;
      for (int k = i; k < 15; ++k) { s0 += pim.sqrt_info[i][k] * A0[k][c]; s2 += pim.sqrt_info[i][k] * A2[k][c]; }
      Ji[i][c] = s0; Jj[i][c] = s2;
    }
    for (int c = 0; c < 9; ++c) {
      double s1 = 0, s3 = 0;
      for (int k = i; k < 15; ++k) { s1 += pim.sqrt_info[i][k] * A1[k][c]; s3 += pim.sqrt_info[i][k] * A3[k][c]; }
      Jsi[i][c] = s1; Jsj[i][c] = s3;
    }
  }
}

// ---- lidar factor -------------------------------------------------------------------------------
static inline V3 row_times(const V3 &w, const M3 &m) {
  return V3(w.x * m(0, 0) + w.y * m(1, 0) + w.z * m(2, 0), w.x * m(0, 1) + w.y * m(1, 1) + w.z * m(2, 1),
            w.x * m(0, 2) + w.y * m(1, 2) + w.z * m(2, 2));
}

void ppp_evaluate_single(const double point[3], const double coeff[4], const double *pose_pivot, const double *pose_i,
                         const double *pose_ex, double *residual, double *J0, double *J1, double *J2) {
  const V3 P_pivot(pose_pivot), Pi(pose_i), tlb(pose_ex), p(point), w(coeff);
  const Q Q_pivot = pose_q(pose_pivot), Qi = pose_q(pose_i), qlb = pose_q(pose_ex);
  const Q Qlpivot = Q_pivot * conj(qlb);
  const V3 Plpivot = P_pivot - rotate(Qlpivot, tlb);
  const Q Qli = Qi * conj(qlb);
  const V3 Pli = Pi - rotate(Qli, tlb);
  const Q Qlpi = conj(Qlpivot) * Qli;
  const V3 Plpi = rotate(conj(Qlpivot), Pli - Plpivot);
  *residual = dot(w, rotate(Qlpi, p) + Plpi) + coeff[3];
  if (!J0 && !J1 && !J2) return;
  const M3 Ri = toR(Qi), Rp = toR(Q_pivot), rlb = toR(qlb);
  auto put = [](double *J, const V3 &a, const V3 &b) { J[0] = a.x; J[1] = a.y; J[2] = a.z; J[3] = b.x; J[4] = b.y; J[5] = b.z; J[6] = 0.0; };
  if (J0) {
    M3 S = skew(T(Rp) * (Ri * (T(rlb) * (p - tlb)))) + skew(T(Rp) * (Pi - P_pivot));
    put(J0, -row_times(w, rlb * T(Rp)), row_times(w, rlb * S));
  }
  if (J1) {
    M3 A = rlb * T(Rp);
    M3 S = skew(T(rlb) * tlb) - skew(T(rlb) * p);
    put(J1, row_times(w, A), row_times(w, A * Ri * S));
  }
  if (J2) {
    M3 S = T(Rp) * Ri * skew(T(rlb) * (p - tlb)) - skew(T(Rp) * (Ri * (T(rlb) * (p - tlb)))) - skew(T(Rp) * (Pi - P_pivot));
    put(J2, row_times(w, M3::I() - rlb * T(Rp) * Ri * T(rlb)), row_times(w, rlb * S));
  }
}

void ppp_frame_terms(const double *pose_pivot, const double *pose_i, const double *pose_ex, double Rout[9], double tout[3], double Mout[6 * 18]) {
  const V3 P_pivot(pose_pivot), Pi(pose_i), tlb(pose_ex);
  const M3 Rp = toR(pose_q(pose_pivot)), Ri = toR(pose_q(pose_i)), rlb = toR(pose_q(pose_ex));
  const M3 Rlpi = rlb * T(Rp) * Ri * T(rlb);
  // P_lpi = rlb Rp^T (Pi - Pp) - R_lpi tlb + tlb
  const V3 Plpi = rlb * (T(Rp) * (Pi - P_pivot)) - Rlpi * tlb + tlb;
  const V3 t = T(Rlpi) * Plpi;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rout[i * 3 + j] = Rlpi(i, j);
  tout[0] = t.x; tout[1] = t.y; tout[2] = t.z;
  // closed form of the 6x18 map (derivation in DESIGN.md): rows 0-2 multiply a, rows 3-5 multiply p x a
  const M3 B = rlb * T(Ri);           // w^T rlb Rp^T = a^T B
  const M3 C = T(Rp) * Ri * T(rlb);   // rlb^T w = C a
  const M3 Ct = T(C);
  const V3 v = T(Rp) * (Pi - P_pivot);
  const M3 St = skew(tlb), Sv = skew(v);
  const M3 Z;  // zero
  const M3 a_blocks[6] = {-B, Ct * Sv - St * Ct, B, St * rlb, T(Rlpi) - M3::I(), St * Ct - St * rlb - Ct * Sv};
  const M3 x_blocks[6] = {Z, -Ct, Z, rlb, Z, Ct - rlb};
  for (int blk = 0; blk < 6; ++blk)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        Mout[i * 18 + blk * 3 + j] = a_blocks[blk](i, j);
        Mout[(3 + i) * 18 + blk * 3 + j] = x_blocks[blk](i, j);
      }
}

void prior_factor_evaluate(const V3 &pos0, const Q &rot0, const double *pose_ex, double r[6], double (*J)[6]) {
  const V3 P(pose_ex);
  const Q Qx = pose_q(pose_ex);
  const double wp = 1000.0, wr = 0.1;
  const V3 rp = P - pos0;
  const V3 rr = 2.0 * (inverse(rot0) * Qx).vec();
  for (int k = 0; k < 3; ++k) { r[k] = wp * rp[k]; r[3 + k] = wr * rr[k]; }
  if (J) {
    std::memset(J, 0, sizeof(double) * 36);
    const M3 br = left_tl(inverse(Qx) * rot0);
    for (int i = 0; i < 3; ++i) {
      J[i][i] = wp;
      for (int j = 0; j < 3; ++j) J[3 + i][3 + j] = wr * br(i, j);
    }
  }
}

void pose_plus(const double *x, const double *delta, double *out) {
  const Q q = pose_q(x);
  const Q qp = normalized(q * deltaQ(V3(delta[3], delta[4], delta[5])));
  for (int k = 0; k < 3; ++k) out[k] = x[k] + delta[k];
  out[3] = qp.x; out[4] = qp.y; out[5] = qp.z; out[6] = qp.w;
}

V3 R2ypr(const M3 &R) {
  const V3 n(R(0, 0), R(1, 0), R(2, 0)), o(R(0, 1), R(1, 1), R(2, 1)), a(R(0, 2), R(1, 2), R(2, 2));
  const double y = std::atan2(n.y, n.x);
  const double p = std::atan2(-n.z, n.x * std::cos(y) + n.y * std::sin(y));
  const double r = std::atan2(a.x * std::sin(y) - a.y * std::cos(y), -o.x * std::sin(y) + o.y * std::cos(y));
  return V3(y, p, r) * (180.0 / M_PI);
}
M3 ypr2R(const V3 &ypr) {
  const double y = ypr.x / 180.0 * M_PI, p = ypr.y / 180.0 * M_PI, r = ypr.z / 180.0 * M_PI;
  M3 Rz, Ry, Rx;
  Rz(0, 0) = std::cos(y); Rz(0, 1) = -std::sin(y); Rz(1, 0) = std::sin(y); Rz(1, 1) = std::cos(y); Rz(2, 2) = 1;
  Ry(0, 0) = std::cos(p); Ry(0, 2) = std::sin(p); Ry(1, 1) = 1; Ry(2, 0) = -std::sin(p); Ry(2, 2) = std::cos(p);
  Rx(0, 0) = 1; Rx(1, 1) = std::cos(r); Rx(1, 2) = -std::sin(r); Rx(2, 1) = std::sin(r); Rx(2, 2) = std::cos(r);
  return Rz * Ry * Rx;
}

}  // namespace lio
