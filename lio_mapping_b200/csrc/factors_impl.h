// lio_mapping_b200 — factor math shared by the host shell and the device-resident solver
// (__host__ __device__ inline; reference files cited per function in factors_host.h).
#pragma once
#include "hostmath.h"

namespace lio {

struct PimData {  // plain-data view of a finished IntegrationBase, uploadable to the device
  double delta_p[3], delta_q[4] /* x y z w */, delta_v[3], lin_ba[3], lin_bg[3];
  double sum_dt, g_norm;
  double jac[15][15];
  double sqrt_info[15][15];  // upper triangular: LLT(cov^-1).matrixL()^T
};

namespace fi {
using namespace hm;
enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };
LIO_HD inline void put33(double *dst, int ld, int r0, int c0, const M3 &m) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) dst[(r0 + i) * ld + c0 + j] = m(i, j);
}
LIO_HD inline M3 get33(const double *src, int ld, int r0, int c0) {
  M3 m;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(i, j) = src[(r0 + i) * ld + c0 + j];
  return m;
}
LIO_HD inline Q pose_q(const double *p) { return Q(p[6], p[3], p[4], p[5]); }
LIO_HD inline M3 left_tl(const Q &q) { return M3::I() * q.w + skew(q.vec()); }
LIO_HD inline M3 right_tl(const Q &p) { return M3::I() * p.w - skew(p.vec()); }
}  // namespace fi

// Raw (un-whitened) ImuFactor residual and Jacobian: raw[15]; A (15 x 30 over [pose_i(6) | sb_i(9) | pose_j(6) | sb_j(9)]
// tangent columns, row stride ld, or nullptr).  Only the structurally non-zero 3x3 blocks of A are written: the caller
// zeroes the buffer first.
LIO_HD inline void imu_factor_raw(const PimData &pim, const double *pose_i, const double *sb_i, const double *pose_j, const double *sb_j,
                                  double raw[15], double *A, int ld) {
  using namespace hm;
  using namespace fi;
  const V3 Pi(pose_i), Pj(pose_j), Vi(sb_i), Bai(sb_i + 3), Bgi(sb_i + 6), Vj(sb_j), Baj(sb_j + 3), Bgj(sb_j + 6);
  const Q Qi = pose_q(pose_i), Qj = pose_q(pose_j);
  const V3 g_vec(0, 0, -pim.g_norm);
  const double sum_dt = pim.sum_dt;
  const M3 dp_dba = get33(&pim.jac[0][0], 15, O_P, O_BA), dp_dbg = get33(&pim.jac[0][0], 15, O_P, O_BG);
  const M3 dq_dbg = get33(&pim.jac[0][0], 15, O_R, O_BG);
  const M3 dv_dba = get33(&pim.jac[0][0], 15, O_V, O_BA), dv_dbg = get33(&pim.jac[0][0], 15, O_V, O_BG);
  const V3 dba = Bai - V3(pim.lin_ba), dbg = Bgi - V3(pim.lin_bg);
  const Q pim_delta_q(pim.delta_q[3], pim.delta_q[0], pim.delta_q[1], pim.delta_q[2]);
  const Q corrected_delta_q = pim_delta_q * deltaQ(dq_dbg * dbg);
  const V3 corrected_delta_v = V3(pim.delta_v) + dv_dba * dba + dv_dbg * dbg;
  const V3 corrected_delta_p = V3(pim.delta_p) + dp_dba * dba + dp_dbg * dbg;
  const Q Qi_inv = inverse(Qi);
  const V3 rP = rotate(Qi_inv, -0.5 * g_vec * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt) - corrected_delta_p;
  const V3 rR = 2.0 * (inverse(corrected_delta_q) * (Qi_inv * Qj)).vec();
  const V3 rV = rotate(Qi_inv, -1.0 * g_vec * sum_dt + Vj - Vi) - corrected_delta_v;
  for (int k = 0; k < 3; ++k) { raw[O_P + k] = rP[k]; raw[O_R + k] = rR[k]; raw[O_V + k] = rV[k]; raw[O_BA + k] = Baj[k] - Bai[k]; raw[O_BG + k] = Bgj[k] - Bgi[k]; }
  if (!A) return;
  const M3 RiT = toR(Qi_inv);
  put33(A, ld, O_P, 0, -RiT);
  put33(A, ld, O_P, 3, skew(rotate(Qi_inv, -0.5 * g_vec * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt)));
  {  // -(L(Qj^-1 Qi) R(corrected_delta_q)) top-left 3x3
    const Q ql = inverse(Qj) * Qi;
    M3 m = left_tl(ql) * right_tl(corrected_delta_q);
    const V3 qv = ql.vec(), pv = corrected_delta_q.vec();
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) m(a, b) -= qv[a] * pv[b];
    put33(A, ld, O_R, 3, -m);
  }
  put33(A, ld, O_V, 3, skew(rotate(Qi_inv, -1.0 * g_vec * sum_dt + Vj - Vi)));
  put33(A, ld, O_P, 6, -RiT * sum_dt);
  put33(A, ld, O_P, 9, -dp_dba);
  put33(A, ld, O_P, 12, -dp_dbg);
  put33(A, ld, O_R, 12, -(left_tl(inverse(Qj) * Qi * corrected_delta_q) * dq_dbg));
  put33(A, ld, O_V, 6, -RiT);
  put33(A, ld, O_V, 9, -dv_dba);
  put33(A, ld, O_V, 12, -dv_dbg);
  put33(A, ld, O_BA, 9, -M3::I());
  put33(A, ld, O_BG, 12, -M3::I());
  put33(A, ld, O_P, 15, RiT);
  put33(A, ld, O_R, 18, left_tl(inverse(corrected_delta_q) * Qi_inv * Qj));
  put33(A, ld, O_V, 21, RiT);
  put33(A, ld, O_BA, 24, M3::I());
  put33(A, ld, O_BG, 27, M3::I());
}

// whiten = false returns the raw residual and raw Jacobian blocks (the caller applies sqrt_info, e.g. in parallel)
// Combined layout: J is 15 x 30 row-major over [pose_i(6) | sb_i(9) | pose_j(6) | sb_j(9)] tangent columns (or nullptr).
LIO_HD inline void imu_factor_eval30(const PimData &pim, const double *pose_i, const double *sb_i, const double *pose_j, const double *sb_j,
                                     double r[15], double (*J)[30], bool whiten = true) {
  double raw[15];
  double A[15][30];
  if (J) for (int a = 0; a < 15; ++a) for (int c = 0; c < 30; ++c) A[a][c] = 0;
  imu_factor_raw(pim, pose_i, sb_i, pose_j, sb_j, raw, J ? &A[0][0] : nullptr, 30);
  if (whiten) { for (int i = 0; i < 15; ++i) { double s = 0; for (int j = i; j < 15; ++j) s += pim.sqrt_info[i][j] * raw[j]; r[i] = s; } }
  else { for (int i = 0; i < 15; ++i) r[i] = raw[i]; }
  if (!J) return;
  if (!whiten) {
    for (int i = 0; i < 15; ++i) for (int c = 0; c < 30; ++c) J[i][c] = A[i][c];
    return;
  }
  // J = sqrt_info (upper triangular) * A, row-wise axpy (stride-1 inner loop)
  for (int i = 0; i < 15; ++i) {
    for (int c = 0; c < 30; ++c) J[i][c] = 0;
    for (int k = i; k < 15; ++k) {
      const double s = pim.sqrt_info[i][k];
      for (int c = 0; c < 30; ++c) J[i][c] += s * A[k][c];
    }
  }
}

LIO_HD inline void imu_factor_eval_impl(const PimData &pim, const double *pose_i, const double *sb_i, const double *pose_j, const double *sb_j,
                         double r[15], double (*Ji)[6], double (*Jsi)[9], double (*Jj)[6], double (*Jsj)[9], bool whiten = true) {
  if (!Ji) { imu_factor_eval30(pim, pose_i, sb_i, pose_j, sb_j, r, nullptr, whiten); return; }
  double J[15][30];
  imu_factor_eval30(pim, pose_i, sb_i, pose_j, sb_j, r, J, whiten);
  for (int i = 0; i < 15; ++i) {
    for (int c = 0; c < 6; ++c) { Ji[i][c] = J[i][c]; Jj[i][c] = J[i][15 + c]; }
    for (int c = 0; c < 9; ++c) { Jsi[i][c] = J[i][6 + c]; Jsj[i][c] = J[i][21 + c]; }
  }
}

LIO_HD inline void ppp_frame_terms_impl(const double *pose_pivot, const double *pose_i, const double *pose_ex, double Rout[9], double tout[3], double Mout[6 * 18]) {
  using namespace hm;
  using namespace fi;
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

LIO_HD inline void prior_factor_impl(const hm::V3 &pos0, const hm::Q &rot0, const double *pose_ex, double r[6], double (*J)[6]) {
  using namespace hm;
  using namespace fi;
  const V3 P(pose_ex);
  const Q Qx = pose_q(pose_ex);
  const double wp = 1000.0, wr = 0.1;
  const V3 rp = P - pos0;
  const V3 rr = 2.0 * (inverse(rot0) * Qx).vec();
  for (int k = 0; k < 3; ++k) { r[k] = wp * rp[k]; r[3 + k] = wr * rr[k]; }
  if (J) {
    for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) J[a][c] = 0;
    const M3 br = left_tl(inverse(Qx) * rot0);
    for (int i = 0; i < 3; ++i) {
      J[i][i] = wp;
      for (int j = 0; j < 3; ++j) J[3 + i][3 + j] = wr * br(i, j);
    }
  }
}

LIO_HD inline void pose_plus_impl(const double *x, const double *delta, double *out) {
  using namespace hm;
  using namespace fi;
  const Q q = pose_q(x);
  const Q qp = normalized(q * deltaQ(V3(delta[3], delta[4], delta[5])));
  for (int k = 0; k < 3; ++k) out[k] = x[k] + delta[k];
  out[3] = qp.x; out[4] = qp.y; out[5] = qp.z; out[6] = qp.w;
}

}  // namespace lio
