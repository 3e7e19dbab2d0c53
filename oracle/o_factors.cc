// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
//
// CPU restatement (fp64) of the reference's factors:
//   PivotPointPlaneFactor::Evaluate    src/factor/PivotPointPlaneFactor.cc:43-137
//   IntegrationBase                    include/imu_processor/IntegrationBase.h:72-388
//   ImuFactor::Evaluate                include/factor/ImuFactor.h:53-167
//   PriorFactor::Evaluate              src/factor/PriorFactor.cc:35-67
//   PoseLocalParameterization::Plus    src/factor/PoseLocalParameterization.cc:35-52
//   ResidualBlockInfo / MarginalizationInfo / MarginalizationFactor
//                                      src/factor/MarginalizationFactor.cc:37-392
#include "o_factors.h"
#include <cmath>
#include <cstring>
#include <thread>

namespace orc {

// ---- helpers ------------------------------------------------------------------------------------
static inline void set33(double *J, int ld, int r0, int c0, const M3 &m) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J[(r0 + i) * ld + c0 + j] = m(i, j);
}
static inline void setrow3(double *J, int c0, const V3 &v) { J[c0] = v.x; J[c0 + 1] = v.y; J[c0 + 2] = v.z; }
static inline V3 rowvec_times(const V3 &w, const M3 &m) {  // w^T * m
  return V3(w.x * m(0, 0) + w.y * m(1, 0) + w.z * m(2, 0), w.x * m(0, 1) + w.y * m(1, 1) + w.z * m(2, 1),
            w.x * m(0, 2) + w.y * m(1, 2) + w.z * m(2, 2));
}
// mathutils::LeftQuatMatrix / RightQuatMatrix top-left 3x3 (math_utils.h:139-162)
static inline M3 LeftQuatTL(const Qd &q) { return M3::Identity() * q.w + Skew(q.vec()); }
static inline M3 RightQuatTL(const Qd &p) { return M3::Identity() * p.w - Skew(p.vec()); }
// full 4x4 products only ever use the top-left 3x3 of L(q)*R(p):
static inline M3 LeftTimesRightTL(const Qd &q, const Qd &p) {
  // L(q) = [[q4 I + [qv]x, qv], [-qv^T, q4]],  R(p) = [[p4 I - [pv]x, pv], [-pv^T, p4]]
  M3 a = LeftQuatTL(q) * RightQuatTL(p);
  V3 qv = q.vec(), pv = p.vec();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a(i, j) += qv[i] * (-pv[j]);
  return a;
}

// ---- PivotPointPlaneFactor ----------------------------------------------------------------------
PivotPointPlaneFactor::PivotPointPlaneFactor(const double p[3], const double c[4]) : point(p[0], p[1], p[2]) {
  for (int k = 0; k < 4; ++k) coeff[k] = c[k];
  num_residuals = 1;
  block_sizes = {7, 7, 7};
}

bool PivotPointPlaneFactor::Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
  V3 P_pivot(parameters[0][0], parameters[0][1], parameters[0][2]);
  Qd Q_pivot(parameters[0][6], parameters[0][3], parameters[0][4], parameters[0][5]);
  V3 Pi(parameters[1][0], parameters[1][1], parameters[1][2]);
  Qd Qi(parameters[1][6], parameters[1][3], parameters[1][4], parameters[1][5]);
  V3 tlb(parameters[2][0], parameters[2][1], parameters[2][2]);
  Qd qlb(parameters[2][6], parameters[2][3], parameters[2][4], parameters[2][5]);

  Qd Qlpivot = Q_pivot * qlb.conjugate();
  V3 Plpivot = P_pivot - Qlpivot * tlb;
  Qd Qli = Qi * qlb.conjugate();
  V3 Pli = Pi - Qli * tlb;
  Qd Qlpi = Qlpivot.conjugate() * Qli;
  V3 Plpi = Qlpivot.conjugate() * (Pli - Plpivot);

  V3 w(coeff[0], coeff[1], coeff[2]);
  double b = coeff[3];
  double residual = w.dot(Qlpi * point + Plpi) + b;
  const double sqrt_info = 1.0;  // sqrt_info_static (:35)
  residuals[0] = sqrt_info * residual;

  if (jacobians) {
    M3 Ri = Qi.toRotationMatrix();
    M3 Rp = Q_pivot.toRotationMatrix();
    M3 rlb = qlb.toRotationMatrix();
    if (jacobians[0]) {
      double *J = jacobians[0];
      for (int k = 0; k < 7; ++k) J[k] = 0;
      V3 jl = -rowvec_times(w, rlb * Rp.transpose());
      M3 S = Skew(Rp.transpose() * (Ri * (rlb.transpose() * (point - tlb)))) + Skew(Rp.transpose() * (Pi - P_pivot));
      V3 jr = rowvec_times(w, rlb * S);
      setrow3(J, 0, jl * sqrt_info);
      setrow3(J, 3, jr * sqrt_info);
    }
    if (jacobians[1]) {
      double *J = jacobians[1];
      for (int k = 0; k < 7; ++k) J[k] = 0;
      M3 A = rlb * Rp.transpose();
      V3 jl = rowvec_times(w, A);
      M3 S = -Skew(rlb.transpose() * point) + Skew(rlb.transpose() * tlb);
      V3 jr = rowvec_times(w, A * Ri * S);
      setrow3(J, 0, jl * sqrt_info);
      setrow3(J, 3, jr * sqrt_info);
    }
    if (jacobians[2]) {
      double *J = jacobians[2];
      for (int k = 0; k < 7; ++k) J[k] = 0;
      M3 I3 = M3::Identity();
      V3 jl = rowvec_times(w, I3 - rlb * Rp.transpose() * Ri * rlb.transpose());
      M3 S = -Skew(Rp.transpose() * (Ri * (rlb.transpose() * (point - tlb)))) +
             Rp.transpose() * Ri * Skew(rlb.transpose() * (point - tlb)) - Skew(Rp.transpose() * (Pi - P_pivot));
      V3 jr = rowvec_times(w, rlb * S);
      setrow3(J, 0, jl * sqrt_info);
      setrow3(J, 3, jr * sqrt_info);
    }
  }
  return true;
}

// ---- IntegrationBase ----------------------------------------------------------------------------
static void set_block(MatX &m, int r0, int c0, const M3 &b) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m(r0 + i, c0 + j) = b(i, j);
}
static M3 get_block(const MatX &m, int r0, int c0) {
  M3 b;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) b(i, j) = m(r0 + i, c0 + j);
  return b;
}
enum { O_P = 0, O_R = 3, O_V = 6, O_BA = 9, O_BG = 12 };

IntegrationBase::IntegrationBase(const V3 &acc0, const V3 &gyr0, const V3 &ba, const V3 &bg, const IntegrationBaseConfig &cfg)
    : acc0_(acc0), gyr0_(gyr0), linearized_acc_(acc0), linearized_gyr_(gyr0), linearized_ba_(ba), linearized_bg_(bg),
      jacobian_(15, 15), covariance_(15, 15), noise_(18, 18), config_(cfg) {
  for (int i = 0; i < 15; ++i) jacobian_(i, i) = 1.0;
  g_vec_ = V3(0, 0, -config_.g_norm);
  for (int k = 0; k < 3; ++k) {
    noise_(0 + k, 0 + k) = config_.acc_n * config_.acc_n;
    noise_(3 + k, 3 + k) = config_.gyr_n * config_.gyr_n;
    noise_(6 + k, 6 + k) = config_.acc_n * config_.acc_n;
    noise_(9 + k, 9 + k) = config_.gyr_n * config_.gyr_n;
    noise_(12 + k, 12 + k) = config_.acc_w * config_.acc_w;
    noise_(15 + k, 15 + k) = config_.gyr_w * config_.gyr_w;
  }
}

void IntegrationBase::push_back(double dt, const V3 &acc, const V3 &gyr) {
  dt_buf_.push_back(dt); acc_buf_.push_back(acc); gyr_buf_.push_back(gyr);
  Propagate(dt, acc, gyr);
}

void IntegrationBase::Repropagate(const V3 &ba, const V3 &bg) {
  sum_dt_ = 0.0;
  acc0_ = linearized_acc_; gyr0_ = linearized_gyr_;
  delta_p_ = V3(); delta_q_ = Qd(); delta_v_ = V3();
  linearized_ba_ = ba; linearized_bg_ = bg;
  jacobian_.setZero();
  for (int i = 0; i < 15; ++i) jacobian_(i, i) = 1.0;
  covariance_.setZero();
  for (size_t i = 0; i < dt_buf_.size(); ++i) Propagate(dt_buf_[i], acc_buf_[i], gyr_buf_[i]);
}

void IntegrationBase::MidPointIntegration(double dt, const V3 &acc0, const V3 &gyr0, const V3 &acc1, const V3 &gyr1,
                                          const V3 &delta_p, const Qd &delta_q, const V3 &delta_v, const V3 &lba, const V3 &lbg,
                                          V3 &rp, Qd &rq, V3 &rv, bool update_jacobian) {
  V3 un_acc_0 = delta_q * (acc0 - lba);
  V3 un_gyr = 0.5 * (gyr0 + gyr1) - lbg;
  rq = delta_q * Qd(1, un_gyr.x * dt / 2, un_gyr.y * dt / 2, un_gyr.z * dt / 2);
  V3 un_acc_1 = rq * (acc1 - lba);
  V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
  rp = delta_p + delta_v * dt + 0.5 * un_acc * dt * dt;
  rv = delta_v + un_acc * dt;
  if (update_jacobian) {
    V3 w_x = 0.5 * (gyr0 + gyr1) - lbg;
    V3 a_0_x = acc0 - lba;
    V3 a_1_x = acc1 - lba;
    M3 R_w_x = Skew(w_x), R_a_0_x = Skew(a_0_x), R_a_1_x = Skew(a_1_x);
    M3 I3 = M3::Identity();
    M3 dR = delta_q.toRotationMatrix(), rR = rq.toRotationMatrix();
    MatX F(15, 15);
    set_block(F, 0, 0, I3);
    set_block(F, 0, 3, dR * R_a_0_x * (-0.25 * dt * dt) + rR * R_a_1_x * (I3 - R_w_x * dt) * (-0.25 * dt * dt));
    set_block(F, 0, 6, I3 * dt);
    set_block(F, 0, 9, (dR + rR) * (-0.25 * dt * dt));
    set_block(F, 0, 12, rR * R_a_1_x * (-0.1667 * dt * dt * -dt));
    set_block(F, 3, 3, I3 - R_w_x * dt);
    set_block(F, 3, 12, I3 * (-1.0 * dt));
    set_block(F, 6, 3, dR * R_a_0_x * (-0.5 * dt) + rR * R_a_1_x * (I3 - R_w_x * dt) * (-0.5 * dt));
    set_block(F, 6, 6, I3);
    set_block(F, 6, 9, (dR + rR) * (-0.5 * dt));
    set_block(F, 6, 12, rR * R_a_1_x * (-0.5 * dt * -dt));
    set_block(F, 9, 9, I3);
    set_block(F, 12, 12, I3);
    MatX V(15, 18);
    set_block(V, 0, 0, dR * (0.5 * dt * dt));
    M3 v03 = (-rR) * R_a_1_x * (0.25 * dt * dt * 0.5 * dt);
    set_block(V, 0, 3, v03);
    set_block(V, 0, 6, rR * (0.5 * dt * dt));
    set_block(V, 0, 9, v03);
    set_block(V, 3, 3, I3 * (0.5 * dt));
    set_block(V, 3, 9, I3 * (0.5 * dt));
    set_block(V, 6, 0, dR * (0.5 * dt));
    M3 v63 = (-rR) * R_a_1_x * (0.5 * dt * 0.5 * dt);
    set_block(V, 6, 3, v63);
    set_block(V, 6, 6, rR * (0.5 * dt));
    set_block(V, 6, 9, v63);
    set_block(V, 9, 12, I3 * dt);
    set_block(V, 12, 15, I3 * dt);
    jacobian_ = matmul(F, jacobian_);
    MatX FP = matmul(matmul(F, covariance_), F.transpose());
    MatX VN = matmul(matmul(V, noise_), V.transpose());
    for (size_t k = 0; k < FP.d.size(); ++k) FP.d[k] += VN.d[k];
    covariance_ = FP;
  }
}

void IntegrationBase::Propagate(double dt, const V3 &acc1, const V3 &gyr1) {
  dt_ = dt; acc1_ = acc1; gyr1_ = gyr1;
  V3 rp, rv;
  Qd rq;
  MidPointIntegration(dt, acc0_, gyr0_, acc1, gyr1, delta_p_, delta_q_, delta_v_, linearized_ba_, linearized_bg_, rp, rq, rv, true);
  delta_p_ = rp; delta_q_ = rq; delta_v_ = rv;
  delta_q_.normalize();
  sum_dt_ += dt_;
  acc0_ = acc1_; gyr0_ = gyr1_;
}

void IntegrationBase::Evaluate(const V3 &Pi, const Qd &Qi, const V3 &Vi, const V3 &Bai, const V3 &Bgi, const V3 &Pj, const Qd &Qj,
                               const V3 &Vj, const V3 &Baj, const V3 &Bgj, double res[15]) const {
  M3 dp_dba = get_block(jacobian_, O_P, O_BA), dp_dbg = get_block(jacobian_, O_P, O_BG);
  M3 dq_dbg = get_block(jacobian_, O_R, O_BG);
  M3 dv_dba = get_block(jacobian_, O_V, O_BA), dv_dbg = get_block(jacobian_, O_V, O_BG);
  V3 dba = Bai - linearized_ba_, dbg = Bgi - linearized_bg_;
  Qd corrected_delta_q = delta_q_ * DeltaQ(dq_dbg * dbg);
  V3 corrected_delta_v = delta_v_ + dv_dba * dba + dv_dbg * dbg;
  V3 corrected_delta_p = delta_p_ + dp_dba * dba + dp_dbg * dbg;
  V3 rP = Qi.inverse() * (-0.5 * g_vec_ * sum_dt_ * sum_dt_ + Pj - Pi - Vi * sum_dt_) - corrected_delta_p;
  V3 rR = 2.0 * (corrected_delta_q.inverse() * (Qi.inverse() * Qj)).vec();
  V3 rV = Qi.inverse() * (-g_vec_ * sum_dt_ + Vj - Vi) - corrected_delta_v;
  V3 rBa = Baj - Bai, rBg = Bgj - Bgi;
  for (int k = 0; k < 3; ++k) { res[O_P + k] = rP[k]; res[O_R + k] = rR[k]; res[O_V + k] = rV[k]; res[O_BA + k] = rBa[k]; res[O_BG + k] = rBg[k]; }
}

// ---- ImuFactor ----------------------------------------------------------------------------------
ImuFactor::ImuFactor(std::shared_ptr<IntegrationBase> pi) : pre_integration_(pi) {
  g_vec_ = pre_integration_->g_vec_;
  num_residuals = 15;
  block_sizes = {7, 9, 7, 9};
}

// sqrt_info = LLT(covariance^-1).matrixL().transpose()  (ImuFactor.h:74-75): upper U with U^T U = cov^-1
static MatX imu_sqrt_info(const MatX &cov) {
  const int n = 15;
  MatX L = cov;
  bool ok = cholesky_lower(L);  // cov = L L^T
  MatX inv(n, n);
  if (ok) {
    for (int c = 0; c < n; ++c) {
      VecX e(n, 0.0);
      e[c] = 1.0;
      chol_solve(L, e);
      for (int r = 0; r < n; ++r) inv(r, c) = e[r];
    }
    for (int i = 0; i < n; ++i) for (int j = i + 1; j < n; ++j) { double s = 0.5 * (inv(i, j) + inv(j, i)); inv(i, j) = inv(j, i) = s; }
  }
  MatX Li = inv;
  cholesky_lower(Li);
  return Li.transpose();
}

bool ImuFactor::Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
  V3 Pi(parameters[0][0], parameters[0][1], parameters[0][2]);
  Qd Qi(parameters[0][6], parameters[0][3], parameters[0][4], parameters[0][5]);
  V3 Vi(parameters[1][0], parameters[1][1], parameters[1][2]);
  V3 Bai(parameters[1][3], parameters[1][4], parameters[1][5]);
  V3 Bgi(parameters[1][6], parameters[1][7], parameters[1][8]);
  V3 Pj(parameters[2][0], parameters[2][1], parameters[2][2]);
  Qd Qj(parameters[2][6], parameters[2][3], parameters[2][4], parameters[2][5]);
  V3 Vj(parameters[3][0], parameters[3][1], parameters[3][2]);
  V3 Baj(parameters[3][3], parameters[3][4], parameters[3][5]);
  V3 Bgj(parameters[3][6], parameters[3][7], parameters[3][8]);
  double res[15];
  pre_integration_->Evaluate(Pi, Qi, Vi, Bai, Bgi, Pj, Qj, Vj, Baj, Bgj, res);
  MatX sqrt_info = imu_sqrt_info(pre_integration_->covariance_);
  for (int i = 0; i < 15; ++i) {
    double s = 0;
    for (int j = 0; j < 15; ++j) s += sqrt_info(i, j) * res[j];
    residuals[i] = s;
  }
  if (jacobians) {
    const IntegrationBase &pim = *pre_integration_;
    double sum_dt = pim.sum_dt_;
    M3 dp_dba = get_block(pim.jacobian_, O_P, O_BA), dp_dbg = get_block(pim.jacobian_, O_P, O_BG);
    M3 dq_dbg = get_block(pim.jacobian_, O_R, O_BG);
    M3 dv_dba = get_block(pim.jacobian_, O_V, O_BA), dv_dbg = get_block(pim.jacobian_, O_V, O_BG);
    Qd corrected_delta_q = pim.delta_q_ * DeltaQ(dq_dbg * (Bgi - pim.linearized_bg_));
    M3 RiT = Qi.inverse().toRotationMatrix();
    auto whiten = [&](const MatX &Jin, double *Jout, int cols) {
      for (int i = 0; i < 15; ++i)
        for (int j = 0; j < cols; ++j) {
          double s = 0;
          for (int k = 0; k < 15; ++k) s += sqrt_info(i, k) * Jin(k, j);
          Jout[i * cols + j] = s;
        }
    };
    if (jacobians[0]) {
      MatX J(15, 7);
      set_block(J, O_P, O_P, -RiT);
      set_block(J, O_P, O_R, Skew(Qi.inverse() * (-0.5 * g_vec_ * sum_dt * sum_dt + Pj - Pi - Vi * sum_dt)));
      set_block(J, O_R, O_R, -LeftTimesRightTL(Qj.inverse() * Qi, corrected_delta_q));
      set_block(J, O_V, O_R, Skew(Qi.inverse() * (-g_vec_ * sum_dt + Vj - Vi)));
      whiten(J, jacobians[0], 7);
    }
    if (jacobians[1]) {
      MatX J(15, 9);
      set_block(J, O_P, O_V - O_V, -RiT * sum_dt);
      set_block(J, O_P, O_BA - O_V, -dp_dba);
      set_block(J, O_P, O_BG - O_V, -dp_dbg);
      set_block(J, O_R, O_BG - O_V, -(LeftQuatTL(Qj.inverse() * Qi * corrected_delta_q) * dq_dbg));
      set_block(J, O_V, O_V - O_V, -RiT);
      set_block(J, O_V, O_BA - O_V, -dv_dba);
      set_block(J, O_V, O_BG - O_V, -dv_dbg);
      set_block(J, O_BA, O_BA - O_V, -M3::Identity());
      set_block(J, O_BG, O_BG - O_V, -M3::Identity());
      whiten(J, jacobians[1], 9);
    }
    if (jacobians[2]) {
      MatX J(15, 7);
      set_block(J, O_P, O_P, RiT);
      set_block(J, O_R, O_R, LeftQuatTL(corrected_delta_q.inverse() * Qi.inverse() * Qj));
      whiten(J, jacobians[2], 7);
    }
    if (jacobians[3]) {
      MatX J(15, 9);
      set_block(J, O_V, O_V - O_V, RiT);
      set_block(J, O_BA, O_BA - O_V, M3::Identity());
      set_block(J, O_BG, O_BG - O_V, M3::Identity());
      whiten(J, jacobians[3], 9);
    }
  }
  return true;
}

// ---- PriorFactor --------------------------------------------------------------------------------
PriorFactor::PriorFactor(const V3 &pos, const Qd &rot) : pos_(pos), rot_(rot) {
  num_residuals = 6;
  block_sizes = {7};
}
bool PriorFactor::Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
  V3 P(parameters[0][0], parameters[0][1], parameters[0][2]);
  Qd Q(parameters[0][6], parameters[0][3], parameters[0][4], parameters[0][5]);
  const double wp = 1000.0, wr = 0.1;
  V3 rp = P - pos_;
  V3 rr = 2.0 * (rot_.inverse() * Q).vec();
  for (int k = 0; k < 3; ++k) { residuals[k] = wp * rp[k]; residuals[3 + k] = wr * rr[k]; }
  if (jacobians && jacobians[0]) {
    double *J = jacobians[0];
    for (int k = 0; k < 42; ++k) J[k] = 0;
    M3 br = LeftQuatTL(Q.inverse() * rot_);
    for (int i = 0; i < 3; ++i) {
      J[i * 7 + i] = wp;
      for (int j = 0; j < 3; ++j) J[(3 + i) * 7 + 3 + j] = wr * br(i, j);
    }
  }
  return true;
}

void PosePlus(const double *x, const double *delta, double *x_plus_delta) {
  Qd q(x[6], x[3], x[4], x[5]);
  Qd dq = DeltaQ(V3(delta[3], delta[4], delta[5]));
  Qd qp = (q * dq).normalized();
  for (int k = 0; k < 3; ++k) x_plus_delta[k] = x[k] + delta[k];
  x_plus_delta[3] = qp.x; x_plus_delta[4] = qp.y; x_plus_delta[5] = qp.z; x_plus_delta[6] = qp.w;
}

// ---- symmetric eigen-solver (Householder tridiagonalisation + implicit QL) ----------------------
void SymEigen(const MatX &A, VecX &d, MatX &V) {
  const int n = A.r;
  V = A;
  d.assign(n, 0.0);
  VecX e(n, 0.0);
  if (n == 0) return;
  for (int j = 0; j < n; ++j) d[j] = V(n - 1, j);
  for (int i = n - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) { d[j] = V(i - 1, j); V(i, j) = 0.0; V(j, i) = 0.0; }
    } else {
      for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1];
      double g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h = h - f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j];
        V(j, i) = f;
        g = e[j] + V(j, j) * f;
        for (int k = j + 1; k <= i - 1; ++k) { g += V(k, j) * d[k]; e[k] += V(k, j) * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
      double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j]; g = e[j];
        for (int k = j; k <= i - 1; ++k) V(k, j) -= (f * e[k] + g * d[k]);
        d[j] = V(i - 1, j);
        V(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; ++i) {
    V(n - 1, i) = V(i, i);
    V(i, i) = 1.0;
    double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; ++k) d[k] = V(k, i + 1) / h;
      for (int j = 0; j <= i; ++j) {
        double g = 0.0;
        for (int k = 0; k <= i; ++k) g += V(k, i + 1) * V(k, j);
        for (int k = 0; k <= i; ++k) V(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) V(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; ++j) { d[j] = V(n - 1, j); V(n - 1, j) = 0.0; }
  V(n - 1, n - 1) = 1.0;
  e[0] = 0.0;
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = std::pow(2.0, -52.0);
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) { if (std::fabs(e[m]) <= eps * tst1) break; ++m; }
    if (m > l) {
      int iter = 0;
      do {
        ++iter;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c, el1 = e[l + 1], s = 0.0, s2 = 0.0;
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
            h = V(k, i + 1);
            V(k, i + 1) = s * V(k, i) + c * h;
            V(k, i) = c * V(k, i) - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] = d[l] + f;
    e[l] = 0.0;
  }
  for (int i = 0; i < n - 1; ++i) {
    int k = i;
    double p = d[i];
    for (int j = i + 1; j < n; ++j) if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i]; d[i] = p;
      for (int j = 0; j < n; ++j) { double t = V(j, i); V(j, i) = V(j, k); V(j, k) = t; }
    }
  }
}

// ---- marginalisation ----------------------------------------------------------------------------
void ResidualBlockInfo::Evaluate() {
  const int nr = cost_function->num_residuals;
  residuals.assign(nr, 0.0);
  const std::vector<int> &bs = cost_function->block_sizes;
  jacobians.clear();
  std::vector<double *> raw(bs.size());
  for (size_t i = 0; i < bs.size(); ++i) jacobians.push_back(MatX(nr, bs[i]));
  for (size_t i = 0; i < bs.size(); ++i) raw[i] = jacobians[i].d.data();
  cost_function->Evaluate(parameter_blocks.data(), residuals.data(), raw.data());
  if (loss_function) {  // :69-95 (ceres corrector restated by the reference)
    double residual_scaling_, alpha_sq_norm_;
    double sq_norm = 0, rho[3];
    for (double v : residuals) sq_norm += v * v;
    loss_function->Evaluate(sq_norm, rho);
    double sqrt_rho1_ = std::sqrt(rho[1]);
    if ((sq_norm == 0.0) || (rho[2] <= 0.0)) {
      residual_scaling_ = sqrt_rho1_;
      alpha_sq_norm_ = 0.0;
    } else {
      const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
      const double alpha = 1.0 - std::sqrt(D);
      residual_scaling_ = sqrt_rho1_ / (1 - alpha);
      alpha_sq_norm_ = alpha / sq_norm;
    }
    for (size_t i = 0; i < jacobians.size(); ++i) {
      MatX &J = jacobians[i];
      VecX rtJ(J.c, 0.0);
      for (int c = 0; c < J.c; ++c) { double s = 0; for (int r = 0; r < nr; ++r) s += residuals[r] * J(r, c); rtJ[c] = s; }
      for (int r = 0; r < nr; ++r)
        for (int c = 0; c < J.c; ++c) J(r, c) = sqrt_rho1_ * (J(r, c) - alpha_sq_norm_ * residuals[r] * rtJ[c]);
    }
    for (double &v : residuals) v *= residual_scaling_;
  }
}

void MarginalizationInfo::AddResidualBlockInfo(std::shared_ptr<ResidualBlockInfo> rbi) {
  factors.push_back(rbi);
  const std::vector<int> &sizes = rbi->cost_function->block_sizes;
  for (size_t i = 0; i < rbi->parameter_blocks.size(); ++i)
    parameter_block_size[reinterpret_cast<long>(rbi->parameter_blocks[i])] = sizes[i];
  for (size_t i = 0; i < rbi->drop_set.size(); ++i)
    parameter_block_idx[reinterpret_cast<long>(rbi->parameter_blocks[rbi->drop_set[i]])] = 0;
}

void MarginalizationInfo::PreMarginalize() {
  for (auto &it : factors) {
    it->Evaluate();
    const std::vector<int> &bs = it->cost_function->block_sizes;
    for (size_t i = 0; i < bs.size(); ++i) {
      long addr = reinterpret_cast<long>(it->parameter_blocks[i]);
      if (parameter_block_data.find(addr) == parameter_block_data.end())
        parameter_block_data[addr] = std::vector<double>(it->parameter_blocks[i], it->parameter_blocks[i] + bs[i]);
    }
  }
}

namespace {
struct ThreadsStruct {
  std::vector<std::shared_ptr<ResidualBlockInfo>> sub_factors;
  MatX A;
  VecX b;
  const std::map<long, int> *parameter_block_size, *parameter_block_idx;
};
// ThreadsConstructA (:157-183)
void ConstructA(ThreadsStruct *p) {
  for (auto &it : p->sub_factors) {
    const int nr = it->cost_function->num_residuals;
    for (size_t i = 0; i < it->parameter_blocks.size(); ++i) {
      int idx_i = p->parameter_block_idx->at(reinterpret_cast<long>(it->parameter_blocks[i]));
      int size_i = MarginalizationInfo::LocalSize(p->parameter_block_size->at(reinterpret_cast<long>(it->parameter_blocks[i])));
      const MatX &Ji = it->jacobians[i];
      for (size_t j = i; j < it->parameter_blocks.size(); ++j) {
        int idx_j = p->parameter_block_idx->at(reinterpret_cast<long>(it->parameter_blocks[j]));
        int size_j = MarginalizationInfo::LocalSize(p->parameter_block_size->at(reinterpret_cast<long>(it->parameter_blocks[j])));
        const MatX &Jj = it->jacobians[j];
        for (int a = 0; a < size_i; ++a)
          for (int c = 0; c < size_j; ++c) {
            double s = 0;
            for (int r = 0; r < nr; ++r) s += Ji(r, a) * Jj(r, c);
            p->A(idx_i + a, idx_j + c) += s;
          }
        if (i != j)
          for (int a = 0; a < size_i; ++a)
            for (int c = 0; c < size_j; ++c) p->A(idx_j + c, idx_i + a) = p->A(idx_i + a, idx_j + c);
      }
      for (int a = 0; a < size_i; ++a) {
        double s = 0;
        for (int r = 0; r < nr; ++r) s += Ji(r, a) * it->residuals[r];
        p->b[idx_i + a] += s;
      }
    }
  }
}
}  // namespace

void MarginalizationInfo::Marginalize() {
  int pos = 0;
  for (auto &it : parameter_block_idx) { it.second = pos; pos += LocalSize(parameter_block_size[it.first]); }
  m = pos;
  for (const auto &it : parameter_block_size)
    if (parameter_block_idx.find(it.first) == parameter_block_idx.end()) { parameter_block_idx[it.first] = pos; pos += LocalSize(it.second); }
  n = pos - m;
  MatX A(pos, pos);
  VecX b(pos, 0.0);
  const int NUM_THREADS = 4;  // include/factor/MarginalizationFactor.h:50
  ThreadsStruct ts[NUM_THREADS];
  int i = 0;
  for (auto &it : factors) { ts[i].sub_factors.push_back(it); i = (i + 1) % NUM_THREADS; }
  std::vector<std::thread> th;
  for (int t = 0; t < NUM_THREADS; ++t) {
    ts[t].A = MatX(pos, pos);
    ts[t].b.assign(pos, 0.0);
    ts[t].parameter_block_size = &parameter_block_size;
    ts[t].parameter_block_idx = &parameter_block_idx;
    th.emplace_back(ConstructA, &ts[t]);
  }
  for (int t = NUM_THREADS - 1; t >= 0; --t) {
    th[t].join();
    for (size_t k = 0; k < A.d.size(); ++k) A.d[k] += ts[t].A.d[k];
    for (int k = 0; k < pos; ++k) b[k] += ts[t].b[k];
  }
  A_dbg = A;
  b_dbg = b;
  // Schur complement with eigen pseudo-inverse (:270-289)
  MatX Amm(m, m);
  for (int r = 0; r < m; ++r) for (int c = 0; c < m; ++c) Amm(r, c) = 0.5 * (A(r, c) + A(c, r));
  VecX ev;
  MatX evec;
  SymEigen(Amm, ev, evec);
  MatX Amm_inv(m, m);
  for (int r = 0; r < m; ++r)
    for (int c = 0; c < m; ++c) {
      double s = 0;
      for (int k = 0; k < m; ++k) s += evec(r, k) * (ev[k] > eps ? 1.0 / ev[k] : 0.0) * evec(c, k);
      Amm_inv(r, c) = s;
    }
  MatX Amr(m, n), Arm(n, m), Arr(n, n);
  VecX bmm(b.begin(), b.begin() + m), brr(b.begin() + m, b.end());
  for (int r = 0; r < m; ++r) for (int c = 0; c < n; ++c) { Amr(r, c) = A(r, m + c); Arm(c, r) = A(m + c, r); }
  for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) Arr(r, c) = A(m + r, m + c);
  MatX T = matmul(Arm, Amm_inv);
  MatX TA = matmul(T, Amr);
  MatX A2(n, n);
  for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) A2(r, c) = Arr(r, c) - TA(r, c);
  VecX Tb = matvec(T, bmm);
  VecX b2(n);
  for (int r = 0; r < n; ++r) b2[r] = brr[r] - Tb[r];
  VecX ev2;
  MatX evec2;
  SymEigen(A2, ev2, evec2);
  linearized_jacobians = MatX(n, n);
  linearized_residuals.assign(n, 0.0);
  for (int k = 0; k < n; ++k) {
    double S = ev2[k] > eps ? ev2[k] : 0.0;
    double S_inv = ev2[k] > eps ? 1.0 / ev2[k] : 0.0;
    double ss = std::sqrt(S), sis = std::sqrt(S_inv);
    double vb = 0;
    for (int r = 0; r < n; ++r) { linearized_jacobians(k, r) = ss * evec2(r, k); vb += evec2(r, k) * b2[r]; }
    linearized_residuals[k] = sis * vb;
  }
}

std::vector<double *> MarginalizationInfo::GetParameterBlocks(std::map<long, double *> &addr_shift) {
  std::vector<double *> keep_block_addr;
  keep_block_size.clear(); keep_block_idx.clear(); keep_block_data.clear();
  for (const auto &it : parameter_block_idx) {
    if (it.second >= m) {
      keep_block_size.push_back(parameter_block_size[it.first]);
      keep_block_idx.push_back(parameter_block_idx[it.first]);
      keep_block_data.push_back(parameter_block_data[it.first]);
      keep_block_addr.push_back(addr_shift[it.first]);
    }
  }
  return keep_block_addr;
}

MarginalizationFactor::MarginalizationFactor(std::shared_ptr<MarginalizationInfo> mi) : marginalization_info(mi) {
  for (int s : mi->keep_block_size) block_sizes.push_back(s);
  num_residuals = mi->n;
}

bool MarginalizationFactor::Evaluate(double const *const *parameters, double *residuals, double **jacobians) const {
  const MarginalizationInfo &mi = *marginalization_info;
  int n = mi.n, m = mi.m;
  VecX dx(n, 0.0);
  for (size_t i = 0; i < mi.keep_block_size.size(); ++i) {
    int size = mi.keep_block_size[i];
    int idx = mi.keep_block_idx[i] - m;
    const double *x = parameters[i];
    const double *x0 = mi.keep_block_data[i].data();
    if (size != 7) {
      for (int k = 0; k < size; ++k) dx[idx + k] = x[k] - x0[k];
    } else {
      for (int k = 0; k < 3; ++k) dx[idx + k] = x[k] - x0[k];
      Qd q0(x0[6], x0[3], x0[4], x0[5]), q(x[6], x[3], x[4], x[5]);
      Qd dq = q0.inverse() * q;
      V3 v = dq.normalized().vec() * 2.0;
      if (dq.w < 0) v = -v;
      for (int k = 0; k < 3; ++k) dx[idx + 3 + k] = v[k];
    }
  }
  for (int r = 0; r < n; ++r) {
    double s = mi.linearized_residuals[r];
    for (int c = 0; c < n; ++c) s += mi.linearized_jacobians(r, c) * dx[c];
    residuals[r] = s;
  }
  if (jacobians) {
    for (size_t i = 0; i < mi.keep_block_size.size(); ++i) {
      if (!jacobians[i]) continue;
      int size = mi.keep_block_size[i], local_size = MarginalizationInfo::LocalSize(size);
      int idx = mi.keep_block_idx[i] - m;
      double *J = jacobians[i];
      for (int r = 0; r < n; ++r)
        for (int c = 0; c < size; ++c) J[r * size + c] = (c < local_size) ? mi.linearized_jacobians(r, idx + c) : 0.0;
    }
  }
  return true;
}

}  // namespace orc
