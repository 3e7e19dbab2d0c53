// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
// fp64 factors of the sliding-window problem, restated from the reference with the Ceres
// CostFunction calling convention (row-major num_residuals x global_size Jacobian blocks).
#pragma once
#include "o_linalg.h"
#include <map>
#include <memory>
#include <vector>

namespace orc {

typedef Vec3<double> V3;
typedef Mat3<double> M3;
typedef Quat<double> Qd;

struct CostFunction {  // ceres::CostFunction contract (SURVEY.md §8b "Factor operator")
  int num_residuals = 0;
  std::vector<int> block_sizes;
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
};

struct CauchyLoss {  // ceres::CauchyLoss(a) (Estimator.cc:1664 uses a = 1.0)
  double b, c;
  explicit CauchyLoss(double a) : b(a * a), c(1.0 / (a * a)) {}
  void Evaluate(double s, double rho[3]) const {
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c * (inv * inv);
  }
};

// src/factor/PivotPointPlaneFactor.cc:43-137
struct PivotPointPlaneFactor : CostFunction {
  V3 point;
  double coeff[4];
  PivotPointPlaneFactor(const double p[3], const double c[4]);
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};

struct IntegrationBaseConfig {  // include/imu_processor/IntegrationBase.h:64-70
  double acc_n = 0.1, gyr_n = 0.01, acc_w = 0.0002, gyr_w = 2.0e-5, g_norm = 9.805;
};

// include/imu_processor/IntegrationBase.h:72-388 (mid-point pre-integration)
struct IntegrationBase {
  double dt_ = 0;
  V3 acc0_, gyr0_, acc1_, gyr1_;
  V3 linearized_acc_, linearized_gyr_;
  V3 linearized_ba_, linearized_bg_;
  MatX jacobian_, covariance_, noise_;  // 15x15, 15x15, 18x18
  double sum_dt_ = 0;
  V3 delta_p_, delta_v_;
  Qd delta_q_;
  std::vector<double> dt_buf_;
  std::vector<V3> acc_buf_, gyr_buf_;
  IntegrationBaseConfig config_;
  V3 g_vec_;
  IntegrationBase(const V3 &acc0, const V3 &gyr0, const V3 &ba, const V3 &bg, const IntegrationBaseConfig &cfg);
  void push_back(double dt, const V3 &acc, const V3 &gyr);
  void Repropagate(const V3 &ba, const V3 &bg);
  void Propagate(double dt, const V3 &acc1, const V3 &gyr1);
  void MidPointIntegration(double dt, const V3 &acc0, const V3 &gyr0, const V3 &acc1, const V3 &gyr1, const V3 &delta_p,
                           const Qd &delta_q, const V3 &delta_v, const V3 &lba, const V3 &lbg, V3 &rp, Qd &rq, V3 &rv,
                           bool update_jacobian);
  void Evaluate(const V3 &Pi, const Qd &Qi, const V3 &Vi, const V3 &Bai, const V3 &Bgi, const V3 &Pj, const Qd &Qj,
                const V3 &Vj, const V3 &Baj, const V3 &Bgj, double res[15]) const;
};

// include/factor/ImuFactor.h:44-177
struct ImuFactor : CostFunction {
  std::shared_ptr<IntegrationBase> pre_integration_;
  V3 g_vec_;
  explicit ImuFactor(std::shared_ptr<IntegrationBase> pi);
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};

// src/factor/PriorFactor.cc:35-67
struct PriorFactor : CostFunction {
  V3 pos_;
  Qd rot_;
  PriorFactor(const V3 &pos, const Qd &rot);
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};

// src/factor/PoseLocalParameterization.cc:35-52
void PosePlus(const double *x, const double *delta, double *x_plus_delta);

// src/factor/MarginalizationFactor.cc (ResidualBlockInfo :37-96, MarginalizationInfo :98-341)
struct ResidualBlockInfo {
  std::shared_ptr<CostFunction> cost_function;
  const CauchyLoss *loss_function;  // may be null
  std::vector<double *> parameter_blocks;
  std::vector<int> drop_set;
  std::vector<MatX> jacobians;  // row-major num_res x block size
  VecX residuals;
  ResidualBlockInfo(std::shared_ptr<CostFunction> cf, const CauchyLoss *loss, std::vector<double *> pb, std::vector<int> ds)
      : cost_function(cf), loss_function(loss), parameter_blocks(pb), drop_set(ds) {}
  void Evaluate();
};

struct MarginalizationInfo {
  std::vector<std::shared_ptr<ResidualBlockInfo>> factors;
  int m = 0, n = 0;
  // the reference keys these by pointer value in std::unordered_map<long,...> (iteration order =
  // parameter ordering, unspecified); the oracle uses an ordered map for determinism.
  std::map<long, int> parameter_block_size;
  std::map<long, int> parameter_block_idx;
  std::map<long, std::vector<double>> parameter_block_data;
  std::vector<int> keep_block_size, keep_block_idx;
  std::vector<std::vector<double>> keep_block_data;
  MatX linearized_jacobians;
  VecX linearized_residuals;
  MatX A_dbg;  // the assembled (pos x pos) A before the Schur complement (test aid)
  VecX b_dbg;
  const double eps = 1e-8;
  void AddResidualBlockInfo(std::shared_ptr<ResidualBlockInfo> rbi);
  void PreMarginalize();
  void Marginalize();
  std::vector<double *> GetParameterBlocks(std::map<long, double *> &addr_shift);
  static int LocalSize(int size) { return size == 7 ? 6 : size; }
};

struct MarginalizationFactor : CostFunction {
  std::shared_ptr<MarginalizationInfo> marginalization_info;
  explicit MarginalizationFactor(std::shared_ptr<MarginalizationInfo> mi);
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override;
};

// Dense symmetric eigen-decomposition used where the reference calls Eigen::SelfAdjointEigenSolver
// on MatrixXd (MarginalizationFactor.cc:276,293): ascending eigenvalues, eigenvectors in columns.
void SymEigen(const MatX &A, VecX &evals, MatX &evecs);

}  // namespace orc
