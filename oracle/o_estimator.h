// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
// Sliding-window estimator (steady state, stage_flag_ == INITED) restated from
// src/imu_processor/Estimator.cc: ProcessImu :338-427, ProcessLaserOdom :430-774 (INITED branch),
// TransformToEnd :62-103, BuildLocalMap :1361-1646, SolveOptimization :1648-2438,
// VectorToDouble / DoubleToVector :2440-2568, SlideWindow :2570-2666.
// ROS I/O, visualisation, IMU initialisation and the cube-map database are out of scope.
#pragma once
#include "o_api.h"
#include "o_solver.h"

namespace orc {

struct EstimatorConfig {  // include/imu_processor/Estimator.h:77-108 (lidar/solver subset)
  int window_size = 10, opt_window_size = 10;
  StageBConfig b;
  int estimate_extrinsic = 1;
  bool opt_extrinsic = true;
  bool imu_factor = true, point_distance_factor = true, prior_factor = false, marginalization_factor = true;
  bool enable_deskew = true, cutoff_deskew = true;
  IntegrationBaseConfig pim;
  SolverOptions solver;
};

struct ImuStamped { double time; Transform transform; };

struct Estimator {
  EstimatorConfig cfg;
  int W = 0, O = 0;
  std::vector<V3> Ps, Vs, Bas, Bgs;
  std::vector<M3> Rs;
  std::vector<std::shared_ptr<IntegrationBase>> pre_integrations;
  std::vector<Cloud> surf_stack;
  std::vector<int> size_surf_stack;
  std::shared_ptr<IntegrationBase> tmp_pre_integration;
  V3 acc_last, gyr_last, g_vec;
  bool first_imu = false;
  Transform transform_lb;  // float
  std::vector<ImuStamped> imu_stampedtransforms;
  Transform transform_es;
  // optimisation parameter arrays (Estimator.h:282-284)
  // one contiguous block [pose_0..pose_O | sb_0..sb_O | ex]: the marginalisation orders parameter blocks by
  // address (std::map), so this fixes the prior's block order to poses, speed-biases, extrinsic
  std::vector<double> para_storage;
  std::vector<double *> para_pose, para_speed_bias;
  double *para_ex_pose = nullptr;
  std::shared_ptr<MarginalizationInfo> last_marginalization_info;
  std::vector<double *> last_marginalization_parameter_blocks;
  bool convergence_flag = false, init_local_map = false;
  int extrinsic_stage = 1;
  CauchyLoss loss{1.0};
  // results of the last SolveOptimization (inspection)
  Cloud local_surf_points, local_surf_points_filtered;
  std::vector<std::vector<PointPlaneFeature>> feature_frames;
  std::vector<Transform> local_transforms;
  SolverSummary summary;
  double cost_pim = 0, cost_ppp = 0, cost_marg = 0;
  bool turn_off = true;
  int laser_odom_iters = 0;
  // timing probes (seconds) for the cpu_baseline leg
  double t_build_map = 0, t_features = 0, t_solve = 0, t_marg = 0, t_total = 0;

  explicit Estimator(const EstimatorConfig &c);
  // warm start: frames 0..W-1 (state + own down-sampled surf cloud + pre-integration ending at the frame)
  void InitFrame(int k, const V3 &P, const Qd &Q, const V3 &V, const V3 &Ba, const V3 &Bg, const Cloud &surf_ds,
                 std::shared_ptr<IntegrationBase> pim);
  void FinishInit(const V3 &acc_last, const V3 &gyr_last);
  void ProcessImu(double dt, const V3 &acc, const V3 &gyr, double stamp);
  void ProcessScan(const Cloud &laser_cloud_surf_last);  // ProcessLaserOdom, INITED branch
  void BuildLocalMap();
  void SolveOptimization();
  void SlideWindow();
  void VectorToDouble();
  void DoubleToVector();
};

size_t TransformToEnd(Cloud &cloud, const Transform &transform_es, float time_factor);

}  // namespace orc
