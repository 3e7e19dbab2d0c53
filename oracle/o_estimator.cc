// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
// Steady-state sliding-window estimator restated from src/imu_processor/Estimator.cc (see
// o_estimator.h for the line map).
#include "o_estimator.h"
#include <chrono>
#include <cmath>

namespace orc {

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// mathutils::R2ypr / ypr2R (include/utils/math_utils.h:188-232), degrees
static V3 R2ypr(const M3 &R) {
  V3 n = R.col(0), o = R.col(1), a = R.col(2);
  double y = std::atan2(n.y, n.x);
  double p = std::atan2(-n.z, n.x * std::cos(y) + n.y * std::sin(y));
  double r = std::atan2(a.x * std::sin(y) - a.y * std::cos(y), -o.x * std::sin(y) + o.y * std::cos(y));
  return V3(y, p, r) / M_PI * 180.0;
}
static M3 ypr2R(const V3 &ypr) {
  double y = ypr.x / 180.0 * M_PI, p = ypr.y / 180.0 * M_PI, r = ypr.z / 180.0 * M_PI;
  M3 Rz, Ry, Rx;
  Rz(0, 0) = std::cos(y); Rz(0, 1) = -std::sin(y); Rz(1, 0) = std::sin(y); Rz(1, 1) = std::cos(y); Rz(2, 2) = 1;
  Ry(0, 0) = std::cos(p); Ry(0, 2) = std::sin(p); Ry(1, 1) = 1; Ry(2, 0) = -std::sin(p); Ry(2, 2) = std::cos(p);
  Rx(0, 0) = 1; Rx(1, 1) = std::cos(r); Rx(1, 2) = -std::sin(r); Rx(2, 1) = std::sin(r); Rx(2, 2) = std::cos(r);
  return Rz * Ry * Rx;
}

// Estimator.cc:62-103
size_t TransformToEnd(Cloud &cloud, const Transform &transform_es, float time_factor) {
  size_t cloud_size = cloud.size();
  for (size_t i = 0; i < cloud_size; i++) {
    PointXYZI &point = cloud[i];
    float s = time_factor * (point.intensity - int(point.intensity));
    point.x -= s * transform_es.pos.x;
    point.y -= s * transform_es.pos.y;
    point.z -= s * transform_es.pos.z;
    point.intensity -= int(point.intensity);
    Quat<float> q_id, q_e = transform_es.rot;
    Quat<float> q_s = q_id.slerp(s, q_e);
    Vec3<float> v(point.x, point.y, point.z);
    v = q_s.conjugate().normalized() * v;
    v = q_e * v;
    point.x = v.x + transform_es.pos.x;
    point.y = v.y + transform_es.pos.y;
    point.z = v.z + transform_es.pos.z;
  }
  return cloud_size;
}

Estimator::Estimator(const EstimatorConfig &c) : cfg(c) {
  W = cfg.window_size; O = cfg.opt_window_size;
  Ps.assign(W + 1, V3()); Vs.assign(W + 1, V3()); Bas.assign(W + 1, V3()); Bgs.assign(W + 1, V3());
  Rs.assign(W + 1, M3::Identity());
  pre_integrations.assign(W + 1, nullptr);
  surf_stack.assign(W + 1, Cloud());
  size_surf_stack.assign(W + 1, 0);
  para_storage.assign(16 * (O + 1) + 7, 0.0);
  para_pose.resize(O + 1); para_speed_bias.resize(O + 1);
  for (int i = 0; i <= O; ++i) { para_pose[i] = &para_storage[7 * i]; para_speed_bias[i] = &para_storage[7 * (O + 1) + 9 * i]; }
  para_ex_pose = &para_storage[16 * (O + 1)];
  g_vec = V3(0, 0, -cfg.pim.g_norm);
  extrinsic_stage = cfg.estimate_extrinsic;
  transform_lb = Transform(Quat<float>(1, 0, 0, 0), Vec3<float>(0, 0, -0.1f));  // Estimator.h:89
}

void Estimator::InitFrame(int k, const V3 &P, const Qd &Q, const V3 &V, const V3 &Ba, const V3 &Bg, const Cloud &surf_ds,
                          std::shared_ptr<IntegrationBase> pim) {
  // frames 0..W-1 land one slot to the right for clouds / pre-integrations: the first ProcessScan
  // pushes (drops slot 0) exactly like the reference's CircularBuffers.
  Ps[k] = P; Rs[k] = Q.normalized().toRotationMatrix(); Vs[k] = V; Bas[k] = Ba; Bgs[k] = Bg;
  surf_stack[k + 1] = surf_ds;
  size_surf_stack[k + 1] = (int)surf_ds.size();
  pre_integrations[k + 1] = pim;
}

void Estimator::FinishInit(const V3 &a_last, const V3 &g_last) {
  Ps[W] = Ps[W - 1]; Rs[W] = Rs[W - 1]; Vs[W] = Vs[W - 1]; Bas[W] = Bas[W - 1]; Bgs[W] = Bgs[W - 1];  // SlideWindow :2651-2655
  acc_last = a_last; gyr_last = g_last;
  first_imu = true;
  tmp_pre_integration = std::make_shared<IntegrationBase>(acc_last, gyr_last, Bas[W], Bgs[W], cfg.pim);
  imu_stampedtransforms.clear();
}

void Estimator::ProcessImu(double dt, const V3 &acc, const V3 &gyr, double stamp) {
  if (!first_imu) { first_imu = true; acc_last = acc; gyr_last = gyr; }
  const int j = W;
  tmp_pre_integration->push_back(dt, acc, gyr);
  V3 un_acc_0 = Rs[j] * (acc_last - Bas[j]) + g_vec;
  V3 un_gyr = 0.5 * (gyr_last + gyr) - Bgs[j];
  Rs[j] = Rs[j] * DeltaQ(un_gyr * dt).toRotationMatrix();
  V3 un_acc_1 = Rs[j] * (acc - Bas[j]) + g_vec;
  V3 un_acc = 0.5 * (un_acc_0 + un_acc_1);
  Ps[j] += dt * Vs[j] + 0.5 * dt * dt * un_acc;
  Vs[j] += dt * un_acc;
  ImuStamped tt;
  tt.time = stamp;
  tt.transform.pos = Ps[j].cast<float>();
  tt.transform.rot = Quat<float>::fromRotationMatrix(Rs[j].cast<float>());
  imu_stampedtransforms.push_back(tt);
  if (imu_stampedtransforms.size() > 100) imu_stampedtransforms.erase(imu_stampedtransforms.begin());  // CircularBuffer
  acc_last = acc; gyr_last = gyr;
}

template <typename T> static void push_shift(std::vector<T> &v, T x) { v.erase(v.begin()); v.push_back(x); }

void Estimator::ProcessScan(const Cloud &surf_last_in) {
  double t0 = now_s();
  push_shift(pre_integrations, tmp_pre_integration);
  tmp_pre_integration = std::make_shared<IntegrationBase>(acc_last, gyr_last, Bas[W], Bgs[W], cfg.pim);
  Cloud surf_last = surf_last_in;
  if (cfg.enable_deskew || cfg.cutoff_deskew) {  // :628-693
    if (!imu_stampedtransforms.empty()) {
      double time_e = imu_stampedtransforms.back().time;
      Transform transform_e = imu_stampedtransforms.back().transform;
      double time_s = time_e;
      Transform transform_s = transform_e;
      for (int i = int(imu_stampedtransforms.size()) - 1; i >= 0; --i) {
        time_s = imu_stampedtransforms[i].time;
        transform_s = imu_stampedtransforms[i].transform;
        if (time_e - imu_stampedtransforms[i].time >= 0.1) break;
      }
      Transform transform_body_es = transform_e.inverse() * transform_s;
      {
        float s = 0.1 / (time_e - time_s);
        Quat<float> q_id, q_e = transform_body_es.rot;
        transform_body_es.rot = q_id.slerp(s, q_e);
        transform_body_es.pos = s * transform_body_es.pos;
      }
      transform_es = transform_lb * transform_body_es * transform_lb.inverse();
      if (!cfg.cutoff_deskew) TransformToEnd(surf_last, transform_es, 10);
    }
  }
  Cloud ds;
  VoxelGridFilter(surf_last, cfg.b.surf_filter_size, ds);
  push_shift(surf_stack, ds);
  push_shift(size_surf_stack, (int)ds.size());
  SolveOptimization();
  SlideWindow();
  t_total = now_s() - t0;
}

static Twist<double> lidar_pose(const V3 &P, const M3 &R, const Twist<double> &transform_lb) {
  // Quaterniond rot_li(Rs_i * transform_lb.rot.inverse()); pos_li = Ps_i - rot_li * transform_lb.pos  (:1387-1390)
  Qd rot = Qd::fromRotationMatrix(R * transform_lb.rot.inverse().toRotationMatrix());
  V3 pos = P - rot * transform_lb.pos;
  return Twist<double>(rot, pos);
}

void Estimator::BuildLocalMap() {
  double t0 = now_s();
  feature_frames.assign(W + 1, std::vector<PointPlaneFeature>());
  local_surf_points.clear();
  local_surf_points_filtered.clear();
  local_transforms.clear();
  const int pivot_idx = W - O;
  Twist<double> tlb = transform_lb.cast<double>();
  Twist<double> transform_pivot = lidar_pose(Ps[pivot_idx], Rs[pivot_idx], tlb);
  if (!init_local_map) {  // :1409-1441
    Cloud tmp;
    for (int i = 0; i <= pivot_idx; ++i) {
      Twist<double> transform_li = lidar_pose(Ps[i], Rs[i], tlb);
      Twist<float> tpi = (transform_pivot.inverse() * transform_li).cast<float>();
      Cloud tc;
      TransformCloudAffine(surf_stack[i], tpi.linear(), tpi.pos, tc);
      tmp.insert(tmp.end(), tc.begin(), tc.end());
    }
    surf_stack[pivot_idx] = tmp;
    init_local_map = true;
  }
  for (int i = 0; i < W + 1; ++i) {
    Twist<double> transform_li = lidar_pose(Ps[i], Rs[i], tlb);
    Twist<float> tpi = (transform_pivot.inverse() * transform_li).cast<float>();
    Mat3<float> lin = tpi.linear();
    Transform local_transform = Transform::fromAffine(lin, tpi.pos);  // Twist(Affine3f)
    local_transforms.push_back(local_transform);
    if (i < pivot_idx) continue;
    if (i != W) {
      if (i == pivot_idx) {
        local_surf_points.insert(local_surf_points.end(), surf_stack[i].begin(), surf_stack[i].end());
        continue;
      }
      Cloud tc;
      TransformCloudAffine(surf_stack[i], lin, tpi.pos, tc);
      for (PointXYZI &p : tc) p.intensity = (float)i;
      local_surf_points.insert(local_surf_points.end(), tc.begin(), tc.end());
    }
  }
  VoxelGridFilter(local_surf_points, cfg.b.surf_filter_size, local_surf_points_filtered);
  KdTree kd;
  kd.Build(local_surf_points_filtered);
  t_build_map = now_s() - t0;
  double t1 = now_s();
  for (int idx = 0; idx < W + 1; ++idx) {
    std::vector<PointPlaneFeature> features;
    if (idx > pivot_idx) {
      if (idx != W || !cfg.imu_factor) {
        CalculateFeatures(kd, local_surf_points_filtered, surf_stack[idx], local_transforms[idx], cfg.b, features);
      } else {
        CalculateLaserOdom(kd, local_surf_points_filtered, surf_stack[idx], local_transforms[idx], cfg.b, features, &laser_odom_iters);
      }
    }
    feature_frames[idx] = features;
  }
  t_features = now_s() - t1;
}

void Estimator::VectorToDouble() {
  const int pivot_idx = W - O;
  for (int i = 0, opt_i = pivot_idx; i < O + 1; ++i, ++opt_i) {
    para_pose[i][0] = Ps[opt_i].x; para_pose[i][1] = Ps[opt_i].y; para_pose[i][2] = Ps[opt_i].z;
    Qd q = Qd::fromRotationMatrix(Rs[opt_i]);
    para_pose[i][3] = q.x; para_pose[i][4] = q.y; para_pose[i][5] = q.z; para_pose[i][6] = q.w;
    for (int k = 0; k < 3; ++k) { para_speed_bias[i][k] = Vs[opt_i][k]; para_speed_bias[i][3 + k] = Bas[opt_i][k]; para_speed_bias[i][6 + k] = Bgs[opt_i][k]; }
  }
  para_ex_pose[0] = transform_lb.pos.x; para_ex_pose[1] = transform_lb.pos.y; para_ex_pose[2] = transform_lb.pos.z;
  para_ex_pose[3] = transform_lb.rot.x; para_ex_pose[4] = transform_lb.rot.y; para_ex_pose[5] = transform_lb.rot.z;
  para_ex_pose[6] = transform_lb.rot.w;
}

void Estimator::DoubleToVector() {
  const int pivot_idx = W - O;
  V3 origin_P0 = Ps[pivot_idx];
  V3 origin_R0 = R2ypr(Rs[pivot_idx]);
  auto qpose = [&](int i) { return Qd(para_pose[i][6], para_pose[i][3], para_pose[i][4], para_pose[i][5]).normalized().toRotationMatrix(); };
  V3 origin_R00 = R2ypr(qpose(0));
  double y_diff = origin_R0.x - origin_R00.x;
  M3 rot_diff = ypr2R(V3(y_diff, 0, 0));
  if (std::fabs(std::fabs(origin_R0.y) - 90) < 1.0 || std::fabs(std::fabs(origin_R00.y) - 90) < 1.0)
    rot_diff = Rs[pivot_idx] * qpose(0).transpose();
  {
    Twist<double> trans_pivot(Qd::fromRotationMatrix(Rs[pivot_idx]), Ps[pivot_idx]);
    M3 R_opt_pivot = rot_diff * qpose(0);
    Twist<double> trans_opt_pivot(Qd::fromRotationMatrix(R_opt_pivot), origin_P0);
    for (int idx = 0; idx < pivot_idx; ++idx) {
      Twist<double> trans_idx(Qd::fromRotationMatrix(Rs[idx]), Ps[idx]);
      Twist<double> trans_opt_idx = trans_opt_pivot * trans_pivot.inverse() * trans_idx;
      Ps[idx] = trans_opt_idx.pos;
      Rs[idx] = trans_opt_idx.rot.normalized().toRotationMatrix();
    }
  }
  for (int i = 0, opt_i = pivot_idx; i < O + 1; ++i, ++opt_i) {
    Rs[opt_i] = rot_diff * qpose(i);
    Ps[opt_i] = rot_diff * V3(para_pose[i][0] - para_pose[0][0], para_pose[i][1] - para_pose[0][1], para_pose[i][2] - para_pose[0][2]) + origin_P0;
    Vs[opt_i] = rot_diff * V3(para_speed_bias[i][0], para_speed_bias[i][1], para_speed_bias[i][2]);
    Bas[opt_i] = V3(para_speed_bias[i][3], para_speed_bias[i][4], para_speed_bias[i][5]);
    Bgs[opt_i] = V3(para_speed_bias[i][6], para_speed_bias[i][7], para_speed_bias[i][8]);
  }
  transform_lb.pos = Vec3<float>((float)para_ex_pose[0], (float)para_ex_pose[1], (float)para_ex_pose[2]);
  transform_lb.rot = Quat<float>((float)para_ex_pose[6], (float)para_ex_pose[3], (float)para_ex_pose[4], (float)para_ex_pose[5]);
}

void Estimator::SolveOptimization() {
  turn_off = true;
  Problem problem;
  BuildLocalMap();
  double t0 = now_s();
  std::vector<double *> para_ids;
  for (int i = 0; i < O + 1; ++i) {
    problem.AddParameterBlock(para_pose[i], 7, true);
    problem.AddParameterBlock(para_speed_bias[i], 9, false);
  }
  problem.AddParameterBlock(para_ex_pose, 7, true);
  if (extrinsic_stage == 0 || cfg.opt_extrinsic == false) problem.SetParameterBlockConstant(para_ex_pose);
  VectorToDouble();
  std::vector<int> res_ids_marg, res_ids_pim, res_ids_proj;
  int res_id_marg = -1;
  if (cfg.marginalization_factor && last_marginalization_info) {
    auto f = std::make_shared<MarginalizationFactor>(last_marginalization_info);
    res_id_marg = problem.AddResidualBlock(f, nullptr, last_marginalization_parameter_blocks);
    res_ids_marg.push_back(res_id_marg);
  }
  const int pivot_idx = W - O;
  if (cfg.imu_factor) {
    for (int i = 0; i < O; ++i) {
      int j = i + 1;
      int opt_j = pivot_idx + i + 1;
      if (pre_integrations[opt_j]->sum_dt_ > 10.0) continue;
      auto f = std::make_shared<ImuFactor>(pre_integrations[opt_j]);
      res_ids_pim.push_back(problem.AddResidualBlock(
          f, nullptr, {para_pose[i], para_speed_bias[i], para_pose[j], para_speed_bias[j]}));
    }
  }
  if (cfg.point_distance_factor) {
    for (int i = 0; i < O + 1; ++i) {
      int opt_i = pivot_idx + i;
      const std::vector<PointPlaneFeature> &features = feature_frames[opt_i];
      if (i == 0) continue;
      for (const PointPlaneFeature &fj : features) {
        auto f = std::make_shared<PivotPointPlaneFactor>(fj.point, fj.coeffs);
        res_ids_proj.push_back(problem.AddResidualBlock(f, &loss, {para_pose[0], para_pose[i], para_ex_pose}));
      }
    }
  }
  if (cfg.prior_factor) {
    Twist<double> tt = transform_lb.cast<double>();
    problem.AddResidualBlock(std::make_shared<PriorFactor>(tt.pos, tt.rot), nullptr, {para_ex_pose});
  }
  // residuals before optimisation + gates (:1924-1985)
  cost_pim = cost_ppp = cost_marg = 0.0;
  if (cfg.imu_factor) {
    cost_pim = problem.EvaluateCost(&res_ids_pim);
    turn_off = cost_pim > 1e3;
  }
  if (cfg.point_distance_factor) cost_ppp = problem.EvaluateCost(&res_ids_proj);
  if (cfg.marginalization_factor && last_marginalization_info) cost_marg = problem.EvaluateCost(&res_ids_marg);
  {
    double ratio = cost_marg / (cost_ppp + cost_pim);
    if (!convergence_flag && !turn_off && ratio <= 2 && ratio != 0) convergence_flag = true;
    if (!convergence_flag) {
      problem.SetParameterBlockConstant(para_ex_pose);
      if (last_marginalization_info) last_marginalization_info.reset();
      if (res_id_marg >= 0) { problem.RemoveResidualBlock(res_id_marg); res_ids_marg.clear(); }
    }
  }
  Solve(cfg.solver, &problem, &summary);
  t_solve = now_s() - t0;
  DoubleToVector();
  double t1 = now_s();
  if (cfg.marginalization_factor && !turn_off) {  // :2040-2275
    auto marginalization_info = std::make_shared<MarginalizationInfo>();
    VectorToDouble();
    if (last_marginalization_info) {
      std::vector<int> drop_set;
      for (int i = 0; i < (int)last_marginalization_parameter_blocks.size(); i++)
        if (last_marginalization_parameter_blocks[i] == para_pose[0] || last_marginalization_parameter_blocks[i] == para_speed_bias[0])
          drop_set.push_back(i);
      auto mf = std::make_shared<MarginalizationFactor>(last_marginalization_info);
      marginalization_info->AddResidualBlockInfo(std::make_shared<ResidualBlockInfo>(mf, nullptr, last_marginalization_parameter_blocks, drop_set));
    }
    if (cfg.imu_factor) {
      if (pre_integrations[pivot_idx + 1]->sum_dt_ < 10.0) {
        auto imu_factor = std::make_shared<ImuFactor>(pre_integrations[pivot_idx + 1]);
        marginalization_info->AddResidualBlockInfo(std::make_shared<ResidualBlockInfo>(
            imu_factor, nullptr,
            std::vector<double *>{para_pose[0], para_speed_bias[0], para_pose[1], para_speed_bias[1]},
            std::vector<int>{0, 1}));
      }
    }
    if (cfg.point_distance_factor) {
      for (int i = 1; i < O + 1; ++i) {
        int opt_i = pivot_idx + i;
        for (const PointPlaneFeature &fj : feature_frames[opt_i]) {
          auto f = std::make_shared<PivotPointPlaneFactor>(fj.point, fj.coeffs);
          marginalization_info->AddResidualBlockInfo(std::make_shared<ResidualBlockInfo>(
              f, &loss, std::vector<double *>{para_pose[0], para_pose[i], para_ex_pose}, std::vector<int>{0}));
        }
      }
    }
    marginalization_info->PreMarginalize();
    marginalization_info->Marginalize();
    std::map<long, double *> addr_shift;
    for (int i = 1; i < O + 1; ++i) {
      addr_shift[reinterpret_cast<long>(para_pose[i])] = para_pose[i - 1];
      addr_shift[reinterpret_cast<long>(para_speed_bias[i])] = para_speed_bias[i - 1];
    }
    addr_shift[reinterpret_cast<long>(para_ex_pose)] = para_ex_pose;
    std::vector<double *> parameter_blocks = marginalization_info->GetParameterBlocks(addr_shift);
    last_marginalization_info = marginalization_info;
    last_marginalization_parameter_blocks = parameter_blocks;
  }
  t_marg = now_s() - t1;
}

void Estimator::SlideWindow() {
  if (init_local_map) {
    const int pivot_idx = W - O;
    Twist<double> tlb = transform_lb.cast<double>();
    Twist<double> transform_pivot = lidar_pose(Ps[pivot_idx], Rs[pivot_idx], tlb);
    int i = pivot_idx + 1;
    Twist<double> transform_li = lidar_pose(Ps[i], Rs[i], tlb);
    Twist<float> tip = (transform_li.inverse() * transform_pivot).cast<float>();
    Cloud transformed;
    TransformCloudAffine(surf_stack[pivot_idx], tip.linear(), tip.pos, transformed);
    Cloud filtered;
    for (size_t k = (size_t)size_surf_stack[0]; k < transformed.size(); ++k) filtered.push_back(transformed[k]);
    filtered.insert(filtered.end(), surf_stack[i].begin(), surf_stack[i].end());
    surf_stack[i] = filtered;
  }
  push_shift(Ps, Ps[W]); push_shift(Vs, Vs[W]); push_shift(Rs, Rs[W]); push_shift(Bas, Bas[W]); push_shift(Bgs, Bgs[W]);
}

}  // namespace orc
