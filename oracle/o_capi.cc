// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
// ctypes-facing C API of the CPU restatement (loaded by tests/ and bench.py's cpu_baseline leg).
#include "o_api.h"
#include <cstring>

using namespace orc;

extern "C" {

// ---------------- stage A ----------------
void *orc_a_create(float lower_bound, float upper_bound, int num_rings, double scan_period) {
  StageA *a = new StageA();
  a->cfg.lower_bound = lower_bound;
  a->cfg.upper_bound = upper_bound;
  a->cfg.num_rings = num_rings;
  a->cfg.scan_period = scan_period;
  return a;
}
void orc_a_destroy(void *h) { delete (StageA *)h; }
void orc_a_run(void *h, const float *xyzi, int n) {
  StageA *a = (StageA *)h;
  a->PointToRing((const PointXYZI *)xyzi, (size_t)n);
  a->ExtractFeaturePoints();
}
void orc_a_run_ring(void *h, const float *xyzi, const unsigned short *rings, int n) {  // PointXYZIR input
  StageA *a = (StageA *)h;
  a->PointToRingWithRingField((const PointXYZI *)xyzi, rings, (size_t)n);
  a->ExtractFeaturePoints();
}
// which: 0 laser_scans (ring-ordered, intensity=ring+rel_time) 1 cloud_in_rings 2 sharp 3 less_sharp 4 flat 5 less_flat
static const Cloud *a_cloud(StageA *a, int which, Cloud &tmp) {
  switch (which) {
    case 0: tmp.clear(); for (auto &c : a->laser_scans) tmp.insert(tmp.end(), c.begin(), c.end()); return &tmp;
    case 1: return &a->cloud_in_rings;
    case 2: return &a->corner_sharp;
    case 3: return &a->corner_less_sharp;
    case 4: return &a->surf_flat;
    case 5: return &a->surf_less_flat;
  }
  return nullptr;
}
int orc_a_cloud_size(void *h, int which) { Cloud t; const Cloud *c = a_cloud((StageA *)h, which, t); return c ? (int)c->size() : -1; }
void orc_a_cloud_copy(void *h, int which, float *out) {
  Cloud t; const Cloud *c = a_cloud((StageA *)h, which, t);
  if (c && !c->empty()) std::memcpy(out, c->data(), c->size() * sizeof(PointXYZI));
}
// which: 0 idx_sharp 1 idx_less_sharp 2 idx_flat 3 less_flat_prevoxel 4 orig_index (ring-ordered)
static const std::vector<int> *a_idx(StageA *a, int which, std::vector<int> &tmp) {
  switch (which) {
    case 0: return &a->idx_sharp;
    case 1: return &a->idx_less_sharp;
    case 2: return &a->idx_flat;
    case 3: return &a->less_flat_prevoxel;
    case 4: tmp.clear(); for (auto &c : a->orig_index) tmp.insert(tmp.end(), c.begin(), c.end()); return &tmp;
  }
  return nullptr;
}
int orc_a_idx_size(void *h, int which) { std::vector<int> t; auto *v = a_idx((StageA *)h, which, t); return v ? (int)v->size() : -1; }
void orc_a_idx_copy(void *h, int which, int *out) {
  std::vector<int> t; auto *v = a_idx((StageA *)h, which, t);
  if (v && !v->empty()) std::memcpy(out, v->data(), v->size() * sizeof(int));
}
void orc_a_scan_ranges(void *h, int *out /* 2*R */) {
  StageA *a = (StageA *)h;
  for (size_t i = 0; i < a->scan_ranges.size(); ++i) { out[2 * i] = (int)a->scan_ranges[i].first; out[2 * i + 1] = (int)a->scan_ranges[i].second; }
}
void orc_a_mask_labels(void *h, unsigned char *mask, signed char *labels) {
  StageA *a = (StageA *)h;
  if (!a->final_mask.empty()) std::memcpy(mask, a->final_mask.data(), a->final_mask.size());
  if (!a->label_all.empty()) std::memcpy(labels, a->label_all.data(), a->label_all.size());
}
float orc_a_start_ori(void *h) { return ((StageA *)h)->start_ori; }

// ---------------- math KATs (test/test_point_processor/test_point_processor.cc:55-63) ----------------
double orc_normalize_rad(double rad) {  // math_utils.h:43-50
  rad = fmod(rad + M_PI, 2 * M_PI);
  if (rad < 0) rad += 2 * M_PI;
  return rad - M_PI;
}
double orc_normalize_deg(double deg) {  // math_utils.h:57-64
  deg = fmod(deg + 180.0, 360.0);
  if (deg < 0) deg += 360.0;
  return deg - 180.0;
}

// ---------------- cloud utilities ----------------
int orc_voxel_grid(const float *in, int n, float leaf, float *out) {
  Cloud ci((const PointXYZI *)in, (const PointXYZI *)in + n), co;
  VoxelGridFilter(ci, leaf, co);
  if (!co.empty()) std::memcpy(out, co.data(), co.size() * sizeof(PointXYZI));
  return (int)co.size();
}
void orc_transform_cloud(const float *in, int n, const float *R9, const float *t3, float *out) {
  Cloud ci((const PointXYZI *)in, (const PointXYZI *)in + n), co;
  Mat3<float> R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = R9[i * 3 + j];
  TransformCloudAffine(ci, R, Vec3<float>(t3[0], t3[1], t3[2]), co);
  if (!co.empty()) std::memcpy(out, co.data(), co.size() * sizeof(PointXYZI));
}
void orc_knn(const float *map, int K, const float *queries, int M, int k, int *idx_out, float *d2_out) {
  Cloud cm((const PointXYZI *)map, (const PointXYZI *)map + K);
  KdTree kd; kd.Build(cm);
  for (int i = 0; i < M; ++i) kd.Knn(((const PointXYZI *)queries)[i], k, idx_out + (size_t)i * k, d2_out + (size_t)i * k);
}

static Transform make_tf(const float *t7) {  // (qx,qy,qz,qw, px,py,pz)
  return Transform(Quat<float>(t7[3], t7[0], t7[1], t7[2]), Vec3<float>(t7[4], t7[5], t7[6]));
}
static void copy_feats(const std::vector<PointPlaneFeature> &f, float *pts4, float *coef4, int *src) {
  for (size_t i = 0; i < f.size(); ++i) {
    pts4[4 * i + 0] = (float)f[i].point[0]; pts4[4 * i + 1] = (float)f[i].point[1]; pts4[4 * i + 2] = (float)f[i].point[2];
    pts4[4 * i + 3] = (float)f[i].score;
    for (int k = 0; k < 4; ++k) coef4[4 * i + k] = (float)f[i].coeffs[k];
    if (src) src[i] = f[i].src_index;
  }
}
// Estimator::CalculateFeatures on explicit arrays; returns the feature count (outputs sized M).
int orc_calculate_features(const float *map, int K, const float *surf, int M, const float *tf7, float min_match_sq_dis,
                           float min_plane_dis, float *pts4, float *coef4, int *src) {
  Cloud cm((const PointXYZI *)map, (const PointXYZI *)map + K), cs((const PointXYZI *)surf, (const PointXYZI *)surf + M);
  KdTree kd; kd.Build(cm);
  StageBConfig cfg; cfg.min_match_sq_dis = min_match_sq_dis; cfg.min_plane_dis = min_plane_dis;
  std::vector<PointPlaneFeature> feats;
  CalculateFeatures(kd, cm, cs, make_tf(tf7), cfg, feats);
  copy_feats(feats, pts4, coef4, src);
  return (int)feats.size();
}
// Point-to-line branch (Estimator.cc:1101-1227 / PointMapping.cc:381-512); outputs sized 2*M.
int orc_calculate_line_features(const float *map, int K, const float *corner, int M, const float *tf7, float min_match_sq_dis,
                                float *pts4, float *coef4, int *src) {
  Cloud cm((const PointXYZI *)map, (const PointXYZI *)map + K), cs((const PointXYZI *)corner, (const PointXYZI *)corner + M);
  KdTree kd; kd.Build(cm);
  StageBConfig cfg; cfg.min_match_sq_dis = min_match_sq_dis;
  std::vector<PointPlaneFeature> feats;
  CalculateLineFeatures(kd, cm, cs, make_tf(tf7), cfg, feats);
  copy_feats(feats, pts4, coef4, src);
  return (int)feats.size();
}
// PointMapping::OptimizeTransformTobeMapped on explicit arrays; tf7 in/out; feature outputs (last pass) sized Mc + Ms.
int orc_scan_to_map(const float *cmap, int Kc, const float *smap, int Ks, const float *corner, int Mc, const float *surf, int Ms,
                    float *tf7, float min_match_sq_dis, float min_plane_dis, int max_iter, double delta_r_abort, double delta_t_abort,
                    float *pts4, float *coef4, int *src, int *iters, int variant) {
  Cloud cm((const PointXYZI *)cmap, (const PointXYZI *)cmap + Kc), sm((const PointXYZI *)smap, (const PointXYZI *)smap + Ks),
      cs((const PointXYZI *)corner, (const PointXYZI *)corner + Mc), ss((const PointXYZI *)surf, (const PointXYZI *)surf + Ms);
  StageBConfig cfg; cfg.min_match_sq_dis = min_match_sq_dis; cfg.min_plane_dis = min_plane_dis; cfg.num_max_iterations = max_iter;
  cfg.delta_r_abort = delta_r_abort; cfg.delta_t_abort = delta_t_abort;
  Transform t = make_tf(tf7);
  std::vector<PointPlaneFeature> feats;
  OptimizeTransformTobeMapped(cm, sm, cs, ss, t, cfg, iters, &feats, variant);
  tf7[0] = t.rot.x; tf7[1] = t.rot.y; tf7[2] = t.rot.z; tf7[3] = t.rot.w; tf7[4] = t.pos.x; tf7[5] = t.pos.y; tf7[6] = t.pos.z;
  copy_feats(feats, pts4, coef4, src);
  return (int)feats.size();
}
// Estimator::CalculateLaserOdom; tf7 is in/out; outputs sized M*(keep_features? max_iter : 1).
int orc_laser_odom(const float *map, int K, const float *surf, int M, float *tf7, float min_match_sq_dis, float min_plane_dis,
                   int keep_features, int max_iter, float *pts4, float *coef4, int *src, int *iters) {
  Cloud cm((const PointXYZI *)map, (const PointXYZI *)map + K), cs((const PointXYZI *)surf, (const PointXYZI *)surf + M);
  KdTree kd; kd.Build(cm);
  StageBConfig cfg; cfg.min_match_sq_dis = min_match_sq_dis; cfg.min_plane_dis = min_plane_dis;
  cfg.keep_features = keep_features; cfg.num_max_iterations = max_iter;
  std::vector<PointPlaneFeature> feats;
  Transform t = make_tf(tf7);
  CalculateLaserOdom(kd, cm, cs, t, cfg, feats, iters);
  tf7[0] = t.rot.x; tf7[1] = t.rot.y; tf7[2] = t.rot.z; tf7[3] = t.rot.w; tf7[4] = t.pos.x; tf7[5] = t.pos.y; tf7[6] = t.pos.z;
  copy_feats(feats, pts4, coef4, src);
  return (int)feats.size();
}

}  // extern "C"

// =================================================================================================
// fp64 factors / solver / estimator
#include "o_estimator.h"

extern "C" {

// PivotPointPlaneFactor::Evaluate (src/factor/PivotPointPlaneFactor.cc:43-137); J* are 1x7 row-major
void orc_ppp_evaluate(const double *point3, const double *coeff4, const double *pose_pivot, const double *pose_i,
                      const double *pose_ex, double *residual, double *J0, double *J1, double *J2) {
  PivotPointPlaneFactor f(point3, coeff4);
  const double *params[3] = {pose_pivot, pose_i, pose_ex};
  double *jac[3] = {J0, J1, J2};
  f.Evaluate(params, residual, (J0 || J1 || J2) ? jac : nullptr);
}

void orc_prior_evaluate(const double *pos3, const double *quat_xyzw, const double *pose, double *res6, double *J6x7) {
  PriorFactor f(V3(pos3[0], pos3[1], pos3[2]), Qd(quat_xyzw[3], quat_xyzw[0], quat_xyzw[1], quat_xyzw[2]));
  const double *params[1] = {pose};
  double *jac[1] = {J6x7};
  f.Evaluate(params, res6, J6x7 ? jac : nullptr);
}

void orc_pose_plus(const double *x7, const double *delta6, double *out7) { PosePlus(x7, delta6, out7); }

// IntegrationBase (include/imu_processor/IntegrationBase.h)
void *orc_pim_create(const double *acc0, const double *gyr0, const double *ba, const double *bg, const double *cfg5) {
  IntegrationBaseConfig c;
  c.acc_n = cfg5[0]; c.gyr_n = cfg5[1]; c.acc_w = cfg5[2]; c.gyr_w = cfg5[3]; c.g_norm = cfg5[4];
  return new std::shared_ptr<IntegrationBase>(new IntegrationBase(V3(acc0[0], acc0[1], acc0[2]), V3(gyr0[0], gyr0[1], gyr0[2]),
                                                                  V3(ba[0], ba[1], ba[2]), V3(bg[0], bg[1], bg[2]), c));
}
void orc_pim_destroy(void *h) { delete (std::shared_ptr<IntegrationBase> *)h; }
void orc_pim_push_back(void *h, double dt, const double *acc, const double *gyr) {
  (*(std::shared_ptr<IntegrationBase> *)h)->push_back(dt, V3(acc[0], acc[1], acc[2]), V3(gyr[0], gyr[1], gyr[2]));
}
// out: delta_p(3) delta_q(xyzw 4) delta_v(3) sum_dt(1) | jacobian 225 | covariance 225
void orc_pim_get(void *h, double *state11, double *jac225, double *cov225) {
  IntegrationBase &p = **(std::shared_ptr<IntegrationBase> *)h;
  state11[0] = p.delta_p_.x; state11[1] = p.delta_p_.y; state11[2] = p.delta_p_.z;
  state11[3] = p.delta_q_.x; state11[4] = p.delta_q_.y; state11[5] = p.delta_q_.z; state11[6] = p.delta_q_.w;
  state11[7] = p.delta_v_.x; state11[8] = p.delta_v_.y; state11[9] = p.delta_v_.z; state11[10] = p.sum_dt_;
  if (jac225) std::memcpy(jac225, p.jacobian_.d.data(), 225 * sizeof(double));
  if (cov225) std::memcpy(cov225, p.covariance_.d.data(), 225 * sizeof(double));
}
// ImuFactor::Evaluate (include/factor/ImuFactor.h:53-167): params = pose_i(7) sb_i(9) pose_j(7) sb_j(9)
void orc_imu_factor_evaluate(void *h, const double *pose_i, const double *sb_i, const double *pose_j, const double *sb_j,
                             double *res15, double *J0, double *J1, double *J2, double *J3) {
  ImuFactor f(*(std::shared_ptr<IntegrationBase> *)h);
  const double *params[4] = {pose_i, sb_i, pose_j, sb_j};
  double *jac[4] = {J0, J1, J2, J3};
  f.Evaluate(params, res15, (J0 || J1 || J2 || J3) ? jac : nullptr);
}

void orc_sym_eigen(const double *A, int n, double *evals, double *evecs) {
  MatX a(n, n);
  std::memcpy(a.d.data(), A, sizeof(double) * n * n);
  VecX ev; MatX V;
  SymEigen(a, ev, V);
  std::memcpy(evals, ev.data(), sizeof(double) * n);
  std::memcpy(evecs, V.d.data(), sizeof(double) * n * n);
}

// ---------------- estimator ----------------
// cfg (doubles): [0] window_size [1] opt_window_size [2] min_match_sq_dis [3] min_plane_dis [4] surf_filter_size
// [5] keep_features [6] estimate_extrinsic [7] opt_extrinsic [8] imu_factor [9] point_distance_factor [10] prior_factor
// [11] marginalization_factor [12] enable_deskew [13] cutoff_deskew [14..18] acc_n gyr_n acc_w gyr_w g_norm
// [19] max_num_iterations [20] laser-odom max iterations
void *orc_est_create(const double *c) {
  EstimatorConfig cfg;
  cfg.window_size = (int)c[0]; cfg.opt_window_size = (int)c[1];
  cfg.b.min_match_sq_dis = (float)c[2]; cfg.b.min_plane_dis = (float)c[3]; cfg.b.surf_filter_size = (float)c[4];
  cfg.b.keep_features = (int)c[5]; cfg.estimate_extrinsic = (int)c[6]; cfg.opt_extrinsic = c[7] != 0;
  cfg.imu_factor = c[8] != 0; cfg.point_distance_factor = c[9] != 0; cfg.prior_factor = c[10] != 0;
  cfg.marginalization_factor = c[11] != 0; cfg.enable_deskew = c[12] != 0; cfg.cutoff_deskew = c[13] != 0;
  cfg.pim.acc_n = c[14]; cfg.pim.gyr_n = c[15]; cfg.pim.acc_w = c[16]; cfg.pim.gyr_w = c[17]; cfg.pim.g_norm = c[18];
  cfg.solver.max_num_iterations = (int)c[19];
  cfg.b.num_max_iterations = (int)c[20];
  return new Estimator(cfg);
}
void orc_est_destroy(void *h) { delete (Estimator *)h; }
void orc_est_set_extrinsic(void *h, const float *tf7 /* qx qy qz qw px py pz */) {
  ((Estimator *)h)->transform_lb = Transform(Quat<float>(tf7[3], tf7[0], tf7[1], tf7[2]), Vec3<float>(tf7[4], tf7[5], tf7[6]));
}
// state16 = P(3) Q(xyzw 4) V(3) Ba(3) Bg(3); pim may be NULL for frame 0
void orc_est_init_frame(void *h, int k, const double *s, const float *surf_ds, int n, void *pim) {
  Estimator *e = (Estimator *)h;
  Cloud c((const PointXYZI *)surf_ds, (const PointXYZI *)surf_ds + n);
  std::shared_ptr<IntegrationBase> p = pim ? *(std::shared_ptr<IntegrationBase> *)pim : nullptr;
  e->InitFrame(k, V3(s[0], s[1], s[2]), Qd(s[6], s[3], s[4], s[5]), V3(s[7], s[8], s[9]), V3(s[10], s[11], s[12]), V3(s[13], s[14], s[15]), c, p);
}
void orc_est_finish_init(void *h, const double *acc_last, const double *gyr_last) {
  ((Estimator *)h)->FinishInit(V3(acc_last[0], acc_last[1], acc_last[2]), V3(gyr_last[0], gyr_last[1], gyr_last[2]));
}
void orc_est_process_imu(void *h, double dt, const double *acc, const double *gyr, double stamp) {
  ((Estimator *)h)->ProcessImu(dt, V3(acc[0], acc[1], acc[2]), V3(gyr[0], gyr[1], gyr[2]), stamp);
}
void orc_est_process_scan(void *h, const float *surf_last, int n) {
  Cloud c((const PointXYZI *)surf_last, (const PointXYZI *)surf_last + n);
  ((Estimator *)h)->ProcessScan(c);
}
// window states: (W+1) x 16 doubles, same layout as init_frame
void orc_est_get_states(void *h, double *out) {
  Estimator *e = (Estimator *)h;
  for (int k = 0; k <= e->W; ++k) {
    double *s = out + 16 * k;
    Qd q = Qd::fromRotationMatrix(e->Rs[k]);
    s[0] = e->Ps[k].x; s[1] = e->Ps[k].y; s[2] = e->Ps[k].z; s[3] = q.x; s[4] = q.y; s[5] = q.z; s[6] = q.w;
    for (int a = 0; a < 3; ++a) { s[7 + a] = e->Vs[k][a]; s[10 + a] = e->Bas[k][a]; s[13 + a] = e->Bgs[k][a]; }
  }
}
void orc_est_get_extrinsic(void *h, float *tf7) {
  Estimator *e = (Estimator *)h;
  tf7[0] = e->transform_lb.rot.x; tf7[1] = e->transform_lb.rot.y; tf7[2] = e->transform_lb.rot.z; tf7[3] = e->transform_lb.rot.w;
  tf7[4] = e->transform_lb.pos.x; tf7[5] = e->transform_lb.pos.y; tf7[6] = e->transform_lb.pos.z;
}
// summary: [0] iterations [1] successful [2] termination [3] initial_cost [4] final_cost [5] cost_pim [6] cost_ppp
// [7] cost_marg [8] turn_off [9] convergence_flag [10] map size [11] total features [12] laser odom iters
// [13..17] t_build_map t_features t_solve t_marg t_total [18] has prior [19] linearizations [20] cost evals
void orc_est_summary(void *h, double *out) {
  Estimator *e = (Estimator *)h;
  out[0] = e->summary.num_iterations; out[1] = e->summary.num_successful_steps; out[2] = e->summary.termination;
  out[3] = e->summary.initial_cost; out[4] = e->summary.final_cost; out[5] = e->cost_pim; out[6] = e->cost_ppp; out[7] = e->cost_marg;
  out[8] = e->turn_off; out[9] = e->convergence_flag; out[10] = (double)e->local_surf_points_filtered.size();
  size_t nf = 0;
  for (auto &f : e->feature_frames) nf += f.size();
  out[11] = (double)nf; out[12] = e->laser_odom_iters;
  out[13] = e->t_build_map; out[14] = e->t_features; out[15] = e->t_solve; out[16] = e->t_marg; out[17] = e->t_total;
  out[18] = e->last_marginalization_info ? 1 : 0; out[19] = e->summary.num_linearizations; out[20] = e->summary.num_cost_evaluations;
}
int orc_est_feature_count(void *h, int frame) { return (int)((Estimator *)h)->feature_frames[frame].size(); }
void orc_est_get_features(void *h, int frame, float *pts4, float *coef4, int *src) {
  copy_feats(((Estimator *)h)->feature_frames[frame], pts4, coef4, src);
}
int orc_est_map_size(void *h) { return (int)((Estimator *)h)->local_surf_points_filtered.size(); }
void orc_est_get_map(void *h, float *out) {
  Estimator *e = (Estimator *)h;
  if (!e->local_surf_points_filtered.empty())
    std::memcpy(out, e->local_surf_points_filtered.data(), e->local_surf_points_filtered.size() * sizeof(PointXYZI));
}
int orc_est_frame_size(void *h, int frame) { return (int)((Estimator *)h)->surf_stack[frame].size(); }
void orc_est_get_frame(void *h, int frame, float *out) {
  Estimator *e = (Estimator *)h;
  if (!e->surf_stack[frame].empty()) std::memcpy(out, e->surf_stack[frame].data(), e->surf_stack[frame].size() * sizeof(PointXYZI));
}
void orc_est_get_local_transform(void *h, int frame, float *tf7) {
  const Transform &t = ((Estimator *)h)->local_transforms[frame];
  tf7[0] = t.rot.x; tf7[1] = t.rot.y; tf7[2] = t.rot.z; tf7[3] = t.rot.w; tf7[4] = t.pos.x; tf7[5] = t.pos.y; tf7[6] = t.pos.z;
}
// marginalisation prior of the last solve: n, then linearized_jacobians (n x n row-major), residuals (n)
// kept blocks of the prior, in its own order: kind (0 pose, 1 speed-bias, 2 extrinsic), window index (already
// shifted to the NEXT window's numbering), tangent offset in the prior
int orc_est_prior_blocks(void *h, int *kind, int *index, int *offset) {
  Estimator *e = (Estimator *)h;
  if (!e->last_marginalization_info) return 0;
  auto &mi = *e->last_marginalization_info;
  int nb = (int)e->last_marginalization_parameter_blocks.size();
  for (int b = 0; b < nb; ++b) {
    double *p = e->last_marginalization_parameter_blocks[b];
    kind[b] = -1; index[b] = -1;
    for (int i = 0; i <= e->O; ++i) {
      if (p == e->para_pose[i]) { kind[b] = 0; index[b] = i; }
      if (p == e->para_speed_bias[i]) { kind[b] = 1; index[b] = i; }
    }
    if (p == e->para_ex_pose) { kind[b] = 2; index[b] = 0; }
    offset[b] = mi.keep_block_idx[b] - mi.m;
  }
  return nb;
}
int orc_est_normal_dim(void *h) { return ((Estimator *)h)->summary.H_initial.r; }
void orc_est_get_normal(void *h, double *H, double *g) {
  Estimator *e = (Estimator *)h;
  int n = e->summary.H_initial.r;
  std::memcpy(H, e->summary.H_initial.d.data(), sizeof(double) * n * n);
  std::memcpy(g, e->summary.g_initial.data(), sizeof(double) * n);
}
int orc_est_prior_dim(void *h) { Estimator *e = (Estimator *)h; return e->last_marginalization_info ? e->last_marginalization_info->n : 0; }
void orc_est_get_prior(void *h, double *J, double *r) {
  Estimator *e = (Estimator *)h;
  if (!e->last_marginalization_info) return;
  auto &mi = *e->last_marginalization_info;
  std::memcpy(J, mi.linearized_jacobians.d.data(), sizeof(double) * mi.n * mi.n);
  std::memcpy(r, mi.linearized_residuals.data(), sizeof(double) * mi.n);
}

}  // extern "C"

extern "C" void orc_transform_to_end(float *cloud, int n, const float *tf7, float time_factor) {
  Cloud c((const PointXYZI *)cloud, (const PointXYZI *)cloud + n);
  Transform t = make_tf(tf7);
  TransformToEnd(c, t, time_factor);
  std::memcpy(cloud, c.data(), sizeof(PointXYZI) * n);
}

// ---- toy nonlinear least squares through the oracle's Ceres-style Problem / Solve (controller parity on the CPU) --------
// residual k:  r_k = a_k . x + amp * sin(b_k . x) - y_k   (1-dim blocks over ONE n-dim Euclidean parameter block),
// optional CauchyLoss(1.0).  summary = {iterations, successful steps, termination, initial cost, final cost, linearizations}.
namespace {
struct ToyCost : orc::CostFunction {
  const double *a, *b;
  double y, amp;
  int n;
  ToyCost(const double *a_, const double *b_, double y_, double amp_, int n_) : a(a_), b(b_), y(y_), amp(amp_), n(n_) {
    num_residuals = 1;
    block_sizes = {n_};
  }
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override {
    const double *x = parameters[0];
    double s = 0, t = 0;
    for (int j = 0; j < n; ++j) { s += a[j] * x[j]; t += b[j] * x[j]; }
    residuals[0] = s + amp * std::sin(t) - y;
    if (jacobians && jacobians[0]) {
      const double cb = amp * std::cos(t);
      for (int j = 0; j < n; ++j) jacobians[0][j] = a[j] + cb * b[j];
    }
    return true;
  }
};
}  // namespace

extern "C" int orc_toy_solve(int n, int m, const double *A, const double *B, const double *y, double amp, int use_cauchy, double *x,
                             int max_iter, double *summary) {
  orc::Problem P;
  P.AddParameterBlock(x, n, false);
  orc::CauchyLoss loss(1.0);
  for (int k = 0; k < m; ++k)
    P.AddResidualBlock(std::make_shared<ToyCost>(A + (size_t)k * n, B + (size_t)k * n, y[k], amp, n), use_cauchy ? &loss : nullptr, {x});
  orc::SolverOptions opt;
  opt.max_num_iterations = max_iter;
  orc::SolverSummary sum;
  orc::Solve(opt, &P, &sum);
  summary[0] = sum.num_iterations; summary[1] = sum.num_successful_steps; summary[2] = sum.termination;
  summary[3] = sum.initial_cost; summary[4] = sum.final_cost; summary[5] = sum.num_linearizations;
  return 0;
}

// ---- toy marginalisation through the oracle's MarginalizationInfo (pins the Schur / eigen square-root algebra against numpy) ----
// nb Euclidean parameter blocks of size bs in one contiguous array x; factor k: r = Wi x_i + Wj x_j - y (nr rows) between blocks
// (bi[k], bj[k]); every factor that touches block 0 is added with block 0 in its drop set (all others are left out, like the
// reference's marginalisation step which only takes the factors of the dropped states).  Outputs: m, n, the order of the kept
// blocks, linearized_jacobians (n x n), linearized_residuals (n), and the assembled A (pos x pos), b (pos).
namespace {
struct ToyLinear2 : orc::CostFunction {
  const double *wi, *wj, *y;
  int nr, bs;
  ToyLinear2(const double *wi_, const double *wj_, const double *y_, int nr_, int bs_) : wi(wi_), wj(wj_), y(y_), nr(nr_), bs(bs_) {
    num_residuals = nr_;
    block_sizes = {bs_, bs_};
  }
  bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const override {
    for (int r = 0; r < nr; ++r) {
      double s = -y[r];
      for (int c = 0; c < bs; ++c) s += wi[r * bs + c] * parameters[0][c] + wj[r * bs + c] * parameters[1][c];
      residuals[r] = s;
    }
    if (jacobians) {
      if (jacobians[0]) std::memcpy(jacobians[0], wi, sizeof(double) * nr * bs);
      if (jacobians[1]) std::memcpy(jacobians[1], wj, sizeof(double) * nr * bs);
    }
    return true;
  }
};
}  // namespace

extern "C" int orc_toy_marginalize(int nb, int bs, int nf, int nr, const int *bi, const int *bj, const double *W /*nf x 2 x nr x bs*/,
                                   const double *y /*nf x nr*/, double *x /*nb x bs*/, int use_cauchy, int *mn /*2*/, int *kept_blocks /*nb*/,
                                   double *lin_J, double *lin_r, double *A_out, double *b_out) {
  orc::MarginalizationInfo mi;
  orc::CauchyLoss loss(1.0);
  for (int k = 0; k < nf; ++k) {
    if (bi[k] != 0 && bj[k] != 0) continue;
    std::vector<int> drop;
    if (bi[k] == 0) drop.push_back(0);
    if (bj[k] == 0) drop.push_back(1);
    auto cost = std::make_shared<ToyLinear2>(W + (size_t)k * 2 * nr * bs, W + (size_t)k * 2 * nr * bs + nr * bs, y + (size_t)k * nr, nr, bs);
    mi.AddResidualBlockInfo(std::make_shared<orc::ResidualBlockInfo>(cost, use_cauchy ? &loss : nullptr,
                                                                     std::vector<double *>{x + (size_t)bi[k] * bs, x + (size_t)bj[k] * bs}, drop));
  }
  mi.PreMarginalize();
  mi.Marginalize();
  mn[0] = mi.m; mn[1] = mi.n;
  int nk = 0;
  for (const auto &it : mi.parameter_block_idx)
    if (it.second >= mi.m) kept_blocks[(it.second - mi.m) / bs] = (int)((reinterpret_cast<double *>(it.first) - x) / bs), ++nk;
  std::memcpy(lin_J, mi.linearized_jacobians.d.data(), sizeof(double) * mi.n * mi.n);
  std::memcpy(lin_r, mi.linearized_residuals.data(), sizeof(double) * mi.n);
  const int pos = mi.m + mi.n;
  std::memcpy(A_out, mi.A_dbg.d.data(), sizeof(double) * pos * pos);
  std::memcpy(b_out, mi.b_dbg.data(), sizeof(double) * pos);
  return nk;
}
