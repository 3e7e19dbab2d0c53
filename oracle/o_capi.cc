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
