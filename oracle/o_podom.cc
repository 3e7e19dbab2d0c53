// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
//
// CPU restatement of lio::PointOdometry, the scan-to-scan odometry that runs before IMU initialisation and that turns into a
// pass-through once the estimator switches it off (/enable_odom):
//   constructor defaults     src/point_processor/PointOdometry.cc:66-86   (time_factor = 1 / scan_period, abort 0.1 deg / 0.1 cm)
//   TransformToStart         :237-259
//   TransformToEnd           :261-292     (intensity <- int(intensity))
//   Process                  :294-708     first sweep :302-310; corner matching :338-441; surf matching :443-549;
//                                         6 x 6 float Gauss-Newton :551-664; transform_sum_ :667-669; de-skew + swap :673-690
//   PublishResults           :710-766     io_ratio gate :726, TransformToEnd(full_cloud_) :728-730, /compact_data :732-764
// ROS message synchronisation (HasNewData :227-235) and the topic handlers are the caller's business.
// k-NN: exact nearest neighbour, ties by index (o_cloud.cc KdTree).  AtA / AtB are accumulated sequentially in float (the
// reference's Eigen GEMM order is unspecified; the device kernel accumulates the float products in double).
// Device counterpart: lio_mapping_b200/csrc/podom.cu (lio_po_*), compared in tests/test_point_odometry_gpu.py.
#include "o_api.h"
#include <cmath>
#include <cstring>

namespace orc {

void CompactEncode(const Transform &transform_sum, const Cloud &last_corner_cloud, const Cloud &last_surf_cloud, const Cloud &full_cloud,
                   Cloud &compact_data);  // o_wire.cc

static inline float CalcSquaredDiff(const PointXYZI &a, const PointXYZI &b) {  // include/utils/math_utils.h:85-91
  float diff_x = a.x - b.x, diff_y = a.y - b.y, diff_z = a.z - b.z;
  return diff_x * diff_x + diff_y * diff_y + diff_z * diff_z;
}
static inline float CalcPointDistance(const PointXYZI &p) { return std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z); }  // :103-105
static inline void RotatePoint(const Quat<float> &q, PointXYZI &p) {  // include/utils/geometry_utils.h:289-298
  Vec3<float> v = q * Vec3<float>(p.x, p.y, p.z);
  p.x = v.x; p.y = v.y; p.z = v.z;
}

struct PointOdometry {
  float scan_period, time_factor;
  int io_ratio;
  size_t num_max_iterations;
  double delta_r_abort = 0.1, delta_t_abort = 0.1;
  bool system_inited = false, enable_odom = true, no_deskew = false;
  long frame_count = 0;
  Cloud corner_points_sharp, corner_points_less_sharp, surf_points_flat, surf_points_less_flat, full_cloud;
  Cloud last_corner_cloud, last_surf_cloud;
  Transform transform_es, transform_sum;
  KdTree kdtree_corner_last, kdtree_surf_last;
  std::vector<int> idx_corner1, idx_corner2, idx_surf1, idx_surf2, idx_surf3;
  int iters_done = 0, published = 0, num_point_sel_last = 0;
  Cloud compact_data;

  PointOdometry(float sp, int io, size_t it) : scan_period(sp), time_factor(1 / sp), io_ratio(io), num_max_iterations(it) {}

  void TransformToStart(const PointXYZI &pi, PointXYZI &po) const {
    float s = time_factor * (pi.intensity - int(pi.intensity));
    if (no_deskew) s = 0;
    if (s < 0 || s > 1.001) { po = pi; return; }
    po.x = pi.x - s * transform_es.pos.x;
    po.y = pi.y - s * transform_es.pos.y;
    po.z = pi.z - s * transform_es.pos.z;
    po.intensity = pi.intensity;
    Quat<float> q_id, q_e = transform_es.rot;
    Quat<float> q_s = q_id.slerp(s, q_e);
    RotatePoint(q_s.conjugate(), po);
  }

  size_t TransformToEnd(Cloud &cloud) const {
    size_t cloud_size = cloud.size();
    for (size_t i = 0; i < cloud_size; i++) {
      PointXYZI &point = cloud[i];
      float s = time_factor * (point.intensity - int(point.intensity));
      if (no_deskew) s = 0;
      point.x -= s * transform_es.pos.x;
      point.y -= s * transform_es.pos.y;
      point.z -= s * transform_es.pos.z;
      point.intensity = int(point.intensity);
      Quat<float> q_id, q_e = transform_es.rot;
      Quat<float> q_s = q_id.slerp(s, q_e);
      RotatePoint(q_s.conjugate(), point);
      RotatePoint(q_e, point);
      point.x += transform_es.pos.x;
      point.y += transform_es.pos.y;
      point.z += transform_es.pos.z;
    }
    return cloud_size;
  }

  void Process() {
    iters_done = 0; published = 0; num_point_sel_last = 0;
    if (!system_inited) {  // :302-310
      corner_points_less_sharp.swap(last_corner_cloud);
      surf_points_less_flat.swap(last_surf_cloud);
      kdtree_corner_last.Build(last_corner_cloud);
      kdtree_surf_last.Build(last_surf_cloud);
      system_inited = true;
      return;
    }
    PointXYZI coeff;
    bool is_degenerate = false;
    float mat_P[6][6] = {};
    ++frame_count;
    size_t last_corner_size = last_corner_cloud.size();
    size_t last_surf_size = last_surf_cloud.size();
    if (enable_odom) {
      if (last_corner_size > 10 && last_surf_size > 100) {
        int point_search_idx[1];
        float point_search_sq_dis[1];
        int num_curr_corner_points_sharp = (int)corner_points_sharp.size();
        int num_curr_surf_points_flat = (int)surf_points_flat.size();
        idx_corner1.assign(num_curr_corner_points_sharp, 0); idx_corner2.assign(num_curr_corner_points_sharp, 0);
        idx_surf1.assign(num_curr_surf_points_flat, 0); idx_surf2.assign(num_curr_surf_points_flat, 0); idx_surf3.assign(num_curr_surf_points_flat, 0);
        Cloud laser_cloud_ori, coeff_sel;
        for (size_t iter_count = 0; iter_count < num_max_iterations; ++iter_count) {
          iters_done = (int)iter_count + 1;
          PointXYZI point_sel, tripod1, tripod2, tripod3;
          laser_cloud_ori.clear();
          coeff_sel.clear();
          for (int i = 0; i < num_curr_corner_points_sharp; ++i) {
            TransformToStart(corner_points_sharp[i], point_sel);
            if (iter_count % 5 == 0) {
              kdtree_corner_last.Knn(point_sel, 1, point_search_idx, point_search_sq_dis);
              int closest_point_idx = -1, second_closet_point_idx = -1;
              if (point_search_sq_dis[0] < 25) {
                closest_point_idx = point_search_idx[0];
                int closest_point_scan = int(last_corner_cloud[closest_point_idx].intensity);
                float point_sq_dis, second_point_sq_dis = 25;
                for (int j = closest_point_idx + 1; j < (int)last_corner_size; j++) {
                  if (int(last_corner_cloud[j].intensity) > closest_point_scan + 2.5) break;
                  point_sq_dis = CalcSquaredDiff(last_corner_cloud[j], point_sel);
                  if (int(last_corner_cloud[j].intensity) > closest_point_scan) {
                    if (point_sq_dis < second_point_sq_dis) { second_point_sq_dis = point_sq_dis; second_closet_point_idx = j; }
                  }
                }
                for (int j = closest_point_idx - 1; j >= 0; j--) {
                  if (int(last_corner_cloud[j].intensity) < closest_point_scan - 2.5) break;
                  point_sq_dis = CalcSquaredDiff(last_corner_cloud[j], point_sel);
                  if (int(last_corner_cloud[j].intensity) < closest_point_scan) {
                    if (point_sq_dis < second_point_sq_dis) { second_point_sq_dis = point_sq_dis; second_closet_point_idx = j; }
                  }
                }
              }
              idx_corner1[i] = closest_point_idx;
              idx_corner2[i] = second_closet_point_idx;
            }
            if (idx_corner2[i] >= 0) {
              tripod1 = last_corner_cloud[idx_corner1[i]];
              tripod2 = last_corner_cloud[idx_corner2[i]];
              float x0 = point_sel.x, y0 = point_sel.y, z0 = point_sel.z;
              float x1 = tripod1.x, y1 = tripod1.y, z1 = tripod1.z;
              float x2 = tripod2.x, y2 = tripod2.y, z2 = tripod2.z;
              float a012 = std::sqrt(((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1)) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1))
                                     + ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1)) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))
                                     + ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1)));
              float l12 = std::sqrt((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2) + (z1 - z2) * (z1 - z2));
              float la = ((y1 - y2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1))
                          + (z1 - z2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))) / a012 / l12;
              float lb = -((x1 - x2) * ((x0 - x1) * (y0 - y2) - (x0 - x2) * (y0 - y1))
                           - (z1 - z2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
              float lc = -((x1 - x2) * ((x0 - x1) * (z0 - z2) - (x0 - x2) * (z0 - z1))
                           + (y1 - y2) * ((y0 - y1) * (z0 - z2) - (y0 - y2) * (z0 - z1))) / a012 / l12;
              float ld2 = a012 / l12;
              float s = 1;
              if (iter_count >= 5) s = 1 - 1.8f * std::fabs(ld2);
              coeff.x = s * la; coeff.y = s * lb; coeff.z = s * lc; coeff.intensity = s * ld2;
              if (s > 0.1 && ld2 != 0) { laser_cloud_ori.push_back(corner_points_sharp[i]); coeff_sel.push_back(coeff); }
            }
          }
          for (int i = 0; i < num_curr_surf_points_flat; ++i) {
            TransformToStart(surf_points_flat[i], point_sel);
            if (iter_count % 5 == 0) {
              kdtree_surf_last.Knn(point_sel, 1, point_search_idx, point_search_sq_dis);
              int closest_point_idx = -1, second_closet_point_idx = -1, third_clost_point_idx = -1;
              if (point_search_sq_dis[0] < 25) {
                closest_point_idx = point_search_idx[0];
                int closestPointScan = int(last_surf_cloud[closest_point_idx].intensity);
                float point_sq_dis, point_sq_dis2 = 25, point_sq_dis3 = 25;
                for (int j = closest_point_idx + 1; j < (int)last_surf_size; j++) {
                  if (int(last_surf_cloud[j].intensity) > closestPointScan + 2.5) break;
                  point_sq_dis = CalcSquaredDiff(last_surf_cloud[j], point_sel);
                  if (int(last_surf_cloud[j].intensity) <= closestPointScan) {
                    if (point_sq_dis < point_sq_dis2) { point_sq_dis2 = point_sq_dis; second_closet_point_idx = j; }
                  } else {
                    if (point_sq_dis < point_sq_dis3) { point_sq_dis3 = point_sq_dis; third_clost_point_idx = j; }
                  }
                }
                for (int j = closest_point_idx - 1; j >= 0; j--) {
                  if (int(last_surf_cloud[j].intensity) < closestPointScan - 2.5) break;
                  point_sq_dis = CalcSquaredDiff(last_surf_cloud[j], point_sel);
                  if (int(last_surf_cloud[j].intensity) >= closestPointScan) {
                    if (point_sq_dis < point_sq_dis2) { point_sq_dis2 = point_sq_dis; second_closet_point_idx = j; }
                  } else {
                    if (point_sq_dis < point_sq_dis3) { point_sq_dis3 = point_sq_dis; third_clost_point_idx = j; }
                  }
                }
              }
              idx_surf1[i] = closest_point_idx;
              idx_surf2[i] = second_closet_point_idx;
              idx_surf3[i] = third_clost_point_idx;
            }
            if (idx_surf2[i] >= 0 && idx_surf3[i] >= 0) {
              tripod1 = last_surf_cloud[idx_surf1[i]];
              tripod2 = last_surf_cloud[idx_surf2[i]];
              tripod3 = last_surf_cloud[idx_surf3[i]];
              float pa = (tripod2.y - tripod1.y) * (tripod3.z - tripod1.z) - (tripod3.y - tripod1.y) * (tripod2.z - tripod1.z);
              float pb = (tripod2.z - tripod1.z) * (tripod3.x - tripod1.x) - (tripod3.z - tripod1.z) * (tripod2.x - tripod1.x);
              float pc = (tripod2.x - tripod1.x) * (tripod3.y - tripod1.y) - (tripod3.x - tripod1.x) * (tripod2.y - tripod1.y);
              float pd = -(pa * tripod1.x + pb * tripod1.y + pc * tripod1.z);
              float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
              pa /= ps; pb /= ps; pc /= ps; pd /= ps;
              float pd2 = pa * point_sel.x + pb * point_sel.y + pc * point_sel.z + pd;
              float s = 1;
              if (iter_count >= 5) s = 1 - 1.8f * std::fabs(pd2) / std::sqrt(CalcPointDistance(point_sel));
              coeff.x = s * pa; coeff.y = s * pb; coeff.z = s * pc; coeff.intensity = s * pd2;
              if (s > 0.1 && pd2 != 0) { laser_cloud_ori.push_back(surf_points_flat[i]); coeff_sel.push_back(coeff); }
            }
          }
          int num_point_sel = (int)laser_cloud_ori.size();
          num_point_sel_last = num_point_sel;
          if (num_point_sel < 10) continue;
          Quat<float> R_SO3 = transform_es.rot;  // SO3 R_SO3(transform_es_.rot): Sophus normalises
          R_SO3.normalize();
          float AtA[6][6] = {}, AtB[6] = {};
          Mat3<float> Rm = transform_es.rot.toRotationMatrix();
          for (int i = 0; i < num_point_sel; ++i) {
            const PointXYZI &point_ori = laser_cloud_ori[i];
            coeff = coeff_sel[i];
            Vec3<float> p(point_ori.x, point_ori.y, point_ori.z), w(coeff.x, coeff.y, coeff.z);
            Vec3<float> p_minus_t = p - transform_es.pos;
            Mat3<float> S = Skew(transform_es.rot.conjugate() * p_minus_t);
            float row[6];
            for (int c = 0; c < 3; ++c) row[c] = w.x * S(0, c) + w.y * S(1, c) + w.z * S(2, c);            // J_r = w^T [.]x
            for (int c = 0; c < 3; ++c) row[3 + c] = (-w.x) * Rm(c, 0) + (-w.y) * Rm(c, 1) + (-w.z) * Rm(c, 2);  // J_t = -w^T R^T
            float d2 = coeff.intensity;
            float b = (float)(-0.1 * d2);
            for (int a = 0; a < 6; ++a) {
              for (int c = 0; c < 6; ++c) AtA[a][c] += row[a] * row[c];
              AtB[a] += row[a] * b;
            }
          }
          float Aw[6][6], Bw[6], X[6];
          for (int a = 0; a < 6; ++a) { for (int c = 0; c < 6; ++c) Aw[a][c] = AtA[a][c]; Bw[a] = AtB[a]; }
          colpiv_householder_qr_solve<float, 6, 6>(Aw, Bw, X);
          if (iter_count == 0) {
            float E[6], V[36], V2[36];
            sym_eigen_jacobi<float>(6, &AtA[0][0], E, V);
            for (int k = 0; k < 36; ++k) V2[k] = V[k];
            is_degenerate = false;
            const float eign_thre[6] = {10, 10, 10, 10, 10, 10};
            for (int i = 0; i < 6; i++) {
              if (E[i] < eign_thre[i]) {
                for (int j = 0; j < 6; j++) V2[i * 6 + j] = 0;
                is_degenerate = true;
              } else break;
            }
            for (int a = 0; a < 6; ++a)  // mat_P = mat_V2 * mat_V.inverse(); V orthogonal
              for (int c = 0; c < 6; ++c) { float s = 0; for (int k = 0; k < 6; ++k) s += V2[a * 6 + k] * V[c * 6 + k]; mat_P[a][c] = s; }
          }
          if (is_degenerate) {
            float X2[6];
            for (int a = 0; a < 6; ++a) { float s = 0; for (int c = 0; c < 6; ++c) s += mat_P[a][c] * X[c]; X2[a] = s; }
            for (int a = 0; a < 6; ++a) X[a] = X2[a];
          }
          transform_es.pos.x += X[3]; transform_es.pos.y += X[4]; transform_es.pos.z += X[5];
          transform_es.rot = transform_es.rot * DeltaQ(Vec3<float>(X[0], X[1], X[2]));
          if (!std::isfinite(transform_es.pos.x)) transform_es.pos.x = 0.0f;
          if (!std::isfinite(transform_es.pos.y)) transform_es.pos.y = 0.0f;
          if (!std::isfinite(transform_es.pos.z)) transform_es.pos.z = 0.0f;
          float delta_r = (float)((double)R_SO3.angularDistance(transform_es.rot) * 180.0 / M_PI);
          float delta_t = (float)std::sqrt(std::pow((double)(X[3] * 100), 2) + std::pow((double)(X[4] * 100), 2) + std::pow((double)(X[5] * 100), 2));
          if (delta_r < delta_r_abort && delta_t < delta_t_abort) break;
        }
      }
      Transform transform_se = transform_es.inverse();
      Transform transform_sum_tmp = transform_sum * transform_se;
      transform_sum = transform_sum_tmp;
      TransformToEnd(corner_points_less_sharp);
      TransformToEnd(surf_points_less_flat);
      transform_es.rot.normalize();
    }
    corner_points_less_sharp.swap(last_corner_cloud);
    surf_points_less_flat.swap(last_surf_cloud);
    last_corner_size = last_corner_cloud.size();
    last_surf_size = last_surf_cloud.size();
    if (last_corner_size > 10 && last_surf_size > 100) {
      kdtree_corner_last.Build(last_corner_cloud);
      kdtree_surf_last.Build(last_surf_cloud);
    }
    // PublishResults :726-765
    if (io_ratio < 2 || frame_count % io_ratio == 1) {
      if (enable_odom) TransformToEnd(full_cloud);
      CompactEncode(transform_sum, last_corner_cloud, last_surf_cloud, full_cloud, compact_data);
      published = 1;
    }
  }
};

}  // namespace orc

using namespace orc;
extern "C" {
void *orc_po_create(float scan_period, int io_ratio, int num_max_iterations) { return new PointOdometry(scan_period, io_ratio, (size_t)num_max_iterations); }
void orc_po_destroy(void *h) { delete (PointOdometry *)h; }
void orc_po_set_enable_odom(void *h, int en) { ((PointOdometry *)h)->enable_odom = en != 0; }
// info4: iterations executed, published, frame_count, matches of the last executed iteration
void orc_po_process(void *h, const float *sharp, int n_sharp, const float *less_sharp, int n_less_sharp, const float *flat, int n_flat,
                    const float *less_flat, int n_less_flat, const float *full, int n_full, float *transform_sum7, float *transform_es7, int *info4) {
  PointOdometry &o = *(PointOdometry *)h;
  auto load = [](Cloud &c, const float *p, int n) { c.assign((const PointXYZI *)p, (const PointXYZI *)p + n); };
  load(o.corner_points_sharp, sharp, n_sharp); load(o.corner_points_less_sharp, less_sharp, n_less_sharp);
  load(o.surf_points_flat, flat, n_flat); load(o.surf_points_less_flat, less_flat, n_less_flat); load(o.full_cloud, full, n_full);
  o.Process();
  const Transform &s = o.transform_sum, &e = o.transform_es;
  const float ts[7] = {s.rot.x, s.rot.y, s.rot.z, s.rot.w, s.pos.x, s.pos.y, s.pos.z};
  const float te[7] = {e.rot.x, e.rot.y, e.rot.z, e.rot.w, e.pos.x, e.pos.y, e.pos.z};
  std::memcpy(transform_sum7, ts, sizeof(ts)); std::memcpy(transform_es7, te, sizeof(te));
  info4[0] = o.iters_done; info4[1] = o.published; info4[2] = (int)o.frame_count; info4[3] = o.num_point_sel_last;
}
// which: 0 last_corner_cloud_, 1 last_surf_cloud_, 2 full_cloud_, 3 the /compact_data payload of the last published sweep
int orc_po_cloud_size(void *h, int which) {
  PointOdometry &o = *(PointOdometry *)h;
  const Cloud &c = which == 0 ? o.last_corner_cloud : which == 1 ? o.last_surf_cloud : which == 2 ? o.full_cloud : o.compact_data;
  return (int)c.size();
}
void orc_po_cloud_copy(void *h, int which, float *out) {
  PointOdometry &o = *(PointOdometry *)h;
  const Cloud &c = which == 0 ? o.last_corner_cloud : which == 1 ? o.last_surf_cloud : which == 2 ? o.full_cloud : o.compact_data;
  std::memcpy(out, c.data(), sizeof(PointXYZI) * c.size());
}
// match indices of the last search (iteration 0, 5, ...): kind 0 corner (2 per query), 1 surf (3 per query); returns the query count
int orc_po_matches(void *h, int kind, int *out) {
  PointOdometry &o = *(PointOdometry *)h;
  if (kind == 0) { for (size_t i = 0; i < o.idx_corner1.size(); ++i) { out[2 * i] = o.idx_corner1[i]; out[2 * i + 1] = o.idx_corner2[i]; } return (int)o.idx_corner1.size(); }
  for (size_t i = 0; i < o.idx_surf1.size(); ++i) { out[3 * i] = o.idx_surf1[i]; out[3 * i + 1] = o.idx_surf2[i]; out[3 * i + 2] = o.idx_surf3[i]; }
  return (int)o.idx_surf1.size();
}
}
