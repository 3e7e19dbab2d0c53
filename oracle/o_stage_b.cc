// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
//
// CPU restatement of stage B of hyye/lio-mapping (surf / point-to-plane branch; the corner branch
// is compiled out in the reference: USE_CORNER undefined, include/imu_processor/Estimator.h:55-56):
//   PointMapping::PointAssociateToMap   src/point_processor/PointMapping.cc:303-314
//   RotatePoint                         include/utils/geometry_utils.h:288-299
//   Estimator::CalculateFeatures        src/imu_processor/Estimator.cc:970-1097
//   Estimator::CalculateLaserOdom       src/imu_processor/Estimator.cc:1242-1359
//   PointMapping::OptimizeTransformTobeMapped  src/point_processor/PointMapping.cc:325-753 (scan-to-map GN)
// Eigen pieces restated in o_linalg.h (colPivHouseholderQr, SelfAdjointEigenSolver, quaternion).
#include "o_api.h"
#include <cmath>

namespace orc {

void PointAssociateToMap(const PointXYZI &pi, PointXYZI &po, const Transform &t) {
  Vec3<float> v(pi.x, pi.y, pi.z);
  Vec3<float> o = t.rot * v;  // RotatePoint: vec_out = q * vec
  po.x = o.x + t.pos.x;
  po.y = o.y + t.pos.y;
  po.z = o.z + t.pos.z;
  po.intensity = pi.intensity;
}

static inline float SqDiff(const PointXYZI &a, const PointXYZI &b) {  // math_utils.h:84-91
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}

void CalculateFeatures(const KdTree &kd, const Cloud &map, const Cloud &surf_stack, const Transform &local_transform,
                       const StageBConfig &cfg, std::vector<PointPlaneFeature> &features) {
  if (!cfg.keep_features) features.clear();  // Estimator.cc:978-980
  const int num_neighbors = 5;
  int point_search_idx[5];
  float point_search_sq_dis[5];
  PointXYZI point_ori, point_sel;
  const size_t surf_points_size = surf_stack.size();
  for (size_t i = 0; i < surf_points_size; i++) {
    point_ori = surf_stack[i];
    PointAssociateToMap(point_ori, point_sel, local_transform);
    kd.Knn(point_sel, num_neighbors, point_search_idx, point_search_sq_dis);
    if (point_search_sq_dis[num_neighbors - 1] < cfg.min_match_sq_dis) {
      float A[5][3], B[5], X[3];
      for (int j = 0; j < num_neighbors; j++) {
        A[j][0] = map[point_search_idx[j]].x;
        A[j][1] = map[point_search_idx[j]].y;
        A[j][2] = map[point_search_idx[j]].z;
        B[j] = -1.f;
      }
      colpiv_householder_qr_solve<float, 5, 3>(A, B, X);  // :1027
      float pa = X[0], pb = X[1], pc = X[2], pd = 1;
      float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
      pa /= ps; pb /= ps; pc /= ps; pd /= ps;
      bool planeValid = true;
      for (int j = 0; j < num_neighbors; j++) {
        const PointXYZI &m = map[point_search_idx[j]];
        if (std::fabs(pa * m.x + pb * m.y + pc * m.z + pd) > cfg.min_plane_dis) { planeValid = false; break; }
      }
      if (planeValid) {
        float pd2 = pa * point_sel.x + pb * point_sel.y + pc * point_sel.z + pd;
        // s = 1 - 0.9f*fabs(pd2)/sqrt(CalcPointDistance(point_sel))  (:1056, sqrt of the norm)
        float dist = std::sqrt(point_sel.x * point_sel.x + point_sel.y * point_sel.y + point_sel.z * point_sel.z);
        float s = 1 - 0.9f * std::fabs(pd2) / std::sqrt(dist);
        PointXYZI coeff1;
        coeff1.x = s * pa; coeff1.y = s * pb; coeff1.z = s * pc; coeff1.intensity = s * pd;
        bool is_in_laser_fov = false;
        PointXYZI transform_pos, point_on_z_axis;
        point_on_z_axis.x = 0.0f; point_on_z_axis.y = 0.0f; point_on_z_axis.z = 10.0f; point_on_z_axis.intensity = 0.f;
        PointAssociateToMap(point_on_z_axis, point_on_z_axis, local_transform);
        transform_pos.x = local_transform.pos.x; transform_pos.y = local_transform.pos.y; transform_pos.z = local_transform.pos.z;
        float squared_side1 = SqDiff(transform_pos, point_sel);
        float squared_side2 = SqDiff(point_on_z_axis, point_sel);
        float check1 = 100.0f + squared_side1 - squared_side2 - 10.0f * std::sqrt(3.0f) * std::sqrt(squared_side1);
        float check2 = 100.0f + squared_side1 - squared_side2 + 10.0f * std::sqrt(3.0f) * std::sqrt(squared_side1);
        if (check1 < 0 && check2 > 0) is_in_laser_fov = true;
        if (s > 0.1 && is_in_laser_fov) {
          PointPlaneFeature f;
          f.score = s;
          f.point[0] = point_ori.x; f.point[1] = point_ori.y; f.point[2] = point_ori.z;
          f.coeffs[0] = coeff1.x; f.coeffs[1] = coeff1.y; f.coeffs[2] = coeff1.z; f.coeffs[3] = coeff1.intensity;
          f.src_index = (int)i;
          features.push_back(f);
        }
      }
    }
  }
}

// Point-to-line branch of Estimator::CalculateFeatures (src/imu_processor/Estimator.cc:1101-1227; compiled out there
// because USE_CORNER is undefined, include/imu_processor/Estimator.h:55-56) == the live corner matching of
// PointMapping::OptimizeTransformTobeMapped (src/point_processor/PointMapping.cc:381-512): 5-NN in the corner map,
// centroid + 3x3 covariance, SelfAdjointEigenSolver, line iff lambda_3 > 3 lambda_2, and the line encoded as TWO
// half-weight plane-like features (normal_to_point and normal_cross_point).  The FOV test uses the query's own
// transform like the surf branch (:1063-1086; PointMapping keeps point_on_z_axis_ = T * (0,0,10) the same way).
// Mixed precision kept as written: `0.1 * mat_V1(k,2)` is a double product rounded to float on assignment.
// SelfAdjointEigenSolver<Matrix3f> is restated by cyclic Jacobi (o_linalg.h, parity unpinned w.r.t. Eigen's QL path;
// the eigenvector sign only flips coeff2, whose squared residual is unchanged).
void CalculateLineFeatures(const KdTree &kd, const Cloud &map, const Cloud &corner_stack, const Transform &local_transform,
                           const StageBConfig &cfg, std::vector<PointPlaneFeature> &features) {
  if (!cfg.keep_features) features.clear();
  int point_search_idx[5];
  float point_search_sq_dis[5];
  PointXYZI point_ori, point_sel, point_proj;
  PointXYZI point_on_z_axis;
  point_on_z_axis.x = 0.0f; point_on_z_axis.y = 0.0f; point_on_z_axis.z = 10.0f; point_on_z_axis.intensity = 0.f;
  PointAssociateToMap(point_on_z_axis, point_on_z_axis, local_transform);
  for (size_t i = 0; i < corner_stack.size(); i++) {
    point_ori = corner_stack[i];
    PointAssociateToMap(point_ori, point_sel, local_transform);
    kd.Knn(point_sel, 5, point_search_idx, point_search_sq_dis);
    if (!(point_search_sq_dis[4] < cfg.min_match_sq_dis)) continue;
    float vc[3] = {0, 0, 0};
    for (int j = 0; j < 5; j++) {
      const PointXYZI &m = map[point_search_idx[j]];
      vc[0] += m.x; vc[1] += m.y; vc[2] += m.z;
    }
    vc[0] /= 5.0f; vc[1] /= 5.0f; vc[2] /= 5.0f;  // Vector3f /= 5.0: the scalar is cast to float
    float a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
    for (int j = 0; j < 5; j++) {
      const PointXYZI &m = map[point_search_idx[j]];
      float ax = m.x - vc[0], ay = m.y - vc[1], az = m.z - vc[2];
      a00 += ax * ax; a10 += ax * ay; a20 += ax * az; a11 += ay * ay; a21 += ay * az; a22 += az * az;
    }
    // mat_A1 = mat_a / 5.0; only the lower triangle is written and SelfAdjointEigenSolver reads only it
    float A1[9];
    A1[0] = a00 / 5.0f; A1[4] = a11 / 5.0f; A1[8] = a22 / 5.0f;
    A1[3] = A1[1] = a10 / 5.0f; A1[6] = A1[2] = a20 / 5.0f; A1[7] = A1[5] = a21 / 5.0f;
    float D1[3], V1[9];
    sym_eigen_jacobi<float>(3, A1, D1, V1);
    if (!(D1[2] > 3 * D1[1])) continue;
    float x0 = point_sel.x, y0 = point_sel.y, z0 = point_sel.z;
    float x1 = (float)((double)vc[0] + 0.1 * (double)V1[0 * 3 + 2]);
    float y1 = (float)((double)vc[1] + 0.1 * (double)V1[1 * 3 + 2]);
    float z1 = (float)((double)vc[2] + 0.1 * (double)V1[2 * 3 + 2]);
    float x2 = (float)((double)vc[0] - 0.1 * (double)V1[0 * 3 + 2]);
    float y2 = (float)((double)vc[1] - 0.1 * (double)V1[1 * 3 + 2]);
    float z2 = (float)((double)vc[2] - 0.1 * (double)V1[2 * 3 + 2]);
    Vec3<float> X0(x0, y0, z0), X1(x1, y1, z1), X2(x2, y2, z2);
    Vec3<float> a012_vec = (X0 - X1).cross(X0 - X2);
    Vec3<float> l12_vec = X1 - X2;
    Vec3<float> ntp = l12_vec.cross(a012_vec);
    {  // .normalized(): v / sqrt(squaredNorm) when squaredNorm > 0
      float z = ntp.x * ntp.x + ntp.y * ntp.y + ntp.z * ntp.z;
      if (z > 0.f) { float nn = std::sqrt(z); ntp = Vec3<float>(ntp.x / nn, ntp.y / nn, ntp.z / nn); }
    }
    Vec3<float> ncp = l12_vec.cross(ntp);
    float a012 = std::sqrt(a012_vec.x * a012_vec.x + a012_vec.y * a012_vec.y + a012_vec.z * a012_vec.z);
    float l12 = std::sqrt(l12_vec.x * l12_vec.x + l12_vec.y * l12_vec.y + l12_vec.z * l12_vec.z);
    float la = ntp.x, lb = ntp.y, lc = ntp.z;
    float ld2 = a012 / l12;
    point_proj = point_sel;
    point_proj.x -= la * ld2; point_proj.y -= lb * ld2; point_proj.z -= lc * ld2;
    float ld_p1 = -(ntp.x * point_proj.x + ntp.y * point_proj.y + ntp.z * point_proj.z);
    float ld_p2 = -(ncp.x * point_proj.x + ncp.y * point_proj.y + ncp.z * point_proj.z);
    float s = 1 - 0.9f * std::fabs(ld2);
    PointXYZI coeff1, coeff2;
    coeff1.x = s * la; coeff1.y = s * lb; coeff1.z = s * lc; coeff1.intensity = s * ld_p1;
    coeff2.x = s * ncp.x; coeff2.y = s * ncp.y; coeff2.z = s * ncp.z; coeff2.intensity = s * ld_p2;
    PointXYZI transform_pos;
    transform_pos.x = local_transform.pos.x; transform_pos.y = local_transform.pos.y; transform_pos.z = local_transform.pos.z;
    float squared_side1 = SqDiff(transform_pos, point_sel);
    float squared_side2 = SqDiff(point_on_z_axis, point_sel);
    float check1 = 100.0f + squared_side1 - squared_side2 - 10.0f * std::sqrt(3.0f) * std::sqrt(squared_side1);
    float check2 = 100.0f + squared_side1 - squared_side2 + 10.0f * std::sqrt(3.0f) * std::sqrt(squared_side1);
    bool is_in_laser_fov = (check1 < 0 && check2 > 0);
    if (s > 0.1 && is_in_laser_fov) {
      for (int h = 0; h < 2; ++h) {
        const PointXYZI &c = h == 0 ? coeff1 : coeff2;
        PointPlaneFeature f;
        f.score = s * 0.5;
        f.point[0] = point_ori.x; f.point[1] = point_ori.y; f.point[2] = point_ori.z;
        f.coeffs[0] = c.x * 0.5; f.coeffs[1] = c.y * 0.5; f.coeffs[2] = c.z * 0.5; f.coeffs[3] = c.intensity * 0.5;
        f.src_index = (int)i;
        features.push_back(f);
      }
    }
  }
}

void CalculateLaserOdom(const KdTree &kd, const Cloud &map, const Cloud &surf_stack, Transform &local_transform,
                        const StageBConfig &cfg, std::vector<PointPlaneFeature> &features, int *iters_done) {
  bool is_degenerate = false;
  float matP[6][6] = {};
  int it_done = 0;
  for (size_t iter_count = 0; iter_count < (size_t)cfg.num_max_iterations; ++iter_count) {
    it_done = (int)iter_count + 1;
    CalculateFeatures(kd, map, surf_stack, local_transform, cfg, features);
    size_t n = features.size();
    Quat<float> R_SO3 = local_transform.rot;  // SO3 R_SO3(local_transform.rot): Sophus normalises
    R_SO3.normalize();
    // mat_A rows [J_r, J_t], mat_B = -d2 (:1276-1300); AtA/AtB accumulated sequentially in float
    float AtA[6][6] = {}, AtB[6] = {};
    Mat3<float> Rm = local_transform.rot.toRotationMatrix();
    for (size_t i = 0; i < n; i++) {
      const PointPlaneFeature &f = features[i];
      Vec3<float> p((float)f.point[0], (float)f.point[1], (float)f.point[2]);
      Vec3<float> w((float)f.coeffs[0], (float)f.coeffs[1], (float)f.coeffs[2]);
      float b = (float)f.coeffs[3];
      // J_r = -w^T (R * skew(p)),  J_t = w^T
      Mat3<float> RS = Rm * Skew(p);
      float row[6];
      for (int c = 0; c < 3; ++c) row[c] = -(w.x * RS(0, c) + w.y * RS(1, c) + w.z * RS(2, c));
      row[3] = w.x; row[4] = w.y; row[5] = w.z;
      Vec3<float> rp = local_transform.rot * p + local_transform.pos;  // quaternion transform (:1282)
      float d2 = w.dot(rp) + b;
      for (int a = 0; a < 6; ++a) {
        for (int c = 0; c < 6; ++c) AtA[a][c] += row[a] * row[c];
        AtB[a] += row[a] * (-d2);
      }
    }
    float Aw[6][6], Bw[6], X[6];
    for (int a = 0; a < 6; ++a) { for (int c = 0; c < 6; ++c) Aw[a][c] = AtA[a][c]; Bw[a] = AtB[a]; }
    colpiv_householder_qr_solve<float, 6, 6>(Aw, Bw, X);  // :1306
    if (iter_count == 0) {  // :1308-1339
      float E[6], V[36];
      sym_eigen_jacobi<float>(6, &AtA[0][0], E, V);
      // mat_V = esolver.eigenvectors() (columns); mat_V2 = mat_V with ROW i zeroed for small E[i]
      float V2[36];
      for (int k = 0; k < 36; ++k) V2[k] = V[k];
      is_degenerate = false;
      const float eignThre[6] = {100, 100, 100, 100, 100, 100};
      for (int i = 0; i < 6; ++i) {
        if (E[i] < eignThre[i]) {
          for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0;
          is_degenerate = true;
        } else break;
      }
      // matP = mat_V2 * mat_V.inverse(); V orthogonal => inverse = transpose
      for (int a = 0; a < 6; ++a)
        for (int c = 0; c < 6; ++c) {
          float s = 0;
          for (int k = 0; k < 6; ++k) s += V2[a * 6 + k] * V[c * 6 + k];
          matP[a][c] = s;
        }
    }
    if (is_degenerate) {
      float X2[6];
      for (int a = 0; a < 6; ++a) { float s = 0; for (int c = 0; c < 6; ++c) s += matP[a][c] * X[c]; X2[a] = s; }
      for (int a = 0; a < 6; ++a) X[a] = X2[a];
    }
    local_transform.pos.x += X[3];
    local_transform.pos.y += X[4];
    local_transform.pos.z += X[5];
    local_transform.rot = local_transform.rot * DeltaQ(Vec3<float>(X[0], X[1], X[2]));
    if (!std::isfinite(local_transform.pos.x)) local_transform.pos.x = 0.0f;
    if (!std::isfinite(local_transform.pos.y)) local_transform.pos.y = 0.0f;
    if (!std::isfinite(local_transform.pos.z)) local_transform.pos.z = 0.0f;
    float ad = R_SO3.angularDistance(local_transform.rot);
    float delta_r = (float)(ad * 180.0 / M_PI);
    float delta_t = std::sqrt(std::pow(X[3] * 100, 2) + std::pow(X[4] * 100, 2) + std::pow(X[5] * 100, 2));
    if (delta_r < cfg.delta_r_abort && delta_t < cfg.delta_t_abort) break;
  }
  if (iters_done) *iters_done = it_done;
}

// PointMapping::OptimizeTransformTobeMapped (src/point_processor/PointMapping.cc:325-753), scan-to-map 6-DoF float
// Gauss-Newton against explicit corner / surf maps (the cube-map store that selects them is outside this operator):
//   corner matching :381-512  (ONE feature per line: coeff = s (la, lb, lc), intensity = s ld2)
//   surf matching   :514-606  (sign-normalised plane: coeff.intensity = s |pd2|)
//   6x6 GN          :608-715  (rows [-w^T (R [p]x), w^T], rhs -coeff.intensity; first-iteration degeneracy projection)
// point_on_z_axis_ is set ONCE from the initial transform (PointMapping.cc:803-806) and not moved by the iterations.
// features_out (optional): the (point_ori, coeff) pairs of the LAST executed feature pass, corner then surf.
// variant 1 = MapBuilder::OptimizeMap (src/map_builder/MapBuilder.cc:624-1014), the same loop with a rotation
// information matrix on the Jacobian, J_r = -w^T (R [p]x) R^-1 diag(5e-3, 5e-3, 1) (:905-911, non-DEBUG branch), and a
// LEFT-multiplicative update rot = DeltaQ(x) * rot (:984-985).
void OptimizeTransformTobeMapped(const Cloud &corner_map, const Cloud &surf_map, const Cloud &corner_stack, const Cloud &surf_stack,
                                 Transform &tobe, const StageBConfig &cfg, int *iters_done, std::vector<PointPlaneFeature> *features_out,
                                 int variant) {
  if (iters_done) *iters_done = 0;
  if (corner_map.size() <= 10 || surf_map.size() <= 100) return;  // :327-329
  KdTree kd_corner, kd_surf;
  kd_corner.Build(corner_map);
  kd_surf.Build(surf_map);
  PointXYZI point_on_z_axis;
  point_on_z_axis.x = 0.0f; point_on_z_axis.y = 0.0f; point_on_z_axis.z = 10.0f; point_on_z_axis.intensity = 0.f;
  PointAssociateToMap(point_on_z_axis, point_on_z_axis, tobe);
  bool is_degenerate = false;
  float matP[6][6] = {};
  for (int a = 0; a < 6; ++a) matP[a][a] = 1.f;
  int point_search_idx[5];
  float point_search_sq_dis[5];
  Cloud laser_cloud_ori, coeff_sel;
  std::vector<int> src;
  int it_done = 0;
  for (size_t iter_count = 0; iter_count < (size_t)cfg.num_max_iterations; ++iter_count) {
    it_done = (int)iter_count + 1;
    laser_cloud_ori.clear(); coeff_sel.clear(); src.clear();
    PointXYZI point_ori, point_sel, coeff;
    auto in_fov = [&](const PointXYZI &ps) {
      PointXYZI transform_pos;
      transform_pos.x = tobe.pos.x; transform_pos.y = tobe.pos.y; transform_pos.z = tobe.pos.z;
      float squared_side1 = SqDiff(transform_pos, ps);
      float squared_side2 = SqDiff(point_on_z_axis, ps);
      float check1 = 100.0f + squared_side1 - squared_side2 - 10.0f * std::sqrt(3.0f) * std::sqrt(squared_side1);
      float check2 = 100.0f + squared_side1 - squared_side2 + 10.0f * std::sqrt(3.0f) * std::sqrt(squared_side1);
      return check1 < 0 && check2 > 0;
    };
    for (size_t i = 0; i < corner_stack.size(); ++i) {
      point_ori = corner_stack[i];
      PointAssociateToMap(point_ori, point_sel, tobe);
      kd_corner.Knn(point_sel, 5, point_search_idx, point_search_sq_dis);
      if (!(point_search_sq_dis[4] < cfg.min_match_sq_dis)) continue;
      float vc[3] = {0, 0, 0};
      for (int j = 0; j < 5; j++) { const PointXYZI &m = corner_map[point_search_idx[j]]; vc[0] += m.x; vc[1] += m.y; vc[2] += m.z; }
      vc[0] /= 5.0f; vc[1] /= 5.0f; vc[2] /= 5.0f;
      float a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
      for (int j = 0; j < 5; j++) {
        const PointXYZI &m = corner_map[point_search_idx[j]];
        float ax = m.x - vc[0], ay = m.y - vc[1], az = m.z - vc[2];
        a00 += ax * ax; a10 += ax * ay; a20 += ax * az; a11 += ay * ay; a21 += ay * az; a22 += az * az;
      }
      float A1[9], D1[3], V1[9];
      A1[0] = a00 / 5.0f; A1[4] = a11 / 5.0f; A1[8] = a22 / 5.0f;
      A1[3] = A1[1] = a10 / 5.0f; A1[6] = A1[2] = a20 / 5.0f; A1[7] = A1[5] = a21 / 5.0f;
      sym_eigen_jacobi<float>(3, A1, D1, V1);
      if (!(D1[2] > 3 * D1[1])) continue;
      float x1 = (float)((double)vc[0] + 0.1 * (double)V1[2]), y1 = (float)((double)vc[1] + 0.1 * (double)V1[5]),
            z1 = (float)((double)vc[2] + 0.1 * (double)V1[8]);
      float x2 = (float)((double)vc[0] - 0.1 * (double)V1[2]), y2 = (float)((double)vc[1] - 0.1 * (double)V1[5]),
            z2 = (float)((double)vc[2] - 0.1 * (double)V1[8]);
      Vec3<float> X0(point_sel.x, point_sel.y, point_sel.z), X1(x1, y1, z1), X2(x2, y2, z2);
      Vec3<float> a012_vec = (X0 - X1).cross(X0 - X2);
      Vec3<float> l12_vec = X1 - X2;
      Vec3<float> ntp = l12_vec.cross(a012_vec);
      {
        float z = ntp.x * ntp.x + ntp.y * ntp.y + ntp.z * ntp.z;
        if (z > 0.f) { float nn = std::sqrt(z); ntp = Vec3<float>(ntp.x / nn, ntp.y / nn, ntp.z / nn); }
      }
      float a012 = std::sqrt(a012_vec.x * a012_vec.x + a012_vec.y * a012_vec.y + a012_vec.z * a012_vec.z);
      float l12 = std::sqrt(l12_vec.x * l12_vec.x + l12_vec.y * l12_vec.y + l12_vec.z * l12_vec.z);
      float ld2 = a012 / l12;
      float sc = 1 - 0.9f * std::fabs(ld2);
      coeff.x = sc * ntp.x; coeff.y = sc * ntp.y; coeff.z = sc * ntp.z; coeff.intensity = sc * ld2;
      if (sc > 0.1 && in_fov(point_sel)) { laser_cloud_ori.push_back(point_ori); coeff_sel.push_back(coeff); src.push_back((int)i); }
    }
    for (size_t i = 0; i < surf_stack.size(); ++i) {
      point_ori = surf_stack[i];
      PointAssociateToMap(point_ori, point_sel, tobe);
      kd_surf.Knn(point_sel, 5, point_search_idx, point_search_sq_dis);
      if (!(point_search_sq_dis[4] < cfg.min_match_sq_dis)) continue;
      float A[5][3], B[5], X[3];
      for (int j = 0; j < 5; j++) {
        A[j][0] = surf_map[point_search_idx[j]].x; A[j][1] = surf_map[point_search_idx[j]].y; A[j][2] = surf_map[point_search_idx[j]].z;
        B[j] = -1.f;
      }
      colpiv_householder_qr_solve<float, 5, 3>(A, B, X);
      float pa = X[0], pb = X[1], pc = X[2], pd = 1;
      float ps = std::sqrt(pa * pa + pb * pb + pc * pc);
      pa /= ps; pb /= ps; pc /= ps; pd /= ps;
      bool planeValid = true;
      for (int j = 0; j < 5; j++) {
        const PointXYZI &m = surf_map[point_search_idx[j]];
        if (std::fabs(pa * m.x + pb * m.y + pc * m.z + pd) > cfg.min_plane_dis) { planeValid = false; break; }
      }
      if (!planeValid) continue;
      float pd2 = pa * point_sel.x + pb * point_sel.y + pc * point_sel.z + pd;
      float dist = std::sqrt(point_sel.x * point_sel.x + point_sel.y * point_sel.y + point_sel.z * point_sel.z);
      float sc = 1 - 0.9f * std::fabs(pd2) / std::sqrt(dist);
      if (pd2 > 0) { coeff.x = sc * pa; coeff.y = sc * pb; coeff.z = sc * pc; coeff.intensity = sc * pd2; }
      else { coeff.x = -sc * pa; coeff.y = -sc * pb; coeff.z = -sc * pc; coeff.intensity = -sc * pd2; }
      if (sc > 0.1 && in_fov(point_sel)) { laser_cloud_ori.push_back(point_ori); coeff_sel.push_back(coeff); src.push_back((int)i); }
    }
    const size_t n = laser_cloud_ori.size();
    if (n < 50) continue;  // :609-611
    Quat<float> R_SO3 = tobe.rot;
    R_SO3.normalize();
    float AtA[6][6] = {}, AtB[6] = {};
    Mat3<float> Rm = tobe.rot.toRotationMatrix();
    Mat3<float> Rinv = tobe.rot.inverse().toRotationMatrix();
    for (size_t i = 0; i < n; i++) {
      Vec3<float> p(laser_cloud_ori[i].x, laser_cloud_ori[i].y, laser_cloud_ori[i].z);
      Vec3<float> w(coeff_sel[i].x, coeff_sel[i].y, coeff_sel[i].z);
      Mat3<float> RS = Rm * Skew(p);
      float row[6];
      for (int c = 0; c < 3; ++c) row[c] = (-w.x) * RS(0, c) + (-w.y) * RS(1, c) + (-w.z) * RS(2, c);
      if (variant == 1) {
        float t3[3];
        for (int c = 0; c < 3; ++c) t3[c] = row[0] * Rinv(0, c) + row[1] * Rinv(1, c) + row[2] * Rinv(2, c);
        row[0] = t3[0] * 5e-3f; row[1] = t3[1] * 5e-3f; row[2] = t3[2] * 1.0f;
      }
      row[3] = w.x; row[4] = w.y; row[5] = w.z;
      float d2 = coeff_sel[i].intensity;
      for (int a = 0; a < 6; ++a) {
        for (int c = 0; c < 6; ++c) AtA[a][c] += row[a] * row[c];
        AtB[a] += row[a] * (-d2);
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
      for (int i = 0; i < 6; ++i) {
        if (E[i] < 100.f) { for (int j = 0; j < 6; ++j) V2[i * 6 + j] = 0; is_degenerate = true; }
        else break;
      }
      for (int a = 0; a < 6; ++a)
        for (int c = 0; c < 6; ++c) { float sm = 0; for (int k = 0; k < 6; ++k) sm += V2[a * 6 + k] * V[c * 6 + k]; matP[a][c] = sm; }
    }
    if (is_degenerate) {
      float X2[6];
      for (int a = 0; a < 6; ++a) { float sm = 0; for (int c = 0; c < 6; ++c) sm += matP[a][c] * X[c]; X2[a] = sm; }
      for (int a = 0; a < 6; ++a) X[a] = X2[a];
    }
    tobe.pos.x += X[3]; tobe.pos.y += X[4]; tobe.pos.z += X[5];
    if (variant == 1) tobe.rot = DeltaQ(Vec3<float>(X[0], X[1], X[2])) * tobe.rot;
    else tobe.rot = tobe.rot * DeltaQ(Vec3<float>(X[0], X[1], X[2]));
    if (!std::isfinite(tobe.pos.x)) tobe.pos.x = 0.0f;
    if (!std::isfinite(tobe.pos.y)) tobe.pos.y = 0.0f;
    if (!std::isfinite(tobe.pos.z)) tobe.pos.z = 0.0f;
    float ad = R_SO3.angularDistance(tobe.rot);
    float delta_r = (float)(ad * 180.0 / M_PI);
    float delta_t = std::sqrt(std::pow(X[3] * 100, 2) + std::pow(X[4] * 100, 2) + std::pow(X[5] * 100, 2));
    if (delta_r < cfg.delta_r_abort && delta_t < cfg.delta_t_abort) break;
  }
  if (iters_done) *iters_done = it_done;
  if (features_out) {
    features_out->clear();
    for (size_t i = 0; i < laser_cloud_ori.size(); ++i) {
      PointPlaneFeature f;
      f.score = 0;
      f.point[0] = laser_cloud_ori[i].x; f.point[1] = laser_cloud_ori[i].y; f.point[2] = laser_cloud_ori[i].z;
      f.coeffs[0] = coeff_sel[i].x; f.coeffs[1] = coeff_sel[i].y; f.coeffs[2] = coeff_sel[i].z; f.coeffs[3] = coeff_sel[i].intensity;
      f.src_index = src[i];
      features_out->push_back(f);
    }
  }
}

}  // namespace orc
