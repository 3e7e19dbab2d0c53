// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
//
// CPU restatement of the rolling cube map of lio::PointMapping (the store that feeds OptimizeTransformTobeMapped):
//   constants               src/point_processor/PointMapping.cc:77-82, :121-122  (21 x 21 x 11 cubes of 50 m, centre (10,10,5))
//   ToIndex / FromIndex     include/point_processor/PointMapping.h:150-159
//   re-centring             PointMapping.cc:809-931   (shift the cube arrays until the sensor cube is >= 3 from every face)
//   cube selection          PointMapping.cc:944-1003  (5 x 5 x 5 neighbourhood; a cube is "valid" when one of its eight
//                                                      corners lies within +-60 deg of the sensor's z axis)
//   map extraction          PointMapping.cc:1005-1011 (concatenate the valid cubes, corner and surf separately)
//   UpdateMapDatabase       PointMapping.cc:1112-1208 (insert the down-sampled stacks, re-filter the touched valid cubes)
//   PointMapping::Process   PointMapping.cc:765-1052  (imu_inited_ == false path, num_stack_frames_ == 1)
// Device counterpart: lio_mapping_b200/csrc/cubemap.cu (lio_pm_*), compared cube by cube in tests/test_point_mapping_gpu.py.
#include "o_api.h"
#include <cmath>

namespace orc {

struct CubeMap {
  static constexpr int kL = 21, kW = 21, kH = 11;
  int cen_l = 10, cen_w = 10, cen_h = 5;
  std::vector<Cloud> corner, surf;
  float corner_leaf = 0.2f, surf_leaf = 0.4f;
  CubeMap() : corner(kL * kW * kH), surf(kL * kW * kH) {}
  static size_t ToIndex(int i, int j, int k) { return (size_t)i + (size_t)kL * j + (size_t)kL * kW * k; }
  static void FromIndex(size_t index, int &i, int &j, int &k) {
    int residual = (int)(index % (kL * kW));
    k = (int)(index / (kL * kW));
    j = residual / kL;
    i = residual % kL;
  }
  static int CubeOf(float v, int cen) {  // int((v + 25.0) / 50.0) + cen, minus one for negatives (:812-819)
    int c = int(((double)v + 25.0) / 50.0) + cen;
    if ((double)v + 25.0 < 0) --c;
    return c;
  }
  // :821-931; returns the sensor's cube after the shifts
  void Recentre(const Vec3<float> &pos, int &ci, int &cj, int &ck) {
    ci = CubeOf(pos.x, cen_l); cj = CubeOf(pos.y, cen_w); ck = CubeOf(pos.z, cen_h);
    auto shift = [&](int axis, int dir) {  // dir +1: contents move towards higher indices (the low face is cleared)
      const int n[3] = {kL, kW, kH};
      int idx[3];
      for (idx[(axis + 1) % 3] = 0; idx[(axis + 1) % 3] < n[(axis + 1) % 3]; ++idx[(axis + 1) % 3])
        for (idx[(axis + 2) % 3] = 0; idx[(axis + 2) % 3] < n[(axis + 2) % 3]; ++idx[(axis + 2) % 3]) {
          if (dir > 0) {
            for (int a = n[axis] - 1; a >= 1; --a) {
              idx[axis] = a; const size_t ia = ToIndex(idx[0], idx[1], idx[2]);
              idx[axis] = a - 1; const size_t ib = ToIndex(idx[0], idx[1], idx[2]);
              std::swap(corner[ia], corner[ib]); std::swap(surf[ia], surf[ib]);
            }
            idx[axis] = 0;
          } else {
            for (int a = 0; a < n[axis] - 1; ++a) {
              idx[axis] = a; const size_t ia = ToIndex(idx[0], idx[1], idx[2]);
              idx[axis] = a + 1; const size_t ib = ToIndex(idx[0], idx[1], idx[2]);
              std::swap(corner[ia], corner[ib]); std::swap(surf[ia], surf[ib]);
            }
            idx[axis] = n[axis] - 1;
          }
          const size_t ic = ToIndex(idx[0], idx[1], idx[2]);
          corner[ic].clear(); surf[ic].clear();
        }
    };
    while (ci < 3) { shift(0, +1); ++ci; ++cen_l; }
    while (ci >= kL - 3) { shift(0, -1); --ci; --cen_l; }
    while (cj < 3) { shift(1, +1); ++cj; ++cen_w; }
    while (cj >= kW - 3) { shift(1, -1); --cj; --cen_w; }
    while (ck < 3) { shift(2, +1); ++ck; ++cen_h; }
    while (ck >= kH - 3) { shift(2, -1); --ck; --cen_h; }
  }
  // :944-1003
  void Select(const Vec3<float> &pos, const PointXYZI &point_on_z_axis, int ci, int cj, int ck, std::vector<size_t> &valid,
              std::vector<size_t> &surround) const {
    valid.clear(); surround.clear();
    for (int i = ci - 2; i <= ci + 2; ++i)
      for (int j = cj - 2; j <= cj + 2; ++j)
        for (int k = ck - 2; k <= ck + 2; ++k) {
          if (!(i >= 0 && i < kL && j >= 0 && j < kW && k >= 0 && k < kH)) continue;
          float center_x = 50.0f * (i - cen_l), center_y = 50.0f * (j - cen_w), center_z = 50.0f * (k - cen_h);
          bool is_in_laser_fov = false;
          for (int ii = -1; ii <= 1; ii += 2)
            for (int jj = -1; jj <= 1; jj += 2)
              for (int kk = -1; kk <= 1; kk += 2) {
                float cx = center_x + 25.0f * ii, cy = center_y + 25.0f * jj, cz = center_z + 25.0f * kk;
                float d0 = pos.x - cx, d1 = pos.y - cy, d2 = pos.z - cz;
                float squared_side1 = d0 * d0 + d1 * d1 + d2 * d2;
                float e0 = point_on_z_axis.x - cx, e1 = point_on_z_axis.y - cy, e2 = point_on_z_axis.z - cz;
                float squared_side2 = e0 * e0 + e1 * e1 + e2 * e2;
                float check1 = 100.0f + squared_side1 - squared_side2 - 10.0f * std::sqrt(3.0f) * std::sqrt(squared_side1);
                float check2 = 100.0f + squared_side1 - squared_side2 + 10.0f * std::sqrt(3.0f) * std::sqrt(squared_side1);
                if (check1 < 0 && check2 > 0) is_in_laser_fov = true;
              }
          const size_t cube_idx = ToIndex(i, j, k);
          if (is_in_laser_fov) valid.push_back(cube_idx);
          surround.push_back(cube_idx);
        }
  }
  // :1005-1011
  void FromMap(const std::vector<size_t> &valid, Cloud &corner_from_map, Cloud &surf_from_map) const {
    corner_from_map.clear(); surf_from_map.clear();
    for (size_t v : valid) {
      corner_from_map.insert(corner_from_map.end(), corner[v].begin(), corner[v].end());
      surf_from_map.insert(surf_from_map.end(), surf[v].begin(), surf[v].end());
    }
  }
  // :1112-1208 (margin_cube_center = the centre the valid indices were computed with)
  void UpdateMapDatabase(const Cloud &corner_ds, const Cloud &surf_ds, const std::vector<size_t> &valid_idx, const Transform &t,
                         int m_cen_l, int m_cen_w, int m_cen_h) {
    PointXYZI point_sel;
    auto insert = [&](const Cloud &src, std::vector<Cloud> &dst) {
      for (const PointXYZI &p : src) {
        PointAssociateToMap(p, point_sel, t);
        int cube_i = CubeOf(point_sel.x, cen_l), cube_j = CubeOf(point_sel.y, cen_w), cube_k = CubeOf(point_sel.z, cen_h);
        if (cube_i >= 0 && cube_i < kL && cube_j >= 0 && cube_j < kW && cube_k >= 0 && cube_k < kH)
          dst[ToIndex(cube_i, cube_j, cube_k)].push_back(point_sel);
      }
    };
    insert(corner_ds, corner);
    insert(surf_ds, surf);
    for (size_t index : valid_idx) {
      int last_i, last_j, last_k;
      FromIndex(index, last_i, last_j, last_k);
      float center_x = 50.0f * (last_i - m_cen_l), center_y = 50.0f * (last_j - m_cen_w), center_z = 50.0f * (last_k - m_cen_h);
      int cube_i = CubeOf(center_x, cen_l), cube_j = CubeOf(center_y, cen_w), cube_k = CubeOf(center_z, cen_h);
      if (!(cube_i >= 0 && cube_i < kL && cube_j >= 0 && cube_j < kW && cube_k >= 0 && cube_k < kH)) continue;
      index = ToIndex(cube_i, cube_j, cube_k);
      Cloud c2, s2;
      VoxelGridFilter(corner[index], corner_leaf, c2);
      VoxelGridFilter(surf[index], surf_leaf, s2);
      corner[index].swap(c2);
      surf[index].swap(s2);
    }
  }
};

// PointMapping::Process with imu_inited_ == false and num_stack_frames_ == 1 (PointMapping.cc:765-1052): associate the
// odometry increment, bring the last features to the map frame, re-centre / select cubes, pull the map, take the stacks
// back to the sensor frame and down-sample them, optimise against the map, update the map database.
struct PointMappingOracle {
  CubeMap map;
  Transform sum, bef, aft, tobe;   // transform_sum_, transform_bef_mapped_, transform_aft_mapped_, transform_tobe_mapped_
  StageBConfig cfg;
  int last_iters = 0;
  size_t last_corner_from_map = 0, last_surf_from_map = 0;
  static void PointAssociateTobeMapped(const PointXYZI &pi, PointXYZI &po, const Transform &t) {  // :316-323
    Vec3<float> v(pi.x - t.pos.x, pi.y - t.pos.y, pi.z - t.pos.z);
    Vec3<float> o = t.rot.conjugate() * v;
    po.x = o.x; po.y = o.y; po.z = o.z; po.intensity = pi.intensity;
  }
  void Process(const Cloud &corner_last, const Cloud &surf_last, const Transform &transform_sum) {
    sum = transform_sum;
    {  // TransformAssociateToMap :753-756
      Transform incre = bef.inverse() * sum;
      tobe = tobe * incre;
    }
    Cloud corner_stack, surf_stack;
    PointXYZI point_sel;
    for (const PointXYZI &p : corner_last) { PointAssociateToMap(p, point_sel, tobe); corner_stack.push_back(point_sel); }
    for (const PointXYZI &p : surf_last) { PointAssociateToMap(p, point_sel, tobe); surf_stack.push_back(point_sel); }
    PointXYZI point_on_z_axis;
    point_on_z_axis.x = 0.0f; point_on_z_axis.y = 0.0f; point_on_z_axis.z = 10.0f; point_on_z_axis.intensity = 0.f;
    PointAssociateToMap(point_on_z_axis, point_on_z_axis, tobe);
    int ci, cj, ck;
    map.Recentre(tobe.pos, ci, cj, ck);
    std::vector<size_t> valid, surround;
    map.Select(tobe.pos, point_on_z_axis, ci, cj, ck, valid, surround);
    Cloud corner_from_map, surf_from_map;
    map.FromMap(valid, corner_from_map, surf_from_map);
    last_corner_from_map = corner_from_map.size(); last_surf_from_map = surf_from_map.size();
    for (PointXYZI &p : corner_stack) PointAssociateTobeMapped(p, p, tobe);
    for (PointXYZI &p : surf_stack) PointAssociateTobeMapped(p, p, tobe);
    Cloud corner_ds, surf_ds;
    VoxelGridFilter(corner_stack, map.corner_leaf, corner_ds);
    VoxelGridFilter(surf_stack, map.surf_leaf, surf_ds);
    const bool optimised = !(corner_from_map.size() <= 10 || surf_from_map.size() <= 100);
    last_iters = 0;
    OptimizeTransformTobeMapped(corner_from_map, surf_from_map, corner_ds, surf_ds, tobe, cfg, &last_iters, nullptr, 0);
    if (optimised) { bef = sum; aft = tobe; }   // TransformUpdate() sits behind the early return of the optimiser (:327-329, :716)
    map.UpdateMapDatabase(corner_ds, surf_ds, valid, tobe, map.cen_l, map.cen_w, map.cen_h);
  }
};

}  // namespace orc

using namespace orc;
extern "C" {
void *orc_pm_create() { return new PointMappingOracle(); }
void orc_pm_destroy(void *h) { delete (PointMappingOracle *)h; }
// transform_sum as tf7 (qx,qy,qz,qw,px,py,pz); out: tobe tf7, then {iterations, corner_from_map, surf_from_map}
void orc_pm_process(void *h, const float *corner, int nc, const float *surf, int ns, const float *sum7, float *tobe7, int *info3) {
  PointMappingOracle *m = (PointMappingOracle *)h;
  Cloud c((const PointXYZI *)corner, (const PointXYZI *)corner + nc), s((const PointXYZI *)surf, (const PointXYZI *)surf + ns);
  Transform t(Quat<float>(sum7[3], sum7[0], sum7[1], sum7[2]), Vec3<float>(sum7[4], sum7[5], sum7[6]));
  m->Process(c, s, t);
  tobe7[0] = m->tobe.rot.x; tobe7[1] = m->tobe.rot.y; tobe7[2] = m->tobe.rot.z; tobe7[3] = m->tobe.rot.w;
  tobe7[4] = m->tobe.pos.x; tobe7[5] = m->tobe.pos.y; tobe7[6] = m->tobe.pos.z;
  info3[0] = m->last_iters; info3[1] = (int)m->last_corner_from_map; info3[2] = (int)m->last_surf_from_map;
}
int orc_pm_cube_size(void *h, long long index, int which) {
  PointMappingOracle *m = (PointMappingOracle *)h;
  return (int)(which == 0 ? m->map.corner[(size_t)index] : m->map.surf[(size_t)index]).size();
}
void orc_pm_cube_copy(void *h, long long index, int which, float *out) {
  PointMappingOracle *m = (PointMappingOracle *)h;
  const Cloud &c = which == 0 ? m->map.corner[(size_t)index] : m->map.surf[(size_t)index];
  std::memcpy(out, c.data(), sizeof(PointXYZI) * c.size());
}
void orc_pm_centre(void *h, int *out3) {
  PointMappingOracle *m = (PointMappingOracle *)h;
  out3[0] = m->map.cen_l; out3[1] = m->map.cen_w; out3[2] = m->map.cen_h;
}
void *orc_cm_create() { return new CubeMap(); }
void orc_cm_destroy(void *h) { delete (CubeMap *)h; }
// out: centre cube (3), map centre after the shifts (3)
void orc_cm_recentre(void *h, const float *pos3, int *out6) {
  CubeMap *m = (CubeMap *)h;
  int ci, cj, ck;
  m->Recentre(Vec3<float>(pos3[0], pos3[1], pos3[2]), ci, cj, ck);
  out6[0] = ci; out6[1] = cj; out6[2] = ck; out6[3] = m->cen_l; out6[4] = m->cen_w; out6[5] = m->cen_h;
}
// valid / surround sized 125; returns counts through n2
void orc_cm_select(void *h, const float *pos3, const float *zaxis3, const int *centre3, long long *valid, long long *surround, int *n2) {
  CubeMap *m = (CubeMap *)h;
  PointXYZI z; z.x = zaxis3[0]; z.y = zaxis3[1]; z.z = zaxis3[2]; z.intensity = 0;
  std::vector<size_t> v, s;
  m->Select(Vec3<float>(pos3[0], pos3[1], pos3[2]), z, centre3[0], centre3[1], centre3[2], v, s);
  for (size_t i = 0; i < v.size(); ++i) valid[i] = (long long)v[i];
  for (size_t i = 0; i < s.size(); ++i) surround[i] = (long long)s[i];
  n2[0] = (int)v.size(); n2[1] = (int)s.size();
}
int orc_cm_cube_size(void *h, long long index, int which) {
  CubeMap *m = (CubeMap *)h;
  return (int)(which == 0 ? m->corner[(size_t)index] : m->surf[(size_t)index]).size();
}
void orc_cm_cube_copy(void *h, long long index, int which, float *out) {
  CubeMap *m = (CubeMap *)h;
  const Cloud &c = which == 0 ? m->corner[(size_t)index] : m->surf[(size_t)index];
  std::memcpy(out, c.data(), sizeof(PointXYZI) * c.size());
}
void orc_cm_update(void *h, const float *corner, int nc, const float *surf, int ns, const long long *valid, int nv, const float *tf7,
                   const int *margin_centre3) {
  CubeMap *m = (CubeMap *)h;
  Cloud c((const PointXYZI *)corner, (const PointXYZI *)corner + nc), s((const PointXYZI *)surf, (const PointXYZI *)surf + ns);
  std::vector<size_t> v(valid, valid + nv);
  Transform t(Quat<float>(tf7[3], tf7[0], tf7[1], tf7[2]), Vec3<float>(tf7[4], tf7[5], tf7[6]));
  m->UpdateMapDatabase(c, s, v, t, margin_centre3[0], margin_centre3[1], margin_centre3[2]);
}
}
