// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
//
// CPU restatement of stage A of hyye/lio-mapping:
//   PointProcessor::PointToRing        src/point_processor/PointProcessor.cc:185-426
//   PointProcessor::PrepareRing        :542-585
//   PointProcessor::PrepareSubregion   :587-622
//   PointProcessor::MaskPickedInRing   :624-645
//   PointProcessor::ExtractFeaturePoints :647-783
//   ElevationToRing                    include/point_processor/PointProcessor.h:153-156
//   math helpers                       include/utils/math_utils.h:38-110
//   pcl::VoxelGrid<PointXYZI> (PCL 1.8, un-vendored): restated in o_cloud.cc
// Build flags mirror the reference (no -march, no FMA contraction): every float expression is
// IEEE binary32 in source order; mixed float/double expressions promote as in C++.
#include "o_api.h"
#include <cmath>
#include <cstdint>
#include <utility>
#include <algorithm>

namespace orc {

static inline bool finite3(const PointXYZI &p) {
  return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z);
}

// mathutils::RadToDeg<float> (math_utils.h:38-41): double arithmetic, rounded to float on return.
static inline float RadToDegF(float rad) { return (float)(rad * 180.0 / M_PI); }

// PointProcessor::ElevationToRing (PointProcessor.h:153-156)
static inline int ElevationToRing(float rad, float lower_bound, float factor) {
  return int((RadToDegF(rad) - lower_bound) * factor + 0.5);
}

// math_utils.h:84-101
static inline float CalcSquaredDiff(const PointXYZI &a, const PointXYZI &b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return dx * dx + dy * dy + dz * dz;
}
static inline float CalcSquaredDiffW(const PointXYZI &a, const PointXYZI &b, float wb) {
  float dx = a.x - b.x * wb, dy = a.y - b.y * wb, dz = a.z - b.z * wb;
  return dx * dx + dy * dy + dz * dz;
}
// math_utils.h:102-110
static inline float CalcPointDistance(const PointXYZI &p) { return std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z); }
static inline float CalcSquaredPointDistance(const PointXYZI &p) { return p.x * p.x + p.y * p.y + p.z * p.z; }

void StageA::PointToRing(const PointXYZI *points, size_t cloud_size) {
  // PointProcessor.cc:207-426 (non-DEBUG_ORIGIN branch), wrapper :185-205
  const int R = cfg.num_rings;
  const float factor = (cfg.num_rings - 1) / (cfg.upper_bound - cfg.lower_bound);  // PointProcessor.cc:80
  laser_scans.assign(R, Cloud());
  intensity_scans.assign(R, Cloud());
  orig_index.assign(R, std::vector<int>());
  bool start_flag = false;
  start_ori = 0.f;
  for (size_t i = 0; i < cloud_size; ++i) {
    PointXYZI p = points[i];
    PointXYZI p_with_intensity = points[i];
    if (!finite3(p)) continue;  // :240-244
    float dis = std::sqrt(p.x * p.x + p.y * p.y);
    float ele_rad = std::atan2(p.z, dis);                      // atan2f (using namespace std)
    float azi_rad = (float)(2 * M_PI - std::atan2(p.y, p.x));  // double - float -> float
    if (azi_rad >= 2 * M_PI) azi_rad = (float)(azi_rad - 2 * M_PI);
    int scan_id = ElevationToRing(ele_rad, cfg.lower_bound, factor);
    if (scan_id >= R || scan_id < 0) continue;
    if (!start_flag) { start_ori = azi_rad; start_flag = true; }
    p.intensity = azi_rad;
    laser_scans[scan_id].push_back(p);
    intensity_scans[scan_id].push_back(p_with_intensity);
    orig_index[scan_id].push_back((int)i);
  }
  // second pass :393-423
  for (int ring = 0; ring < R; ++ring) {
    Cloud &pts = laser_scans[ring];
    Cloud &pti = intensity_scans[ring];
    for (size_t i = 0; i < pts.size(); ++i) {
      float azi_rad_rel = pts[i].intensity - start_ori;
      if (azi_rad_rel < 0) azi_rad_rel = (float)(azi_rad_rel + 2 * M_PI);
      float rel_time = (float)(cfg.scan_period * azi_rad_rel / (2 * M_PI));
      pts[i].intensity = ring + rel_time;
      pti[i].intensity = int(pti[i].intensity) + rel_time;
    }
  }
  // wrapper :191-201
  cloud_in_rings.clear();
  scan_ranges.clear();
  size_t cs = 0;
  for (int i = 0; i < R; ++i) {
    cloud_in_rings.insert(cloud_in_rings.end(), intensity_scans[i].begin(), intensity_scans[i].end());
    std::pair<size_t, size_t> range(cs, 0);
    cs += laser_scans[i].size();
    range.second = (cs > 0 ? cs - 1 : 0);
    scan_ranges.push_back(range);
  }
}

// PointToRing for sensors that deliver the ring index (lio::PointXYZIR, include/point_processor/point_types.h:37-52):
// PointProcessor.cc:428-536.  Ring id from the field; the azimuth gets +2 pi when it lies before start_ori_ (the
// half_passed branch can never be entered: `i > 3 * cloud_size / 2` is false for every i < cloud_size); end_ori_ is the
// running maximum (from 0) and rel_time = scan_period * (azi - start_ori_) / (end_ori_ - start_ori_).
void StageA::PointToRingWithRingField(const PointXYZI *points, const unsigned short *rings, size_t cloud_size) {
  const int R = cfg.num_rings;
  laser_scans.assign(R, Cloud());
  intensity_scans.assign(R, Cloud());
  orig_index.assign(R, std::vector<int>());
  bool start_flag = false;
  start_ori = 0.f;
  float end_ori = 0.f;
  bool half_passed = false;
  for (size_t i = 0; i < cloud_size; ++i) {
    PointXYZI p = points[i];
    PointXYZI p_with_intensity = p;
    if (!finite3(p)) continue;  // :456-460
    float azi_rad = (float)(2 * M_PI - std::atan2(p.y, p.x));
    if (azi_rad >= 2 * M_PI) azi_rad = (float)(azi_rad - 2 * M_PI);
    int scan_id = rings[i];
    if (scan_id >= R || scan_id < 0) continue;
    if (!start_flag) { start_ori = azi_rad; start_flag = true; }
    float azi_rad_rel = azi_rad - start_ori;
    if (!half_passed) {
      if (azi_rad_rel < 0) azi_rad = (float)(azi_rad + 2 * M_PI);
      if (azi_rad_rel > M_PI && i > 3 * cloud_size / 2) half_passed = true;
    } else {
      if (azi_rad_rel < M_PI / 2) azi_rad = (float)(azi_rad + 2 * M_PI);
    }
    if (end_ori < azi_rad) end_ori = azi_rad;
    p.intensity = azi_rad;
    laser_scans[scan_id].push_back(p);
    intensity_scans[scan_id].push_back(p_with_intensity);
    orig_index[scan_id].push_back((int)i);
  }
  const float range_ori = end_ori - start_ori;
  for (int ring = 0; ring < R; ++ring) {
    Cloud &pts = laser_scans[ring];
    Cloud &pti = intensity_scans[ring];
    for (size_t i = 0; i < pts.size(); ++i) {
      float azi_rad_rel = pts[i].intensity - start_ori;
      float rel_time = (float)(cfg.scan_period * azi_rad_rel / range_ori);
      pts[i].intensity = ring + rel_time;
      pti[i].intensity = int(pti[i].intensity) + rel_time;
    }
  }
  cloud_in_rings.clear();
  scan_ranges.clear();
  size_t cs = 0;
  for (int i = 0; i < R; ++i) {   // wrapper :185-205 (shared by both variants)
    cloud_in_rings.insert(cloud_in_rings.end(), intensity_scans[i].begin(), intensity_scans[i].end());
    std::pair<size_t, size_t> range(cs, 0);
    cs += laser_scans[i].size();
    range.second = (cs > 0 ? cs - 1 : 0);
    scan_ranges.push_back(range);
  }
}

void StageA::PrepareRing(const Cloud &scan) {
  // PointProcessor.cc:542-585.  The reference writes scan_ring_mask_[scan_size] (one past the
  // end) for i = scan_size-d-1 in the "else" branch; we allocate one spare slot.
  const size_t scan_size = scan.size();
  const size_t d = (size_t)cfg.num_curvature_regions;
  mask.assign(scan_size + 1, 0);
  for (size_t i = 0 + d; i < scan_size - d; ++i) {
    const PointXYZI &p_prev = scan[i - 1];
    const PointXYZI &p_curr = scan[i];
    const PointXYZI &p_next = scan[i + 1];
    float diff_next2 = CalcSquaredDiff(p_curr, p_next);
    if (diff_next2 > 0.1) {
      float depth = CalcPointDistance(p_curr);
      float depth_next = CalcPointDistance(p_next);
      if (depth > depth_next) {
        float weighted_diff = std::sqrt(CalcSquaredDiffW(p_next, p_curr, depth_next / depth)) / depth_next;
        if (weighted_diff < 0.1) {
          for (size_t k = 0; k < d + 1; ++k) mask[i - d + k] = 1;
          continue;
        }
      } else {
        float weighted_diff = std::sqrt(CalcSquaredDiffW(p_curr, p_next, depth / depth_next)) / depth;
        if (weighted_diff < 0.1) {
          for (size_t k = 0; k < d + 1; ++k) mask[i + 1 + k] = 1;
          continue;
        }
      }
    }
    float diff_prev2 = CalcSquaredDiff(p_curr, p_prev);
    float dis2 = CalcSquaredPointDistance(p_curr);
    if (diff_next2 > 0.0002 * dis2 && diff_prev2 > 0.0002 * dis2) mask[i] = 1;
  }
}

void StageA::PrepareSubregion(const Cloud &scan, size_t idx_start, size_t idx_end) {
  // PointProcessor.cc:587-622
  size_t region_size = idx_end - idx_start + 1;
  curvature_idx_pairs.resize(region_size);
  subregion_labels.assign(region_size, 0 /*SURFACE_LESS_FLAT*/);
  int num_point_neighbors = 2 * cfg.num_curvature_regions;
  for (size_t i = idx_start, in_region_idx = 0; i <= idx_end; ++i, ++in_region_idx) {
    float diff_x = -num_point_neighbors * scan[i].x;
    float diff_y = -num_point_neighbors * scan[i].y;
    float diff_z = -num_point_neighbors * scan[i].z;
    for (int j = 1; j <= cfg.num_curvature_regions; j++) {
      diff_x += scan[i + j].x + scan[i - j].x;
      diff_y += scan[i + j].y + scan[i - j].y;
      diff_z += scan[i + j].z + scan[i - j].z;
    }
    float curvature = diff_x * diff_x + diff_y * diff_y + diff_z * diff_z;
    curvature_idx_pairs[in_region_idx] = std::pair<float, size_t>(curvature, i);
  }
  std::sort(curvature_idx_pairs.begin(), curvature_idx_pairs.end());
}

void StageA::MaskPickedInRing(const Cloud &scan, size_t in_scan_idx) {
  // PointProcessor.cc:624-645
  mask[in_scan_idx] = 1;
  for (int i = 1; i <= cfg.num_curvature_regions; ++i) {
    if (CalcSquaredDiff(scan[in_scan_idx + i], scan[in_scan_idx + i - 1]) > 0.05) break;
    mask[in_scan_idx + i] = 1;
  }
  for (int i = 1; i <= cfg.num_curvature_regions; ++i) {
    if (CalcSquaredDiff(scan[in_scan_idx - i], scan[in_scan_idx - i + 1]) > 0.05) break;
    mask[in_scan_idx - i] = 1;
  }
}

void StageA::ExtractFeaturePoints() {
  // PointProcessor.cc:647-783
  corner_sharp.clear(); corner_less_sharp.clear(); surf_flat.clear(); surf_less_flat.clear();
  idx_sharp.clear(); idx_less_sharp.clear(); idx_flat.clear();
  less_flat_prevoxel.clear();
  final_mask.assign(num_ring_points(), 0);
  label_all.assign(num_ring_points(), 0);
  const size_t d = (size_t)cfg.num_curvature_regions;
  const int S = cfg.num_scan_subregions;
  for (size_t i = 0; i < (size_t)cfg.num_rings; ++i) {
    Cloud less_flat_ring;
    size_t start_idx = scan_ranges[i].first;
    size_t end_idx = scan_ranges[i].second;
    if (end_idx <= start_idx + 2 * d) continue;  // :660
    const Cloud &scan_ring = laser_scans[i];
    size_t scan_size = scan_ring.size();
    PrepareRing(scan_ring);
    for (int j = 0; j < S; ++j) {
      size_t sp = ((0 + d) * (S - j) + (scan_size - d) * j) / S;
      size_t ep = ((0 + d) * (S - 1 - j) + (scan_size - d) * (j + 1)) / S - 1;
      if (ep <= sp) continue;
      size_t region_size = ep - sp + 1;
      PrepareSubregion(scan_ring, sp, ep);
      int num_largest_picked = 0;
      for (size_t k = region_size; k > 0 && num_largest_picked < cfg.max_corner_less_sharp;) {
        const std::pair<float, size_t> &ci = curvature_idx_pairs[--k];
        float curvature = ci.first;
        size_t idx = ci.second;
        size_t in_scan_idx = idx;
        size_t in_region_idx = idx - sp;
        if (mask[in_scan_idx] == 0 && curvature > cfg.surf_curv_th) {
          ++num_largest_picked;
          if (num_largest_picked <= cfg.max_corner_sharp) {
            subregion_labels[in_region_idx] = 2;  // CORNER_SHARP
            corner_sharp.push_back(scan_ring[in_scan_idx]);
            idx_sharp.push_back((int)(start_idx + in_scan_idx));
          } else {
            subregion_labels[in_region_idx] = 1;  // CORNER_LESS_SHARP
          }
          corner_less_sharp.push_back(scan_ring[in_scan_idx]);
          idx_less_sharp.push_back((int)(start_idx + in_scan_idx));
          MaskPickedInRing(scan_ring, in_scan_idx);
        }
      }
      int num_smallest_picked = 0;
      for (int k = 0; k < (int)region_size && num_smallest_picked < cfg.max_surf_flat; ++k) {
        const std::pair<float, size_t> &ci = curvature_idx_pairs[k];
        float curvature = ci.first;
        size_t idx = ci.second;
        size_t in_scan_idx = idx;
        size_t in_region_idx = idx - sp;
        if (mask[in_scan_idx] == 0 && curvature < cfg.surf_curv_th) {
          ++num_smallest_picked;
          subregion_labels[in_region_idx] = -1;  // SURFACE_FLAT
          surf_flat.push_back(scan_ring[in_scan_idx]);
          idx_flat.push_back((int)(start_idx + in_scan_idx));
          MaskPickedInRing(scan_ring, in_scan_idx);
        }
      }
      for (int k = 0; k < (int)region_size; ++k) {
        label_all[start_idx + sp + k] = (signed char)subregion_labels[k];
        if (subregion_labels[k] <= 0) {
          less_flat_ring.push_back(scan_ring[sp + k]);
          less_flat_prevoxel.push_back((int)(start_idx + sp + k));
        }
      }
    }
    for (size_t k = 0; k < scan_size; ++k) final_mask[start_idx + k] = (unsigned char)mask[k];
    if (less_flat_ring.empty()) continue;
    Cloud ds;
    VoxelGridFilter(less_flat_ring, cfg.less_flat_filter_size, ds);
    surf_less_flat.insert(surf_less_flat.end(), ds.begin(), ds.end());
  }
  // :753-778
  for (size_t i = 0; i < surf_less_flat.size(); ++i) {
    PointXYZI &p = surf_less_flat[i];
    float azi_rad = (float)(2 * M_PI - std::atan2(p.y, p.x));
    if (azi_rad >= 2 * M_PI) azi_rad = (float)(azi_rad - 2 * M_PI);
    float azi_rad_rel = azi_rad - start_ori;
    if (azi_rad_rel < 0) azi_rad_rel = (float)(azi_rad_rel + 2 * M_PI);
    float rel_time = (float)(cfg.scan_period * azi_rad_rel / (2 * M_PI));
    p.intensity = int(p.intensity) + rel_time;
  }
}

}  // namespace orc
