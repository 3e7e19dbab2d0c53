// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
// Internal C++ declarations of the CPU restatement; the ctypes-facing C API is in o_capi.cc.
#pragma once
#include "o_linalg.h"
#include <vector>
#include <utility>
#include <memory>
#include <cstdint>

namespace orc {

struct PointXYZI { float x, y, z, intensity; };  // pcl::PointXYZI payload (16 useful bytes)
typedef std::vector<PointXYZI> Cloud;
typedef Twist<float> Transform;  // include/point_processor/PointMapping.h: typedef Twist<float> Transform

// ---- stage A --------------------------------------------------------------------------------
struct StageAConfig {  // PointProcessorConfig, include/point_processor/PointProcessor.h:104-120 + ctor args
  float lower_bound = -15.f, upper_bound = 15.f;
  int num_rings = 16;
  double scan_period = 0.1;
  int num_scan_subregions = 8;
  int num_curvature_regions = 5;
  float surf_curv_th = 0.1f;
  int max_corner_sharp = 2;
  int max_corner_less_sharp = 20;
  int max_surf_flat = 4;
  float less_flat_filter_size = 0.2f;
};

struct StageA {
  StageAConfig cfg;
  std::vector<Cloud> laser_scans, intensity_scans;
  std::vector<std::vector<int>> orig_index;   // input index of every ring point (test aid)
  std::vector<std::pair<size_t, size_t>> scan_ranges;
  Cloud cloud_in_rings;
  float start_ori = 0.f;
  Cloud corner_sharp, corner_less_sharp, surf_flat, surf_less_flat;
  // index sets (global ring-ordered indices = scan_ranges[r].first + in-ring index)
  std::vector<int> idx_sharp, idx_less_sharp, idx_flat, less_flat_prevoxel;
  std::vector<unsigned char> final_mask;
  std::vector<signed char> label_all;
  // scratch
  std::vector<int> mask;
  std::vector<std::pair<float, size_t>> curvature_idx_pairs;
  std::vector<int> subregion_labels;

  size_t num_ring_points() const { size_t n = 0; for (auto &c : laser_scans) n += c.size(); return n; }
  void PointToRing(const PointXYZI *points, size_t n);
  void PointToRingWithRingField(const PointXYZI *points, const unsigned short *rings, size_t n);  // PointProcessor.cc:428-536
  void PrepareRing(const Cloud &scan);
  void PrepareSubregion(const Cloud &scan, size_t idx_start, size_t idx_end);
  void MaskPickedInRing(const Cloud &scan, size_t in_scan_idx);
  void ExtractFeaturePoints();
};

// ---- cloud utilities (PCL restatements) -----------------------------------------------------
void VoxelGridFilter(const Cloud &in, float leaf, Cloud &out);
void TransformCloudAffine(const Cloud &in, const Mat3<float> &R, const Vec3<float> &t, Cloud &out);

struct KdTree {  // exact k-NN, squared L2 on xyz, FLANN KDTreeSingleIndex(leaf 15)-like
  struct Node { int left, right, lo, hi, dim; float split_lo, split_hi; };
  const Cloud *cloud = nullptr;
  std::vector<int> idx;
  std::vector<Node> nodes;
  float bb_min[3], bb_max[3];
  void Build(const Cloud &c);
  // results sorted ascending by (d2, index)
  void Knn(const PointXYZI &q, int k, int *out_idx, float *out_d2) const;
 private:
  int BuildRec(int lo, int hi, float *bmin, float *bmax);
  void Search(int node, const float *q, float mindist, float *dists, int k, int *bi, float *bd, int &cnt) const;
};

// ---- stage B ----------------------------------------------------------------------------------
struct PointPlaneFeature {  // include/feature_manager/FeatureManager.h:84-109
  double score;
  double point[3];
  double coeffs[4];
  int src_index;  // index of the originating surf point in its frame (test aid)
};

struct StageBConfig {  // lidar subset of EstimatorConfig, include/imu_processor/Estimator.h:77-108
  float min_match_sq_dis = 1.0f;
  float min_plane_dis = 0.2f;
  float surf_filter_size = 0.4f;
  int keep_features = 0;
  int num_max_iterations = 10;     // PointMapping.h (num_max_iterations_)
  double delta_r_abort = 0.05;     // PointMapping.cc:75 (double delta_r_abort_)
  double delta_t_abort = 0.05;     // PointMapping.cc:76
};

void PointAssociateToMap(const PointXYZI &pi, PointXYZI &po, const Transform &t);
void CalculateFeatures(const KdTree &kd, const Cloud &map, const Cloud &surf_stack, const Transform &local_transform,
                       const StageBConfig &cfg, std::vector<PointPlaneFeature> &features);
void CalculateLineFeatures(const KdTree &kd, const Cloud &map, const Cloud &corner_stack, const Transform &local_transform,
                           const StageBConfig &cfg, std::vector<PointPlaneFeature> &features);
void OptimizeTransformTobeMapped(const Cloud &corner_map, const Cloud &surf_map, const Cloud &corner_stack, const Cloud &surf_stack,
                                 Transform &tobe, const StageBConfig &cfg, int *iters_done, std::vector<PointPlaneFeature> *features_out,
                                 int variant = 0);
void CalculateLaserOdom(const KdTree &kd, const Cloud &map, const Cloud &surf_stack, Transform &local_transform,
                        const StageBConfig &cfg, std::vector<PointPlaneFeature> &features, int *iters_done);

}  // namespace orc
