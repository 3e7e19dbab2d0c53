/* lio_b200.h — C-ABI of liblio_b200.so: the B200-native (sm_100a) replacement of the compute hot
 * path of hyye/lio-mapping.  Plain pointers and sizes only; no C++/torch types.
 *
 * Each entry point names the reference interface (file:line under the reference tree) it stands
 * in for.  The reference has no FFI layer of its own: the seams are its C++ member functions and
 * the ceres::CostFunction contract; INTEGRATION.md shows the shim a maintainer
 * adds at each seam.
 *
 * Conventions
 *   - point clouds are arrays of float4 {x, y, z, intensity}  (pcl::PointXYZI payload);
 *   - poses are double[7] {px,py,pz,qx,qy,qz,qw} (para_pose_ layout, Estimator.cc:2445-2452),
 *     float transforms are float[7] {qx,qy,qz,qw,px,py,pz} (Twist<float>);
 *   - every function returns LIO_OK (0) or a negative lio_status; lio_last_error() gives text;
 *   - "_host" variants take host buffers and copy in/out on the context's stream (synchronous
 *     on return); "_dev" variants take device pointers and are stream-ordered (asynchronous);
 *   - one thread drives a given handle at a time (same rule as the reference objects).
 */
#ifndef LIO_B200_H_
#define LIO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum lio_status {
  LIO_OK = 0,
  LIO_ERR_CUDA = -1,        /* a CUDA runtime call failed (text in lio_last_error) */
  LIO_ERR_INVALID = -2,     /* bad argument */
  LIO_ERR_CAPACITY = -3,    /* input exceeds the capacity given at create time */
  LIO_ERR_NO_DEVICE = -4,   /* no CUDA device: there is NO CPU fallback in this library */
  LIO_ERR_NUMERIC = -5      /* solver breakdown (non-finite / not positive definite) */
} lio_status;

const char *lio_last_error(void);
int lio_version(void);
/* Number of usable CUDA devices (0 => every compute entry point returns LIO_ERR_NO_DEVICE). */
int lio_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Stage A — lio::PointProcessor  (include/point_processor/PointProcessor.h:122-229,
 *           src/point_processor/PointProcessor.cc:185-783)
 * ---------------------------------------------------------------------------------------- */
typedef struct lio_pp_config {      /* PointProcessorConfig, PointProcessor.h:104-120 + ctor :76-80 */
  float lower_bound;                /* deg */
  float upper_bound;                /* deg */
  int num_rings;
  double scan_period;
  int num_scan_subregions;
  int num_curvature_regions;
  float surf_curv_th;
  int max_corner_sharp;
  int max_corner_less_sharp;
  int max_surf_flat;
  float less_flat_filter_size;
} lio_pp_config;

typedef struct lio_pp lio_pp;

/* Which output of the processor (names follow the reference members, PointProcessor.h:170-194). */
typedef enum lio_pp_cloud {
  LIO_PP_LASER_SCANS = 0,           /* laser_scans concatenated: intensity = ring + rel_time      */
  LIO_PP_CLOUD_IN_RINGS = 1,        /* cloud_in_rings_: intensity = int(I) + rel_time             */
  LIO_PP_CORNER_SHARP = 2,          /* corner_points_sharp_                                       */
  LIO_PP_CORNER_LESS_SHARP = 3,     /* corner_points_less_sharp_                                  */
  LIO_PP_SURF_FLAT = 4,             /* surface_points_flat_                                       */
  LIO_PP_SURF_LESS_FLAT = 5,        /* surface_points_less_flat_ (per-ring VoxelGrid(0.2) output) */
  LIO_PP_NUM_CLOUDS = 6
} lio_pp_cloud;

typedef enum lio_pp_index {         /* index sets into the ring-ordered cloud (parity / debug)   */
  LIO_PP_IDX_SHARP = 0,
  LIO_PP_IDX_LESS_SHARP = 1,
  LIO_PP_IDX_FLAT = 2,
  LIO_PP_IDX_ORIG = 3               /* input index of every ring-ordered point                   */
} lio_pp_index;

void lio_pp_default_config(lio_pp_config *cfg);                 /* PointProcessor.h:104-120 defaults */
/* PointProcessor::PointProcessor(lower, upper, rings) + SetupConfig  (PointProcessor.cc:74-95) */
int lio_pp_create(const lio_pp_config *cfg, int max_points, int device, void *cuda_stream, lio_pp **out);
int lio_pp_destroy(lio_pp *pp);
/* SetInputCloud + PointToRing + ExtractFeaturePoints (PointProcessor.cc:96-100,140-172,185-783);
 * xyzi: n x float4 on the HOST (pinned or pageable).  Synchronous. */
int lio_pp_process_host(lio_pp *pp, const float *xyzi, int n);
/* Same, input already on the device; asynchronous on the stream. */
int lio_pp_process_dev(lio_pp *pp, const float *xyzi_dev, int n);
/* Sizes of the six output clouds (synchronises the stream). */
int lio_pp_cloud_sizes(lio_pp *pp, int sizes[LIO_PP_NUM_CLOUDS]);
/* Copy one output cloud to the host (cap in points).  Returns the point count via *n. */
int lio_pp_download_cloud(lio_pp *pp, int which, float *out, int cap, int *n);
/* Device pointer of one output cloud (float4 array, valid until the next process call). */
int lio_pp_cloud_dev(lio_pp *pp, int which, const float **ptr);
int lio_pp_download_index(lio_pp *pp, int which, int32_t *out, int cap, int *n);
/* scan_ranges (PointProcessor.h:172): R pairs (start,end) exactly as the reference stores them. */
int lio_pp_download_scan_ranges(lio_pp *pp, int32_t *out_2R);
/* scan_ring_mask_ after the last subregion and PointLabel of every ring-ordered point. */
int lio_pp_download_mask_labels(lio_pp *pp, uint8_t *mask, int8_t *labels, int cap);
int lio_pp_start_ori(lio_pp *pp, float *start_ori);
/* Number of kernels launched by the last process call (bench bookkeeping). */
int lio_pp_last_launches(lio_pp *pp);

/* ------------------------------------------------------------------------------------------
 * Stage B primitives on explicit host arrays (parity entries; the estimator context below keeps
 * the same kernels device-resident).
 * ---------------------------------------------------------------------------------------- */
/* pcl::VoxelGrid<PointXYZI>::filter with setLeafSize(leaf,leaf,leaf)
 * (call sites Estimator.cc:679-687, :1518-1519; PointProcessor.cc:737-751).  out sized cap points. */
int lio_voxel_grid_host(const float *cloud, int n, float leaf, float *out, int cap, int *n_out, int device);
/* Estimator::CalculateFeatures (Estimator.cc:970-1097) with the kd-tree replaced by the voxel-hash
 * k-NN: map (K x float4) = local_surf_points_filtered_ptr_, surf (M x float4) = surf_stack_[idx],
 * tf7 = local_transform {qx,qy,qz,qw,px,py,pz}.  Outputs (sized M): pts4 = {point.xyz, score},
 * coef4 = coeffs, src = index of the originating surf point. */
int lio_calculate_features_host(const float *map, int K, const float *surf, int M, const float *tf7,
                                float min_match_sq_dis, float min_plane_dis, float *pts4, float *coef4, int32_t *src,
                                int *n_out, int device);

#ifdef __cplusplus
}
#endif
#endif /* LIO_B200_H_ */
