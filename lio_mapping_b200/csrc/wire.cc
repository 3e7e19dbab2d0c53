// lio_mapping_b200 — /compact_data wire format (host only; SURVEY.md §8b "Topic wire format", §8f-4).
// Encoder: PointOdometry.cc:732-762 (reference), decoder: PointMapping::CompactDataHandler, PointMapping.cc:171-238.
//   point 0 = transform_sum_.pos (intensity 0), point 1 = (rot.x, rot.y, rot.z | intensity = rot.w),
//   point 2 = (corner_size, surf_size, full_size) AS FLOATS (its intensity keeps rot.w: the encoder reuses one PointT
//   and only overwrites x, y, z), then corner || surf || full.  Sizes are exact only below 2^24.
// The payload of the PointCloud2 is pcl::PointXYZI memory: point_step 32, x/y/z at 0/4/8, data[3] = 1.0f at 12,
// intensity at 16, bytes 20..31 padding.  These functions work on packed (x, y, z, intensity) float4 arrays and convert
// to / from the 32-byte layout.
#include <cmath>
#include <cstdint>
#include <cstring>
#include "../../include/lio_b200.h"

void lio_set_last_error(const char *file, int line, const char *msg);  // capi_common.cu

extern "C" int lio_compact_encode(const float tf7[7], const float *corner, int nc, const float *surf, int ns, const float *full,
                                  int nf, float *out_xyzi, int cap_points, int *n_points) {
  if (!tf7 || !out_xyzi || !n_points || nc < 0 || ns < 0 || nf < 0 || (nc > 0 && !corner) || (ns > 0 && !surf) || (nf > 0 && !full))
    return LIO_ERR_INVALID;
  const long long total = 3LL + nc + ns + nf;
  if (nc >= (1 << 24) || ns >= (1 << 24) || nf >= (1 << 24)) {
    lio_set_last_error(__FILE__, __LINE__, "cloud size not representable as a float (>= 2^24)");
    return LIO_ERR_CAPACITY;
  }
  if (total > cap_points) return LIO_ERR_CAPACITY;
  float *o = out_xyzi;
  o[0] = tf7[4]; o[1] = tf7[5]; o[2] = tf7[6]; o[3] = 0.f;
  o[4] = tf7[0]; o[5] = tf7[1]; o[6] = tf7[2]; o[7] = tf7[3];
  o[8] = (float)nc; o[9] = (float)ns; o[10] = (float)nf; o[11] = tf7[3];
  o += 12;
  if (nc) std::memcpy(o, corner, sizeof(float) * 4 * nc);
  o += 4 * (size_t)nc;
  if (ns) std::memcpy(o, surf, sizeof(float) * 4 * ns);
  o += 4 * (size_t)ns;
  if (nf) std::memcpy(o, full, sizeof(float) * 4 * nf);
  *n_points = (int)total;
  return LIO_OK;
}

extern "C" int lio_compact_sizes(const float *xyzi, int n_points, int sizes[3]) {
  if (!xyzi || !sizes) return LIO_ERR_INVALID;
  if (n_points < 4) {  // "compact_points not enough"
    lio_set_last_error(__FILE__, __LINE__, "compact_points not enough");
    return LIO_ERR_INVALID;
  }
  const int corner_size = (int)xyzi[8], surf_size = (int)xyzi[9], full_size = (int)xyzi[10];
  if (corner_size < 0 || surf_size < 0 || full_size < 0 || 3LL + corner_size + surf_size + full_size != (long long)n_points) {
    lio_set_last_error(__FILE__, __LINE__, "compact data error: 3 + corner + surf + full != size");
    return LIO_ERR_INVALID;
  }
  sizes[0] = corner_size; sizes[1] = surf_size; sizes[2] = full_size;
  return LIO_OK;
}

extern "C" int lio_compact_decode(const float *xyzi, int n_points, float tf7[7], float *corner, float *surf, float *full) {
  int sz[3];
  const int rc = lio_compact_sizes(xyzi, n_points, sz);
  if (rc != LIO_OK) return rc;
  if (!tf7 || (sz[0] > 0 && !corner) || (sz[1] > 0 && !surf) || (sz[2] > 0 && !full)) return LIO_ERR_INVALID;
  tf7[4] = xyzi[0]; tf7[5] = xyzi[1]; tf7[6] = xyzi[2];
  tf7[0] = xyzi[4]; tf7[1] = xyzi[5]; tf7[2] = xyzi[6]; tf7[3] = xyzi[7];
  const float *p = xyzi + 12;
  if (sz[0]) std::memcpy(corner, p, sizeof(float) * 4 * sz[0]);
  p += 4 * (size_t)sz[0];
  if (sz[1]) std::memcpy(surf, p, sizeof(float) * 4 * sz[1]);
  p += 4 * (size_t)sz[1];
  if (sz[2]) std::memcpy(full, p, sizeof(float) * 4 * sz[2]);
  return LIO_OK;
}

extern "C" int lio_xyzi_to_pcl32(const float *xyzi, int n, uint8_t *out32) {
  if (n < 0 || (n > 0 && (!xyzi || !out32))) return LIO_ERR_INVALID;
  for (int i = 0; i < n; ++i) {
    float rec[8] = {xyzi[4 * i], xyzi[4 * i + 1], xyzi[4 * i + 2], 1.0f, xyzi[4 * i + 3], 0.f, 0.f, 0.f};
    std::memcpy(out32 + 32 * (size_t)i, rec, 32);
  }
  return LIO_OK;
}

extern "C" int lio_pcl32_to_xyzi(const uint8_t *in32, int n, float *xyzi) {
  if (n < 0 || (n > 0 && (!xyzi || !in32))) return LIO_ERR_INVALID;
  for (int i = 0; i < n; ++i) {
    float rec[8];
    std::memcpy(rec, in32 + 32 * (size_t)i, 32);
    xyzi[4 * i] = rec[0]; xyzi[4 * i + 1] = rec[1]; xyzi[4 * i + 2] = rec[2]; xyzi[4 * i + 3] = rec[4];
  }
  return LIO_OK;
}
