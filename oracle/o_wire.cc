// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
//
// CPU restatement of the /compact_data wire format of hyye/lio-mapping:
//   encoder  PointOdometry::Process            src/point_processor/PointOdometry.cc:732-762
//   decoder  PointMapping::CompactDataHandler  src/point_processor/PointMapping.cc:171-238
// written with the reference's own control flow (one reused PointT, push_back / operator+=, the decoder's index loop).
#include "o_api.h"

namespace orc {

void CompactEncode(const Transform &transform_sum, const Cloud &last_corner_cloud, const Cloud &last_surf_cloud, const Cloud &full_cloud,
                   Cloud &compact_data) {
  compact_data.clear();
  PointXYZI compact_point;  // pcl::PointXYZI(): x = y = z = 0, intensity = 0
  compact_point.x = compact_point.y = compact_point.z = 0.f; compact_point.intensity = 0.f;
  {
    compact_point.x = transform_sum.pos.x;
    compact_point.y = transform_sum.pos.y;
    compact_point.z = transform_sum.pos.z;
    compact_data.push_back(compact_point);
    compact_point.x = transform_sum.rot.x;
    compact_point.y = transform_sum.rot.y;
    compact_point.z = transform_sum.rot.z;
    compact_point.intensity = transform_sum.rot.w;
    compact_data.push_back(compact_point);
  }
  {
    compact_point.x = last_corner_cloud.size();
    compact_point.y = last_surf_cloud.size();
    compact_point.z = full_cloud.size();
    compact_data.push_back(compact_point);
    compact_data.insert(compact_data.end(), last_corner_cloud.begin(), last_corner_cloud.end());
    compact_data.insert(compact_data.end(), last_surf_cloud.begin(), last_surf_cloud.end());
    compact_data.insert(compact_data.end(), full_cloud.begin(), full_cloud.end());
  }
}

// returns false where the reference logs an error and returns
bool CompactDecode(const Cloud &compact_points, Transform &transform_sum, Cloud &corner, Cloud &surf, Cloud &full) {
  size_t compact_point_size = compact_points.size();
  if (compact_point_size < 4) return false;
  PointXYZI compact_point = compact_points[2];
  int corner_size = int(compact_point.x);
  int surf_size = int(compact_point.y);
  int full_size = int(compact_point.z);
  if ((size_t)(3 + corner_size + surf_size + full_size) != compact_point_size) return false;
  compact_point = compact_points[0];
  transform_sum.pos.x = compact_point.x; transform_sum.pos.y = compact_point.y; transform_sum.pos.z = compact_point.z;
  compact_point = compact_points[1];
  transform_sum.rot.x = compact_point.x; transform_sum.rot.y = compact_point.y; transform_sum.rot.z = compact_point.z;
  transform_sum.rot.w = compact_point.intensity;
  corner.clear(); surf.clear(); full.clear();
  for (size_t i = 3; i < compact_point_size; ++i) {
    const PointXYZI &p = compact_points[i];
    if (i < (size_t)(3 + corner_size)) corner.push_back(p);
    else if (i >= (size_t)(3 + corner_size) && i < (size_t)(3 + corner_size + surf_size)) surf.push_back(p);
    else full.push_back(p);
  }
  return true;
}

}  // namespace orc

using namespace orc;
extern "C" {
int orc_compact_encode(const float *tf7, const float *corner, int nc, const float *surf, int ns, const float *full, int nf, float *out) {
  Transform t(Quat<float>(tf7[3], tf7[0], tf7[1], tf7[2]), Vec3<float>(tf7[4], tf7[5], tf7[6]));
  Cloud c((const PointXYZI *)corner, (const PointXYZI *)corner + nc), s((const PointXYZI *)surf, (const PointXYZI *)surf + ns),
      f((const PointXYZI *)full, (const PointXYZI *)full + nf), o;
  CompactEncode(t, c, s, f, o);
  std::memcpy(out, o.data(), sizeof(PointXYZI) * o.size());
  return (int)o.size();
}
// returns 0 on the reference's error paths; sizes[3] out; clouds sized n
int orc_compact_decode(const float *data, int n, float *tf7, float *corner, float *surf, float *full, int *sizes) {
  Cloud in((const PointXYZI *)data, (const PointXYZI *)data + n), c, s, f;
  Transform t;
  if (!CompactDecode(in, t, c, s, f)) return 0;
  tf7[0] = t.rot.x; tf7[1] = t.rot.y; tf7[2] = t.rot.z; tf7[3] = t.rot.w; tf7[4] = t.pos.x; tf7[5] = t.pos.y; tf7[6] = t.pos.z;
  std::memcpy(corner, c.data(), sizeof(PointXYZI) * c.size());
  std::memcpy(surf, s.data(), sizeof(PointXYZI) * s.size());
  std::memcpy(full, f.data(), sizeof(PointXYZI) * f.size());
  sizes[0] = (int)c.size(); sizes[1] = (int)s.size(); sizes[2] = (int)f.size();
  return 1;
}
}
