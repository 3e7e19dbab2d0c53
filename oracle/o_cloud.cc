// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
//
// Restatements of the PCL 1.8 primitives the reference's hot path calls (PCL is an un-vendored
// dependency, absent here: "parity unpinned" against PCL itself; semantics from SURVEY.md App. A):
//   pcl::VoxelGrid<PointXYZI>::applyFilter   call sites PointProcessor.cc:737-751,
//                                            Estimator.cc:679-687, :1518-1519
//   pcl::transformPointCloud(Affine3f)       call sites Estimator.cc:1425, :1498, :2600
//   pcl::KdTreeFLANN<PointXYZI>::nearestKSearch (exact k-NN, L2_Simple<float>)  Estimator.cc:1019
#include "o_api.h"
#include <cmath>
#include <algorithm>
#include <numeric>

namespace orc {

// VoxelGrid semantics restated: bbox from min/max, ijk = floor(p*inv_leaf) - min_b, linear index
// i + j*dx + k*dx*dy, points sorted by index, one centroid (x,y,z,intensity) per occupied voxel in
// ascending index order, float accumulation, division by the point count as float.
// Within-voxel summation order: PCL uses an unstable std::sort so its order is unspecified; the
// oracle defines it as ascending input index (stable sort).
void VoxelGridFilter(const Cloud &in, float leaf, Cloud &out) {
  out.clear();
  if (in.empty()) return;
  const float inv = 1.0f / leaf;
  float mn[3] = {in[0].x, in[0].y, in[0].z}, mx[3] = {in[0].x, in[0].y, in[0].z};
  for (const PointXYZI &p : in) {
    mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
    mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
  }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {  // PCL: warn, output = input
    out = in;
    return;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int)std::floor(mn[a] * inv);
    max_b[a] = (int)std::floor(mx[a] * inv);
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<std::pair<unsigned, int>> iv(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    int ijk0 = (int)(std::floor(in[i].x * inv) - (float)min_b[0]);
    int ijk1 = (int)(std::floor(in[i].y * inv) - (float)min_b[1]);
    int ijk2 = (int)(std::floor(in[i].z * inv) - (float)min_b[2]);
    int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
    iv[i] = std::make_pair((unsigned)idx, (int)i);
  }
  std::sort(iv.begin(), iv.end());  // (voxel, input index): stable order inside a voxel
  size_t first = 0;
  while (first < iv.size()) {
    size_t last = first + 1;
    while (last < iv.size() && iv[last].first == iv[first].first) ++last;
    float sx = 0, sy = 0, sz = 0, si = 0;
    for (size_t k = first; k < last; ++k) {
      const PointXYZI &p = in[iv[k].second];
      sx += p.x; sy += p.y; sz += p.z; si += p.intensity;
    }
    float n = (float)(last - first);
    PointXYZI c;
    c.x = sx / n; c.y = sy / n; c.z = sz / n; c.intensity = si / n;
    out.push_back(c);
    first = last;
  }
}

// pcl::transformPointCloud (1.8): x' = m00*x + m01*y + m02*z + m03, left to right, float.
void TransformCloudAffine(const Cloud &in, const Mat3<float> &R, const Vec3<float> &t, Cloud &out) {
  out.resize(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    const PointXYZI &p = in[i];
    PointXYZI o;
    o.x = R(0, 0) * p.x + R(0, 1) * p.y + R(0, 2) * p.z + t.x;
    o.y = R(1, 0) * p.x + R(1, 1) * p.y + R(1, 2) * p.z + t.y;
    o.z = R(2, 0) * p.x + R(2, 1) * p.y + R(2, 2) * p.z + t.z;
    o.intensity = p.intensity;
    out[i] = o;
  }
}

// ---- exact k-NN kd-tree ------------------------------------------------------------------------
static inline float coord(const PointXYZI &p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }

void KdTree::Build(const Cloud &c) {
  cloud = &c;
  idx.resize(c.size());
  std::iota(idx.begin(), idx.end(), 0);
  nodes.clear();
  nodes.reserve(c.size() / 4 + 16);
  if (c.empty()) return;
  float bmin[3], bmax[3];
  BuildRec(0, (int)c.size(), bmin, bmax);
}

int KdTree::BuildRec(int lo, int hi, float *bmin, float *bmax) {
  const Cloud &c = *cloud;
  int me = (int)nodes.size();
  nodes.push_back(Node());
  for (int d = 0; d < 3; ++d) { bmin[d] = coord(c[idx[lo]], d); bmax[d] = bmin[d]; }
  for (int i = lo + 1; i < hi; ++i)
    for (int d = 0; d < 3; ++d) {
      float v = coord(c[idx[i]], d);
      bmin[d] = std::min(bmin[d], v); bmax[d] = std::max(bmax[d], v);
    }
  if (hi - lo <= 15) {  // PCL: KDTreeSingleIndexParams(15)
    Node n; n.left = n.right = -1; n.lo = lo; n.hi = hi; n.dim = 0; n.split_lo = n.split_hi = 0;
    nodes[me] = n;
    return me;
  }
  int dim = 0;
  float span = bmax[0] - bmin[0];
  for (int d = 1; d < 3; ++d) if (bmax[d] - bmin[d] > span) { span = bmax[d] - bmin[d]; dim = d; }
  int mid = (lo + hi) / 2;
  std::nth_element(idx.begin() + lo, idx.begin() + mid, idx.begin() + hi, [&](int a, int b) {
    float va = coord(c[a], dim), vb = coord(c[b], dim);
    return va < vb || (va == vb && a < b);
  });
  float lmin[3], lmax[3], rmin[3], rmax[3];
  int l = BuildRec(lo, mid, lmin, lmax);
  int r = BuildRec(mid, hi, rmin, rmax);
  Node n; n.left = l; n.right = r; n.lo = lo; n.hi = hi; n.dim = dim;
  n.split_lo = lmax[dim];  // every left point has coord <= split_lo
  n.split_hi = rmin[dim];  // every right point has coord >= split_hi
  nodes[me] = n;
  return me;
}

static inline bool better(float d, int i, float bd, int bi) { return d < bd || (d == bd && i < bi); }

void KdTree::Search(int node, const float *q, float, float *, int k, int *bi, float *bd, int &cnt) const {
  const Node &n = nodes[node];
  const Cloud &c = *cloud;
  if (n.left < 0) {
    for (int t = n.lo; t < n.hi; ++t) {
      int i = idx[t];
      // flann::L2_Simple<float>: result += diff*diff over x,y,z (float, sequential)
      float d0 = q[0] - c[i].x, d1 = q[1] - c[i].y, d2 = q[2] - c[i].z;
      float d = 0.f;
      d += d0 * d0; d += d1 * d1; d += d2 * d2;
      if (cnt < k || better(d, i, bd[k - 1], bi[k - 1])) {
        int pos = (cnt < k) ? cnt : k - 1;
        if (cnt < k) ++cnt;
        while (pos > 0 && better(d, i, bd[pos - 1], bi[pos - 1])) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
        bd[pos] = d; bi[pos] = i;
      }
    }
    return;
  }
  float v = q[n.dim];
  // gap to each child's slab along the split dimension (0 if inside)
  float gl = (v > n.split_lo) ? (v - n.split_lo) : 0.f;
  float gr = (v < n.split_hi) ? (n.split_hi - v) : 0.f;
  int first = n.left, second = n.right;
  float g2 = gr;
  if (gr < gl) { first = n.right; second = n.left; g2 = gl; }
  Search(first, q, 0, nullptr, k, bi, bd, cnt);
  // a point of the far child has |dx_dim| >= g2, hence its float d2 >= fl(g2*g2): prune only if strictly worse
  float lb = g2 * g2;
  if (cnt < k || !(lb > bd[k - 1])) Search(second, q, 0, nullptr, k, bi, bd, cnt);
}

void KdTree::Knn(const PointXYZI &q, int k, int *out_idx, float *out_d2) const {
  float qq[3] = {q.x, q.y, q.z};
  int cnt = 0;
  for (int i = 0; i < k; ++i) { out_idx[i] = -1; out_d2[i] = std::numeric_limits<float>::infinity(); }
  if (!nodes.empty()) Search(0, qq, 0, nullptr, k, out_idx, out_d2, cnt);
}

}  // namespace orc
