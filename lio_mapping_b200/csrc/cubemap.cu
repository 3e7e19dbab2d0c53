// lio::PointMapping's rolling cube map and its Process() step (pre-initialisation scan-to-map path) with the map resident in
// HBM - SURVEY section 8 row f2.  Reference: src/point_processor/PointMapping.cc
//   constants / ToIndex        :77-82, :121-122, include/point_processor/PointMapping.h:150-159   (21 x 21 x 11 cubes of 50 m)
//   re-centring                :809-931      cube descriptors are shifted on the host: no point moves in HBM
//   cube selection             :944-1003     5 x 5 x 5 neighbourhood, FOV test on the eight cube corners (host, <= 125 cubes)
//   map extraction             :1005-1011    one gather kernel over the valid cubes' HBM segments
//   Process                    :765-1052     PointAssociateToMap / TobeMapped kernels, VoxelGrid of the stacks,
//                                            OptimizeTransformTobeMapped (scan_to_map_run: voxel-hash k-NN + 6 x 6 float GN)
//   UpdateMapDatabase          :1112-1208    order-preserving insert into the cubes + VoxelGrid of every touched valid cube
// Per-point work runs in kernels; the 4851-entry cube directory (pointer, count, capacity per cube) lives on the host and is
// the only thing the control logic touches.  Compiled with -fmad=false: the float expressions follow the reference's order,
// clouds, cube contents and the mapped pose are compared with the oracle (oracle/o_cubemap.cc).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>
#include "odom.cuh"
#include "voxel.cuh"
#include "twistf.h"

namespace lio {

constexpr int kCubeL = 21, kCubeW = 21, kCubeH = 11, kCubes = kCubeL * kCubeW * kCubeH;

// q * v (Eigen _transformVector) in float, host copy of the device expression
static void rotate_host(const TwistF &t, float vx, float vy, float vz, float &ox, float &oy, float &oz) {
  volatile float ux = t.qy * vz - t.qz * vy, uy = t.qz * vx - t.qx * vz, uz = t.qx * vy - t.qy * vx;
  volatile float ux2 = ux + ux, uy2 = uy + uy, uz2 = uz + uz;
  volatile float cx = t.qy * uz2 - t.qz * uy2, cy = t.qz * ux2 - t.qx * uz2, cz = t.qx * uy2 - t.qy * ux2;
  volatile float ax = ux2 * t.qw, ay = uy2 * t.qw, az = uz2 * t.qw;
  volatile float rx = vx + ax, ry = vy + ay, rz = vz + az;
  ox = rx + cx; oy = ry + cy; oz = rz + cz;
}

// ---- kernels ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rotate_dev(float qx, float qy, float qz, float qw, float vx, float vy, float vz, float &ox, float &oy, float &oz) {
  float ux = qy * vz - qz * vy, uy = qz * vx - qx * vz, uz = qx * vy - qy * vx;
  ux += ux; uy += uy; uz += uz;
  const float cx = qy * uz - qz * uy, cy = qz * ux - qx * uz, cz = qx * uy - qy * ux;
  ox = vx + ux * qw + cx; oy = vy + uy * qw + cy; oz = vz + uz * qw + cz;
}

// mode 0: PointAssociateToMap (po = q * pi + t, :303-314); mode 1: PointAssociateTobeMapped (po = q^* * (pi - t), :316-323)
__global__ void __launch_bounds__(256)
k_associate(const float4 *__restrict__ in, float4 *__restrict__ out, const int *__restrict__ n_dev, TwistF t, int mode) {
  const int n = *n_dev;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(in + i);
  float x, y, z;
  if (mode == 0) {
    rotate_dev(t.qx, t.qy, t.qz, t.qw, p.x, p.y, p.z, x, y, z);
    x += t.px; y += t.py; z += t.pz;
  } else {
    rotate_dev(-t.qx, -t.qy, -t.qz, t.qw, p.x - t.px, p.y - t.py, p.z - t.pz, x, y, z);
  }
  out[i] = make_float4(x, y, z, p.w);
}

// int((v + 25.0) / 50.0) + cen, minus one for negatives (:812-819) - double arithmetic like the reference
__device__ __forceinline__ int cube_of(float v, int cen) {
  int c = int(((double)v + 25.0) / 50.0) + cen;
  if ((double)v + 25.0 < 0) --c;
  return c;
}

// UpdateMapDatabase insert, phase 1: map-frame point and destination cube of every down-sampled stack point (-1: outside)
__global__ void __launch_bounds__(256)
k_cube_ids(const float4 *__restrict__ in, const int *__restrict__ n_dev, TwistF t, int cen_l, int cen_w, int cen_h, float4 *__restrict__ mapped,
           int *__restrict__ cube) {
  const int n = *n_dev;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(in + i);
  float x, y, z;
  rotate_dev(t.qx, t.qy, t.qz, t.qw, p.x, p.y, p.z, x, y, z);
  x += t.px; y += t.py; z += t.pz;
  mapped[i] = make_float4(x, y, z, p.w);
  const int ci = cube_of(x, cen_l), cj = cube_of(y, cen_w), ck = cube_of(z, cen_h);
  cube[i] = (ci >= 0 && ci < kCubeL && cj >= 0 && cj < kCubeW && ck >= 0 && ck < kCubeH) ? ci + kCubeL * cj + kCubeL * kCubeW * ck : -1;
}

// phase 2: dst[i] is the address the host directory assigned to point i (append position inside its cube, input order kept)
__global__ void __launch_bounds__(256)
k_scatter_to_cubes(const float4 *__restrict__ mapped, float4 *const *__restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 *d = dst[i];
  if (d) *d = __ldg(mapped + i);
}

struct Segment { const float4 *src; int n; int off; };
// concatenation of cube segments (laser_cloud_*_from_map_, :1005-1011); one block range per segment
__global__ void __launch_bounds__(256)
k_gather_segments(const Segment *__restrict__ seg, int nseg, float4 *__restrict__ out) {
  for (int s = blockIdx.y; s < nseg; s += gridDim.y) {
    const Segment sg = seg[s];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += gridDim.x * blockDim.x) out[sg.off + i] = __ldg(sg.src + i);
  }
}

}  // namespace lio

using namespace lio;

struct lio_pm {
  struct Cube { float4 *p = nullptr; int n = 0, cap = 0; };
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm = 148;
  int max_points = 0;
  std::vector<Cube> cube[2];           // [0] corner, [1] surf: kCubes descriptors each
  int cen_l = 10, cen_w = 10, cen_h = 5;
  float leaf[2] = {0.2f, 0.4f};
  float min_match_sq_dis = 1.0f, min_plane_dis = 0.2f;
  int max_iter = 10;
  double delta_r_abort = 0.05, delta_t_abort = 0.05;
  TwistF sum, bef, aft, tobe;          // transform_sum_, transform_bef_mapped_, transform_aft_mapped_, transform_tobe_mapped_
  // device scratch
  float4 *d_in[2] = {nullptr, nullptr}, *d_stack[2] = {nullptr, nullptr}, *d_ds[2] = {nullptr, nullptr}, *d_mapped = nullptr, *d_tmp = nullptr;
  float4 *d_map[2] = {nullptr, nullptr};
  int map_cap[2] = {0, 0};
  int *d_cnt = nullptr;                // [0,1] input sizes, [2,3] down-sampled sizes, [4..] voxel-grid outputs
  int *d_cube = nullptr;
  float4 **d_dst = nullptr;
  Segment *d_seg = nullptr;
  int *d_vgout = nullptr;              // per re-filtered cube: output count
  VoxelGrid vg;
  int vg_cap = 0;
  ScanToMapWork stm;
  int stm_cap[3] = {0, 0, 0};
  int last_iters = 0, last_from_map[2] = {0, 0};
  std::vector<int> h_cube;
  std::vector<float4 *> h_dst;
};

static size_t to_index(int i, int j, int k) { return (size_t)i + (size_t)kCubeL * j + (size_t)kCubeL * kCubeW * k; }
static int cube_of_host(float v, int cen) {
  int c = int(((double)v + 25.0) / 50.0) + cen;
  if ((double)v + 25.0 < 0) --c;
  return c;
}

extern "C" int lio_pm_destroy(lio_pm *m) {
  if (!m) return LIO_OK;
  cudaSetDevice(m->device);
  for (int w = 0; w < 2; ++w) {
    for (lio_pm::Cube &c : m->cube[w]) if (c.p) cudaFree(c.p);
    void *fr[] = {m->d_in[w], m->d_stack[w], m->d_ds[w], m->d_map[w]};
    for (void *q : fr) if (q) cudaFree(q);
  }
  void *fr[] = {m->d_mapped, m->d_tmp, m->d_cnt, m->d_cube, m->d_dst, m->d_seg, m->d_vgout};
  for (void *q : fr) if (q) cudaFree(q);
  m->vg.destroy();
  m->stm.destroy();
  delete m;
  return LIO_OK;
}

extern "C" int lio_pm_create(int max_points, float corner_filter_size, float surf_filter_size, float min_match_sq_dis, float min_plane_dis,
                             int max_iterations, int device, void *cuda_stream, lio_pm **out) {
  if (!out || max_points < 16 || !(corner_filter_size > 0) || !(surf_filter_size > 0) || max_iterations < 0) return LIO_ERR_INVALID;
  if (lio_device_count() <= 0) return LIO_ERR_NO_DEVICE;
  LIO_CUDA_OK(cudaSetDevice(device));
  lio_pm *m = new (std::nothrow) lio_pm();
  if (!m) return LIO_ERR_INVALID;
  m->device = device; m->stream = (cudaStream_t)cuda_stream; m->max_points = max_points;
  m->leaf[0] = corner_filter_size; m->leaf[1] = surf_filter_size;
  m->min_match_sq_dis = min_match_sq_dis; m->min_plane_dis = min_plane_dis; m->max_iter = max_iterations;
  cudaDeviceGetAttribute(&m->sm, cudaDevAttrMultiProcessorCount, device);
  m->cube[0].assign(kCubes, lio_pm::Cube());
  m->cube[1].assign(kCubes, lio_pm::Cube());
  bool ok = true;
  for (int w = 0; w < 2 && ok; ++w) {
    ok = ok && cudaMalloc(&m->d_in[w], sizeof(float4) * max_points) == cudaSuccess;
    ok = ok && cudaMalloc(&m->d_stack[w], sizeof(float4) * max_points) == cudaSuccess;
    ok = ok && cudaMalloc(&m->d_ds[w], sizeof(float4) * max_points) == cudaSuccess;
  }
  ok = ok && cudaMalloc(&m->d_mapped, sizeof(float4) * max_points) == cudaSuccess;
  ok = ok && cudaMalloc(&m->d_cnt, sizeof(int) * 8) == cudaSuccess;
  ok = ok && cudaMalloc(&m->d_cube, sizeof(int) * max_points) == cudaSuccess;
  ok = ok && cudaMalloc(&m->d_dst, sizeof(float4 *) * max_points) == cudaSuccess;
  ok = ok && cudaMalloc(&m->d_seg, sizeof(Segment) * 256) == cudaSuccess;
  ok = ok && cudaMalloc(&m->d_vgout, sizeof(int) * 256) == cudaSuccess;
  m->vg_cap = max_points;
  ok = ok && cudaMalloc(&m->d_tmp, sizeof(float4) * m->vg_cap) == cudaSuccess;
  ok = ok && m->vg.init(m->vg_cap) == 0;
  if (!ok) { lio_set_last_error(__FILE__, __LINE__, "lio_pm_create: device allocation failed"); lio_pm_destroy(m); return LIO_ERR_CUDA; }
  m->h_cube.resize(max_points);
  m->h_dst.resize(max_points);
  *out = m;
  return LIO_OK;
}

// ---- host-side directory logic (the reference's own index arithmetic) ---------------------------------------------------
static void pm_recentre(lio_pm *m, float px, float py, float pz, int &ci, int &cj, int &ck) {   // :812-931
  ci = cube_of_host(px, m->cen_l); cj = cube_of_host(py, m->cen_w); ck = cube_of_host(pz, m->cen_h);
  auto shift = [&](int axis, int dir) {  // dir +1: contents move towards higher indices, the low face is cleared
    const int n[3] = {kCubeL, kCubeW, kCubeH};
    int idx[3];
    for (idx[(axis + 1) % 3] = 0; idx[(axis + 1) % 3] < n[(axis + 1) % 3]; ++idx[(axis + 1) % 3])
      for (idx[(axis + 2) % 3] = 0; idx[(axis + 2) % 3] < n[(axis + 2) % 3]; ++idx[(axis + 2) % 3]) {
        if (dir > 0) {
          for (int a = n[axis] - 1; a >= 1; --a) {
            idx[axis] = a; const size_t ia = to_index(idx[0], idx[1], idx[2]);
            idx[axis] = a - 1; const size_t ib = to_index(idx[0], idx[1], idx[2]);
            std::swap(m->cube[0][ia], m->cube[0][ib]); std::swap(m->cube[1][ia], m->cube[1][ib]);
          }
          idx[axis] = 0;
        } else {
          for (int a = 0; a < n[axis] - 1; ++a) {
            idx[axis] = a; const size_t ia = to_index(idx[0], idx[1], idx[2]);
            idx[axis] = a + 1; const size_t ib = to_index(idx[0], idx[1], idx[2]);
            std::swap(m->cube[0][ia], m->cube[0][ib]); std::swap(m->cube[1][ia], m->cube[1][ib]);
          }
          idx[axis] = n[axis] - 1;
        }
        const size_t ic = to_index(idx[0], idx[1], idx[2]);
        m->cube[0][ic].n = 0; m->cube[1][ic].n = 0;   // clear(): the HBM segment is kept for re-use
      }
  };
  while (ci < 3) { shift(0, +1); ++ci; ++m->cen_l; }
  while (ci >= kCubeL - 3) { shift(0, -1); --ci; --m->cen_l; }
  while (cj < 3) { shift(1, +1); ++cj; ++m->cen_w; }
  while (cj >= kCubeW - 3) { shift(1, -1); --cj; --m->cen_w; }
  while (ck < 3) { shift(2, +1); ++ck; ++m->cen_h; }
  while (ck >= kCubeH - 3) { shift(2, -1); --ck; --m->cen_h; }
}

static void pm_select(const lio_pm *m, float px, float py, float pz, const float z[3], int ci, int cj, int ck, std::vector<size_t> &valid) {   // :944-1003
  valid.clear();
  for (int i = ci - 2; i <= ci + 2; ++i)
    for (int j = cj - 2; j <= cj + 2; ++j)
      for (int k = ck - 2; k <= ck + 2; ++k) {
        if (!(i >= 0 && i < kCubeL && j >= 0 && j < kCubeW && k >= 0 && k < kCubeH)) continue;
        const float center_x = 50.0f * (i - m->cen_l), center_y = 50.0f * (j - m->cen_w), center_z = 50.0f * (k - m->cen_h);
        bool is_in_laser_fov = false;
        for (int ii = -1; ii <= 1; ii += 2)
          for (int jj = -1; jj <= 1; jj += 2)
            for (int kk = -1; kk <= 1; kk += 2) {
              const float cx = center_x + 25.0f * ii, cy = center_y + 25.0f * jj, cz = center_z + 25.0f * kk;
              const float d0 = px - cx, d1 = py - cy, d2 = pz - cz;
              volatile float s1 = d0 * d0; s1 = s1 + d1 * d1; s1 = s1 + d2 * d2;
              const float e0 = z[0] - cx, e1 = z[1] - cy, e2 = z[2] - cz;
              volatile float s2 = e0 * e0; s2 = s2 + e1 * e1; s2 = s2 + e2 * e2;
              const float squared_side1 = s1, squared_side2 = s2;
              const float check1 = 100.0f + squared_side1 - squared_side2 - 10.0f * std::sqrt(3.0f) * std::sqrt(squared_side1);
              const float check2 = 100.0f + squared_side1 - squared_side2 + 10.0f * std::sqrt(3.0f) * std::sqrt(squared_side1);
              if (check1 < 0 && check2 > 0) is_in_laser_fov = true;
            }
        if (is_in_laser_fov) valid.push_back(to_index(i, j, k));
      }
}

static int pm_grow(lio_pm *m, lio_pm::Cube &c, int need) {
  if (need <= c.cap) return LIO_OK;
  int cap = std::max(1024, c.cap);
  while (cap < need) cap *= 2;
  float4 *p = nullptr;
  LIO_CUDA_OK(cudaMalloc(&p, sizeof(float4) * cap));
  if (c.p && c.n > 0) LIO_CUDA_OK(cudaMemcpyAsync(p, c.p, sizeof(float4) * c.n, cudaMemcpyDeviceToDevice, m->stream));
  if (c.p) { LIO_CUDA_OK(cudaStreamSynchronize(m->stream)); cudaFree(c.p); }
  c.p = p; c.cap = cap;
  return LIO_OK;
}

// laser_cloud_*_from_map_: concatenate the valid cubes (in `valid` order) into d_map[w]
static int pm_from_map(lio_pm *m, const std::vector<size_t> &valid, int w, int &total) {
  std::vector<Segment> seg;
  total = 0;
  for (size_t v : valid) {
    const lio_pm::Cube &c = m->cube[w][v];
    if (c.n > 0) { seg.push_back(Segment{c.p, c.n, total}); total += c.n; }
  }
  if (total > m->map_cap[w]) {
    if (m->d_map[w]) cudaFree(m->d_map[w]);
    m->map_cap[w] = std::max(2 * total, 1 << 16);
    LIO_CUDA_OK(cudaMalloc(&m->d_map[w], sizeof(float4) * m->map_cap[w]));
  }
  if (seg.empty()) return LIO_OK;
  if (seg.size() > 256) return LIO_ERR_CAPACITY;
  LIO_CUDA_OK(cudaMemcpyAsync(m->d_seg, seg.data(), sizeof(Segment) * seg.size(), cudaMemcpyHostToDevice, m->stream));
  LIO_CUDA_OK(cudaStreamSynchronize(m->stream));   // seg is a stack vector
  k_gather_segments<<<dim3(16, (unsigned)seg.size()), 256, 0, m->stream>>>(m->d_seg, (int)seg.size(), m->d_map[w]);
  return LIO_OK;
}

// UpdateMapDatabase (:1112-1208) with margin centre == current centre (the valid list was computed in this call)
static int pm_update(lio_pm *m, const std::vector<size_t> &valid, const int n_ds[2]) {
  cudaStream_t st = m->stream;
  for (int w = 0; w < 2; ++w) {
    const int n = n_ds[w];
    if (n == 0) continue;
    k_cube_ids<<<(n + 255) / 256, 256, 0, st>>>(m->d_ds[w], m->d_cnt + 2 + w, m->tobe, m->cen_l, m->cen_w, m->cen_h, m->d_mapped, m->d_cube);
    LIO_CUDA_OK(cudaMemcpyAsync(m->h_cube.data(), m->d_cube, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    LIO_CUDA_OK(cudaStreamSynchronize(st));
    // append positions in input order (push_back order): first the per-cube totals to size the segments, then the addresses
    std::vector<int> add(kCubes, 0);
    for (int i = 0; i < n; ++i) if (m->h_cube[i] >= 0) ++add[m->h_cube[i]];
    for (int c = 0; c < kCubes; ++c)
      if (add[c]) { int rc = pm_grow(m, m->cube[w][c], m->cube[w][c].n + add[c]); if (rc != LIO_OK) return rc; }
    for (int i = 0; i < n; ++i) {
      const int c = m->h_cube[i];
      if (c < 0) { m->h_dst[i] = nullptr; continue; }
      lio_pm::Cube &cb = m->cube[w][c];
      m->h_dst[i] = cb.p + cb.n;
      ++cb.n;
    }
    LIO_CUDA_OK(cudaMemcpyAsync(m->d_dst, m->h_dst.data(), sizeof(float4 *) * n, cudaMemcpyHostToDevice, st));
    k_scatter_to_cubes<<<(n + 255) / 256, 256, 0, st>>>(m->d_mapped, m->d_dst, n);
    LIO_CUDA_OK(cudaStreamSynchronize(st));   // h_dst is reused by the next cloud
  }
  // re-filter every valid cube (corner then surf), each with its own bounding box like pcl::VoxelGrid on that cube's cloud
  struct Job { int w; size_t idx; };
  std::vector<Job> jobs;
  for (size_t index : valid) {
    int li, lj, lk;
    { int residual = (int)(index % (kCubeL * kCubeW)); lk = (int)(index / (kCubeL * kCubeW)); lj = residual / kCubeL; li = residual % kCubeL; }
    const float center_x = 50.0f * (li - m->cen_l), center_y = 50.0f * (lj - m->cen_w), center_z = 50.0f * (lk - m->cen_h);
    const int ci = cube_of_host(center_x, m->cen_l), cj = cube_of_host(center_y, m->cen_w), ck = cube_of_host(center_z, m->cen_h);
    if (!(ci >= 0 && ci < kCubeL && cj >= 0 && cj < kCubeW && ck >= 0 && ck < kCubeH)) continue;
    const size_t idx = to_index(ci, cj, ck);
    for (int w = 0; w < 2; ++w) if (m->cube[w][idx].n > 0) jobs.push_back(Job{w, idx});
  }
  for (size_t b0 = 0; b0 < jobs.size(); b0 += 256) {
    const size_t b1 = std::min(jobs.size(), b0 + 256);
    std::vector<int> hn(b1 - b0);
    for (size_t j = b0; j < b1; ++j) {
      lio_pm::Cube &c = m->cube[jobs[j].w][jobs[j].idx];
      if (c.n > m->vg_cap) return LIO_ERR_CAPACITY;
      // input count through d_vgout[j] itself (read before the filter overwrites it with the output count)
      hn[j - b0] = c.n;
    }
    LIO_CUDA_OK(cudaMemcpyAsync(m->d_vgout, hn.data(), sizeof(int) * hn.size(), cudaMemcpyHostToDevice, st));
    LIO_CUDA_OK(cudaStreamSynchronize(st));
    for (size_t j = b0; j < b1; ++j) {
      lio_pm::Cube &c = m->cube[jobs[j].w][jobs[j].idx];
      int rc = m->vg.run(c.p, m->d_vgout + (j - b0), c.n, m->leaf[jobs[j].w], m->d_tmp, m->vg_cap, m->d_vgout + (j - b0), nullptr, st, nullptr);
      if (rc != LIO_OK) return rc;
      LIO_CUDA_OK(cudaMemcpyAsync(c.p, m->d_tmp, sizeof(float4) * c.n, cudaMemcpyDeviceToDevice, st));   // output <= input count
    }
    LIO_CUDA_OK(cudaMemcpyAsync(hn.data(), m->d_vgout, sizeof(int) * hn.size(), cudaMemcpyDeviceToHost, st));
    LIO_CUDA_OK(cudaStreamSynchronize(st));
    for (size_t j = b0; j < b1; ++j) m->cube[jobs[j].w][jobs[j].idx].n = hn[j - b0];
  }
  return LIO_OK;
}

// PointMapping::Process (:765-1052), imu_inited_ == false, num_stack_frames_ == 1.  Clouds: HOST arrays of n x 4 floats.
extern "C" int lio_pm_process_host(lio_pm *m, const float *corner_last, int nc, const float *surf_last, int ns, const float transform_sum7[7],
                                   float transform_tobe_mapped7[7], int info3[3]) {
  if (!m || !transform_sum7 || nc < 0 || ns < 0 || (nc > 0 && !corner_last) || (ns > 0 && !surf_last)) return LIO_ERR_INVALID;
  if (nc > m->max_points || ns > m->max_points) return LIO_ERR_CAPACITY;
  LIO_CUDA_OK(cudaSetDevice(m->device));
  cudaStream_t st = m->stream;
  const float *src[2] = {corner_last, surf_last};
  const int nin[2] = {nc, ns};
  m->sum = TwistF{transform_sum7[0], transform_sum7[1], transform_sum7[2], transform_sum7[3], transform_sum7[4], transform_sum7[5], transform_sum7[6]};
  m->tobe = twist_mul(m->tobe, twist_mul(twist_inverse(m->bef), m->sum));   // TransformAssociateToMap :753-756
  int hcnt[4] = {nc, ns, 0, 0};
  LIO_CUDA_OK(cudaMemcpyAsync(m->d_cnt, hcnt, sizeof(int) * 2, cudaMemcpyHostToDevice, st));
  for (int w = 0; w < 2; ++w) {
    if (nin[w] == 0) continue;
    LIO_CUDA_OK(cudaMemcpyAsync(m->d_in[w], src[w], sizeof(float4) * nin[w], cudaMemcpyHostToDevice, st));
    // to the map frame with the predicted pose, and back (the reference stacks in the map frame first, :782-800, :1013-1016)
    k_associate<<<(nin[w] + 255) / 256, 256, 0, st>>>(m->d_in[w], m->d_stack[w], m->d_cnt + w, m->tobe, 0);
    k_associate<<<(nin[w] + 255) / 256, 256, 0, st>>>(m->d_stack[w], m->d_stack[w], m->d_cnt + w, m->tobe, 1);
  }
  LIO_CUDA_OK(cudaStreamSynchronize(st));   // hcnt is a stack array
  float z[3];
  {  // point_on_z_axis_ = tobe * (0, 0, 10)
    rotate_host(m->tobe, 0.0f, 0.0f, 10.0f, z[0], z[1], z[2]);
    z[0] += m->tobe.px; z[1] += m->tobe.py; z[2] += m->tobe.pz;
  }
  int ci, cj, ck;
  pm_recentre(m, m->tobe.px, m->tobe.py, m->tobe.pz, ci, cj, ck);
  std::vector<size_t> valid;
  pm_select(m, m->tobe.px, m->tobe.py, m->tobe.pz, z, ci, cj, ck, valid);
  int K[2] = {0, 0};
  for (int w = 0; w < 2; ++w) { int rc = pm_from_map(m, valid, w, K[w]); if (rc != LIO_OK) return rc; }
  m->last_from_map[0] = K[0]; m->last_from_map[1] = K[1];
  // down-sample the stacks
  int n_ds[2] = {0, 0};
  for (int w = 0; w < 2; ++w) {
    if (nin[w] == 0) { LIO_CUDA_OK(cudaMemsetAsync(m->d_cnt + 2 + w, 0, sizeof(int), st)); continue; }
    int rc = m->vg.run(m->d_stack[w], m->d_cnt + w, nin[w], m->leaf[w], m->d_ds[w], m->max_points, m->d_cnt + 2 + w, nullptr, st, nullptr);
    if (rc != LIO_OK) return rc;
  }
  LIO_CUDA_OK(cudaMemcpyAsync(hcnt + 2, m->d_cnt + 2, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
  LIO_CUDA_OK(cudaStreamSynchronize(st));
  n_ds[0] = hcnt[2]; n_ds[1] = hcnt[3];
  // OptimizeTransformTobeMapped against the pulled map
  const bool optimised = !(K[0] <= 10 || K[1] <= 100);
  m->last_iters = 0;
  if (optimised && m->max_iter > 0) {
    if (K[0] > m->stm_cap[0] || K[1] > m->stm_cap[1] || n_ds[0] + n_ds[1] > m->stm_cap[2]) {
      m->stm.destroy();
      m->stm_cap[0] = std::max(2 * K[0], 1 << 15); m->stm_cap[1] = std::max(2 * K[1], 1 << 16); m->stm_cap[2] = std::max(2 * (n_ds[0] + n_ds[1]), 1 << 15);
      if (m->stm.init(m->stm_cap[0], m->stm_cap[1], m->stm_cap[2]) != 0) { lio_set_last_error(__FILE__, __LINE__, "scan-to-map workspace allocation failed"); return LIO_ERR_CUDA; }
    }
    float tf7[7] = {m->tobe.qx, m->tobe.qy, m->tobe.qz, m->tobe.qw, m->tobe.px, m->tobe.py, m->tobe.pz};
    int rc = scan_to_map_run(m->stm, m->d_map[0], K[0], m->d_map[1], K[1], m->d_ds[0], m->d_cnt + 2, std::max(n_ds[0], 1), m->d_ds[1], m->d_cnt + 3,
                             std::max(n_ds[1], 1), tf7, m->min_match_sq_dis, m->min_plane_dis, m->max_iter, m->delta_r_abort, m->delta_t_abort, 0, nullptr,
                             &m->last_iters, m->sm, st);
    if (rc != LIO_OK) return rc;
    m->tobe = TwistF{tf7[0], tf7[1], tf7[2], tf7[3], tf7[4], tf7[5], tf7[6]};
  }
  if (optimised) { m->bef = m->sum; m->aft = m->tobe; }   // TransformUpdate sits behind the optimiser's early return (:327-329, :716)
  int rc = pm_update(m, valid, n_ds);
  if (rc != LIO_OK) return rc;
  if (transform_tobe_mapped7) {
    transform_tobe_mapped7[0] = m->tobe.qx; transform_tobe_mapped7[1] = m->tobe.qy; transform_tobe_mapped7[2] = m->tobe.qz; transform_tobe_mapped7[3] = m->tobe.qw;
    transform_tobe_mapped7[4] = m->tobe.px; transform_tobe_mapped7[5] = m->tobe.py; transform_tobe_mapped7[6] = m->tobe.pz;
  }
  if (info3) { info3[0] = m->last_iters; info3[1] = K[0]; info3[2] = K[1]; }
  return LIO_OK;
}

extern "C" int lio_pm_map_centre(lio_pm *m, int centre3[3]) {
  if (!m || !centre3) return LIO_ERR_INVALID;
  centre3[0] = m->cen_l; centre3[1] = m->cen_w; centre3[2] = m->cen_h;
  return LIO_OK;
}

extern "C" int lio_pm_cube_size(lio_pm *m, int cube_index, int which, int *n) {
  if (!m || !n || cube_index < 0 || cube_index >= kCubes || which < 0 || which > 1) return LIO_ERR_INVALID;
  *n = m->cube[which][cube_index].n;
  return LIO_OK;
}

extern "C" int lio_pm_cube_download(lio_pm *m, int cube_index, int which, float *out_xyzi, int cap) {
  if (!m || !out_xyzi || cube_index < 0 || cube_index >= kCubes || which < 0 || which > 1) return LIO_ERR_INVALID;
  const lio_pm::Cube &c = m->cube[which][cube_index];
  if (c.n > cap) return LIO_ERR_CAPACITY;
  LIO_CUDA_OK(cudaSetDevice(m->device));
  if (c.n > 0) LIO_CUDA_OK(cudaMemcpyAsync(out_xyzi, c.p, sizeof(float4) * c.n, cudaMemcpyDeviceToHost, m->stream));
  LIO_CUDA_OK(cudaStreamSynchronize(m->stream));
  return LIO_OK;
}
