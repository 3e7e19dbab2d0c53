// Device pcl::VoxelGrid<PointXYZI> (PCL 1.8 semantics) for the estimator's
// call sites (reference: src/imu_processor/Estimator.cc:679-687, :1518-1519).  Compiled with
// -fmad=false: voxel indices, output order (ascending voxel index) and centroids (float sums in
// input order inside a voxel, divided by the float count) are bit-identical to the CPU path.
//
//   vg_bbox   : min/max of the cloud (order-preserving uint encoding + atomics)
//   vg_keys   : ijk = floor(p*inv_leaf) - min_b, key = i + j*dx + k*dx*dy, value = input index
//   radix sort: stable, 4 x 8-bit passes (sort.cu)
//   vg_emit   : heads of equal-key runs -> ordered compaction by decoupled look-back, one centroid
//               per occupied voxel
#include "voxel.cuh"

namespace lio {

__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Also publishes n_eff = min(*n_dev, n_max): every later kernel of the filter (and the radix sort) reads the clamped
// count, so a caller-side count above the grid / scratch bound cannot run past the buffers.
__global__ void vg_reset(unsigned *bbox, int *ticket, const int *__restrict__ n_dev, int n_max, int *__restrict__ n_eff) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = 0u;
  if (threadIdx.x == 6) *ticket = 0;
  if (threadIdx.x == 7) { int n = *n_dev; *n_eff = n < 0 ? 0 : (n > n_max ? n_max : n); }
}

__global__ void __launch_bounds__(256)
vg_bbox(const float4 *__restrict__ in, const int *__restrict__ n_dev, unsigned *__restrict__ bbox) {
  const int n = *n_dev;
  float mn0 = INFINITY, mn1 = INFINITY, mn2 = INFINITY, mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(in + i);
    mn0 = fminf(mn0, p.x); mn1 = fminf(mn1, p.y); mn2 = fminf(mn2, p.z);
    mx0 = fmaxf(mx0, p.x); mx1 = fmaxf(mx1, p.y); mx2 = fmaxf(mx2, p.z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o)); mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
    mn2 = fminf(mn2, __shfl_xor_sync(0xffffffffu, mn2, o)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o)); mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
  }
  // one set of atomics per CTA, not per warp: with hundreds of CTAs the six addresses serialise (17 us measured)
  __shared__ float sm[256 / 32][6];
  if (lane_id() == 0) { float *r = sm[warp_id()]; r[0] = mn0; r[1] = mn1; r[2] = mn2; r[3] = mx0; r[4] = mx1; r[5] = mx2; }
  __syncthreads();
  if (threadIdx.x < 6) {
    const bool is_min = threadIdx.x < 3;
    float v = sm[0][threadIdx.x];
#pragma unroll
    for (int w = 1; w < 256 / 32; ++w) v = is_min ? fminf(v, sm[w][threadIdx.x]) : fmaxf(v, sm[w][threadIdx.x]);
    if (is_min) { if (v != INFINITY) atomicMin(bbox + threadIdx.x, f2ord(v)); }
    else { if (v != -INFINITY) atomicMax(bbox + threadIdx.x, f2ord(v)); }
  }
}

__global__ void __launch_bounds__(256)
vg_keys(const float4 *__restrict__ in, const int *__restrict__ n_dev, const unsigned *__restrict__ bbox, float leaf,
        unsigned *__restrict__ keys, unsigned *__restrict__ vals, int *__restrict__ overflow) {
  __shared__ int sp[6];
  const int n = *n_dev;
  if (blockIdx.x * blockDim.x >= n) return;
  const float inv = 1.0f / leaf;
  if (threadIdx.x == 0) {
    float a[6];
    for (int q = 0; q < 6; ++q) a[q] = ord2f(bbox[q]);
    long long ddx = (long long)((a[3] - a[0]) * inv) + 1, ddy = (long long)((a[4] - a[1]) * inv) + 1,
              ddz = (long long)((a[5] - a[2]) * inv) + 1;
    int ovf = (ddx * ddy * ddz > 2147483647LL) ? 1 : 0;
    int mb0 = (int)floorf(a[0] * inv), mb1 = (int)floorf(a[1] * inv), mb2 = (int)floorf(a[2] * inv);
    int xb0 = (int)floorf(a[3] * inv), xb1 = (int)floorf(a[4] * inv);
    sp[0] = mb0; sp[1] = mb1; sp[2] = mb2;
    sp[3] = xb0 - mb0 + 1;
    sp[4] = (xb0 - mb0 + 1) * (xb1 - mb1 + 1);
    sp[5] = ovf;
    if (blockIdx.x == 0 && ovf) *overflow = 1;   // sticky until the host reads and clears it
  }
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float4 p = __ldg(in + i);
    int ijk0 = (int)(floorf(p.x * inv) - (float)sp[0]);
    int ijk1 = (int)(floorf(p.y * inv) - (float)sp[1]);
    int ijk2 = (int)(floorf(p.z * inv) - (float)sp[2]);
    keys[i] = sp[5] ? (unsigned)i : (unsigned)(ijk0 + ijk1 * sp[3] + ijk2 * sp[4]);
    vals[i] = (unsigned)i;
  }
}

constexpr int kEmitThreads = 256;

__global__ void __launch_bounds__(kEmitThreads)
vg_emit(const float4 *__restrict__ in, const int *__restrict__ n_dev, const unsigned *__restrict__ keys,
        const unsigned *__restrict__ vals, float4 *__restrict__ out, int out_cap, int *__restrict__ nout_dev,
        unsigned *__restrict__ vox_key_out, unsigned long long *__restrict__ status, int *__restrict__ ticket) {
  __shared__ int sscan[40];
  __shared__ int stile, sbc;
  const int n = *n_dev;
  const int ntiles = (n + kEmitThreads - 1) / kEmitThreads;
  if (threadIdx.x == 0) stile = atomicAdd(ticket, 1);
  __syncthreads();
  const int tile = stile;
  if (tile >= ntiles) {
    if (n == 0 && tile == 0 && threadIdx.x == 0) *nout_dev = 0;
    return;
  }
  const int c = tile * kEmitThreads + threadIdx.x;
  int head = 0;
  unsigned v = 0;
  if (c < n) {
    v = keys[c];
    head = (c == 0) || (v != keys[c - 1]);
  }
  int tot;
  int lpos = block_scan_excl(head, sscan, &tot);
  int excl = lookback_exclusive(status, tile, tot, &sbc);
  if (head) {
    float ax = 0.f, ay = 0.f, az = 0.f, ai = 0.f;
    int cnt = 0;
    for (int c2 = c; c2 < n && keys[c2] == v; ++c2) {
      float4 p = __ldg(in + vals[c2]);
      ax += p.x; ay += p.y; az += p.z; ai += p.w;
      ++cnt;
    }
    float fn = (float)cnt;
    if (excl + lpos < out_cap) {   // the count below keeps growing: the host compares it with the capacity
      out[excl + lpos] = make_float4(ax / fn, ay / fn, az / fn, ai / fn);
      if (vox_key_out) vox_key_out[excl + lpos] = v;
    }
  }
  if (tile == ntiles - 1 && threadIdx.x == 0) *nout_dev = excl + tot;
}

int VoxelGrid::init(int cap_) {
  cap = cap_;
  if (cudaMalloc(&keys_a, sizeof(unsigned) * cap) != cudaSuccess) return -1;
  if (cudaMalloc(&vals_a, sizeof(unsigned) * cap) != cudaSuccess) return -1;
  if (cudaMalloc(&keys_b, sizeof(unsigned) * cap) != cudaSuccess) return -1;
  if (cudaMalloc(&vals_b, sizeof(unsigned) * cap) != cudaSuccess) return -1;
  if (rs.init(cap) != 0) return -1;
  nstatus = (cap + kEmitThreads - 1) / kEmitThreads + 1;
  if (cudaMalloc(&status, sizeof(unsigned long long) * nstatus) != cudaSuccess) return -1;
  if (cudaMalloc(&bbox, sizeof(unsigned) * 8) != cudaSuccess) return -1;
  if (cudaMalloc(&ticket, sizeof(int) * 4) != cudaSuccess) return -1;
  if (cudaMemset(ticket, 0, sizeof(int) * 4) != cudaSuccess) return -1;
  return 0;
}

void VoxelGrid::destroy() {
  void *p[] = {keys_a, vals_a, keys_b, vals_b, status, bbox, ticket};
  for (void *q : p) if (q) cudaFree(q);
  rs.destroy();
  keys_a = vals_a = keys_b = vals_b = nullptr; status = nullptr; bbox = nullptr; ticket = nullptr;
}

int VoxelGrid::run(const float4 *in, const int *n_dev_in, int n_max, float leaf, float4 *out, int out_cap, int *nout_dev,
                   unsigned *vox_key_out, cudaStream_t st, int *launches) {
  if (n_max > cap) return LIO_ERR_CAPACITY;
  if (n_max <= 0) n_max = 1;
  int *overflow = ticket + 1;
  int *n_dev = ticket + 2;   // clamped copy of the caller's count
  vg_reset<<<1, 32, 0, st>>>(bbox, ticket, n_dev_in, n_max, n_dev);
  int nblk = (n_max + 255) / 256;
  int bb_blocks = nblk < 592 ? nblk : 592;
  vg_bbox<<<bb_blocks, 256, 0, st>>>(in, n_dev, bbox);
  vg_keys<<<nblk, 256, 0, st>>>(in, n_dev, bbox, leaf, keys_a, vals_a, overflow);
  if (launches) *launches += 3;
  int which = radix_sort_pairs(keys_a, vals_a, keys_b, vals_b, n_dev, n_max, 32, rs, st, launches);
  if (which < 0) return LIO_ERR_CAPACITY;
  const unsigned *k = which ? keys_b : keys_a, *v = which ? vals_b : vals_a;
  int etiles = (n_max + kEmitThreads - 1) / kEmitThreads;
  cudaMemsetAsync(status, 0, sizeof(unsigned long long) * (size_t)(etiles + 1), st);
  vg_emit<<<etiles, kEmitThreads, 0, st>>>(in, n_dev, k, v, out, out_cap, nout_dev, vox_key_out, status, ticket);
  if (launches) *launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { lio_set_last_error(__FILE__, __LINE__, cudaGetErrorString(e)); return LIO_ERR_CUDA; }
  return LIO_OK;
}

}  // namespace lio

// ---- C-ABI: standalone voxel filter on host buffers (parity entry for the PCL call sites) -------
using namespace lio;

extern "C" int lio_voxel_grid_host(const float *cloud, int n, float leaf, float *out, int cap, int *n_out, int device) {
  if ((!cloud && n > 0) || n < 0 || !n_out || !(leaf > 0)) return LIO_ERR_INVALID;
  if (lio_device_count() <= 0) return LIO_ERR_NO_DEVICE;
  LIO_CUDA_OK(cudaSetDevice(device));
  *n_out = 0;
  if (n == 0) return LIO_OK;
  VoxelGrid vg;
  if (vg.init(n) != 0) { vg.destroy(); lio_set_last_error(__FILE__, __LINE__, "cudaMalloc failed"); return LIO_ERR_CUDA; }
  float4 *d_in = nullptr, *d_out = nullptr;
  int *d_n = nullptr;
  int rc = LIO_OK;
  if (cudaMalloc(&d_in, sizeof(float4) * n) != cudaSuccess || cudaMalloc(&d_out, sizeof(float4) * n) != cudaSuccess ||
      cudaMalloc(&d_n, 2 * sizeof(int)) != cudaSuccess) {
    rc = LIO_ERR_CUDA;
  }
  if (rc == LIO_OK) {
    cudaMemcpy(d_in, cloud, sizeof(float4) * n, cudaMemcpyHostToDevice);
    cudaMemcpy(d_n, &n, sizeof(int), cudaMemcpyHostToDevice);
    rc = vg.run(d_in, d_n, n, leaf, d_out, n, d_n + 1, nullptr, 0, nullptr);
    if (rc == LIO_OK) {
      int m = 0;
      cudaError_t e = cudaMemcpy(&m, d_n + 1, sizeof(int), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { lio_set_last_error(__FILE__, __LINE__, cudaGetErrorString(e)); rc = LIO_ERR_CUDA; }
      else {
        *n_out = m;
        if (m > cap) rc = LIO_ERR_CAPACITY;
        else if (m > 0) cudaMemcpy(out, d_out, sizeof(float4) * m, cudaMemcpyDeviceToHost);
      }
    }
  }
  if (d_in) cudaFree(d_in);
  if (d_out) cudaFree(d_out);
  if (d_n) cudaFree(d_n);
  vg.destroy();
  return rc;
}
