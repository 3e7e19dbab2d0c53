// Stage C on sm_100a — one fused kernel per Gauss-Newton iteration evaluates every
// PivotPointPlaneFactor (reference: src/factor/PivotPointPlaneFactor.cc:43-137, residual blocks
// added at src/imu_processor/Estimator.cc:1831-1889 with CauchyLoss(1.0), :1664) and reduces their
// contribution to the normal equations.
//
// Algebra: with T_lpi = (R, P) the lidar pose of frame i in the pivot lidar
// frame, a = R^T w, g = [a ; p x a], r = a.(p + R^T P) + b, the factor's 1x18 Jacobian row over
// (pose_pivot, pose_i, extrinsic) is g^T M_i with a 6x18 matrix M_i that depends on the state only.
// Ceres' Cauchy corrector (rho'' < 0 => alpha = 0) scales residual and Jacobian by sqrt(rho'),
// rho' = 1/(1+r^2).  Hence per frame the lidar part of J^T J, J^T r is M_i^T S_i M_i with
//     S_i = sum_k rho'(r_k^2) [g_k; r_k][g_k; r_k]^T      (7x7 symmetric, 28 numbers)
// and the cost is 1/2 sum_k log(1 + r_k^2).  The kernel streams 32 B per feature (float4 point +
// float4 plane), does the arithmetic in fp64 (inputs are exact fp32 values), and reduces with warp
// shuffles -> shared memory -> per-tile partials; the last CTA to finish sums the partials of each
// frame in tile order, so the result is deterministic and needs no second launch.
#include "assemble.cuh"

namespace lio {

constexpr int kAsmThreads = 256;

__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void accumulate(double (&acc)[29], const double (&R)[9], const double (&t)[3], float4 pf, float4 cf) {
  const double px = pf.x, py = pf.y, pz = pf.z;
  const double wx = cf.x, wy = cf.y, wz = cf.z, b = cf.w;
  double u[7];
  // a = R^T w
  u[0] = R[0] * wx + R[3] * wy + R[6] * wz;
  u[1] = R[1] * wx + R[4] * wy + R[7] * wz;
  u[2] = R[2] * wx + R[5] * wy + R[8] * wz;
  // p x a
  u[3] = py * u[2] - pz * u[1];
  u[4] = pz * u[0] - px * u[2];
  u[5] = px * u[1] - py * u[0];
  // r = a.(p + t) + b
  u[6] = u[0] * (px + t[0]) + u[1] * (py + t[1]) + u[2] * (pz + t[2]) + b;
  const double s = 1.0 + u[6] * u[6];
  const double c = 1.0 / s;  // rho'
  acc[28] += log(s);         // rho
  int k = 0;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const double cu = c * u[i];
#pragma unroll
    for (int j = i; j < 7; ++j) acc[k++] += cu * u[j];
  }
}

__global__ void __launch_bounds__(kAsmThreads)
asm_ppp(const AsmParams P, double *__restrict__ partial, double *__restrict__ out, unsigned *__restrict__ counter) {
  __shared__ double sred[kAsmThreads / 32][29];
  __shared__ bool is_last;
  const int tile = blockIdx.x;
  // frame of this tile
  int fi = 0;
#pragma unroll 1
  for (int k = 1; k < P.nframes; ++k) if (tile >= P.f[k].tile0) fi = k;
  const AsmFrame &F = P.f[fi];
  double R[9], t[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = F.R[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) t[k] = F.t[k];
  double acc[29];
#pragma unroll
  for (int k = 0; k < 29; ++k) acc[k] = 0.0;
  const int begin = (tile - F.tile0) * P.tile_feats;
  const int end = min(begin + P.tile_feats, F.n);
  // two independent loads in flight per thread
  int i = begin + threadIdx.x;
  for (; i + kAsmThreads < end; i += 2 * kAsmThreads) {
    float4 p0 = ld_stream(F.pts + i), c0 = ld_stream(F.coef + i);
    float4 p1 = ld_stream(F.pts + i + kAsmThreads), c1 = ld_stream(F.coef + i + kAsmThreads);
    accumulate(acc, R, t, p0, c0);
    accumulate(acc, R, t, p1, c1);
  }
  if (i < end) accumulate(acc, R, t, ld_stream(F.pts + i), ld_stream(F.coef + i));
  // warp tree, then cross-warp in fixed order
#pragma unroll
  for (int k = 0; k < 29; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane_id() == 0) sred[warp_id()][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 29) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kAsmThreads / 32; ++w) v += sred[w][threadIdx.x];
    partial[(size_t)tile * kAsmStride + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned prev = atomicAdd(counter, 1u);
    is_last = (prev == (unsigned)(P.ntiles - 1));
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int q = threadIdx.x; q < P.nframes * 29; q += kAsmThreads) {
    const int f = q / 29, k = q - f * 29;
    const int t0 = P.f[f].tile0;
    const int t1 = (f + 1 < P.nframes) ? P.f[f + 1].tile0 : P.ntiles;
    double v = 0.0;
    for (int tt = t0; tt < t1; ++tt) v += __ldcg(partial + (size_t)tt * kAsmStride + k);
    out[f * kAsmStride + k] = v;
  }
  if (threadIdx.x == 0) *counter = 0u;
}

int AsmWork::init(int max_features_total) {
  ntiles_max = max_features_total / kAsmThreads + 2 * kMaxOpt + 16;
  if (cudaMalloc(&partial, sizeof(double) * (size_t)ntiles_max * kAsmStride) != cudaSuccess) return -1;
  if (cudaMalloc(&out, sizeof(double) * kMaxOpt * kAsmStride) != cudaSuccess) return -1;
  if (cudaMalloc(&counter, sizeof(unsigned)) != cudaSuccess) return -1;
  cudaMemset(counter, 0, sizeof(unsigned));
  return 0;
}
void AsmWork::destroy() {
  if (partial) cudaFree(partial);
  if (out) cudaFree(out);
  if (counter) cudaFree(counter);
  partial = out = nullptr; counter = nullptr;
}

void asm_plan(AsmParams &p, int sm_count) {
  long long total = 0;
  for (int k = 0; k < p.nframes; ++k) total += p.f[k].n;
  // aim at ~4 tiles per SM, at least 2 loads per thread, tile a multiple of the block size
  long long per = (total + (long long)sm_count * 4 - 1) / ((long long)sm_count * 4);
  int tf = (int)((per + kAsmThreads - 1) / kAsmThreads) * kAsmThreads;
  if (tf < 2 * kAsmThreads) tf = 2 * kAsmThreads;
  p.tile_feats = tf;
  int t = 0;
  for (int k = 0; k < p.nframes; ++k) {
    p.f[k].tile0 = t;
    int nt = (p.f[k].n + tf - 1) / tf;
    if (nt < 1) nt = 1;  // every frame owns at least one (possibly empty) tile
    t += nt;
  }
  p.ntiles = t;
}

int asm_launch(const AsmParams &p, AsmWork &work, cudaStream_t st, int *launches) {
  if (p.ntiles > work.ntiles_max) return LIO_ERR_CAPACITY;
  if (p.nframes <= 0) return LIO_OK;
  asm_ppp<<<p.ntiles, kAsmThreads, 0, st>>>(p, work.partial, work.out, work.counter);
  if (launches) *launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { lio_set_last_error(__FILE__, __LINE__, cudaGetErrorString(e)); return LIO_ERR_CUDA; }
  return LIO_OK;
}

// ---- per-factor evaluation on the device (parity entry for the ceres::CostFunction seam) -------
__global__ void ppp_rows(const float4 *__restrict__ pts, const float4 *__restrict__ coef, int n, const double *__restrict__ Rt12,
                         const double *__restrict__ M /*6x18*/, double *__restrict__ r_out, double *__restrict__ J_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 pf = pts[i], cf = coef[i];
  const double px = pf.x, py = pf.y, pz = pf.z, wx = cf.x, wy = cf.y, wz = cf.z;
  double g[6];
  g[0] = Rt12[0] * wx + Rt12[3] * wy + Rt12[6] * wz;
  g[1] = Rt12[1] * wx + Rt12[4] * wy + Rt12[7] * wz;
  g[2] = Rt12[2] * wx + Rt12[5] * wy + Rt12[8] * wz;
  g[3] = py * g[2] - pz * g[1];
  g[4] = pz * g[0] - px * g[2];
  g[5] = px * g[1] - py * g[0];
  r_out[i] = g[0] * (px + Rt12[9]) + g[1] * (py + Rt12[10]) + g[2] * (pz + Rt12[11]) + (double)cf.w;
  for (int c = 0; c < 18; ++c) {
    double s = 0;
    for (int k = 0; k < 6; ++k) s += g[k] * M[k * 18 + c];
    J_out[(size_t)i * 18 + c] = s;
  }
}

int ppp_rows_launch(const float4 *pts, const float4 *coef, int n, const double *Rt12_dev, const double *M_dev, double *r_out,
                    double *J_out, cudaStream_t st) {
  if (n <= 0) return LIO_OK;
  ppp_rows<<<(n + 127) / 128, 128, 0, st>>>(pts, coef, n, Rt12_dev, M_dev, r_out, J_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { lio_set_last_error(__FILE__, __LINE__, cudaGetErrorString(e)); return LIO_ERR_CUDA; }
  return LIO_OK;
}

}  // namespace lio
