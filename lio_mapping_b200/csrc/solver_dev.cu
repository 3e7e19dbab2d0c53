// Device-resident Gauss-Newton / dogleg loop (stage C shell on the GPU).
//
// One CTA (1024 threads) per evaluation step turns the S blocks of the fused lidar kernel into the full normal
// equations (lidar M^T S M blocks, ImuFactors, marginalisation prior, extrinsic PriorFactor), judges the last
// candidate, and computes the next trust-region step — Ceres-1.14 TrustRegionMinimizer + TRADITIONAL_DOGLEG
// semantics (reference configuration src/imu_processor/Estimator.cc:1909-1921), identical to solver_host.cc.
// The dense Cholesky runs on the packed lower triangle in shared memory.  The host only enqueues
// [asm_ppp, k_solver_step] pairs; there is no host synchronisation inside a solve.  A frozen extrinsic keeps its
// 6 tangent slots with an identity block and zero gradient (its step is exactly zero), so n is fixed.
#include "solver_dev.cuh"
#include <algorithm>

namespace lio {
using namespace hm;

constexpr int kDsThreads = 1024;

__device__ __forceinline__ int d_off_pose(int k) { return 15 * k; }
__device__ __forceinline__ const double *x_pose(const double *x, int k) { return x + 16 * k; }
__device__ __forceinline__ const double *x_sb(const double *x, int k) { return x + 16 * k + 7; }

__device__ double block_sum(double v, double *sred /*33*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane_id() == 0) sred[warp_id()] = v;
  __syncthreads();
  if (warp_id() == 0) {
    double w = (lane_id() < (blockDim.x >> 5)) ? sred[lane_id()] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane_id() == 0) sred[32] = w;
  }
  __syncthreads();
  return sred[32];
}

// y = A x, A n x n row-major in global memory: one warp per row, coalesced.
__device__ void matvec(const double *__restrict__ A, int n, const double *__restrict__ x, double *__restrict__ y) {
  const int w = warp_id(), nw = blockDim.x >> 5, l = lane_id();
  for (int r = w; r < n; r += nw) {
    double s = 0;
    for (int c = l; c < n; c += 32) s += A[(size_t)r * n + c] * x[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (l == 0) y[r] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ double &LP(double *L, int i, int j) { return L[(size_t)i * (i + 1) / 2 + j]; }

// Cholesky of (H + mu D^2) on the packed lower triangle in shared memory, then solve for rhs (in/out, global).
// Returns 1 on success (uniform).  s_flag[0] is scratch.
__device__ int chol_solve_smem(const double *__restrict__ H, int n, double mu, const double *__restrict__ diagonal, double *L,
                               double *rhs, int *s_flag) {
  const int tid = threadIdx.x, T = blockDim.x;
  for (int p = tid; p < n * (n + 1) / 2; p += T) {
    int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= p) ++i;
    while (i * (i + 1) / 2 > p) --i;
    int j = p - i * (i + 1) / 2;
    double v = H[(size_t)i * n + j];
    if (i == j) v += mu * diagonal[i] * diagonal[i];
    L[p] = v;
  }
  if (tid == 0) s_flag[0] = 1;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    if (tid == 0) {
      double d = LP(L, k, k);
      if (!(d > 0.0) || !isfinite(d)) s_flag[0] = 0;
      else LP(L, k, k) = sqrt(d);
    }
    __syncthreads();
    if (!s_flag[0]) break;
    const double lkk = LP(L, k, k);
    for (int i = k + 1 + tid; i < n; i += T) LP(L, i, k) /= lkk;
    __syncthreads();
    // trailing update a(i,j) -= l_ik l_jk for k < j <= i: 16 threads share a row
    for (int i = k + 1 + (tid >> 4); i < n; i += (T >> 4)) {
      const double lik = LP(L, i, k);
      double *row = L + (size_t)i * (i + 1) / 2;
      for (int j = k + 1 + (tid & 15); j <= i; j += 16) row[j] -= lik * LP(L, j, k);
    }
    __syncthreads();
  }
  const int ok = s_flag[0];
  __syncthreads();
  if (!ok) return 0;
  // triangular solves by warp 0 (lane owns entries i == lane mod 32), values in registers
  if (warp_id() == 0) {
    const int l = lane_id();
    double b[(kDsMaxN + 31) / 32];
#pragma unroll
    for (int q = 0; q < (kDsMaxN + 31) / 32; ++q) { int i = l + 32 * q; b[q] = i < n ? rhs[i] : 0.0; }
    for (int k = 0; k < n; ++k) {  // L y = b
      double bk = 0.0;
#pragma unroll
      for (int q = 0; q < (kDsMaxN + 31) / 32; ++q) if (q == (k >> 5)) bk = b[q];
      double yk = bk / LP(L, k, k);
      yk = __shfl_sync(0xffffffffu, yk, k & 31);
#pragma unroll
      for (int q = 0; q < (kDsMaxN + 31) / 32; ++q) {
        int i = l + 32 * q;
        if (i == k) b[q] = yk;
        else if (i > k && i < n) b[q] -= LP(L, i, k) * yk;
      }
    }
    for (int k = n - 1; k >= 0; --k) {  // L^T x = y
      double bk = 0.0;
#pragma unroll
      for (int q = 0; q < (kDsMaxN + 31) / 32; ++q) if (q == (k >> 5)) bk = b[q];
      double xk = bk / LP(L, k, k);
      xk = __shfl_sync(0xffffffffu, xk, k & 31);
#pragma unroll
      for (int q = 0; q < (kDsMaxN + 31) / 32; ++q) {
        int i = l + 32 * q;
        if (i == k) b[q] = xk;
        else if (i < k) b[q] -= LP(L, k, i) * xk;
      }
    }
#pragma unroll
    for (int q = 0; q < (kDsMaxN + 31) / 32; ++q) { int i = l + 32 * q; if (i < n) rhs[i] = b[q]; }
  }
  __syncthreads();
  if (tid == 0) { int okv = 1; for (int i = 0; i < n; ++i) if (!isfinite(rhs[i])) okv = 0; s_flag[0] = okv; }
  __syncthreads();
  const int fin = s_flag[0];
  __syncthreads();
  return fin;
}

__device__ void pose_dx_dev(const double *x, const double *x0, double *out) {  // MarginalizationFactor::Evaluate :347-372
  for (int k = 0; k < 3; ++k) out[k] = x[k] - x0[k];
  Q q0(x0[6], x0[3], x0[4], x0[5]), q(x[6], x[3], x[4], x[5]);
  Q dq = inverse(q0) * q;
  Q dn = normalized(dq);
  const double s = dq.w < 0 ? -2.0 : 2.0;
  out[3] = s * dn.x; out[4] = s * dn.y; out[5] = s * dn.z;
}

// Normal equations at state xe into (Hd, gd); cost components into s_cost[0..3] (ppp, pim, marg, ex prior).
__device__ void build_normal(DevSolveState *S, const double *__restrict__ xe, const double *__restrict__ Sblk, const double *__restrict__ Hp,
                             double *__restrict__ Hd, double *__restrict__ gd, double *sM, double *sSM, double *scratch, double *s_cost,
                             double *sred) {
  const int tid = threadIdx.x, T = blockDim.x;
  const int O = S->O, n = S->n;
  const int oe = 15 * (O + 1);
  const bool ex_free = S->ex_free != 0;
  for (int p = tid; p < n * n; p += T) Hd[p] = 0.0;
  for (int p = tid; p < n; p += T) gd[p] = 0.0;
  if (tid < 4) s_cost[tid] = 0.0;
  // frame terms
  if (tid < O) {
    double R[9], t[3];
    ppp_frame_terms_impl(x_pose(xe, 0), x_pose(xe, tid + 1), xe + 16 * (O + 1), R, t, sM + tid * 108);
  }
  __syncthreads();
  if (S->point_distance_factor) {
    // SM_i = Sgg_i * M_i (6 x 18)
    for (int p = tid; p < O * 108; p += T) {
      const int i = p / 108, q = p - i * 108, a = q / 18, c = q - a * 18;
      const double *Sb = Sblk + i * kAsmStride;
      double s = 0;
      for (int b = 0; b < 6; ++b) {
        const int lo = a < b ? a : b, hi = a < b ? b : a;
        const int idx = lo * 7 - lo * (lo - 1) / 2 + (hi - lo);  // upper-triangle (row-major) index of (lo,hi) in 7x7
        s += Sb[idx] * sM[i * 108 + b * 18 + c];
      }
      sSM[p] = s;
    }
    __syncthreads();
    // shared blocks (pose_0 / ex rows and columns): sum over frames in frame order
    for (int p = tid; p < 144; p += T) {
      const int ra = p / 12, rb = p - ra * 12;
      const int a = ra < 6 ? ra : ra + 6, b = rb < 6 ? rb : rb + 6;  // columns 0..5 (pose_0) or 12..17 (ex) of M
      if ((!ex_free) && (ra >= 6 || rb >= 6)) continue;
      double s = 0;
      for (int i = 0; i < O; ++i) {
        double v = 0;
        for (int k = 0; k < 6; ++k) v += sM[i * 108 + k * 18 + a] * sSM[i * 108 + k * 18 + b];
        s += v;
      }
      const int ia = ra < 6 ? ra : oe + (ra - 6), ib = rb < 6 ? rb : oe + (rb - 6);
      Hd[(size_t)ia * n + ib] = s;
    }
    // frame-specific blocks
    for (int p = tid; p < O * 324; p += T) {
      const int i = p / 324, q = p - i * 324, a = q / 18, b = q - a * 18;
      const int ba = a / 6, bb = b / 6;
      if (ba != 1 && bb != 1) continue;
      if (!ex_free && (ba == 2 || bb == 2)) continue;
      double v = 0;
      for (int k = 0; k < 6; ++k) v += sM[i * 108 + k * 18 + a] * sSM[i * 108 + k * 18 + b];
      const int offs[3] = {0, d_off_pose(i + 1), oe};
      Hd[(size_t)(offs[ba] + a % 6) * n + offs[bb] + b % 6] = v;
    }
    // gradient
    for (int p = tid; p < 12 + O * 6; p += T) {
      if (p < 12) {
        if (!ex_free && p >= 6) continue;
        const int a = p < 6 ? p : p + 6;
        double s = 0;
        for (int i = 0; i < O; ++i) {
          const double *Sb = Sblk + i * kAsmStride;
          double v = 0;
          for (int k = 0; k < 6; ++k) v += sM[i * 108 + k * 18 + a] * Sb[k * 7 - k * (k - 1) / 2 + (6 - k)];
          s += v;
        }
        gd[p < 6 ? p : oe + (p - 6)] = s;
      } else {
        const int i = (p - 12) / 6, a6 = (p - 12) % 6, a = 6 + a6;
        const double *Sb = Sblk + i * kAsmStride;
        double v = 0;
        for (int k = 0; k < 6; ++k) v += sM[i * 108 + k * 18 + a] * Sb[k * 7 - k * (k - 1) / 2 + (6 - k)];
        gd[d_off_pose(i + 1) + a6] = v;
      }
    }
    if (tid == 0) { double c = 0; for (int i = 0; i < O; ++i) c += 0.5 * Sblk[i * kAsmStride + 28]; s_cost[0] = c; }
  }
  __syncthreads();
  // ---- ImuFactors: one thread per factor builds the raw (sparse) Jacobian blocks and residual into shared scratch,
  // all threads whiten with the upper-triangular sqrt_info in parallel, then J^T J / J^T r are accumulated in two
  // phases (consecutive factors overlap on one pose/speed-bias block).
  if (S->imu_factor) {
    double *sA = scratch;                 // O x 15 x 31 raw [J | r]
    double *sJ = scratch + O * 465;       // O x 15 x 31 whitened
    if (tid < O && S->pim_valid[tid]) {
      double r[15], Ji[15][6], Jsi[15][9], Jj[15][6], Jsj[15][9];
      imu_factor_eval_impl(S->pim[tid], x_pose(xe, tid), x_sb(xe, tid), x_pose(xe, tid + 1), x_sb(xe, tid + 1), r, Ji, Jsi, Jj, Jsj, false);
      double *A = sA + tid * 465;
      for (int a = 0; a < 15; ++a) {
        for (int c = 0; c < 6; ++c) { A[a * 31 + c] = Ji[a][c]; A[a * 31 + 15 + c] = Jj[a][c]; }
        for (int c = 0; c < 9; ++c) { A[a * 31 + 6 + c] = Jsi[a][c]; A[a * 31 + 21 + c] = Jsj[a][c]; }
        A[a * 31 + 30] = r[a];
      }
    }
    __syncthreads();
    for (int p = tid; p < O * 465; p += T) {
      const int i = p / 465, q = p - i * 465, a = q / 31, c = q - a * 31;
      if (!S->pim_valid[i]) continue;
      double s = 0;
      for (int k = a; k < 15; ++k) s += S->pim[i].sqrt_info[a][k] * sA[i * 465 + k * 31 + c];
      sJ[p] = s;
    }
    __syncthreads();
    for (int parity = 0; parity < 2; ++parity) {
      for (int p = tid; p < O * 930; p += T) {
        const int i = p / 930, q = p - i * 930;
        if ((i & 1) != parity || !S->pim_valid[i]) continue;
        const int base = 15 * i;
        const double *J = sJ + i * 465;
        if (q < 900) {
          const int a = q / 30, b = q - a * 30;
          double s = 0;
#pragma unroll
          for (int k = 0; k < 15; ++k) s += J[k * 31 + a] * J[k * 31 + b];
          Hd[(size_t)(base + a) * n + base + b] += s;
        } else {
          const int a = q - 900;
          double s = 0;
#pragma unroll
          for (int k = 0; k < 15; ++k) s += J[k * 31 + a] * J[k * 31 + 30];
          gd[base + a] += s;
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      double c = 0;
      for (int i = 0; i < O; ++i) if (S->pim_valid[i]) { double sq = 0; for (int k = 0; k < 15; ++k) sq += sJ[i * 465 + k * 31 + 30] * sJ[i * 465 + k * 31 + 30]; c += 0.5 * sq; }
      s_cost[1] = c;
    }
  }
  __syncthreads();
  // ---- marginalisation prior
  if (S->marginalization_factor && S->prior_valid) {
    const int np = 15 * O + 6;
    if (tid < O) {
      pose_dx_dev(x_pose(xe, tid), S->x0_pose + 7 * tid, S->dx + 15 * tid);
      for (int a = 0; a < 9; ++a) S->dx[15 * tid + 6 + a] = x_sb(xe, tid)[a] - S->x0_sb[9 * tid + a];
    } else if (tid == O) {
      pose_dx_dev(xe + 16 * (O + 1), S->x0_ex, S->dx + 15 * O);
    }
    __syncthreads();
    matvec(Hp, np, S->dx, S->Hdx);
    double part = 0;
    for (int a = tid; a < np; a += T) part += 2.0 * S->bp[a] * S->dx[a] + S->dx[a] * S->Hdx[a];
    const double tot = block_sum(part, sred);
    if (tid == 0) s_cost[2] = 0.5 * (S->c0 + tot);
    for (int p = tid; p < np * np; p += T) {
      const int a = p / np, b = p - a * np;
      const int ta = a < 15 * O ? a : (ex_free ? oe + (a - 15 * O) : -1);
      const int tb = b < 15 * O ? b : (ex_free ? oe + (b - 15 * O) : -1);
      if (ta >= 0 && tb >= 0) Hd[(size_t)ta * n + tb] += Hp[p];
    }
    for (int a = tid; a < np; a += T) {
      const int ta = a < 15 * O ? a : (ex_free ? oe + (a - 15 * O) : -1);
      if (ta >= 0) gd[ta] += S->Hdx[a] + S->bp[a];
    }
  }
  __syncthreads();
  // ---- extrinsic PriorFactor / frozen extrinsic
  if (tid == 0) {
    if (ex_free) {
      if (S->prior_factor) {
        double r[6], J[6][6];
        prior_factor_impl(V3(S->ex0_pos), Q(S->ex0_quat[3], S->ex0_quat[0], S->ex0_quat[1], S->ex0_quat[2]), xe + 16 * (O + 1), r, J);
        double c = 0;
        for (int a = 0; a < 6; ++a) {
          double gs = 0;
          for (int k = 0; k < 6; ++k) gs += J[k][a] * r[k];
          gd[oe + a] += gs;
          for (int b = 0; b < 6; ++b) { double s = 0; for (int k = 0; k < 6; ++k) s += J[k][a] * J[k][b]; Hd[(size_t)(oe + a) * n + oe + b] += s; }
          c += 0.5 * r[a] * r[a];
        }
        s_cost[3] = c;
      }
    } else {
      for (int a = 0; a < 6; ++a) { Hd[(size_t)(oe + a) * n + oe + a] = 1.0; gd[oe + a] = 0.0; }
    }
  }
  __syncthreads();
}

__device__ void plus_state(const DevSolveState *S, const double *x, const double *delta, double *out) {
  const int O = S->O;
  if ((int)threadIdx.x <= O) {
    const int k = threadIdx.x;
    pose_plus_impl(x + 16 * k, delta + 15 * k, out + 16 * k);
    for (int a = 0; a < 9; ++a) out[16 * k + 7 + a] = x[16 * k + 7 + a] + delta[15 * k + 6 + a];
  } else if ((int)threadIdx.x == O + 1) {
    if (S->ex_free) pose_plus_impl(x + 16 * (O + 1), delta + 15 * (O + 1), out + 16 * (O + 1));
    else for (int a = 0; a < 7; ++a) out[16 * (O + 1) + a] = x[16 * (O + 1) + a];
  }
  __syncthreads();
}

__device__ void write_terms(const DevSolveState *S, const double *xe, double *Rt) {
  if ((int)threadIdx.x < S->O) {
    double M[108];
    ppp_frame_terms_impl(x_pose(xe, 0), x_pose(xe, threadIdx.x + 1), xe + 16 * (S->O + 1), Rt + threadIdx.x * kAsmRtStride,
                         Rt + threadIdx.x * kAsmRtStride + 9, M);
  }
}

__global__ void k_solver_terms(DevSolveState *S, double *Rt) { write_terms(S, S->x, Rt); }

__global__ void __launch_bounds__(kDsThreads, 1)
k_solver_step(DevSolveState *S, double *H, double *Hc, const double *__restrict__ Hp, double *H0, double *g0,
              const double *__restrict__ Sblk, double *Rt, int eval_index, size_t lsize) {
  extern __shared__ double dsm[];
  __shared__ double s_cost[4], sred[33];
  __shared__ int s_flag[4];
  if (S->done) return;
  const int tid = threadIdx.x, T = blockDim.x;
  const int O = S->O, n = S->n;
  double *L = dsm;                                   // packed lower triangle, n(n+1)/2
  double *sM = dsm + lsize;                          // O x 108
  double *sSM = sM + O * 108;                        // O x 108
  const double min_diagonal = 1e-6, max_diagonal = 1e32, min_mu = 1e-8, max_mu = 1.0, mu_factor = 10.0;
  const double initial_radius = 1e4, max_radius = 1e16, min_radius = 1e-32, min_relative_decrease = 1e-3;
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const int xdim = 16 * (O + 1) + 7;

  // ---------------- evaluate ----------------
  const double *xe = eval_index == 0 ? S->x : S->cand;
  build_normal(S, xe, Sblk, Hp, Hc, S->gc, sM, sSM, L, s_cost, sred);
  if (eval_index == 0) {
    // residuals before optimisation + gates (Estimator.cc:1924-1985)
    if (tid == 0) {
      S->cost_ppp = s_cost[0]; S->cost_pim = s_cost[1]; S->cost_marg = s_cost[2];
      int turn_off = 1;
      if (S->imu_factor) turn_off = s_cost[1] > 1e3;
      S->turn_off = turn_off;
      const double ratio = s_cost[2] / (s_cost[0] + s_cost[1]);
      if (!S->convergence_flag && !turn_off && ratio <= 2 && ratio != 0) S->convergence_flag = 1;
      int changed = 0;
      if (!S->convergence_flag) {
        if (S->ex_free || S->prior_valid) changed = 1;
        S->ex_free = 0; S->prior_valid = 0;
      }
      s_flag[1] = changed;
    }
    __syncthreads();
    if (s_flag[1]) build_normal(S, xe, Sblk, Hp, Hc, S->gc, sM, sSM, L, s_cost, sred);
    // iteration zero
    for (int p = tid; p < n * n; p += T) { const double v = Hc[p]; H[p] = v; H0[p] = v; }
    for (int p = tid; p < n; p += T) { S->g[p] = S->gc[p]; g0[p] = S->gc[p]; }
    __syncthreads();
    if (tid == 0) {
      const double c = s_cost[0] + s_cost[1] + s_cost[2] + s_cost[3];
      S->x_cost = c; S->initial_cost = c; S->cost0 = c;
      S->evaluations = 1;
      S->radius = initial_radius; S->mu = min_mu; S->reuse = 0; S->invalid = 0; S->iteration = 0; S->successful = 0; S->termination = 0;
      double gm = 0;
      for (int i = 0; i < n; ++i) { S->scale[i] = 1.0 / (1.0 + sqrt(H[(size_t)i * n + i])); gm = fmax(gm, fabs(S->g[i])); }
      double xn = 0;
      const int xd = S->ex_free ? xdim : xdim - 7;
      for (int i = 0; i < xd; ++i) xn += S->x[i] * S->x[i];
      S->x_norm = sqrt(xn);
      s_flag[2] = (gm <= gradient_tolerance) || !isfinite(c);
      if (!isfinite(c)) S->termination = 2;
      else if (gm <= gradient_tolerance) S->termination = 1;
    }
    __syncthreads();
    if (s_flag[2]) { if (tid == 0) { S->done = 1; } return; }
    for (int p = tid; p < n * n; p += T) { const int i = p / n, j = p - i * n; H[p] *= S->scale[i] * S->scale[j]; }
    for (int p = tid; p < n; p += T) S->g[p] *= S->scale[p];
    __syncthreads();
  } else {
    // ---------------- judge the candidate ----------------
    if (tid == 0) {
      double cand_cost = s_cost[0] + s_cost[1] + s_cost[2] + s_cost[3];
      if (!isfinite(cand_cost)) cand_cost = 1e300;
      S->cand_cost = cand_cost;
      S->evaluations += 1;
      double sn = 0;
      const int xd = S->ex_free ? xdim : xdim - 7;
      for (int i = 0; i < xd; ++i) { const double d = S->x[i] - S->cand[i]; sn += d * d; }
      sn = sqrt(sn);
      int verdict = 0;  // 0 reject, 1 accept, 2 terminate (converged)
      const double cost_change = S->x_cost - cand_cost;
      if (sn <= parameter_tolerance * (S->x_norm + parameter_tolerance)) { verdict = 2; S->termination = 1; }
      else if (fabs(cost_change) <= function_tolerance * S->x_cost) { verdict = 2; S->termination = 1; }
      else {
        const double rd = cost_change / S->model_cost_change;
        if (rd > min_relative_decrease) {
          verdict = 1;
          S->x_cost = cand_cost;
          S->successful += 1;
          if (rd < 0.25) S->radius *= 0.5;
          if (rd > 0.75) S->radius = fmin(max_radius, fmax(S->radius, 3.0 * S->dogleg_step_norm));
          S->mu = fmax(min_mu, 2.0 * S->mu / mu_factor);
          S->reuse = 0;
        } else {
          S->radius *= 0.5;
          S->reuse = 1;
        }
      }
      s_flag[1] = verdict;
    }
    __syncthreads();
    const int verdict = s_flag[1];
    if (verdict == 2) { if (tid == 0) S->done = 1; return; }
    if (verdict == 1) {
      for (int p = tid; p < xdim; p += T) S->x[p] = S->cand[p];
      for (int p = tid; p < n * n; p += T) { const int i = p / n, j = p - i * n; H[p] = Hc[p] * S->scale[i] * S->scale[j]; }
      double gm = 0;
      for (int p = tid; p < n; p += T) { gm = fmax(gm, fabs(S->gc[p])); S->g[p] = S->gc[p] * S->scale[p]; }
      __syncthreads();
      // max-norm of the unscaled gradient and the new |x|
      if (tid == 0) {
        double m = 0;
        for (int i = 0; i < n; ++i) m = fmax(m, fabs(S->gc[i]));
        double xn = 0;
        const int xd = S->ex_free ? xdim : xdim - 7;
        for (int i = 0; i < xd; ++i) xn += S->x[i] * S->x[i];
        S->x_norm = sqrt(xn);
        s_flag[2] = m <= gradient_tolerance;
        if (s_flag[2]) { S->termination = 1; S->done = 1; }
      }
      __syncthreads();
      if (s_flag[2]) return;
    }
  }

  // ---------------- next step (loops over invalid steps without a new evaluation) ----------------
  while (true) {
    if (tid == 0) {
      int stop = 0;
      if (S->iteration >= S->max_it) { S->termination = 0; stop = 1; }
      else if (S->radius < min_radius) { S->termination = 1; stop = 1; }
      else S->iteration += 1;
      s_flag[1] = stop;
    }
    __syncthreads();
    if (s_flag[1]) { if (tid == 0) S->done = 1; return; }
    int linear_ok = 1;
    if (!S->reuse) {
      __syncthreads();
      if (tid == 0) S->reuse = 1;
      for (int i = tid; i < n; i += T) {
        const double d = sqrt(fmin(fmax(H[(size_t)i * n + i], min_diagonal), max_diagonal));
        S->diagonal[i] = d;
        S->gradient[i] = S->g[i] / d;
        S->tmp[i] = S->g[i] / (d * d);  // sg
      }
      __syncthreads();
      matvec(H, n, S->tmp, S->tmp2);
      double pa = 0, pb = 0;
      for (int i = tid; i < n; i += T) { pa += S->gradient[i] * S->gradient[i]; pb += S->tmp[i] * S->tmp2[i]; }
      const double ga = block_sum(pa, sred);
      const double gb = block_sum(pb, sred);
      if (tid == 0) S->alpha = ga / gb;
      linear_ok = 0;
      while (true) {
        __syncthreads();
        if (!(S->mu < max_mu)) break;
        for (int i = tid; i < n; i += T) S->tmp[i] = S->g[i];
        __syncthreads();
        const int ok = chol_solve_smem(H, n, S->mu, S->diagonal, L, S->tmp, s_flag);
        if (!ok) { __syncthreads(); if (tid == 0) S->mu *= mu_factor; continue; }
        for (int i = tid; i < n; i += T) S->gn[i] = -S->diagonal[i] * S->tmp[i];
        linear_ok = 1;
        break;
      }
      __syncthreads();
    }
    int step_valid = linear_ok;
    if (linear_ok) {
      double p1 = 0, p2 = 0, p3 = 0;
      for (int i = tid; i < n; i += T) { p1 += S->gradient[i] * S->gradient[i]; p2 += S->gn[i] * S->gn[i]; p3 += S->gradient[i] * S->gn[i]; }
      const double gradient_norm = sqrt(block_sum(p1, sred));
      const double gn_norm = sqrt(block_sum(p2, sred));
      const double g_dot_gn = block_sum(p3, sred);
      const double radius = S->radius, alpha = S->alpha;
      double c_grad, c_gn, dsn;
      if (gn_norm <= radius) { c_grad = 0.0; c_gn = 1.0; dsn = gn_norm; }
      else if (gradient_norm * alpha >= radius) { c_grad = -(radius / gradient_norm); c_gn = 0.0; dsn = radius; }
      else {
        const double b_dot_a = -alpha * g_dot_gn;
        const double a2 = pow(alpha * gradient_norm, 2.0);
        const double bma2 = a2 - 2 * b_dot_a + pow(gn_norm, 2);
        const double c = b_dot_a - a2;
        const double d = sqrt(c * c + bma2 * (pow(radius, 2.0) - a2));
        const double beta = (c <= 0) ? (d - c) / bma2 : (radius * radius - a2) / (d + c);
        c_grad = -alpha * (1.0 - beta); c_gn = beta; dsn = -1.0;
      }
      double pn = 0;
      for (int i = tid; i < n; i += T) {
        const double sv = c_grad * S->gradient[i] + c_gn * S->gn[i];
        pn += sv * sv;
        S->step[i] = sv / S->diagonal[i];
      }
      const double sn2 = block_sum(pn, sred);
      if (dsn < 0) dsn = sqrt(sn2);
      matvec(H, n, S->step, S->tmp2);
      double q1 = 0, q2 = 0;
      for (int i = tid; i < n; i += T) { q1 += S->step[i] * S->g[i]; q2 += S->step[i] * S->tmp2[i]; }
      const double sg = block_sum(q1, sred), shs = block_sum(q2, sred);
      const double mcc = -sg - 0.5 * shs;
      step_valid = mcc > 0.0;
      if (tid == 0) { S->model_cost_change = mcc; S->dogleg_step_norm = dsn; }
    }
    __syncthreads();
    if (!step_valid) {
      if (tid == 0) {
        S->invalid += 1;
        s_flag[1] = S->invalid >= 5;
        if (s_flag[1]) { S->termination = 2; S->done = 1; }
        S->mu *= mu_factor;
        S->reuse = 0;
      }
      __syncthreads();
      if (s_flag[1]) return;
      continue;
    }
    if (tid == 0) S->invalid = 0;
    break;
  }
  // candidate = Plus(x, step .* scale); frame terms of the candidate for the next asm_ppp launch
  for (int i = tid; i < n; i += T) S->tmp[i] = S->step[i] * S->scale[i];
  __syncthreads();
  plus_state(S, S->x, S->tmp, S->cand);
  write_terms(S, S->cand, Rt);
}

bool DevSolver::supports(int O) const {
  const int n = 15 * (O + 1) + 6;
  const size_t lsz = std::max((size_t)n * (n + 1) / 2, (size_t)O * 930);
  const size_t need = sizeof(double) * (lsz + (size_t)O * 216);
  return O <= kMaxOpt && need <= 200 * 1024;
}

int DevSolver::init(int O) {
  const int n = 15 * (O + 1) + 6, np = 15 * O + 6;
  lsize = std::max((size_t)n * (n + 1) / 2, (size_t)O * 930);  // Cholesky area, also the IMU scratch
  smem_bytes = sizeof(double) * (lsize + (size_t)O * 216);
  if (cudaMalloc(&st, sizeof(DevSolveState)) != cudaSuccess) return -1;
  if (cudaMalloc(&H, sizeof(double) * n * n) != cudaSuccess) return -1;
  if (cudaMalloc(&Hc, sizeof(double) * n * n) != cudaSuccess) return -1;
  if (cudaMalloc(&H0, sizeof(double) * n * n) != cudaSuccess) return -1;
  if (cudaMalloc(&g0, sizeof(double) * n) != cudaSuccess) return -1;
  if (cudaMalloc(&Hp, sizeof(double) * np * np) != cudaSuccess) return -1;
  if (cudaMallocHost((void **)&h_st, sizeof(DevSolveState)) != cudaSuccess) return -1;
  if (cudaMemset(st, 0, sizeof(DevSolveState)) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(k_solver_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return -1;
  return 0;
}

void DevSolver::destroy() {
  void *p[] = {st, H, Hc, H0, g0, Hp};
  for (void *q : p) if (q) cudaFree(q);
  if (h_st) cudaFreeHost(h_st);
  st = nullptr; H = Hc = H0 = g0 = Hp = nullptr; h_st = nullptr;
}

int dev_solver_terms(DevSolver &ds, double *Rt_dev, cudaStream_t st, int *launches) {
  k_solver_terms<<<1, 32, 0, st>>>(ds.st, Rt_dev);
  if (launches) *launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { lio_set_last_error(__FILE__, __LINE__, cudaGetErrorString(e)); return LIO_ERR_CUDA; }
  return LIO_OK;
}

int dev_solver_step(DevSolver &ds, const double *S_dev, double *Rt_dev, int eval_index, cudaStream_t st, int *launches) {
  k_solver_step<<<1, kDsThreads, ds.smem_bytes, st>>>(ds.st, ds.H, ds.Hc, ds.Hp, ds.H0, ds.g0, S_dev, Rt_dev, eval_index, ds.lsize);
  if (launches) *launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { lio_set_last_error(__FILE__, __LINE__, cudaGetErrorString(e)); return LIO_ERR_CUDA; }
  return LIO_OK;
}

}  // namespace lio
