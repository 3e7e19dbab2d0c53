// Device-resident trust-region (traditional dogleg) solver — see solver_dev.cu.
#pragma once
#include "assemble.cuh"
#include "factors_impl.h"

namespace lio {

constexpr int kDsMaxN = 15 * (kMaxOpt + 1) + 6;  // 261
constexpr int kDsMaxNp = 15 * kMaxOpt + 6;       // 246
constexpr int kDsXDim = 16 * (kMaxOpt + 1) + 7;

struct DevSolveState {
  // parameters: [pose_k(7) sb_k(9)] k = 0..O, then extrinsic pose(7)
  double x[kDsXDim], cand[kDsXDim];
  // flags
  int O, n, max_it;
  int imu_factor, point_distance_factor, prior_factor, marginalization_factor;
  int ex_free, prior_valid, convergence_flag, turn_off;
  int iteration, successful, evaluations, termination, done, reuse, invalid;
  // trust region
  double radius, mu, alpha, x_cost, cand_cost, model_cost_change, dogleg_step_norm, x_norm;
  double initial_cost, cost_pim, cost_ppp, cost_marg;
  double ex0_pos[3], ex0_quat[4];  // PriorFactor target (transform_lb_ at problem build), quat x y z w
  double scale[kDsMaxN], diagonal[kDsMaxN], gradient[kDsMaxN], gn[kDsMaxN], step[kDsMaxN], g[kDsMaxN], gc[kDsMaxN], tmp[kDsMaxN], tmp2[kDsMaxN];
  // marginalisation prior (canonical order [pose_0,sb_0,...,pose_{O-1},sb_{O-1},ex])
  double bp[kDsMaxNp], dx[kDsMaxNp], Hdx[kDsMaxNp], c0;
  double x0_pose[7 * kMaxOpt], x0_sb[9 * kMaxOpt], x0_ex[7];
  // IMU factors i -> i+1
  PimData pim[kMaxOpt];
  int pim_valid[kMaxOpt];
  double imu_J[kMaxOpt][15][30];
  double imu_r[kMaxOpt][15];
  // first linearisation (parity / debugging)
  double cost0;
};

struct DevSolver {
  DevSolveState *st = nullptr;     // device
  double *H = nullptr, *Hc = nullptr, *Hp = nullptr, *H0 = nullptr;  // device n x n (row-major), prior np x np
  double *g0 = nullptr;
  DevSolveState *h_st = nullptr;   // pinned staging copy (only the scalar / vector part is moved each scan)
  size_t smem_bytes = 0, lsize = 0;
  int init(int O);
  void destroy();
  bool supports(int O) const;
};

// Enqueues one evaluation step of the solver on `st` (no host synchronisation):
//   eval_index 0: build the normal equations at x from S_dev, run the convergence gates, take the first step;
//   eval_index k > 0: judge the candidate evaluated by the preceding asm_ppp launch, then take the next step.
// After each call Rt_dev holds the frame terms of the next state to evaluate.
int dev_solver_step(DevSolver &ds, const double *S_dev, double *Rt_dev, int eval_index, cudaStream_t st, int *launches);
// Frame terms of the CURRENT x into Rt_dev (before the first asm_ppp launch of a solve).
int dev_solver_terms(DevSolver &ds, double *Rt_dev, cudaStream_t st, int *launches);

}  // namespace lio
