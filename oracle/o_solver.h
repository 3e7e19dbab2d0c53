// ORACLE — TEST INFRASTRUCTURE ONLY (see o_linalg.h header).
// Minimal stand-in for the slice of Ceres-Solver 1.14.0 the reference uses (docker/Dockerfile:43-46;
// Estimator.cc:1660-1664, 1747-1769, 1909-1990): ceres::Problem with pose local parameterisation,
// CauchyLoss corrector, Problem::Evaluate on residual subsets, and ceres::Solve configured as
// trust-region / TRADITIONAL_DOGLEG / DENSE_SCHUR with library defaults.
// Ceres is an un-vendored dependency: "parity unpinned" against the library itself.
#pragma once
#include "o_factors.h"
#include <map>
#include <memory>
#include <vector>

namespace orc {

struct Problem {
  struct ParamBlock { double *ptr; int size; bool pose; bool constant; int toff; int aoff; };
  struct ResBlock { std::shared_ptr<CostFunction> cost; const CauchyLoss *loss; std::vector<double *> params; bool removed; };
  std::vector<ParamBlock> params;
  std::map<double *, int> index;
  std::vector<ResBlock> res;
  void AddParameterBlock(double *p, int size, bool pose_local_parameterization);
  void SetParameterBlockConstant(double *p);
  int AddResidualBlock(std::shared_ptr<CostFunction> cost, const CauchyLoss *loss, const std::vector<double *> &params);
  void RemoveResidualBlock(int id) { res[id].removed = true; }
  // Problem::Evaluate(options with residual_blocks = ids): 1/2 sum rho(|r|^2)
  double EvaluateCost(const std::vector<int> *ids) const;
  // reduced program bookkeeping
  int Prepare();  // assigns tangent/ambient offsets to non-constant blocks, returns tangent dim
  int ambient_dim = 0, tangent_dim = 0;
  // cost (of residual blocks touching at least one free block), gradient J^T r and J^T J in the tangent space
  void Linearize(MatX *H, VecX *g, double *cost) const;
  void GetState(VecX &x) const;
  void SetState(const VecX &x);
  void Plus(const VecX &x, const VecX &delta, VecX &out) const;
};

struct SolverOptions {
  int max_num_iterations = 10;                 // Estimator.cc:1916
  double max_solver_time_in_seconds = 1e30;    // the reference's 0.10 s wall-clock cap is lifted for parity
  double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  bool jacobi_scaling = true;
  int max_num_consecutive_invalid_steps = 5;
};

struct SolverSummary {
  double initial_cost = 0, final_cost = 0;
  int num_iterations = 0;            // iterations executed (successful + unsuccessful), excluding iteration 0
  int num_successful_steps = 0;
  int num_linearizations = 0;
  int num_cost_evaluations = 0;
  int termination = 0;               // 0 NO_CONVERGENCE, 1 CONVERGENCE, 2 FAILURE
  std::vector<double> cost_trace;
  MatX H_initial;   // unscaled J^T J / J^T r at the initial point (parity aid)
  VecX g_initial;
};

void Solve(const SolverOptions &opt, Problem *problem, SolverSummary *summary);

}  // namespace orc
