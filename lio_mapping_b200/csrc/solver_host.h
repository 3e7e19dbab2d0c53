// lio_mapping_b200 — dense trust-region (traditional dogleg) step controller that drives the
// device-assembled normal equations.  It stands in for ceres::Solve as the reference configures it
// (src/imu_processor/Estimator.cc:1909-1921, :1989-1990: DENSE_SCHUR + DOGLEG, 10 iterations,
// Ceres 1.14 defaults otherwise).  The wall-clock cap of the reference is not applied.
#pragma once
#include "hostmath.h"
#include <functional>

namespace lio {

struct DoglegOptions {
  int max_num_iterations = 10;
  double initial_radius = 1e4, max_radius = 1e16, min_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  int max_consecutive_invalid_steps = 5;
};

struct DoglegSummary {
  int iterations = 0, successful_steps = 0, evaluations = 0;
  int termination = 0;  // 0 NO_CONVERGENCE (iteration cap), 1 CONVERGENCE, 2 FAILURE
  double initial_cost = 0, final_cost = 0;
};

struct DoglegProblem {
  int n = 0;  // tangent dimension
  // Evaluate cost, gradient J^T r and J^T J at the CURRENT state.  Returns false on failure.
  std::function<bool(hm::Mat &H, hm::Vec &g, double &cost)> linearize;
  std::function<void(hm::Vec &x)> get_state;                // ambient state vector
  std::function<void(const hm::Vec &x)> set_state;
  std::function<void(const hm::Vec &x, const hm::Vec &delta, hm::Vec &out)> plus;
};

void dogleg_solve(const DoglegOptions &opt, DoglegProblem &P, DoglegSummary *sum);

}  // namespace lio
